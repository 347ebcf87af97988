// fp32 SIMT GEMM for PDM_PREC_FP32 (the parity-anchor mode: fp32 operands, fp32 FMA accumulate).
// Same problem description / epilogue as the tcgen05 bf16 kernel (gemm_tc.cu).
// 128x64 tile, BK = 16, 256 threads, 8x4 micro-tile, register-prefetched double buffering.
#include "common.cuh"

namespace pdm {

namespace {
constexpr int TM = 128, TN = 64, TK = 16;

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

struct SimtParams {
    const float* A1;
    const float* A2;
    int K1, K2, a1_bs, a2_bs;
    const float* W;
    const float* bias;
    int N, nb, Lr;
    const float* resid;
    int resid_bs;
    float* out32;
    int out32_bs;
    float* out2;
    int out2_bs;
    int gelu;
};

__global__ void __launch_bounds__(256) gemm_simt_kernel(SimtParams p) {
    __shared__ __align__(16) float As[2][TK][TM + 4];
    __shared__ __align__(16) float Bs[2][TK][TN + 4];
    const int tid = threadIdx.x;
    const int M = p.nb * p.Lr;
    const int K = p.K1 + p.K2;
    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;

    // loader mapping: A tile 128 rows x 16 k = 512 float4 -> 2 per thread; B tile 64 x 16 = 256 float4 -> 1
    const int lr = tid >> 2;         // 0..63
    const int lk = (tid & 3) * 4;    // 0,4,8,12
    const float* a_ptr1[2];
    const float* a_ptr2[2];
    bool a_ok[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int m = m0 + lr + h * 64;
        a_ok[h] = m < M;
        const int b = a_ok[h] ? m / p.Lr : 0, t = a_ok[h] ? m % p.Lr : 0;
        a_ptr1[h] = p.A1 + ((long long)b * p.a1_bs + t) * p.K1;
        a_ptr2[h] = p.A2 ? p.A2 + ((long long)b * p.a2_bs + t) * p.K2 : nullptr;
    }
    const int wn = n0 + lr;
    const bool w_ok = wn < p.N;
    const float* w_ptr = p.W + (long long)(w_ok ? wn : 0) * K;

    float4 ra[2], rb;
    auto gload = [&](int k0) {
        const int k = k0 + lk;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            ra[h] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (a_ok[h] && k < K) {
                ra[h] = k < p.K1 ? __ldg(reinterpret_cast<const float4*>(a_ptr1[h] + k))
                                 : __ldg(reinterpret_cast<const float4*>(a_ptr2[h] + (k - p.K1)));
            }
        }
        rb = make_float4(0.f, 0.f, 0.f, 0.f);
        if (w_ok && k < K) rb = __ldg(reinterpret_cast<const float4*>(w_ptr + k));
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            As[buf][lk + 0][lr + h * 64] = ra[h].x;
            As[buf][lk + 1][lr + h * 64] = ra[h].y;
            As[buf][lk + 2][lr + h * 64] = ra[h].z;
            As[buf][lk + 3][lr + h * 64] = ra[h].w;
        }
        Bs[buf][lk + 0][lr] = rb.x;
        Bs[buf][lk + 1][lr] = rb.y;
        Bs[buf][lk + 2][lr] = rb.z;
        Bs[buf][lk + 3][lr] = rb.w;
    };

    const int ty = tid >> 4, tx = tid & 15;  // 16 x 16 threads; rows ty*8.., cols tx*4..
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int nk = (K + TK - 1) / TK;
    gload(0);
    sstore(0);
    __syncthreads();
    for (int kb = 0; kb < nk; ++kb) {
        const int buf = kb & 1;
        if (kb + 1 < nk) gload((kb + 1) * TK);
#pragma unroll
        for (int k = 0; k < TK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
            const float4 b4 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (kb + 1 < nk) {
            sstore(buf ^ 1);
            __syncthreads();
        }
    }

    const int col = n0 + tx * 4;
    if (col >= p.N) return;
    float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.bias) bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + ty * 8 + i;
        if (m >= M) continue;
        const int b = m / p.Lr, t = m % p.Lr;
        float4 v = make_float4(acc[i][0] + bias4.x, acc[i][1] + bias4.y, acc[i][2] + bias4.z, acc[i][3] + bias4.w);
        if (p.gelu) {
            v.x = gelu_erf(v.x);
            v.y = gelu_erf(v.y);
            v.z = gelu_erf(v.z);
            v.w = gelu_erf(v.w);
        }
        if (p.resid) {
            const float4 r =
                *reinterpret_cast<const float4*>(p.resid + ((long long)b * p.resid_bs + t) * p.N + col);
            v.x += r.x;
            v.y += r.y;
            v.z += r.z;
            v.w += r.w;
        }
        if (p.out32) *reinterpret_cast<float4*>(p.out32 + ((long long)b * p.out32_bs + t) * p.N + col) = v;
        if (p.out2) *reinterpret_cast<float4*>(p.out2 + ((long long)b * p.out2_bs + t) * p.N + col) = v;
    }
}
}  // namespace

void gemm_simt_f32(const GemmProblem& g, cudaStream_t s) {
    PDM_REQUIRE(g.K1 % 4 == 0 && g.K2 % 4 == 0 && g.N % 4 == 0, "gemm_simt: K and N must be multiples of 4");
    PDM_REQUIRE(g.A1 && g.W32 && g.Lr > 0 && g.nb > 0, "gemm_simt: bad problem");
    SimtParams p;
    p.A1 = (const float*)g.A1;
    p.A2 = (const float*)g.A2;
    p.K1 = g.K1;
    p.K2 = g.A2 ? g.K2 : 0;
    p.a1_bs = g.a1_bs ? g.a1_bs : g.Lr;
    p.a2_bs = g.a2_bs ? g.a2_bs : g.Lr;
    p.W = g.W32;
    p.bias = g.bias;
    p.N = g.N;
    p.nb = g.nb;
    p.Lr = g.Lr;
    p.resid = g.resid;
    p.resid_bs = g.resid_bs ? g.resid_bs : g.Lr;
    p.out32 = g.out32;
    p.out32_bs = g.out32_bs ? g.out32_bs : g.Lr;
    p.out2 = (float*)g.out2;
    p.out2_bs = g.out2_bs ? g.out2_bs : g.Lr;
    p.gelu = g.gelu ? 1 : 0;
    const long long M = (long long)g.nb * g.Lr;
    dim3 grid(ceil_div(g.N, TN), (unsigned)ceil_div_ll(M, TM));
    gemm_simt_kernel<<<grid, 256, 0, s>>>(p);
    check_launch("gemm_simt");
}

}  // namespace pdm
