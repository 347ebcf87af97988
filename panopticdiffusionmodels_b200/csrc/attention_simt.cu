// fp32 SIMT flash-style attention (libs/uvit_t2i.py:70-74: fp32 SDPA, non-causal, scale 1/sqrt(64)).
// Used by PDM_PREC_FP32 (parity anchor) and as the bring-up path for bf16 inputs.
// One thread per query row (q and the output accumulator live in registers), K/V tiles staged in
// shared memory as fp32 and read with warp-broadcast loads; online softmax per 32-key tile.
#include "common.cuh"

namespace pdm {
namespace {
constexpr int HD = 64;    // head dim
constexpr int QB = 128;   // queries per block (one per thread)
constexpr int KT = 32;    // keys per tile

template <typename T>
__device__ __forceinline__ float to_f(T v);
template <>
__device__ __forceinline__ float to_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f<bf16>(bf16 v) { return __bfloat162float(v); }
template <typename T>
__device__ __forceinline__ T from_f(float v);
template <>
__device__ __forceinline__ float from_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

template <typename T>
__global__ void __launch_bounds__(QB) attention_simt_kernel(const T* __restrict__ qkv, T* __restrict__ out, int L,
                                                            int H) {
    __shared__ __align__(16) float Ks[KT][HD];
    __shared__ __align__(16) float Vs[KT][HD];
    const int D = H * HD;
    const int b = blockIdx.z, h = blockIdx.y;
    const int q_idx = blockIdx.x * QB + threadIdx.x;
    const bool q_ok = q_idx < L;
    const T* base = qkv + (long long)b * L * 3 * D;
    float q[HD], o[HD];
    {
        const T* qp = base + (long long)(q_ok ? q_idx : 0) * 3 * D + h * HD;
#pragma unroll
        for (int d = 0; d < HD; ++d) {
            q[d] = to_f<T>(qp[d]) * 0.125f;
            o[d] = 0.f;
        }
    }
    float m = -INFINITY, l = 0.f;
    for (int k0 = 0; k0 < L; k0 += KT) {
        __syncthreads();
        for (int i = threadIdx.x; i < KT * HD; i += QB) {
            const int kr = i / HD, d = i % HD;
            const int kk = k0 + kr;
            float kv = 0.f, vv = 0.f;
            if (kk < L) {
                const T* kp = base + (long long)kk * 3 * D + D + h * HD + d;
                kv = to_f<T>(kp[0]);
                vv = to_f<T>(kp[D]);
            }
            Ks[kr][d] = kv;
            Vs[kr][d] = vv;
        }
        __syncthreads();
        float sc[KT];
        float tmax = -INFINITY;
#pragma unroll
        for (int j = 0; j < KT; ++j) {
            float acc = 0.f;
#pragma unroll
            for (int d = 0; d < HD; d += 4) {
                const float4 k4 = *reinterpret_cast<const float4*>(&Ks[j][d]);
                acc = fmaf(q[d], k4.x, acc);
                acc = fmaf(q[d + 1], k4.y, acc);
                acc = fmaf(q[d + 2], k4.z, acc);
                acc = fmaf(q[d + 3], k4.w, acc);
            }
            sc[j] = (k0 + j < L) ? acc : -INFINITY;
            tmax = fmaxf(tmax, sc[j]);
        }
        const float m_new = fmaxf(m, tmax);
        const float corr = expf(m - m_new);  // m = -inf on the first tile -> 0
        l *= corr;
#pragma unroll
        for (int d = 0; d < HD; ++d) o[d] *= corr;
#pragma unroll
        for (int j = 0; j < KT; ++j) {
            const float pj = expf(sc[j] - m_new);
            l += pj;
#pragma unroll
            for (int d = 0; d < HD; d += 4) {
                const float4 v4 = *reinterpret_cast<const float4*>(&Vs[j][d]);
                o[d] = fmaf(pj, v4.x, o[d]);
                o[d + 1] = fmaf(pj, v4.y, o[d + 1]);
                o[d + 2] = fmaf(pj, v4.z, o[d + 2]);
                o[d + 3] = fmaf(pj, v4.w, o[d + 3]);
            }
        }
        m = m_new;
    }
    if (q_ok) {
        const float inv = 1.f / l;
        T* op = out + ((long long)b * L + q_idx) * D + h * HD;
#pragma unroll
        for (int d = 0; d < HD; ++d) op[d] = from_f<T>(o[d] * inv);
    }
}
}  // namespace

void attention_simt(const void* qkv, void* out, int nb, int L, int H, bool is_bf16, cudaStream_t s) {
    dim3 grid(ceil_div(L, QB), H, nb);
    if (is_bf16)
        attention_simt_kernel<bf16><<<grid, QB, 0, s>>>((const bf16*)qkv, (bf16*)out, L, H);
    else
        attention_simt_kernel<float><<<grid, QB, 0, s>>>((const float*)qkv, (float*)out, L, H);
    check_launch("attention_simt");
}

}  // namespace pdm
