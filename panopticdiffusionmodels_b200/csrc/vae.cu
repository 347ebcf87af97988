// VAE decoder of the latent-diffusion autoencoder (libs/autoencoder.py:303-410 Decoder, :446-450 FrozenAutoencoderKL.decode)
// on sm_100a: the step after the sampling loop (eval_t2i_discrete.py:74-84 decode_large_batch, utils.py:627-637).
//
//   z / scale_factor -> post_quant_conv (1x1) -> conv_in (3x3) -> mid { ResnetBlock, AttnBlock, ResnetBlock }
//   -> up levels { (num_res_blocks + 1) x ResnetBlock, nearest 2x Upsample + 3x3 conv } -> GroupNorm, swish, conv_out (3x3)
//
// Layout: activations are NHWC -- the residual stream in fp32, every convolution operand in bf16.  The 3x3 convolutions
// (all but the 4-channel conv_in and the 3-channel conv_out: > 99 % of the FLOPs) run on the tcgen05 GEMM kernel of
// gemm_tc.cu as IMPLICIT GEMMs: 128 consecutive output pixels are one accumulator tile, K = 9 taps x C_in, and the A tile of
// tap (ky, kx) is one TMA box of the activation shifted by (ky - 1, kx - 1) whose out-of-bounds part is zero-filled by the
// TMA unit (= the padding); nothing is ever unfolded in memory.  Bias, the residual add (x += conv2(...)) and the fp32 store
// are the GEMM epilogue.  1x1 convolutions (nin_shortcut, q / k / v / proj_out) are plain GEMMs on the flat [pixels, C] view.
// GroupNorm (32 groups, eps 1e-6): the fp32 partial sums per (image, 32-row slab, group) are emitted by the epilogue of the
// GEMM that produces the tensor (GemmProblem::gn_part; a stand-alone statistics pass remains for the outputs of conv_in and
// of the attention), combined in fp64 in a fixed order, and one fused normalise * gamma + beta -> swish -> bf16 pass writes
// the next convolution's operand.  "Nearest 2x upsample -> 3x3 convolution" runs as four 2x2 phase convolutions over the
// LOW-resolution operand (repack_up_phases_kernel; GemmProblem::conv_up): 16 instead of 36 MACs per output element and the
// upsampled tensor is never materialised as an operand.  conv_out (3 output channels) uses warp-level mma.sync.  The single-head 512-channel attention of the mid block
// (1024 tokens at 256 px) is per image S = q k^T (GEMM) -> row softmax -> O = P V (GEMM against V^T, which a GEMM with the
// roles of weight and activation swapped produces directly; v's bias commutes with the softmax and is folded into
// proj_out's).
#include <cstring>
#include <cstdlib>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/pdm.h"
#include "common.cuh"

namespace pdm {
namespace {

constexpr int GN_GROUPS = 32;
constexpr int GN_PPB = 256;  // pixels per GroupNorm-statistics block

// ---------------------------------------------------------------------------------------------- GroupNorm
// partial (sum, sum of squares) per (image, pixel slab, group): block = 256 threads over a slab of pixels, float4 channel
// vectors.  Every reduction runs in a FIXED order (no floating-point atomics): the decode is bit-reproducible and a sample
// does not depend on which other samples share its batch.
__global__ void __launch_bounds__(256) gn_stats_kernel(const float* __restrict__ x, float* __restrict__ part, int hw, int C,
                                                       int pix_per_block) {
    __shared__ float ts[256][2];
    const int n = blockIdx.y;
    const int c4n = C >> 2;                 // float4 columns
    const int cpg4 = (C / GN_GROUPS) >> 2;  // float4 columns per group (>= 1)
    const int col = threadIdx.x % c4n, prow = threadIdx.x / c4n, pstep = blockDim.x / c4n;
    const int p0 = blockIdx.x * pix_per_block;
    const int p1 = min(hw, p0 + pix_per_block);
    float s1 = 0.f, s2 = 0.f;
    if (prow < pstep) {
        const float4* base = reinterpret_cast<const float4*>(x + ((long long)n * hw) * C) + col;
        auto acc = [&](const float4 v) {
            s1 += (v.x + v.y) + (v.z + v.w);
            s2 = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, s2))));
        };
        int p = p0 + prow;
        for (; p + 3 * pstep < p1; p += 4 * pstep) {  // four 16-byte loads in flight; the summation order stays fixed
            const float4 v0 = __ldg(base + (long long)p * c4n), v1 = __ldg(base + (long long)(p + pstep) * c4n);
            const float4 v2 = __ldg(base + (long long)(p + 2 * pstep) * c4n), v3 = __ldg(base + (long long)(p + 3 * pstep) * c4n);
            acc(v0); acc(v1); acc(v2); acc(v3);
        }
        for (; p < p1; p += pstep) acc(__ldg(base + (long long)p * c4n));
    }
    ts[threadIdx.x][0] = s1;
    ts[threadIdx.x][1] = s2;
    __syncthreads();
    if (threadIdx.x < GN_GROUPS) {
        const int g = threadIdx.x;
        float a1 = 0.f, a2 = 0.f;
        for (int r = 0; r < pstep; ++r)
            for (int c = 0; c < cpg4; ++c) {
                const int t = r * c4n + g * cpg4 + c;
                a1 += ts[t][0];
                a2 += ts[t][1];
            }
        float* o = part + (((long long)n * gridDim.x + blockIdx.x) * GN_GROUPS + g) * 2;
        o[0] = a1;
        o[1] = a2;
    }
}
// one warp per (image, group): lane-strided fp64 partial sums, then a fixed xor tree -- the same bits on every run
__global__ void __launch_bounds__(128) gn_finalize_kernel(const float* __restrict__ part, float* __restrict__ mr, int n_images,
                                                          int nblk, double count, float eps) {
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= n_images * GN_GROUPS) return;
    const int n = i / GN_GROUPS, g = i - n * GN_GROUPS;
    double s1 = 0.0, s2 = 0.0;
    for (int b = lane; b < nblk; b += 32) {
        const float2 o = __ldg(reinterpret_cast<const float2*>(part) + ((long long)n * nblk + b) * GN_GROUPS + g);
        s1 += (double)o.x;
        s2 += (double)o.y;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if (lane == 0) {
        const double mean = s1 / count;
        double var = s2 / count - mean * mean;
        if (var < 0.0) var = 0.0;
        mr[2 * i] = (float)mean;
        mr[2 * i + 1] = (float)(1.0 / sqrt(var + (double)eps));
    }
}
__device__ __forceinline__ float swish(float v) { return __fdividef(v, 1.f + __expf(-v)); }
__device__ __forceinline__ uint2 pack4_bf16(float y0, float y1, float y2, float y3) {
    __nv_bfloat162 a = __floats2bfloat162_rn(y0, y1), b = __floats2bfloat162_rn(y2, y3);
    return make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
}
// y = GN(x) [* sigmoid(.)] -> bf16 NHWC (the next convolution's operand).  Grid (pixel slabs, images); a thread keeps ONE
// float4 channel column for the whole slab, so mean / 1/std / gamma / beta collapse into a per-thread (scale, shift) pair read
// once, and the slab is walked with 32-bit indices, four independent 16-byte loads in flight (HBM-bound: 6 B per element).
__global__ void __launch_bounds__(256) gn_apply_kernel(const float* __restrict__ x, const float* __restrict__ mr,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       bf16* __restrict__ out, int hw, int C, int act, int pix_per_block) {
    const int n = blockIdx.y;
    const int c4n = C >> 2, cpg = C / GN_GROUPS;
    const int col = threadIdx.x % c4n, prow = threadIdx.x / c4n, pstep = blockDim.x / c4n;
    if (prow >= pstep) return;
    const int g = (col * 4) / cpg;
    const float mean = mr[((long long)n * GN_GROUPS + g) * 2], rstd = mr[((long long)n * GN_GROUPS + g) * 2 + 1];
    const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma) + col), be = __ldg(reinterpret_cast<const float4*>(beta) + col);
    const float4 sc = make_float4(rstd * ga.x, rstd * ga.y, rstd * ga.z, rstd * ga.w);
    const float4 sh = make_float4(fmaf(-mean, sc.x, be.x), fmaf(-mean, sc.y, be.y), fmaf(-mean, sc.z, be.z), fmaf(-mean, sc.w, be.w));
    const int p0 = blockIdx.x * pix_per_block, p1 = min(hw, p0 + pix_per_block);
    const float4* src = reinterpret_cast<const float4*>(x) + (long long)n * hw * c4n + col;
    uint2* dst = reinterpret_cast<uint2*>(out) + (long long)n * hw * c4n + col;
    auto one = [&](const float4 v, int p) {
        float y0 = fmaf(v.x, sc.x, sh.x), y1 = fmaf(v.y, sc.y, sh.y), y2 = fmaf(v.z, sc.z, sh.z), y3 = fmaf(v.w, sc.w, sh.w);
        if (act) { y0 = swish(y0); y1 = swish(y1); y2 = swish(y2); y3 = swish(y3); }
        dst[(long long)p * c4n] = pack4_bf16(y0, y1, y2, y3);
    };
    int p = p0 + prow;
    for (; p + 3 * pstep < p1; p += 4 * pstep) {
        const float4 v0 = __ldg(src + (long long)p * c4n), v1 = __ldg(src + (long long)(p + pstep) * c4n);
        const float4 v2 = __ldg(src + (long long)(p + 2 * pstep) * c4n), v3 = __ldg(src + (long long)(p + 3 * pstep) * c4n);
        one(v0, p); one(v1, p + pstep); one(v2, p + 2 * pstep); one(v3, p + 3 * pstep);
    }
    for (; p < p1; p += pstep) one(__ldg(src + (long long)p * c4n), p);
}
// fp32 NHWC -> bf16 NHWC, optionally through a nearest-neighbour 2x upsample (F.interpolate(scale_factor=2, 'nearest')):
// a thread reads one float4 of a SOURCE pixel and writes it to the 1 (or 2 x 2) output pixels it maps to.
// Grid (source rows n * H, ceil(W * C/4 / 256)): 32-bit index arithmetic only.
__global__ void __launch_bounds__(256) convert_up_kernel(const float* __restrict__ x, bf16* __restrict__ out, int H, int W, int C,
                                                         int up) {
    const int c4n = C >> 2;
    const int e = blockIdx.y * blockDim.x + threadIdx.x;  // (w, col) within the source row
    if (e >= W * c4n) return;
    const int w = e / c4n, col = e - w * c4n;
    const int row = blockIdx.x;  // n * H + h
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + (long long)row * W * c4n + e);
    const uint2 o = pack4_bf16(v.x, v.y, v.z, v.w);
    uint2* dst = reinterpret_cast<uint2*>(out);
    if (!up) {
        dst[(long long)row * W * c4n + e] = o;
    } else {
        const int Wo = 2 * W;
        const long long r0 = ((long long)row * 2 * Wo + 2 * w) * c4n + col;  // output row 2 (n H + h) = n Ho + 2 h
        dst[r0] = o;
        dst[r0 + c4n] = o;
        dst[r0 + (long long)Wo * c4n] = o;
        dst[r0 + (long long)Wo * c4n + c4n] = o;
    }
}

// ---------------------------------------------------------------------------------------------- the two narrow convolutions
// z [n, Cz, h, w] (NCHW fp32) -> z / scale -> post_quant_conv (1x1) -> conv_in (3x3, pad 1) -> x [n, h, w, Cout] (NHWC fp32).
// A block owns CIN_PX consecutive pixels of one row: the post_quant output of their 3 x (CIN_PX + 2) neighbourhood is built
// once in shared memory (zero outside the image: the 3x3 pads ITS input), then every thread keeps 4 output channels for all
// CIN_PX pixels and walks the 9 taps with that tap's 4 x Cz weights in registers.
constexpr int CIN_PX = 8;
template <int CZ>
__global__ void __launch_bounds__(128) conv_in_kernel(const float* __restrict__ z, const float* __restrict__ pq_w,
                                                      const float* __restrict__ pq_b, const float* __restrict__ w,
                                                      const float* __restrict__ b, float* __restrict__ out, int n, int H, int W,
                                                      int Cout, float inv_scale) {
    static_assert(CZ == 4, "post_quant output is read back as float4");
    __shared__ float4 pq[3][CIN_PX + 2];
    const int segs = (W + CIN_PX - 1) / CIN_PX;
    const int seg = blockIdx.x % segs, y0 = (blockIdx.x / segs) % H, img = blockIdx.x / (segs * H);
    const int xb = seg * CIN_PX;
    if (threadIdx.x < 3 * (CIN_PX + 2)) {
        const int ky = threadIdx.x / (CIN_PX + 2), j = threadIdx.x % (CIN_PX + 2);
        const int yy = y0 + ky - 1, xx = xb + j - 1;
        float v[CZ] = {0.f, 0.f, 0.f, 0.f};
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
            float zin[CZ];
#pragma unroll
            for (int c = 0; c < CZ; ++c) zin[c] = __ldg(z + (((long long)img * CZ + c) * H + yy) * W + xx) * inv_scale;
#pragma unroll
            for (int o = 0; o < CZ; ++o) {
                float t = __ldg(pq_b + o);
#pragma unroll
                for (int c = 0; c < CZ; ++c) t = fmaf(__ldg(pq_w + o * CZ + c), zin[c], t);
                v[o] = t;
            }
        }
        pq[ky][j] = make_float4(v[0], v[1], v[2], v[3]);
    }
    __syncthreads();
    const int c4n = Cout >> 2;
    for (int col = threadIdx.x; col < c4n; col += blockDim.x) {
        float acc[CIN_PX][4];
        const float4 bias = __ldg(reinterpret_cast<const float4*>(b) + col);
#pragma unroll
        for (int px = 0; px < CIN_PX; ++px) { acc[px][0] = bias.x; acc[px][1] = bias.y; acc[px][2] = bias.z; acc[px][3] = bias.w; }
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const int ky = tap / 3, kx = tap % 3;
            float wr[4][CZ];
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int c = 0; c < CZ; ++c) wr[k][c] = __ldg(w + ((long long)(col * 4 + k) * CZ + c) * 9 + tap);  // [Cout, CZ, 3, 3]
#pragma unroll
            for (int px = 0; px < CIN_PX; ++px) {
                const float4 q = pq[ky][px + kx];
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    acc[px][k] = fmaf(wr[k][0], q.x, fmaf(wr[k][1], q.y, fmaf(wr[k][2], q.z, fmaf(wr[k][3], q.w, acc[px][k]))));
            }
        }
#pragma unroll
        for (int px = 0; px < CIN_PX; ++px)
            if (xb + px < W)
                reinterpret_cast<float4*>(out)[((long long)(img * H + y0) * W + xb + px) * c4n + col] =
                    make_float4(acc[px][0], acc[px][1], acc[px][2], acc[px][3]);
    }
}
// conv_out: bf16 NHWC [n, H, W, C] (already GroupNorm'ed + swish) -> 3x3 -> out [n, Co, H, W] (NCHW fp32), Co <= 4.
// Three output channels are far below a tcgen05 tile (N >= 16 of a 128-row tile would be 80 % padding), so this one runs on
// warp-level mma.sync m16n8k16 (bf16 in, fp32 accumulate): a warp owns 16 consecutive pixels of a row, K walks
// 9 taps x C channels.  A fragments come straight from global memory as 16-byte loads -- lane (g, t) reads channels
// [32 cb + 8 t, + 8) of pixels g and g + 8, and the K slots of two consecutive MMAs are ASSIGNED to those channels
// (slot 2t+j <-> channel 8t+4s+j, slot 2t+8+j <-> channel 8t+4s+2+j); the weights are staged in shared memory already in
// that per-lane fragment order (one 8-byte LDS per MMA).  Block = 8 warps = 4 rows x 32 pixels.
__global__ void __launch_bounds__(256) conv_out_kernel(const bf16* __restrict__ a, const float* __restrict__ w,
                                                       const float* __restrict__ b, float* __restrict__ out, int n, int H, int W,
                                                       int C, int Co) {
    extern __shared__ uint2 bfrag[];  // [9][C / 32][2][32]
    const int CB = C >> 5;
    for (int i = threadIdx.x; i < 9 * CB * 64; i += blockDim.x) {
        const int lane = i & 31, s = (i >> 5) & 1, cb = (i >> 6) % CB, tap = (i >> 6) / CB;
        const int g = lane >> 2, t = lane & 3;
        const int ch0 = cb * 32 + 8 * t + 4 * s;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (g < Co)
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = w[((long long)(g * C + ch0 + j)) * 9 + tap];  // [Co, C, 3, 3]
        bfrag[i] = pack4_bf16(v[0], v[1], v[2], v[3]);
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int xsegs = (W + 31) / 32, ysegs = (H + 3) / 4;
    const int ntiles = n * xsegs * ysegs;
    const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {  // the fragment table above is built once per block
    const int bx = tile % xsegs, by = (tile / xsegs) % ysegs, img = tile / (xsegs * ysegs);
    const int y0 = by * 4 + (warp >> 1), xb = bx * 32 + (warp & 1) * 16;
    if (y0 >= H || xb >= W) continue;
    float d[4] = {0.f, 0.f, 0.f, 0.f};
    for (int ky = 0; ky < 3; ++ky) {
        const int yy = y0 + ky - 1;
        if (yy < 0 || yy >= H) continue;  // warp-uniform
        const bf16* rowp = a + ((long long)(img * H + yy) * W) * C + 8 * t;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const int xa = xb + g + kx - 1, xc = xa + 8;
            const bool oka = xa >= 0 && xa < W, okc = xc >= 0 && xc < W;
            const uint4* pa = reinterpret_cast<const uint4*>(rowp + (long long)xa * C);
            const uint4* pc = reinterpret_cast<const uint4*>(rowp + (long long)xc * C);
            const uint2* bf = bfrag + ((ky * 3 + kx) * CB) * 64 + lane;
#pragma unroll 4
            for (int cb = 0; cb < CB; ++cb) {
                const uint4 va = oka ? __ldg(pa + cb * 4) : zero4, vc = okc ? __ldg(pc + cb * 4) : zero4;
                const uint2 b0 = bf[cb * 64], b1 = bf[cb * 64 + 32];
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                             : "r"(va.x), "r"(vc.x), "r"(va.y), "r"(vc.y), "r"(b0.x), "r"(b0.y));
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                             : "r"(va.z), "r"(vc.z), "r"(va.w), "r"(vc.w), "r"(b1.x), "r"(b1.y));
            }
        }
    }
    // d0, d1: pixel g, output channels 2t, 2t + 1;  d2, d3: pixel g + 8
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int o = 2 * t + j;
        if (o < Co) {
            const float bias = __ldg(b + o);
            float* op = out + ((long long)(img * Co + o) * H + y0) * W;
            if (xb + g < W) op[xb + g] = d[j] + bias;
            if (xb + g + 8 < W) op[xb + g + 8] = d[2 + j] + bias;
        }
    }
    }
}

// ---------------------------------------------------------------------------------------------- attention helpers
// row softmax of S [rows, L] fp32 (scaled) -> P bf16; one warp per row
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ S, bf16* __restrict__ P, int rows, int L,
                                                           float scale) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* s = S + (long long)row * L;
    float mx = -INFINITY;
    for (int i = lane; i < L; i += 32) mx = fmaxf(mx, s[i]);
#pragma unroll
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int i = lane; i < L; i += 32) sum += __expf((s[i] - mx) * scale);
#pragma unroll
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.f / sum;
    for (int i = lane; i < L; i += 32) P[(long long)row * L + i] = __float2bfloat16(__expf((s[i] - mx) * scale) * inv);
}
// conv weight (Cout, Cin, 3, 3) fp32 -> bf16 [Cout][ky][kx][Cin] (the K order of the implicit GEMM)
__global__ void repack_conv3_kernel(const float* __restrict__ w, bf16* __restrict__ out, int Cout, int Cin) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)Cout * Cin * 9) return;
    const int c = (int)(i % Cin);
    const int tap = (int)((i / Cin) % 9);
    const int o = (int)(i / ((long long)Cin * 9));
    out[i] = __float2bfloat16(w[((long long)o * Cin + c) * 9 + tap]);
}
// "nearest 2x upsample -> 3x3 convolution" = four 2x2 convolutions over the low-resolution input, one per output phase
// (a, b) = (row, column) parity: output pixel (2y + a, 2x + b) only ever sees source pixels (y + dy + a - 1, x + dx + b - 1),
// dy, dx in {0, 1}, and the 3x3 taps that land on the same source pixel add up (summed in fp32, rounded to bf16 once):
//   a = 0: dy = 0 <- ky {0},     dy = 1 <- ky {1, 2}          a = 1: dy = 0 <- ky {0, 1},  dy = 1 <- ky {2}      (same for b / kx)
// 16 instead of 36 multiply-adds per output pixel and channel pair, and the 4x larger upsampled operand is never written.
// out: [phase = 2a + b][Cout][dy][dx][Cin] bf16.
__global__ void repack_up_phases_kernel(const float* __restrict__ w, bf16* __restrict__ out, int Cout, int Cin) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long per = (long long)Cout * 4 * Cin;
    if (i >= 4 * per) return;
    const int ph = (int)(i / per), a = ph >> 1, b = ph & 1;
    const long long j = i - ph * per;
    const int c = (int)(j % Cin), tap = (int)((j / Cin) % 4), o = (int)(j / (4LL * Cin));
    const int dy = tap >> 1, dx = tap & 1;
    const int ky0 = a == 0 ? (dy == 0 ? 0 : 1) : (dy == 0 ? 0 : 2), ky1 = a == 0 ? (dy == 0 ? 0 : 2) : (dy == 0 ? 1 : 2);
    const int kx0 = b == 0 ? (dx == 0 ? 0 : 1) : (dx == 0 ? 0 : 2), kx1 = b == 0 ? (dx == 0 ? 0 : 2) : (dx == 0 ? 1 : 2);
    const float* wc = w + ((long long)o * Cin + c) * 9;  // [Cout, Cin, 3, 3]
    float acc = 0.f;
    for (int ky = ky0; ky <= ky1; ++ky)
        for (int kx = kx0; kx <= kx1; ++kx) acc += wc[ky * 3 + kx];
    out[i] = __float2bfloat16(acc);
}
// b_out[o] = b_o[o] + sum_c W_o[o, c] * b_v[c]   (v's bias commutes with the softmax: rows of P sum to one)
__global__ void fold_v_bias_kernel(const float* __restrict__ Wo, const float* __restrict__ bo, const float* __restrict__ bv,
                                   float* __restrict__ out, int C) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= C) return;
    float acc = bo[o];
    for (int c = 0; c < C; ++c) acc = fmaf(Wo[(long long)o * C + c], bv[c], acc);
    out[o] = acc;
}

struct VParam {
    std::vector<int64_t> shape;
    size_t n = 0;
    float* d32 = nullptr;
    bool set = false;
};
struct Conv3 {  // 3x3 convolution as implicit GEMM
    bf16* w16 = nullptr;  // [Cout][9][Cin]
    bf16* wup = nullptr;  // Upsample convolutions only: [4 phases][Cout][2][2][Cin] (repack_up_phases_kernel)
    const float* b = nullptr;
    int Cin = 0, Cout = 0;
};
struct Conv1 {  // 1x1 convolution as plain GEMM
    bf16* w16 = nullptr;  // [Cout][Cin]
    const float* b = nullptr;
    int Cin = 0, Cout = 0;
};
struct ResBlock {
    const float *n1w, *n1b, *n2w, *n2b;
    Conv3 c1, c2;
    Conv1 nin;
    bool has_nin = false;
    int Cin = 0, Cout = 0;
};

}  // namespace
}  // namespace pdm

using namespace pdm;

struct pdm_vae {
    pdm_vae_config cfg;
    int nlev = 0;
    std::map<std::string, VParam> params;
    std::vector<void*> owned;  // derived device buffers
    bool finalized = false;
    // development switch: PDM_VAE_NO_FUSED_GN=1 keeps the stand-alone GroupNorm statistics pass for every tensor
    bool fused_gn_stats = getenv("PDM_VAE_NO_FUSED_GN") == nullptr;
    // graph of layers
    ResBlock mid1, mid2;
    struct {
        const float *nw, *nb;
        Conv1 q, k, v, o;
        float* bo_folded = nullptr;
    } attn;
    struct Level {
        std::vector<ResBlock> blocks;
        bool has_up = false;
        Conv3 up;
    };
    std::vector<Level> up;  // index = i_level
    // workspace (grown on demand)
    size_t ws_bytes = 0;
    uint8_t* ws = nullptr;

    ~pdm_vae() {
        for (auto& kv : params)
            if (kv.second.d32) cudaFree(kv.second.d32);
        for (void* p : owned) cudaFree(p);
        if (ws) cudaFree(ws);
    }
    int ch_at(int level) const { return cfg.ch * cfg.ch_mult[level]; }

    void expect(const std::string& k, std::vector<int64_t> shape) {
        VParam p;
        p.shape = shape;
        p.n = 1;
        for (auto s : shape) p.n *= (size_t)s;
        params[k] = p;
    }
    void expect_res(const std::string& pre, int cin, int cout) {
        expect(pre + "norm1.weight", {cin});
        expect(pre + "norm1.bias", {cin});
        expect(pre + "conv1.weight", {cout, cin, 3, 3});
        expect(pre + "conv1.bias", {cout});
        expect(pre + "norm2.weight", {cout});
        expect(pre + "norm2.bias", {cout});
        expect(pre + "conv2.weight", {cout, cout, 3, 3});
        expect(pre + "conv2.bias", {cout});
        if (cin != cout) {
            expect(pre + "nin_shortcut.weight", {cout, cin, 1, 1});
            expect(pre + "nin_shortcut.bias", {cout});
        }
    }
    void declare() {
        const int zc = cfg.z_channels, ed = cfg.embed_dim;
        expect("post_quant_conv.weight", {zc, ed, 1, 1});
        expect("post_quant_conv.bias", {zc});
        int block_in = ch_at(nlev - 1);
        expect("decoder.conv_in.weight", {block_in, zc, 3, 3});
        expect("decoder.conv_in.bias", {block_in});
        expect_res("decoder.mid.block_1.", block_in, block_in);
        for (const char* n : {"q", "k", "v", "proj_out"}) {
            expect(std::string("decoder.mid.attn_1.") + n + ".weight", {block_in, block_in, 1, 1});
            expect(std::string("decoder.mid.attn_1.") + n + ".bias", {block_in});
        }
        expect("decoder.mid.attn_1.norm.weight", {block_in});
        expect("decoder.mid.attn_1.norm.bias", {block_in});
        expect_res("decoder.mid.block_2.", block_in, block_in);
        for (int lev = nlev - 1; lev >= 0; --lev) {
            const int block_out = ch_at(lev);
            for (int i = 0; i <= cfg.num_res_blocks; ++i) {
                expect_res("decoder.up." + std::to_string(lev) + ".block." + std::to_string(i) + ".", block_in, block_out);
                block_in = block_out;
            }
            if (lev != 0) {
                expect("decoder.up." + std::to_string(lev) + ".upsample.conv.weight", {block_in, block_in, 3, 3});
                expect("decoder.up." + std::to_string(lev) + ".upsample.conv.bias", {block_in});
            }
        }
        expect("decoder.norm_out.weight", {block_in});
        expect("decoder.norm_out.bias", {block_in});
        expect("decoder.conv_out.weight", {cfg.out_ch, block_in, 3, 3});
        expect("decoder.conv_out.bias", {cfg.out_ch});
    }
    static bool ignorable(const std::string& k) {
        return k.rfind("encoder.", 0) == 0 || k.rfind("quant_conv.", 0) == 0 || k.rfind("loss.", 0) == 0;
    }
    void set_param(const std::string& key, const void* dev, const int64_t* shape, int ndim, cudaStream_t s) {
        if (ignorable(key)) return;  // the sampling path only decodes
        auto it = params.find(key);
        PDM_REQUIRE(it != params.end(), "unexpected state_dict key '" + key + "'");
        VParam& p = it->second;
        bool ok = (int)p.shape.size() == ndim;
        for (int i = 0; ok && i < ndim; ++i) ok = p.shape[i] == shape[i];
        PDM_REQUIRE(ok, "size mismatch for " + key);
        if (!p.d32) PDM_CHECK_CUDA(cudaMalloc(&p.d32, p.n * sizeof(float)));
        PDM_CHECK_CUDA(cudaMemcpyAsync(p.d32, dev, p.n * sizeof(float), cudaMemcpyDeviceToDevice, s));
        p.set = true;
        finalized = false;
    }
    template <typename T>
    T* dev_alloc(size_t n) {
        void* p = nullptr;
        PDM_CHECK_CUDA(cudaMalloc(&p, n * sizeof(T)));
        owned.push_back(p);
        return (T*)p;
    }
    Conv3 conv3(const std::string& pre, int cin, int cout, cudaStream_t s) {
        Conv3 c;
        c.Cin = cin; c.Cout = cout;
        c.b = params.at(pre + "bias").d32;
        c.w16 = dev_alloc<bf16>((size_t)cout * cin * 9);
        const long long n = (long long)cout * cin * 9;
        repack_conv3_kernel<<<(unsigned)ceil_div_ll(n, 256), 256, 0, s>>>(params.at(pre + "weight").d32, c.w16, cout, cin);
        check_launch("repack_conv3");
        return c;
    }
    Conv1 conv1(const std::string& pre, int cin, int cout, cudaStream_t s) {
        Conv1 c;
        c.Cin = cin; c.Cout = cout;
        c.b = params.at(pre + "bias").d32;
        c.w16 = dev_alloc<bf16>((size_t)cout * cin);
        convert_f32_bf16(params.at(pre + "weight").d32, c.w16, (long long)cout * cin, s);
        return c;
    }
    ResBlock res(const std::string& pre, int cin, int cout, cudaStream_t s) {
        ResBlock r;
        r.Cin = cin; r.Cout = cout;
        r.n1w = params.at(pre + "norm1.weight").d32; r.n1b = params.at(pre + "norm1.bias").d32;
        r.n2w = params.at(pre + "norm2.weight").d32; r.n2b = params.at(pre + "norm2.bias").d32;
        r.c1 = conv3(pre + "conv1.", cin, cout, s);
        r.c2 = conv3(pre + "conv2.", cout, cout, s);
        r.has_nin = cin != cout;
        if (r.has_nin) r.nin = conv1(pre + "nin_shortcut.", cin, cout, s);
        return r;
    }
    void finalize(cudaStream_t s) {
        std::string missing;
        int nmiss = 0;
        for (auto& kv : params)
            if (!kv.second.set) {
                if (nmiss < 6) missing += kv.first + " ";
                ++nmiss;
            }
        PDM_REQUIRE(nmiss == 0, "missing " + std::to_string(nmiss) + " state_dict keys: " + missing);
        for (void* p : owned) cudaFree(p);
        owned.clear();
        clear_tmap_cache_all();
        int block_in = ch_at(nlev - 1);
        mid1 = res("decoder.mid.block_1.", block_in, block_in, s);
        mid2 = res("decoder.mid.block_2.", block_in, block_in, s);
        attn.nw = params.at("decoder.mid.attn_1.norm.weight").d32;
        attn.nb = params.at("decoder.mid.attn_1.norm.bias").d32;
        attn.q = conv1("decoder.mid.attn_1.q.", block_in, block_in, s);
        attn.k = conv1("decoder.mid.attn_1.k.", block_in, block_in, s);
        attn.v = conv1("decoder.mid.attn_1.v.", block_in, block_in, s);
        attn.o = conv1("decoder.mid.attn_1.proj_out.", block_in, block_in, s);
        attn.bo_folded = dev_alloc<float>(block_in);
        fold_v_bias_kernel<<<ceil_div(block_in, 128), 128, 0, s>>>(params.at("decoder.mid.attn_1.proj_out.weight").d32, attn.o.b,
                                                                  attn.v.b, attn.bo_folded, block_in);
        check_launch("fold_v_bias");
        up.assign(nlev, Level());
        for (int lev = nlev - 1; lev >= 0; --lev) {
            const int block_out = ch_at(lev);
            for (int i = 0; i <= cfg.num_res_blocks; ++i) {
                up[lev].blocks.push_back(
                    res("decoder.up." + std::to_string(lev) + ".block." + std::to_string(i) + ".", block_in, block_out, s));
                block_in = block_out;
            }
            if (lev != 0) {
                up[lev].has_up = true;
                const std::string pre = "decoder.up." + std::to_string(lev) + ".upsample.conv.";
                up[lev].up = conv3(pre, block_in, block_in, s);
                const long long nw = 16LL * block_in * block_in;
                up[lev].up.wup = dev_alloc<bf16>((size_t)nw);
                repack_up_phases_kernel<<<(unsigned)ceil_div_ll(nw, 256), 256, 0, s>>>(params.at(pre + "weight").d32, up[lev].up.wup,
                                                                                       block_in, block_in);
                check_launch("repack_up_phases");
            }
        }
        PDM_CHECK_CUDA(cudaStreamSynchronize(s));
        finalized = true;
    }
    static void clear_tmap_cache_all();

    // ------------------------------------------------------------------ workspace
    struct Buf {
        float *X, *X2, *H1, *S;   // residual stream, its alternate (shortcut / upsample output), conv1 output, attention scores
        bf16 *A16, *Q, *K, *VT, *P, *O;
        float* part;  // GroupNorm partial sums [n][slabs][32][2]
        float* mr;
        // when the GEMM that produced a tensor also emitted its GroupNorm partial sums into `part` (GemmProblem::gn_part):
        const float* stats_of = nullptr;
        int stats_nblk = 0;
    };
    size_t max_act_elems(int n, int h0) const {  // largest [pixels, C] activation of the decoder for latent side h0
        size_t mx = 0;
        int side = h0, block_in = ch_at(nlev - 1);
        for (int lev = nlev - 1; lev >= 0; --lev) {
            mx = std::max(mx, (size_t)n * side * side * std::max(block_in, ch_at(lev)));
            block_in = ch_at(lev);
            if (lev != 0) {
                side *= 2;
                mx = std::max(mx, (size_t)n * side * side * block_in);
            }
        }
        return mx;
    }
    Buf carve(int n, int h0, bool dry, size_t* total) {
        const size_t act = max_act_elems(n, h0);
        const size_t tok = (size_t)h0 * h0, C = ch_at(nlev - 1);
        size_t off = 0;
        auto take = [&](size_t bytes) -> void* {
            off = (off + 255) & ~size_t(255);
            void* p = dry ? nullptr : ws + off;
            off += bytes;
            return p;
        };
        Buf b;
        b.X = (float*)take(act * 4);
        b.X2 = (float*)take(act * 4);
        b.H1 = (float*)take(act * 4);
        b.A16 = (bf16*)take(act * 2);
        b.Q = (bf16*)take((size_t)n * tok * C * 2);
        b.K = (bf16*)take((size_t)n * tok * C * 2);
        b.O = (bf16*)take((size_t)n * tok * C * 2);
        b.VT = (bf16*)take(tok * C * 2);
        b.S = (float*)take(tok * tok * 4);
        b.P = (bf16*)take(tok * tok * 2);
        {
            int side = h0;
            for (int lev = nlev - 1; lev > 0; --lev) side *= 2;
            b.part = (float*)take((size_t)n * ceil_div(side * side, 32) * GN_GROUPS * 2 * 4);  // 32-row slabs of the GEMM epilogue
        }
        b.mr = (float*)take((size_t)n * GN_GROUPS * 2 * 4);
        *total = off + 256;
        return b;
    }
    Buf workspace(int n, int h0) {
        size_t need = 0;
        carve(n, h0, true, &need);
        if (need > ws_bytes) {
            PDM_CHECK_CUDA(cudaDeviceSynchronize());
            if (ws) cudaFree(ws);
            ws = nullptr;
            ws_bytes = 0;
            clear_tmap_cache_all();
            PDM_CHECK_CUDA(cudaMalloc(&ws, need));
            ws_bytes = need;
        }
        return carve(n, h0, false, &need);
    }

    // ------------------------------------------------------------------ layers
    void group_norm(Buf& b, const float* x, const float* gw, const float* gb, bf16* out, int n, int hw, int C, bool act,
                    cudaStream_t s) {
        PDM_REQUIRE(C % (GN_GROUPS * 4) == 0 && C <= 1024, "GroupNorm: channels must be a multiple of 128 and <= 1024");
        int nblk = b.stats_nblk;
        if (b.stats_of != x) {  // producer without the fused statistics (conv_in, the attention's proj_out)
            nblk = ceil_div(hw, GN_PPB);
            gn_stats_kernel<<<dim3(nblk, n), 256, 0, s>>>(x, b.part, hw, C, GN_PPB);
            check_launch("gn_stats");
        }
        b.stats_of = nullptr;
        gn_finalize_kernel<<<ceil_div(n * GN_GROUPS, 4), 128, 0, s>>>(b.part, b.mr, n, nblk, (double)hw * (C / GN_GROUPS), 1e-6f);
        check_launch("gn_finalize");
        const int ppb = 8 * (256 / (C / 4));  // 8 pixels per thread
        gn_apply_kernel<<<dim3(ceil_div(hw, ppb), n), 256, 0, s>>>(x, b.mr, gw, gb, out, hw, C, act ? 1 : 0, ppb);
        check_launch("gn_apply");
    }
    // the GEMM writing `out32` ([n, hw, N], whole images) also leaves the GroupNorm partial sums of its output in b.part
    void emit_gn(GemmProblem& g, Buf& b, int hw, int stride = 1, int slot0 = 0) {
        if (fused_gn_stats && g.N % (GN_GROUPS * 4) == 0 && g.N <= 1024 && hw % 128 == 0) {
            g.gn_part = b.part; g.gn_hw = hw; g.gn_stride = stride; g.gn_slot0 = slot0;
            b.stats_of = g.out32;
            b.stats_nblk = (hw / 32) * stride;
        } else if (b.stats_of == g.out32) {
            b.stats_of = nullptr;
        }
    }
    // out32 (+= if accumulate) = conv3x3(a16) + bias
    void conv3_gemm(Buf& b, const Conv3& c, const bf16* a16, float* out32, bool accumulate, int n, int H, int W, cudaStream_t s) {
        GemmProblem g;
        g.A1 = a16; g.K1 = 9 * c.Cin; g.W16 = c.w16; g.bias = c.b; g.N = c.Cout;
        g.nb = 1; g.Lr = n * H * W;
        g.out32 = out32;
        if (accumulate) g.resid = out32;
        g.conv_N = n; g.conv_H = H; g.conv_W = W; g.conv_C = c.Cin;
        emit_gn(g, b, H * W);  // every 3x3 convolution of the decoder feeds a GroupNorm (directly or through the residual sum)
        gemm_tc_bf16(g, s);
    }
    void conv1_gemm(const Conv1& c, const bf16* a16, const float* bias, float* out32, bool accumulate, bf16* out16, long long rows,
                    cudaStream_t s) {
        GemmProblem g;
        g.A1 = a16; g.K1 = c.Cin; g.W16 = c.w16; g.bias = bias; g.N = c.Cout;
        g.nb = 1; g.Lr = (int)rows;
        g.out32 = out32; g.out2 = out16;
        if (accumulate) g.resid = out32;
        gemm_tc_bf16(g, s);
    }
    void to_bf16(const float* x, bf16* out, int n, int H, int W, int C, bool upsample, cudaStream_t s) {
        convert_up_kernel<<<dim3(n * H, ceil_div(W * (C / 4), 256)), 256, 0, s>>>(x, out, H, W, C, upsample ? 1 : 0);
        check_launch("convert_up");
    }
    // x (b.X, [n, hw, Cin]) -> b.X ([n, hw, Cout])        (libs/autoencoder.py:114-134)
    void res_block(const ResBlock& r, Buf& b, int n, int H, int W, cudaStream_t s) {
        const int hw = H * W;
        group_norm(b, b.X, r.n1w, r.n1b, b.A16, n, hw, r.Cin, true, s);
        conv3_gemm(b, r.c1, b.A16, b.H1, false, n, H, W, s);
        if (r.has_nin) {  // x = nin_shortcut(x): 1x1 on the raw stream
            to_bf16(b.X, b.A16, n, H, W, r.Cin, false, s);
            conv1_gemm(r.nin, b.A16, r.nin.b, b.X2, false, nullptr, (long long)n * hw, s);
            std::swap(b.X, b.X2);
        }
        group_norm(b, b.H1, r.n2w, r.n2b, b.A16, n, hw, r.Cout, true, s);
        conv3_gemm(b, r.c2, b.A16, b.X, true, n, H, W, s);
    }
    // x += proj_out(attention(norm(x)))                    (libs/autoencoder.py:171-195)
    void attn_block(Buf& b, int n, int H, int W, cudaStream_t s) {
        const int L = H * W, C = attn.q.Cin;
        group_norm(b, b.X, attn.nw, attn.nb, b.A16, n, L, C, false, s);
        conv1_gemm(attn.q, b.A16, attn.q.b, nullptr, false, b.Q, (long long)n * L, s);
        conv1_gemm(attn.k, b.A16, attn.k.b, nullptr, false, b.K, (long long)n * L, s);
        const float scale = 1.f / sqrtf((float)C);
        for (int i = 0; i < n; ++i) {
            const bf16* hn = b.A16 + (size_t)i * L * C;
            {  // V^T [C, L] = W_v . hn^T  (roles of weight and activation swapped; the bias is folded into proj_out)
                GemmProblem g;
                g.A1 = attn.v.w16; g.K1 = C; g.W16 = hn; g.N = L; g.nb = 1; g.Lr = C; g.out2 = b.VT;
                gemm_tc_bf16(g, s);
            }
            {  // S [L, L] = q k^T
                GemmProblem g;
                g.A1 = b.Q + (size_t)i * L * C; g.K1 = C; g.W16 = b.K + (size_t)i * L * C; g.N = L; g.nb = 1; g.Lr = L;
                g.out32 = b.S;
                gemm_tc_bf16(g, s);
            }
            softmax_rows_kernel<<<ceil_div(L, 8), 256, 0, s>>>(b.S, b.P, L, L, scale);
            check_launch("softmax_rows");
            {  // O [L, C] = P V
                GemmProblem g;
                g.A1 = b.P; g.K1 = L; g.W16 = b.VT; g.N = C; g.nb = 1; g.Lr = L; g.out2 = b.O + (size_t)i * L * C;
                gemm_tc_bf16(g, s);
            }
        }
        conv1_gemm(attn.o, b.O, attn.bo_folded, b.X, true, nullptr, (long long)n * L, s);
        b.stats_of = nullptr;  // x changed through a form that does not emit GroupNorm sums
    }

    void decode(const float* z, float* out, int n, int h0, cudaStream_t s) {
        PDM_REQUIRE(finalized, "VAE parameters not finalized");
        PDM_REQUIRE(cfg.z_channels == 4 && cfg.embed_dim == 4, "VAE: z_channels = embed_dim = 4 expected");
        PDM_REQUIRE((h0 * h0) % 128 == 0 && (h0 >= 128 ? h0 % 128 == 0 : 128 % h0 == 0), "VAE: latent side must be 16, 32, 64, ...");
        Buf b = workspace(n, h0);
        int H = h0, W = h0;
        const int C0 = ch_at(nlev - 1);
        {
            conv_in_kernel<4><<<(unsigned)(n * H * ceil_div(W, CIN_PX)), 128, 0, s>>>(
                z, params.at("post_quant_conv.weight").d32, params.at("post_quant_conv.bias").d32,
                params.at("decoder.conv_in.weight").d32, params.at("decoder.conv_in.bias").d32, b.X, n, H, W, C0,
                1.f / cfg.scale_factor);
            check_launch("vae_conv_in");
        }
        res_block(mid1, b, n, H, W, s);
        attn_block(b, n, H, W, s);
        res_block(mid2, b, n, H, W, s);
        for (int lev = nlev - 1; lev >= 0; --lev) {
            for (auto& r : up[lev].blocks) res_block(r, b, n, H, W, s);
            if (up[lev].has_up) {  // nearest 2x + 3x3 conv (libs/autoencoder.py:46-50)
                const int C = up[lev].up.Cin;
                to_bf16(b.X, b.A16, n, H, W, C, false, s);
                for (int ph = 0; ph < 4; ++ph) {  // four 2x2 phase convolutions over the low-resolution operand
                    GemmProblem g;
                    g.A1 = b.A16; g.K1 = 4 * C; g.W16 = up[lev].up.wup + (size_t)ph * up[lev].up.Cout * 4 * C;
                    g.bias = up[lev].up.b; g.N = up[lev].up.Cout;
                    g.nb = 1; g.Lr = n * H * W;
                    g.out32 = b.X2;
                    g.conv_N = n; g.conv_H = H; g.conv_W = W; g.conv_C = C; g.conv_up = 1 + ph;
                    emit_gn(g, b, H * W, 4, ph);
                    gemm_tc_bf16(g, s);
                }
                H *= 2; W *= 2;
                std::swap(b.X, b.X2);
            }
        }
        const int Cl = ch_at(0);
        group_norm(b, b.X, params.at("decoder.norm_out.weight").d32, params.at("decoder.norm_out.bias").d32, b.A16, n, H * W, Cl,
                   true, s);
        PDM_REQUIRE(cfg.out_ch <= 4 && Cl % 32 == 0, "VAE: out_ch <= 4, last level channels a multiple of 32");
        conv_out_kernel<<<(unsigned)std::min(n * ceil_div(H, 4) * ceil_div(W, 32), 148 * 8), 256, 9 * (Cl / 32) * 64 * sizeof(uint2), s>>>(
            b.A16, params.at("decoder.conv_out.weight").d32, params.at("decoder.conv_out.bias").d32, out, n, H, W, Cl, cfg.out_ch);
        check_launch("vae_conv_out");
    }
};

namespace pdm {
void clear_tmap_cache();
}
void pdm_vae::clear_tmap_cache_all() { pdm::clear_tmap_cache(); }

namespace {
template <typename F>
int vguard(F&& f) {
    try {
        f();
        return 0;
    } catch (const std::exception& e) {
        set_last_error(e.what());
        return 1;
    } catch (...) {
        set_last_error("unknown error");
        return 1;
    }
}
}  // namespace

extern "C" {

int pdm_vae_create(const pdm_vae_config* c, pdm_vae_handle* out) {
    return vguard([&] {
        PDM_REQUIRE(c && out, "null argument");
        int ndev = 0;
        PDM_REQUIRE(cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0, "no CUDA device: libpdm has no CPU path");
        PDM_REQUIRE(c->num_levels >= 1 && c->num_levels <= 8 && c->ch % 32 == 0 && c->num_res_blocks >= 1, "bad VAE config");
        std::unique_ptr<pdm_vae> h(new pdm_vae());
        h->cfg = *c;
        h->nlev = c->num_levels;
        for (int i = 0; i < c->num_levels; ++i) PDM_REQUIRE((c->ch * c->ch_mult[i]) % 128 == 0, "VAE: every level needs C % 128 == 0");
        h->declare();
        *out = h.release();
    });
}
int pdm_vae_destroy(pdm_vae_handle h) {
    return vguard([&] {
        if (h) {
            cudaDeviceSynchronize();
            delete h;
        }
    });
}
int pdm_vae_set_param(pdm_vae_handle h, const char* key, const void* dev_f32, const int64_t* shape, int32_t ndim, void* stream) {
    return vguard([&] {
        PDM_REQUIRE(h && key && dev_f32 && shape, "null argument");
        h->set_param(key, dev_f32, shape, ndim, (cudaStream_t)stream);
    });
}
int pdm_vae_finalize_params(pdm_vae_handle h, void* stream) {
    return vguard([&] {
        PDM_REQUIRE(h, "null handle");
        h->finalize((cudaStream_t)stream);
    });
}
int pdm_vae_decode(pdm_vae_handle h, const float* z, float* out, int32_t n, int32_t latent_size, void* stream) {
    return vguard([&] {
        PDM_REQUIRE(h && z && out && n > 0 && latent_size > 0, "bad argument");
        h->decode(z, out, n, latent_size, (cudaStream_t)stream);
    });
}
int pdm_vae_workspace_bytes(pdm_vae_handle h, int32_t n, int32_t latent_size, size_t* bytes) {
    return vguard([&] {
        PDM_REQUIRE(h && bytes && n > 0, "bad argument");
        h->carve(n, latent_size, true, bytes);
    });
}

}  // extern "C"
