// HBM-bound kernels of the hot path: LayerNorm, token embed, head decode, concat copies,
// the fused CFG + DPM-Solver++ update (K12) and the analog-bit codec.
// All are plain vectorised SIMT kernels; their roofline is HBM bandwidth.
#include "common.cuh"

namespace pdm {

std::atomic<long long> g_launch_count{0};

// ----------------------------------------------------------------------------------------------
// LayerNorm (libs/uvit_t2i.py:142,166,328 -> nn.LayerNorm(D), eps 1e-5, affine), one warp per row.
// Algorithmic bytes: rows * D * (4 read + 2|4 write).
// ----------------------------------------------------------------------------------------------
template <int NV, typename OutT>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ b, OutT* __restrict__ out,
                                                        long long rows, int D) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float4* xr = reinterpret_cast<const float4*>(x + row * D);
    float4 v[NV];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = lane + 32 * i;  // float4 index
        if (c * 4 < D) {
            v[i] = __ldg(xr + c);
            sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        } else {
            v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / (float)D;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = lane + 32 * i;
        if (c * 4 < D) {
            const float a = v[i].x - mean, bb = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
            sq += (a * a + bb * bb) + (cc * cc + d * d);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq / (float)D + 1e-5f);
    const float4* w4 = reinterpret_cast<const float4*>(w);
    const float4* b4 = reinterpret_cast<const float4*>(b);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = lane + 32 * i;
        if (c * 4 < D) {
            const float4 ww = __ldg(w4 + c), bb = __ldg(b4 + c);
            float4 y;
            y.x = (v[i].x - mean) * rstd * ww.x + bb.x;
            y.y = (v[i].y - mean) * rstd * ww.y + bb.y;
            y.z = (v[i].z - mean) * rstd * ww.z + bb.z;
            y.w = (v[i].w - mean) * rstd * ww.w + bb.w;
            if constexpr (sizeof(OutT) == 4) {
                reinterpret_cast<float4*>(out + row * D)[c] = y;
            } else {
                __nv_bfloat162 lo = __floats2bfloat162_rn(y.x, y.y), hi = __floats2bfloat162_rn(y.z, y.w);
                uint2 pk;
                pk.x = *reinterpret_cast<uint32_t*>(&lo);
                pk.y = *reinterpret_cast<uint32_t*>(&hi);
                reinterpret_cast<uint2*>(out + row * D)[c] = pk;
            }
        }
    }
}

template <int NV>
static void launch_ln(const float* x, const float* w, const float* b, void* out, bool out_bf16, long long rows, int D,
                      cudaStream_t s) {
    const int wpb = 8;
    const unsigned grid = (unsigned)ceil_div_ll(rows, wpb);
    if (out_bf16)
        layernorm_kernel<NV, bf16><<<grid, wpb * 32, 0, s>>>(x, w, b, (bf16*)out, rows, D);
    else
        layernorm_kernel<NV, float><<<grid, wpb * 32, 0, s>>>(x, w, b, (float*)out, rows, D);
    check_launch("layernorm");
}

void layernorm(const float* x, const float* w, const float* b, void* out, bool out_bf16, long long rows, int D,
               cudaStream_t s) {
    PDM_REQUIRE(D % 4 == 0 && D <= 2048, "layernorm: D must be a multiple of 4 and <= 2048");
    const int nv = ceil_div(D, 128);
    if (nv <= 1) launch_ln<1>(x, w, b, out, out_bf16, rows, D, s);
    else if (nv <= 2) launch_ln<2>(x, w, b, out, out_bf16, rows, D, s);
    else if (nv <= 4) launch_ln<4>(x, w, b, out, out_bf16, rows, D, s);
    else if (nv <= 6) launch_ln<6>(x, w, b, out, out_bf16, rows, D, s);
    else if (nv <= 8) launch_ln<8>(x, w, b, out, out_bf16, rows, D, s);
    else launch_ln<16>(x, w, b, out, out_bf16, rows, D, s);
}

// ----------------------------------------------------------------------------------------------
__global__ void convert_kernel(const float4* __restrict__ in, uint2* __restrict__ out, long long n4) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n4; i += stride) {
        const float4 v = __ldg(in + i);
        __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&lo);
        pk.y = *reinterpret_cast<uint32_t*>(&hi);
        out[i] = pk;
    }
}
void convert_f32_bf16(const float* in, bf16* out, long long n, cudaStream_t s) {
    PDM_REQUIRE(n % 4 == 0, "convert: n must be a multiple of 4");
    const long long n4 = n / 4;
    const int grid = (int)std::min<long long>(ceil_div_ll(n4, 256), 148 * 16);
    convert_kernel<<<grid, 256, 0, s>>>((const float4*)in, (uint2*)out, n4);
    check_launch("convert");
}

// ----------------------------------------------------------------------------------------------
// Deferred LayerNorm support (bf16 mode).  LN(x).W^T + b == rstd * (bf16(x).Wf^T) + (b + W.beta), Wf = centred gamma*W:
// the GEMM that consumes LN(x) reads the RAW bf16 rows and applies mean / rstd per row in its epilogue, so the
// LayerNorm kernel (read 4 B + write 2 B per element, 2 per block) disappears; the row sums come from the epilogue of
// the GEMM that produced x.  rowstats_convert is the entry point of that chain (the embed output): one warp per row,
// float4 round i of a warp covers exactly the 128-column slice i.
// ----------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(256) rowstats_convert_kernel(const float* __restrict__ x, bf16* __restrict__ out,
                                                               float* __restrict__ stats, long long rows, int D) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float4* xr = reinterpret_cast<const float4*>(x + row * D);
    const int npart = (D + LN_PART - 1) / LN_PART;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = lane + 32 * i;
        if (i * LN_PART >= D) break;
        float s1 = 0.f, s2 = 0.f;
        if (c * 4 < D) {
            const float4 v = __ldg(xr + c);
            s1 = (v.x + v.y) + (v.z + v.w);
            s2 = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, v.w * v.w)));
            __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
            uint2 pk;
            pk.x = *reinterpret_cast<uint32_t*>(&lo);
            pk.y = *reinterpret_cast<uint32_t*>(&hi);
            reinterpret_cast<uint2*>(out + row * D)[c] = pk;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        if (lane == 0) reinterpret_cast<float2*>(stats)[row * npart + i] = make_float2(s1, s2);
    }
}
void rowstats_convert(const float* x, bf16* out, float* stats, long long rows, int D, cudaStream_t s) {
    PDM_REQUIRE(D % 4 == 0 && D <= 2048, "rowstats_convert: D must be a multiple of 4 and <= 2048");
    const int wpb = 8;
    const unsigned grid = (unsigned)ceil_div_ll(rows, wpb);
    const int nv = ceil_div(D, 128);
    if (nv <= 4) rowstats_convert_kernel<4><<<grid, wpb * 32, 0, s>>>(x, out, stats, rows, D);
    else if (nv <= 8) rowstats_convert_kernel<8><<<grid, wpb * 32, 0, s>>>(x, out, stats, rows, D);
    else rowstats_convert_kernel<16><<<grid, wpb * 32, 0, s>>>(x, out, stats, rows, D);
    check_launch("rowstats_convert");
}

// Wf[n, k] = bf16(gamma[k] * W[n, k] - m[n]),  m[n] = mean_k(gamma[k] * W[n, k]);   d[n] = bias[n] + sum_k beta[k] * W[n, k]
// Centring along K costs nothing mathematically (sum_k (x_k - mean) * const = 0) and makes the GEMM itself subtract the row
// mean: sum_k x_k Wf[n, k] = sum_k (x_k - mean) gamma_k W[n, k], so the consuming epilogue is a single FMA, rstd * acc + d.
// (bf16 rounding leaves sum_k Wf[n, k] = eps_n ~ 2^-9 |w| sqrt(K/3) instead of 0: an error of (mean/std) * 1e-3 relative to
// the output, below the bf16 rounding of the stored activation for any row whose mean is not many times its spread.)
__global__ void __launch_bounds__(128) fold_ln_kernel(const float* __restrict__ W, const float* __restrict__ bias,
                                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                                      bf16* __restrict__ Wf, float* __restrict__ d, int K, float dscale) {
    const int n = blockIdx.x;
    float sm = 0.f, sd = 0.f;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        const float w = W[(size_t)n * K + k];
        sm = fmaf(gamma[k], w, sm);
        sd = fmaf(beta[k], w, sd);
    }
    __shared__ float red[2][4];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sm += __shfl_xor_sync(0xffffffffu, sm, o);
        sd += __shfl_xor_sync(0xffffffffu, sd, o);
    }
    if ((threadIdx.x & 31) == 0) {
        red[0][threadIdx.x >> 5] = sm;
        red[1][threadIdx.x >> 5] = sd;
    }
    __syncthreads();
    const float m = ((red[0][0] + red[0][1]) + (red[0][2] + red[0][3])) / (float)K;
    for (int k = threadIdx.x; k < K; k += blockDim.x)
        Wf[(size_t)n * K + k] = __float2bfloat16_rn(gamma[k] * W[(size_t)n * K + k] - m);
    if (threadIdx.x == 0) d[n] = dscale * ((bias ? bias[n] : 0.f) + ((red[1][0] + red[1][1]) + (red[1][2] + red[1][3])));
}
void fold_ln_weight(const float* W, const float* bias, const float* gamma, const float* beta, bf16* Wf, float* d, int N,
                    int K, bool for_gelu, cudaStream_t s) {
    // for_gelu: the consuming epilogue works on x / 2 (gelu_fast2_half in gemm_tc.cu): the folded bias is stored halved
    fold_ln_kernel<<<N, 128, 0, s>>>(W, bias, gamma, beta, Wf, d, K, for_gelu ? 0.5f : 1.f);
    check_launch("fold_ln_weight");
}

// copy [nb, Lr, row_bytes] between buffers with different batch strides (two-stream concat)
__global__ void copy_rows_kernel(uint4* __restrict__ dst, long long dst_bs16, const uint4* __restrict__ src,
                                 long long src_bs16, long long per_batch16, int nb) {
    const long long total = per_batch16 * nb;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        const long long b = i / per_batch16, r = i - b * per_batch16;
        dst[b * dst_bs16 + r] = __ldg(src + b * src_bs16 + r);
    }
}
void copy_rows(void* dst, int dst_bs, const void* src, int src_bs, int Lr, int nb, int row_bytes, cudaStream_t s) {
    PDM_REQUIRE(row_bytes % 16 == 0, "copy_rows: row_bytes must be a multiple of 16");
    const long long r16 = row_bytes / 16;
    const long long per = (long long)Lr * r16;
    const int grid = (int)std::min<long long>(ceil_div_ll(per * nb, 256), 148 * 16);
    copy_rows_kernel<<<grid, 256, 0, s>>>((uint4*)dst, (long long)dst_bs * r16, (const uint4*)src,
                                          (long long)src_bs * r16, per, nb);
    check_launch("copy_rows");
}

// ----------------------------------------------------------------------------------------------
// Token embed (libs/uvit_t2i.py:382-406): time token | context tokens | image patches | mask patches,
// + positional embedding, written straight into the residual stream(s).  One block per (token, row).
// ----------------------------------------------------------------------------------------------
// Kernel A: time + context tokens (pure copy + positional add), float4 per thread.
__global__ void __launch_bounds__(128) embed_extras_kernel(EmbedArgs a) {
    const int tok = blockIdx.x;  // 0 .. T
    const int b = blockIdx.y;
    const int bi = b % a.Bx;
    float* o = a.out_x + ((long long)b * a.Lx + tok) * a.D;
    const float* pe = a.pos + (long long)tok * a.D;
    if (tok == 0) {
        const float t = a.t_dev ? a.t_dev[bi] : a.t_scalar;
        const int half = a.D / 2;
        for (int d = threadIdx.x; d < a.D; d += blockDim.x) {
            float v = 0.f;
            if (d < 2 * half) {
                const float arg = t * a.freqs[d < half ? d : d - half];
                v = d < half ? cosf(arg) : sinf(arg);
            }
            o[d] = v + pe[d];
        }
    } else {
        const float4* src = reinterpret_cast<const float4*>(a.ctxtok + ((long long)b * a.T + (tok - 1)) * a.D);
        const float4* p4 = reinterpret_cast<const float4*>(pe);
        float4* o4 = reinterpret_cast<float4*>(o);
        for (int d = threadIdx.x; d < a.D / 4; d += blockDim.x) {
            const float4 u = __ldg(src + d), q = __ldg(p4 + d);
            o4[d] = make_float4(u.x + q.x, u.y + q.y, u.z + q.z, u.w + q.w);
        }
    }
}

// Kernel B: patch embedding (Conv2d k = s = p == per-patch dot product, weight (D, C, p, p)) for the image and the
// mask stream as a small register-tiled GEMM.  One block = 64 patches x 128 channels of one stream of one sample; a warp
// owns 8 patches, a lane 4 consecutive channels (float4 row-contiguous stores of 512 B per warp).  Weights arrive
// pre-transposed ([C*p*p, D], built once at finalize) so the shared-memory tile is filled with coalesced float4 loads;
// patch pixels are broadcast reads.  128 FMAs per 12 shared loads: HBM-bound on the [tokens, D] fp32 write.
constexpr int EMB_TOK = 64, EMB_DT = 128;
template <int KK>
__global__ void __launch_bounds__(256) embed_patch_kernel(EmbedArgs a) {
    __shared__ __align__(16) float patch[EMB_TOK][KK];
    __shared__ __align__(16) float wts[KK][EMB_DT];
    const int g = a.S / a.p;
    const int P = g * g;
    const int ext = 1 + a.T;
    const int nstream = a.mask ? 2 : 1;
    const int tile = blockIdx.x, d0 = blockIdx.y * EMB_DT;
    const int is_mask = blockIdx.z % nstream, b = blockIdx.z / nstream;
    const int bi = b % a.Bx;
    const int C = is_mask ? a.Cm : a.C;
    const int pp = a.p * a.p;
    const int kk = C * pp;
    const float* src = (is_mask ? a.mask : a.img) + (long long)bi * C * a.S * a.S;
    for (int i = threadIdx.x; i < EMB_TOK * KK; i += blockDim.x) {
        const int t = i / KK, k = i % KK;
        const int pidx = tile * EMB_TOK + t;
        float v = 0.f;
        if (pidx < P && k < kk) {
            const int ph = pidx / g, pw = pidx % g;
            const int c = k / pp, r = k % pp;
            v = src[((long long)c * a.S + ph * a.p + r / a.p) * a.S + pw * a.p + r % a.p];
        }
        patch[t][k] = v;
    }
    const float* wT = is_mask ? a.wT_msk : a.wT_img;  // [kk, D]
    for (int i = threadIdx.x; i < KK * (EMB_DT / 4); i += blockDim.x) {
        const int k = i / (EMB_DT / 4), c4 = i % (EMB_DT / 4);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < kk && d0 + c4 * 4 < a.D) v = __ldg(reinterpret_cast<const float4*>(wT + (long long)k * a.D + d0) + c4);
        *reinterpret_cast<float4*>(&wts[k][c4 * 4]) = v;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t0 = warp * 8;
    float4 acc[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k4 = 0; k4 < KK / 4; ++k4) {
        float4 pk[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) pk[t] = *reinterpret_cast<const float4*>(&patch[t0 + t][k4 * 4]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float4 w4 = *reinterpret_cast<const float4*>(&wts[k4 * 4 + j][lane * 4]);
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const float pv = j == 0 ? pk[t].x : (j == 1 ? pk[t].y : (j == 2 ? pk[t].z : pk[t].w));
                acc[t].x = fmaf(pv, w4.x, acc[t].x);
                acc[t].y = fmaf(pv, w4.y, acc[t].y);
                acc[t].z = fmaf(pv, w4.z, acc[t].z);
                acc[t].w = fmaf(pv, w4.w, acc[t].w);
            }
        }
    }
    const int d = d0 + lane * 4;
    if (d >= a.D) return;
    const float4 b4 = __ldg(reinterpret_cast<const float4*>((is_mask ? a.b_msk : a.b_img) + d));
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const int pidx = tile * EMB_TOK + t0 + t;
        if (pidx >= P) break;
        const float* pe;
        float* o;
        if (is_mask) {
            pe = a.pos_m + (long long)pidx * a.D + d;
            o = a.out_m + ((long long)b * a.Lm + a.m_off + pidx) * a.D + d;
        } else {
            pe = a.pos + (long long)(ext + pidx) * a.D + d;
            o = a.out_x + ((long long)b * a.Lx + ext + pidx) * a.D + d;
        }
        const float4 q = __ldg(reinterpret_cast<const float4*>(pe));
        // (acc + bias) + pos: the reference adds the conv bias first, then the positional embedding
        *reinterpret_cast<float4*>(o) = make_float4((acc[t].x + b4.x) + q.x, (acc[t].y + b4.y) + q.y,
                                                    (acc[t].z + b4.z) + q.z, (acc[t].w + b4.w) + q.w);
    }
}

__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int R, int Cc) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)R * Cc) return;
    const int r = (int)(i / Cc), c = (int)(i % Cc);
    out[(long long)c * R + r] = in[i];
}
void transpose_f32(const float* in, float* out, int R, int Cc, cudaStream_t s) {
    transpose_kernel<<<(unsigned)ceil_div_ll((long long)R * Cc, 256), 256, 0, s>>>(in, out, R, Cc);
    check_launch("transpose");
}

// ---- bf16 engine path: the patch embedding runs on the tcgen05 GEMM kernel (gemm_tc.cu, plain fp32 form + per-row
// positional table).  Patches become bf16 operand rows in the weight's K order (c, p1, p2); every fp32 pixel is split into
// hi + lo bf16 halves ([hi | lo], K = 2 C p^2, against [W | W]) so that the solver state enters the network with 16
// mantissa bits instead of 8 -- the identity path of the residual stream carries the embedding to the decoder.
__global__ void __launch_bounds__(256) im2col_patch_kernel(const float* __restrict__ img, bf16* __restrict__ out, int Bx, int nb,
                                                           int C, int S, int p) {
    const int g = S / p, P = g * g, pp = p * p, kk = C * pp;
    const long long total = (long long)nb * P * kk;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        const int k = (int)(i % kk);
        const long long row = i / kk;
        const int pidx = (int)(row % P);
        const int b = (int)(row / P), bi = b % Bx;
        const int ph = pidx / g, pw = pidx % g;
        const int c = k / pp, r = k % pp;
        const float v = __ldg(img + (((long long)bi * C + c) * S + ph * p + r / p) * S + pw * p + r % p);
        const bf16 hi = __float2bfloat16_rn(v);
        out[row * (2 * kk) + k] = hi;
        out[row * (2 * kk) + kk + k] = __float2bfloat16_rn(v - __bfloat162float(hi));
    }
}
void im2col_patches(const float* img, bf16* out, int Bx, int nb, int C, int S, int p, cudaStream_t s) {
    const long long total = (long long)nb * (S / p) * (S / p) * C * p * p;
    const int grid = (int)std::min<long long>(ceil_div_ll(total, 256), 148 * 16);
    im2col_patch_kernel<<<grid, 256, 0, s>>>(img, out, Bx, nb, C, S, p);
    check_launch("im2col_patches");
}
// time + context tokens of the bf16 engine path: fp32 row, its bf16 copy and the per-128-column partial row sums (the
// deferred-LayerNorm producer contract of the GEMM epilogues), optionally mirrored into the mask stream's buffers (two-stream
// concat).  128 threads = 4 warps; warp w owns the 128-column slices w, w + 4, ... of the row.
__global__ void __launch_bounds__(128) embed_extras_emit_kernel(EmbedArgs a) {
    const int tok = blockIdx.x;  // 0 .. T
    const int b = blockIdx.y;
    const int bi = b % a.Bx;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int npart = (a.D + LN_PART - 1) / LN_PART;
    const long long r1 = (long long)b * a.Lx + tok, r2 = (long long)b * a.L2rows + tok;
    const float* pe = a.pos + (long long)tok * a.D;
    const float t = tok == 0 ? (a.t_dev ? a.t_dev[bi] : a.t_scalar) : 0.f;
    const int half = a.D / 2;
    for (int part = warp; part < npart; part += 4) {
        const int d = part * LN_PART + lane * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (d < a.D) {
            const float4 q = __ldg(reinterpret_cast<const float4*>(pe + d));
            if (tok == 0) {
                float e[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int dd = d + k;
                    float u = 0.f;
                    if (dd < 2 * half) {
                        const float arg = t * a.freqs[dd < half ? dd : dd - half];
                        u = dd < half ? cosf(arg) : sinf(arg);
                    }
                    e[k] = u;
                }
                v = make_float4(e[0] + q.x, e[1] + q.y, e[2] + q.z, e[3] + q.w);
            } else {
                const float4 u = __ldg(reinterpret_cast<const float4*>(a.ctxtok + ((long long)b * a.T + (tok - 1)) * a.D + d));
                v = make_float4(u.x + q.x, u.y + q.y, u.z + q.z, u.w + q.w);
            }
            __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y), h1 = __floats2bfloat162_rn(v.z, v.w);
            const uint2 hb = make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
            *reinterpret_cast<float4*>(a.out_x + r1 * a.D + d) = v;
            *reinterpret_cast<uint2*>(a.xb + r1 * a.D + d) = hb;
            if (a.out_x2) {
                *reinterpret_cast<float4*>(a.out_x2 + r2 * a.D + d) = v;
                *reinterpret_cast<uint2*>(a.xb2 + r2 * a.D + d) = hb;
            }
        }
        float s1 = (v.x + v.y) + (v.z + v.w);
        float s2 = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, v.w * v.w)));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        if (lane == 0) {
            reinterpret_cast<float2*>(a.stats)[r1 * npart + part] = make_float2(s1, s2);
            if (a.stats2) reinterpret_cast<float2*>(a.stats2)[r2 * npart + part] = make_float2(s1, s2);
        }
    }
}
void embed_extras(const EmbedArgs& a, cudaStream_t s) {
    PDM_REQUIRE(a.xb && a.stats && a.D % 4 == 0, "embed_extras: bf16 copy / row-sum destinations missing");
    embed_extras_emit_kernel<<<dim3(1 + a.T, a.nb), 128, 0, s>>>(a);
    check_launch("embed_extras");
}

void embed_tokens(const EmbedArgs& a, cudaStream_t s) {
    const int kmax = std::max(a.C, a.mask ? a.Cm : 0) * a.p * a.p;
    PDM_REQUIRE(kmax <= 64, "embed: patch dimension > 64 unsupported");
    PDM_REQUIRE(a.D % 4 == 0, "embed: D must be a multiple of 4");
    const int g = a.S / a.p;
    const int P = g * g;
    embed_extras_kernel<<<dim3(1 + a.T, a.nb), 128, 0, s>>>(a);
    check_launch("embed_extras");
    PDM_REQUIRE(a.wT_img && (!a.mask || a.wT_msk), "embed: transposed patch weights missing");
    dim3 grid(ceil_div(P, EMB_TOK), ceil_div(a.D, EMB_DT), a.nb * (a.mask ? 2 : 1));
    if (kmax <= 16)
        embed_patch_kernel<16><<<grid, 256, 0, s>>>(a);
    else if (kmax <= 32)
        embed_patch_kernel<32><<<grid, 256, 0, s>>>(a);
    else
        embed_patch_kernel<64><<<grid, 256, 0, s>>>(a);
    check_launch("embed_patch");
}

// ----------------------------------------------------------------------------------------------
// Head (libs/uvit_t2i.py:477, 499-520): final LayerNorm -> decoder_pred / decoder_pred_mask ->
// unpatchify -> 3x3 conv (+ tanh on the mask).
// Kernel 1: one warp per (row, patch, stream); kernel 2: one thread per output pixel.
// ----------------------------------------------------------------------------------------------
// One warp = HEAD_T consecutive patches of one stream of one sample.  The (normalised) token rows live in registers
// (lane owns float4 chunks lane + 32 i); every decoder weight row is read once per HEAD_T tokens, and the 32-lane partial
// sums of 8 outputs are combined with a transposing butterfly (4 + 2 + 1 + 2 shuffles per token instead of 8 x 5).
// (4 tokens per warp for D <= 512; 2 for wider models, whose rows would otherwise take > 200 registers per thread)
template <int NV>
__global__ void __launch_bounds__(256) head_token_kernel(HeadArgs a) {
    constexpr int HEAD_T = NV > 4 ? 2 : 4;
    const int g = a.S / a.p;
    const int P = g * g;
    const int gpr = (P + HEAD_T - 1) / HEAD_T;  // token groups per (stream, sample)
    const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const int nstream = (a.m && !a.gt) ? 2 : 1;
    if (warp >= a.nb * gpr * nstream) return;
    const int stream = warp / (a.nb * gpr);
    const int rem = warp - stream * a.nb * gpr;
    const int b = rem / gpr, p0 = (rem % gpr) * HEAD_T;
    const bool do_ln = stream == 0 || a.ln_m;
    const int D = a.D;
    float4 v[HEAD_T][NV];
#pragma unroll
    for (int t = 0; t < HEAD_T; ++t) {
        const int pidx = min(p0 + t, P - 1);  // a ragged last group recomputes the last patch (its store is skipped)
        const float* row = stream == 0 ? a.x + ((long long)b * a.Lx + a.x_off + pidx) * D
                                       : a.m + ((long long)b * a.Lm + a.m_off + pidx) * D;
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = lane + 32 * i;
            v[t][i] = c * 4 < D ? __ldg(reinterpret_cast<const float4*>(row) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            sum += (v[t][i].x + v[t][i].y) + (v[t][i].z + v[t][i].w);
        }
        if (do_ln) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            const float mean = sum / (float)D;
            float sq = 0.f;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const int c = lane + 32 * i;
                if (c * 4 < D) {
                    const float e0 = v[t][i].x - mean, e1 = v[t][i].y - mean, e2 = v[t][i].z - mean, e3 = v[t][i].w - mean;
                    sq += (e0 * e0 + e1 * e1) + (e2 * e2 + e3 * e3);
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
            const float rstd = rsqrtf(sq / (float)D + 1e-5f);
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const int c = lane + 32 * i;
                if (c * 4 < D) {
                    const float4 w4 = __ldg(reinterpret_cast<const float4*>(a.ln_w) + c);
                    const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.ln_b) + c);
                    v[t][i].x = (v[t][i].x - mean) * rstd * w4.x + b4.x;
                    v[t][i].y = (v[t][i].y - mean) * rstd * w4.y + b4.y;
                    v[t][i].z = (v[t][i].z - mean) * rstd * w4.z + b4.z;
                    v[t][i].w = (v[t][i].w - mean) * rstd * w4.w + b4.w;
                }
            }
        }
    }
    if (a.gt && a.m) {
        // ground-truth mode: image feature + mask feature (mask tokens normalised only in the single-stream topology)
#pragma unroll
        for (int t = 0; t < HEAD_T; ++t) {
            const int pidx = min(p0 + t, P - 1);
            const float* row = a.m + ((long long)b * a.Lm + a.m_off + pidx) * D;
            float4 u[NV];
            float sum = 0.f;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const int c = lane + 32 * i;
                u[i] = c * 4 < D ? __ldg(reinterpret_cast<const float4*>(row) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                sum += (u[i].x + u[i].y) + (u[i].z + u[i].w);
            }
            float mean = 0.f, rstd = 1.f;
            if (a.ln_m) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
                mean = sum / (float)D;
                float sq = 0.f;
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    const int c = lane + 32 * i;
                    if (c * 4 < D) {
                        const float e0 = u[i].x - mean, e1 = u[i].y - mean, e2 = u[i].z - mean, e3 = u[i].w - mean;
                        sq += (e0 * e0 + e1 * e1) + (e2 * e2 + e3 * e3);
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
                rstd = rsqrtf(sq / (float)D + 1e-5f);
            }
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const int c = lane + 32 * i;
                if (c * 4 < D) {
                    if (a.ln_m) {
                        const float4 w4 = __ldg(reinterpret_cast<const float4*>(a.ln_w) + c);
                        const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.ln_b) + c);
                        u[i].x = (u[i].x - mean) * rstd * w4.x + b4.x;
                        u[i].y = (u[i].y - mean) * rstd * w4.y + b4.y;
                        u[i].z = (u[i].z - mean) * rstd * w4.z + b4.z;
                        u[i].w = (u[i].w - mean) * rstd * w4.w + b4.w;
                    }
                    v[t][i].x += u[i].x; v[t][i].y += u[i].y; v[t][i].z += u[i].z; v[t][i].w += u[i].w;
                }
            }
        }
    }
    const int C = stream == 0 ? a.C : a.Cm;
    const int nout = a.p * a.p * C;
    const float* W = stream == 0 ? a.w_dec : a.w_decm;
    const float* bias = stream == 0 ? a.b_dec : a.b_decm;
    float* dst = (stream == 0 ? a.tmp_img : a.tmp_msk) + (long long)b * C * a.S * a.S;
    for (int o0 = 0; o0 < nout; o0 += 8) {
        float acc[HEAD_T][8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4* wr = reinterpret_cast<const float4*>(W + (long long)min(o0 + j, nout - 1) * D);
#pragma unroll
            for (int t = 0; t < HEAD_T; ++t) acc[t][j] = 0.f;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const int c = lane + 32 * i;
                if (c * 4 < D) {
                    const float4 w4 = __ldg(wr + c);
#pragma unroll
                    for (int t = 0; t < HEAD_T; ++t) {
                        acc[t][j] = fmaf(v[t][i].x, w4.x, acc[t][j]);
                        acc[t][j] = fmaf(v[t][i].y, w4.y, acc[t][j]);
                        acc[t][j] = fmaf(v[t][i].z, w4.z, acc[t][j]);
                        acc[t][j] = fmaf(v[t][i].w, w4.w, acc[t][j]);
                    }
                }
            }
        }
        // transposing butterfly over lane bits 2..0: lane l ends with output o0 + (l & 7) summed over its 8-lane group,
        // then two plain exchanges over bits 3 and 4 complete the sum (every lane of a residue class holds the total)
        const int oo = o0 + (lane & 7);
#pragma unroll
        for (int t = 0; t < HEAD_T; ++t) {
#pragma unroll
            for (int off = 4; off >= 1; off >>= 1) {
                const bool up = (lane & off) != 0;
#pragma unroll
                for (int i = 0; i < off; ++i) {
                    const float send = up ? acc[t][i] : acc[t][i + off];
                    const float recv = __shfl_xor_sync(0xffffffffu, send, off);
                    acc[t][i] = (up ? acc[t][i + off] : acc[t][i]) + recv;
                }
            }
            float r = acc[t][0];
            r += __shfl_xor_sync(0xffffffffu, r, 8);
            r += __shfl_xor_sync(0xffffffffu, r, 16);
            const int pidx = p0 + t;
            if (lane < 8 && oo < nout && pidx < P) {
                // feature oo = (p1 * p + p2) * C + c  ->  pixel (ph*p + p1, pw*p + p2), channel c
                const int ph = pidx / g, pw = pidx % g;
                const int c = oo % C, pq = oo / C;
                const int p1 = pq / a.p, p2 = pq % a.p;
                dst[((long long)c * a.S + ph * a.p + p1) * a.S + pw * a.p + p2] = r + bias[oo];
            }
        }
    }
}

__global__ void __launch_bounds__(256) conv3x3_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                      const float* __restrict__ bias, float* __restrict__ out,
                                                      int nb, int C, int S, int do_tanh) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)nb * C * S * S;
    if (idx >= total) return;
    const int x = (int)(idx % S);
    const int y = (int)((idx / S) % S);
    const int co = (int)((idx / ((long long)S * S)) % C);
    const int b = (int)(idx / ((long long)S * S * C));
    float acc = bias[co];
    for (int ci = 0; ci < C; ++ci) {
        const float* ip = in + ((long long)b * C + ci) * S * S;
        const float* wp = w + ((long long)co * C + ci) * 9;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int yy = y + ky - 1;
            if (yy < 0 || yy >= S) continue;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int xx = x + kx - 1;
                if (xx < 0 || xx >= S) continue;
                acc = fmaf(ip[(long long)yy * S + xx], wp[ky * 3 + kx], acc);
            }
        }
    }
    out[idx] = do_tanh ? tanhf(acc) : acc;
}

// 3x3 head on TOKEN-MAJOR decoder outputs (bf16 engine path: decoder_pred runs on the GEMM kernel and leaves
// tok [nb * P, p * p * C]; feature (p1 * p + p2) * C + c of patch (ph, pw) is pixel (ph p + p1, pw p + p2) of channel c,
// libs/uvit_t2i.py:50) -- unpatchify is the gather of this kernel.  One thread = one pixel, ALL C output channels: the 9
// neighbours are 9 contiguous C-float vectors (float4 loads), 9 C^2 FMAs against weights staged in shared memory as
// [tap][ci][co]; stores are coalesced along x per output channel.
template <int C>
__global__ void __launch_bounds__(256) conv3x3_tok_kernel(const float* __restrict__ tok, const float* __restrict__ w,
                                                          const float* __restrict__ bias, float* __restrict__ out, int nb, int S,
                                                          int p, int do_tanh) {
    __shared__ __align__(16) float wsm[9 * C * C];
    for (int i = threadIdx.x; i < 9 * C * C; i += blockDim.x) {
        const int co = i % C, ci = (i / C) % C, tap = i / (C * C);
        wsm[i] = w[((long long)co * C + ci) * 9 + tap];  // [co][ci][3][3] -> [tap][ci][co]
    }
    __syncthreads();
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)nb * S * S) return;
    const int x = (int)(idx % S), y = (int)((idx / S) % S);
    const long long b = idx / ((long long)S * S);
    const int g = S / p, F = p * p * C;
    float acc[C];
#pragma unroll
    for (int co = 0; co < C; ++co) acc[co] = __ldg(bias + co);
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        const int yy = y + ky - 1;
        if (yy < 0 || yy >= S) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const int xx = x + kx - 1;
            if (xx < 0 || xx >= S) continue;
            const float4* row = reinterpret_cast<const float4*>(tok + (b * g * g + (yy / p) * g + xx / p) * F + ((yy % p) * p + xx % p) * C);
            const float* wt = wsm + (ky * 3 + kx) * C * C;
#pragma unroll
            for (int c4 = 0; c4 < C / 4; ++c4) {
                const float4 v = __ldg(row + c4);
                const float in[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int co = 0; co < C; ++co) acc[co] = fmaf(in[q], wt[(c4 * 4 + q) * C + co], acc[co]);
            }
        }
    }
#pragma unroll
    for (int co = 0; co < C; ++co) out[((b * C + co) * S + y) * S + x] = do_tanh ? tanhf(acc[co]) : acc[co];
}
void conv3x3_tokens(const float* tok, const float* w, const float* bias, float* out, int nb, int C, int S, int p, int do_tanh,
                    cudaStream_t s) {
    const long long total = (long long)nb * S * S;
    const unsigned grid = (unsigned)ceil_div_ll(total, 256);
    if (C == 4) conv3x3_tok_kernel<4><<<grid, 256, 0, s>>>(tok, w, bias, out, nb, S, p, do_tanh);
    else if (C == 8) conv3x3_tok_kernel<8><<<grid, 256, 0, s>>>(tok, w, bias, out, nb, S, p, do_tanh);
    else PDM_REQUIRE(false, "conv3x3_tokens: 4 or 8 channels");
    check_launch("conv3x3_tok");
}

void head_decode(const HeadArgs& a, cudaStream_t s) {
    const int g = a.S / a.p;
    const int P = g * g;
    const int head_t = ceil_div(a.D, 128) > 4 ? 2 : 4;  // = HEAD_T of the instantiation chosen below
    const int nwarps = a.nb * ceil_div(P, head_t) * ((a.m && !a.gt) ? 2 : 1);
    PDM_REQUIRE(a.D % 4 == 0 && a.D <= 1024, "head: D must be a multiple of 4 and <= 1024");
    const int nv = ceil_div(a.D, 128);
    const int hgrid = ceil_div(nwarps, 8);
    if (nv <= 1) head_token_kernel<1><<<hgrid, 256, 0, s>>>(a);
    else if (nv <= 2) head_token_kernel<2><<<hgrid, 256, 0, s>>>(a);
    else if (nv <= 4) head_token_kernel<4><<<hgrid, 256, 0, s>>>(a);
    else if (nv <= 6) head_token_kernel<6><<<hgrid, 256, 0, s>>>(a);
    else head_token_kernel<8><<<hgrid, 256, 0, s>>>(a);
    check_launch("head_token");
    {
        const long long total = (long long)a.nb * a.C * a.S * a.S;
        conv3x3_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, s>>>(a.tmp_img, a.w_fin, a.b_fin, a.out_img, a.nb,
                                                                        a.C, a.S, 0);
        check_launch("conv3x3_img");
    }
    if (a.m && !a.gt) {
        const long long total = (long long)a.nb * a.Cm * a.S * a.S;
        conv3x3_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, s>>>(a.tmp_msk, a.w_finm, a.b_finm, a.out_msk,
                                                                        a.nb, a.Cm, a.S, 1);
        check_launch("conv3x3_msk");
    }
}

// ----------------------------------------------------------------------------------------------
// K12: classifier-free guidance on eps and on the mask prediction (train_t2i_discrete.py:429-431),
// eps -> x0 (dpm_solver_pp.py:316) and one DPM-Solver++ singlestep linear update for both streams
// (dpm_solver_pp.py:444-456, 529-555, 724-764).  Explicit round-to-nearest intrinsics keep the
// reference's operation order (no FMA contraction), so with bit-identical scalars the result is
// bit-identical to the reference arithmetic.
//   X   = (x_in - sigma * eps) / alpha,   eps = c + s * (c - u)
//   out = A * x_base + B * X0 [+ C * (X - X0)]          (signs folded into B, C)
// Algorithmic bytes per element: image reads c,u,x_in,(x_base,X0) writes (X0),out; mask likewise.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ float lin_update(float xb, float x0, float xj, float A, float B, float C, int has_c) {
    float r = __fadd_rn(__fmul_rn(A, xb), __fmul_rn(B, x0));
    if (has_c) r = __fadd_rn(r, __fmul_rn(C, __fsub_rn(xj, x0)));
    return r;
}

__global__ void __launch_bounds__(256) update_kernel(UpdateArgs a) {
    const long long n4i = a.n_img >> 2, n4m = a.n_mask >> 2;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n4i + n4m; i += stride) {
        if (i < n4i) {
            const float4 c = __ldg(reinterpret_cast<const float4*>(a.eps_c) + i);
            float4 e = c;
            if (a.eps_u) {
                const float4 u = __ldg(reinterpret_cast<const float4*>(a.eps_u) + i);
                e.x = __fadd_rn(c.x, __fmul_rn(a.scale, __fsub_rn(c.x, u.x)));
                e.y = __fadd_rn(c.y, __fmul_rn(a.scale, __fsub_rn(c.y, u.y)));
                e.z = __fadd_rn(c.z, __fmul_rn(a.scale, __fsub_rn(c.z, u.z)));
                e.w = __fadd_rn(c.w, __fmul_rn(a.scale, __fsub_rn(c.w, u.w)));
            }
            const float4 xi = __ldg(reinterpret_cast<const float4*>(a.x_in) + i);
            float4 X;
            X.x = __fdiv_rn(__fsub_rn(xi.x, __fmul_rn(a.sigma, e.x)), a.alpha);
            X.y = __fdiv_rn(__fsub_rn(xi.y, __fmul_rn(a.sigma, e.y)), a.alpha);
            X.z = __fdiv_rn(__fsub_rn(xi.z, __fmul_rn(a.sigma, e.z)), a.alpha);
            X.w = __fdiv_rn(__fsub_rn(xi.w, __fmul_rn(a.sigma, e.w)), a.alpha);
            float4 x0, xb;
            if (a.stage == 0) {
                x0 = X;
                xb = xi;  // x_in == x_base at stage 0
                reinterpret_cast<float4*>(a.X0)[i] = X;
            } else {
                x0 = reinterpret_cast<const float4*>(a.X0)[i];
                xb = __ldg(reinterpret_cast<const float4*>(a.x_base) + i);
            }
            float4 o;
            o.x = lin_update(xb.x, x0.x, X.x, a.A, a.B_img, a.C_img, a.has_c);
            o.y = lin_update(xb.y, x0.y, X.y, a.A, a.B_img, a.C_img, a.has_c);
            o.z = lin_update(xb.z, x0.z, X.z, a.A, a.B_img, a.C_img, a.has_c);
            o.w = lin_update(xb.w, x0.w, X.w, a.A, a.B_img, a.C_img, a.has_c);
            reinterpret_cast<float4*>(a.x_out)[i] = o;
        } else {
            const long long j = i - n4i;
            const float4 c = __ldg(reinterpret_cast<const float4*>(a.pm_c) + j);
            float4 P = c;
            if (a.pm_u) {
                const float4 u = __ldg(reinterpret_cast<const float4*>(a.pm_u) + j);
                P.x = __fadd_rn(c.x, __fmul_rn(a.scale, __fsub_rn(c.x, u.x)));
                P.y = __fadd_rn(c.y, __fmul_rn(a.scale, __fsub_rn(c.y, u.y)));
                P.z = __fadd_rn(c.z, __fmul_rn(a.scale, __fsub_rn(c.z, u.z)));
                P.w = __fadd_rn(c.w, __fmul_rn(a.scale, __fsub_rn(c.w, u.w)));
            }
            float4 p0;
            if (a.stage == 0) {
                p0 = P;
                reinterpret_cast<float4*>(a.P0)[j] = P;
            } else {
                p0 = reinterpret_cast<const float4*>(a.P0)[j];
            }
            const float4 mb = __ldg(reinterpret_cast<const float4*>(a.m_base) + j);
            // pass-through stream (enable_mask_opt=False): its own A, no difference term
            const float Am = a.mask_plain ? a.A_msk : a.A;
            const int hc = a.mask_plain ? 0 : a.has_c;
            float4 o;
            o.x = lin_update(mb.x, p0.x, P.x, Am, a.B_msk, a.C_msk, hc);
            o.y = lin_update(mb.y, p0.y, P.y, Am, a.B_msk, a.C_msk, hc);
            o.z = lin_update(mb.z, p0.z, P.z, Am, a.B_msk, a.C_msk, hc);
            o.w = lin_update(mb.w, p0.w, P.w, Am, a.B_msk, a.C_msk, hc);
            reinterpret_cast<float4*>(a.m_out)[j] = o;
        }
    }
}

void cfg_solver_update(const UpdateArgs& a, cudaStream_t s) {
    PDM_REQUIRE(a.n_img % 4 == 0 && a.n_mask % 4 == 0, "cfg_update: element counts must be multiples of 4");
    PDM_REQUIRE(a.n_mask == 0 || (a.pm_c && a.m_base && a.P0 && a.m_out), "cfg_update: mask pointers missing");
    const long long n4 = (a.n_img + a.n_mask) / 4;
    const int grid = (int)std::max<long long>(1, std::min<long long>(ceil_div_ll(n4, 256), 148 * 8));
    update_kernel<<<grid, 256, 0, s>>>(a);
    check_launch("cfg_solver_update");
}

// ----------------------------------------------------------------------------------------------
// Multistep DPM-Solver++ 2M / 3M (dpm_solver_pp.py:602-677, data prediction, solver_type='dpm_solver'), fused with
// the CFG combine and eps -> x0 like K12.  Operation order of the reference:
//   1 : x_t = A x + B X0                                   (dpm_solver_first_update with the cached prediction)
//   2M: D1_0 = (1/r0)(X0 - X1);  x_t = A x - B X0 - (0.5 B) D1_0
//   3M: D1_1 = (1/r1)(X1 - X2);  d = D1_0 - D1_1;  D1 = D1_0 + q d;  D2 = (1/(r0+r1)) d
//       x_t = A x - B X0 + C1 D1 - C2 D2
// The mask stream gets the same update with the mask prediction as its data prediction (no reference behaviour
// exists for it -- SURVEY F2 -- so that sub-case is "parity unpinned").
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ float ms_update(float x, float X0, float X1, float X2, const MultistepArgs& a) {
    if (a.order == 1) return __fadd_rn(__fmul_rn(a.A, x), __fmul_rn(a.B, X0));
    const float base = __fsub_rn(__fmul_rn(a.A, x), __fmul_rn(a.B, X0));
    const float D10 = __fmul_rn(a.inv_r0, __fsub_rn(X0, X1));
    if (a.order == 2) return __fsub_rn(base, __fmul_rn(a.halfB, D10));
    const float D11 = __fmul_rn(a.inv_r1, __fsub_rn(X1, X2));
    const float d = __fsub_rn(D10, D11);
    const float D1 = __fadd_rn(D10, __fmul_rn(a.q, d));
    const float D2 = __fmul_rn(a.inv_r01, d);
    return __fsub_rn(__fadd_rn(base, __fmul_rn(a.C1, D1)), __fmul_rn(a.C2, D2));
}

__global__ void __launch_bounds__(256) multistep_kernel(MultistepArgs a) {
    const long long n = a.n_img + a.n_mask;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        if (i < a.n_img) {
            const float c = a.eps_c[i];
            float e = c;
            if (a.eps_u) e = __fadd_rn(c, __fmul_rn(a.scale, __fsub_rn(c, a.eps_u[i])));
            const float x = a.x[i];
            const float X0 = __fdiv_rn(__fsub_rn(x, __fmul_rn(a.sigma, e)), a.alpha);
            const float X1 = a.order >= 2 ? a.X1[i] : 0.f, X2 = a.order >= 3 ? a.X2[i] : 0.f;
            a.X0[i] = X0;
            a.x_out[i] = ms_update(x, X0, X1, X2, a);
        } else {
            const long long j = i - a.n_img;
            const float c = a.pm_c[j];
            float P0 = c;
            if (a.pm_u) P0 = __fadd_rn(c, __fmul_rn(a.scale, __fsub_rn(c, a.pm_u[j])));
            const float P1 = a.order >= 2 ? a.P1[j] : 0.f, P2 = a.order >= 3 ? a.P2[j] : 0.f;
            a.P0[j] = P0;
            a.m_out[j] = ms_update(a.m[j], P0, P1, P2, a);
        }
    }
}

void multistep_update(const MultistepArgs& a, cudaStream_t s) {
    PDM_REQUIRE(a.order >= 1 && a.order <= 3, "multistep: order must be 1, 2 or 3");
    PDM_REQUIRE(a.n_mask == 0 || (a.pm_c && a.m && a.P0 && a.m_out), "multistep: mask pointers missing");
    const long long n = a.n_img + a.n_mask;
    const int grid = (int)std::max<long long>(1, std::min<long long>(ceil_div_ll(n, 256), 148 * 8));
    multistep_kernel<<<grid, 256, 0, s>>>(a);
    check_launch("multistep_update");
}

// ----------------------------------------------------------------------------------------------
// analog-bit codec (utils.py:475-518): MSB first.
// ----------------------------------------------------------------------------------------------
__global__ void bits2int_kernel(const float* __restrict__ pm, int32_t* __restrict__ labels, int B, int nbits, int hw) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)B * hw) return;
    const int b = (int)(idx / hw), p = (int)(idx % hw);
    int v = 0;
    for (int i = 0; i < nbits; ++i) v = (v << 1) | (pm[((long long)b * nbits + i) * hw + p] > 0.f ? 1 : 0);
    labels[idx] = v;
}
__global__ void int2bits_kernel(const int32_t* __restrict__ ids, float* __restrict__ bits, int B, int nbits, int hw) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)B * nbits * hw) return;
    const int p = (int)(idx % hw);
    const int i = (int)((idx / hw) % nbits);
    const int b = (int)(idx / ((long long)hw * nbits));
    const int bit = (ids[(long long)b * hw + p] >> (nbits - 1 - i)) & 1;
    bits[idx] = bit ? 1.f : -1.f;
}
void bits2int(const float* pm, int32_t* labels, int B, int nbits, int hw, cudaStream_t s) {
    bits2int_kernel<<<(unsigned)ceil_div_ll((long long)B * hw, 256), 256, 0, s>>>(pm, labels, B, nbits, hw);
    check_launch("bits2int");
}
void int2bits(const int32_t* ids, float* bits, int B, int nbits, int hw, cudaStream_t s) {
    int2bits_kernel<<<(unsigned)ceil_div_ll((long long)B * nbits * hw, 256), 256, 0, s>>>(ids, bits, B, nbits, hw);
    check_launch("int2bits");
}

}  // namespace pdm
