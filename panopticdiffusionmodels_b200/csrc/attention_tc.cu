// tcgen05 flash attention (bf16 operands, fp32 softmax / accumulate) for head_dim 64, non-causal.
// Replaces F.scaled_dot_product_attention in libs/uvit_t2i.py:70-74.
//
// One CTA per (batch row, head, 128-query tile); it walks the keys in 128-key tiles:
//   warp 4   TMA producer: Q tile once, then K/V tiles (2-stage ring) straight out of the packed
//            qkv activation [nb, L, 3D] through ONE 3-D tensor map (no head-major repack).
//   warp 5   one thread issues S = Q.K^T (UMMA 128x128x16 x4, S in TMEM) and O_j = P.V
//            (UMMA 128x64x16 x8, V consumed MN-major as it lies in memory).
//   warps 0-3  one query row per thread: two passes over S in TMEM (row max, then exp2 / row sum),
//            P written as bf16 into shared memory in the UMMA K-major SWIZZLE_128B layout, O_j read back
//            from TMEM and folded into the fp32 register accumulator with the online-softmax rescale.
// 2 CTAs are resident per SM (112 KB smem, 256 TMEM columns each), so one CTA's softmax overlaps the
// other's MMAs.
#include <cstdlib>
#include "common.cuh"
#include "ptx.cuh"

namespace pdm {

CUtensorMap make_tmap_bf16_3d(const void* ptr, long long K, long long rows, long long nbatch, long long bs,
                              int box_rows, int box_k);

namespace {

constexpr int QT = 128, KT = 128, HD = 64;
constexpr int TILE_BYTES = 128 * HD * 2;  // 16 KB: one 128 x 64 bf16 tile
constexpr int P_BYTES = QT * KT * 2;      // 32 KB
constexpr int SMEM_BYTES = TILE_BYTES * 5 + P_BYTES + 128;
constexpr int THREADS = 192;
constexpr uint32_t TMEM_COLS = 256;
constexpr uint32_t S_COL = 0, O_COL = 128;

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}

__global__ void __launch_bounds__(THREADS, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, bf16* __restrict__ out, int L, int H, int dbg) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sQ = smem;
    uint8_t* sK = smem + TILE_BYTES;          // [2]
    uint8_t* sV = smem + 3 * TILE_BYTES;      // [2]
    uint8_t* sP = smem + 5 * TILE_BYTES;      // 2 atoms of 128 rows x 64 keys
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 5 * TILE_BYTES + P_BYTES);
    uint64_t* q_full = bars;
    uint64_t* kv_full = bars + 1;   // [2]
    uint64_t* kv_empty = bars + 3;  // [2]
    uint64_t* s_full = bars + 5;
    uint64_t* p_full = bars + 6;
    uint64_t* o_full = bars + 7;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int D = H * HD;
    const int q0 = blockIdx.x * QT, h = blockIdx.y, b = blockIdx.z;
    const int nkv = (L + KT - 1) / KT;

    if (threadIdx.x == 0) {
        if (ptx::smem_u32(smem) & 1023) __trap();  // SWIZZLE_128B tiles need 1024-byte alignment
        ptx::prefetch_tmap(&tmQKV);
        ptx::mbar_init(q_full, 1);
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&kv_full[i], 1);
            ptx::mbar_init(&kv_empty[i], 1);
        }
        ptx::mbar_init(s_full, 1);
        ptx::mbar_init(p_full, 128);
        ptx::mbar_init(o_full, 1);
        ptx::fence_mbar_init();
    }
    if (warp == 5) {
        ptx::tmem_alloc(tmem_slot, TMEM_COLS);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4) {
        if (lane == 0) {
            ptx::mbar_expect_tx(q_full, TILE_BYTES);
            ptx::tma_load_3d(&tmQKV, q_full, sQ, h * HD, q0, b);
            for (int j = 0; j < nkv; ++j) {
                const int st = j & 1;
                const uint32_t ph = (j >> 1) & 1;
                ptx::mbar_wait(&kv_empty[st], ph ^ 1);
                ptx::mbar_expect_tx(&kv_full[st], 2 * TILE_BYTES);
                ptx::tma_load_3d(&tmQKV, &kv_full[st], sK + st * TILE_BYTES, D + h * HD, j * KT, b);
                ptx::tma_load_3d(&tmQKV, &kv_full[st], sV + st * TILE_BYTES, 2 * D + h * HD, j * KT, b);
            }
        }
    } else if (warp == 5) {
        if (lane == 0) {
            constexpr uint32_t idesc_qk = ptx::make_idesc_bf16(QT, KT, 0, 0);  // A=Q K-major, B=K K-major
            constexpr uint32_t idesc_pv = ptx::make_idesc_bf16(QT, HD, 0, 1);  // A=P K-major, B=V MN-major
            const uint64_t qdesc = ptx::make_smem_desc_sw128(ptx::smem_u32(sQ), 1024);
            const uint32_t sp_addr = ptx::smem_u32(sP);
            ptx::mbar_wait(q_full, 0);
            for (int j = 0; j < nkv; ++j) {
                const int st = j & 1;
                const uint32_t ph = (j >> 1) & 1;
                ptx::mbar_wait(&kv_full[st], ph);
                ptx::tc_fence_after();
                const uint64_t kdesc = ptx::make_smem_desc_sw128(ptx::smem_u32(sK + st * TILE_BYTES), 1024);
#pragma unroll
                for (int k = 0; k < HD / 16; ++k)
                    ptx::mma_bf16_ss(tmem_base + S_COL, qdesc + 2 * k, kdesc + 2 * k, idesc_qk, k != 0);
                ptx::mma_commit(s_full);
                ptx::mbar_wait(p_full, j & 1);
                ptx::tc_fence_after();
                // V tile: rows = keys (K dim), 128 bytes of head-dim per row (N contiguous) -> MN-major,
                // 8-key groups are 1024 bytes apart; one UMMA_K step (16 keys) = 2048 bytes.
                const uint64_t vdesc = ptx::make_smem_desc_sw128(ptx::smem_u32(sV + st * TILE_BYTES), 1024, 1024);
#pragma unroll
                for (int kk = 0; kk < KT / 16; ++kk) {
                    const uint64_t pdesc =
                        ptx::make_smem_desc_sw128(sp_addr + (kk >> 2) * (QT * 128) + (kk & 3) * 32, 1024);
                    ptx::mma_bf16_ss(tmem_base + O_COL, pdesc, vdesc + kk * (2048 >> 4), idesc_pv, kk != 0);
                }
                ptx::mma_commit(&kv_empty[st]);
                ptx::mma_commit(o_full);
            }
        }
    } else {
        // ---- softmax / accumulate: thread <-> query row ----
        const int r = threadIdx.x;  // 0..127
        const uint32_t lane_base = uint32_t(warp * 32) << 16;
        const float cs = 0.125f * 1.4426950408889634f;  // softmax scale * log2(e)
        float m = -INFINITY, l = 0.f;
        float o[HD];
#pragma unroll
        for (int i = 0; i < HD; ++i) o[i] = 0.f;
        for (int j = 0; j < nkv; ++j) {
            const int nvalid = min(KT, L - j * KT);
            ptx::mbar_wait(s_full, j & 1);
            ptx::tc_fence_after();
            // pass 1: row max
            float mx = (dbg & 1) ? 8.f : -INFINITY;
#pragma unroll 1
            for (int c = 0; c < ((dbg & 1) ? 0 : 4); ++c) {
                uint32_t v[32];
                ptx::tmem_ld_32x32(tmem_base + lane_base + S_COL + c * 32, v);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (c * 32 + i < nvalid) mx = fmaxf(mx, __uint_as_float(v[i]));
            }
            const float m_new = fmaxf(m, mx);
            const float corr = ex2((m - m_new) * cs);  // first tile: ex2(-inf) = 0
            const float mb = m_new * cs;
            l *= corr;
#pragma unroll
            for (int i = 0; i < HD; ++i) o[i] *= corr;
            m = m_new;
            // pass 2: p = exp2(s * cs - m * cs), row sum, bf16 P into smem (K-major SWIZZLE_128B)
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                uint32_t v[32];
                ptx::tmem_ld_32x32(tmem_base + lane_base + S_COL + c * 32, v);
                ptx::tmem_ld_wait();
                float p[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    p[i] = (c * 32 + i < nvalid) ? ex2(fmaf(__uint_as_float(v[i]), cs, -mb)) : 0.f;
                    l += p[i];
                }
                uint8_t* prow = sP + (c >> 1) * (QT * 128) + r * 128;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint4 pk;
                    pk.x = pack_bf16(p[8 * q + 0], p[8 * q + 1]);
                    pk.y = pack_bf16(p[8 * q + 2], p[8 * q + 3]);
                    pk.z = pack_bf16(p[8 * q + 4], p[8 * q + 5]);
                    pk.w = pack_bf16(p[8 * q + 6], p[8 * q + 7]);
                    const int chunk = (c & 1) * 4 + q;  // 16-byte chunk inside the 128-byte row
                    *reinterpret_cast<uint4*>(prow + ((chunk ^ (r & 7)) << 4)) = pk;
                }
            }
            ptx::fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core
            ptx::tc_fence_before();
            ptx::mbar_arrive(p_full);
            // O_j = P.V  ->  fold into the register accumulator
            ptx::mbar_wait(o_full, j & 1);
            ptx::tc_fence_after();
#pragma unroll
            for (int c = 0; c < ((dbg & 2) ? 0 : 2); ++c) {
                uint32_t v[32];
                ptx::tmem_ld_32x32(tmem_base + lane_base + O_COL + c * 32, v);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) o[c * 32 + i] += __uint_as_float(v[i]);
            }
        }
        const int qi = q0 + r;
        if (qi < L) {
            const float inv = 1.f / l;
            uint4* dst = reinterpret_cast<uint4*>(out + ((long long)b * L + qi) * D + h * HD);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                uint4 pk;
                pk.x = pack_bf16(o[8 * q + 0] * inv, o[8 * q + 1] * inv);
                pk.y = pack_bf16(o[8 * q + 2] * inv, o[8 * q + 3] * inv);
                pk.z = pack_bf16(o[8 * q + 4] * inv, o[8 * q + 5] * inv);
                pk.w = pack_bf16(o[8 * q + 6] * inv, o[8 * q + 7] * inv);
                dst[q] = pk;
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

}  // namespace

void attention_tc_v1(const bf16* qkv, bf16* out, int nb, int L, int H, cudaStream_t s) {
    const int D = H * HD;
    const CUtensorMap tm = make_tmap_bf16_3d(qkv, 3LL * D, L, nb, L, 128, HD);
    static bool attr_set = false;
    if (!attr_set) {
        PDM_CHECK_CUDA(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        attr_set = true;
    }
    dim3 grid(ceil_div(L, QT), H, nb);
    static const int dbg = getenv("PDM_ATTN_DBG") ? atoi(getenv("PDM_ATTN_DBG")) : 0;
    attention_tc_kernel<<<grid, THREADS, SMEM_BYTES, s>>>(tm, out, L, H, dbg);
    check_launch("attention_tc");
}

}  // namespace pdm
