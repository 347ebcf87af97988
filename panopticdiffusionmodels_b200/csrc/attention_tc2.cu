// tcgen05 flash attention v2 (bf16 operands, fp32 softmax / accumulate), head_dim 64, non-causal.
// Replaces F.scaled_dot_product_attention in libs/uvit_t2i.py:70-74.
//
// One CTA per (batch row, head, PAIR of 128-query tiles), one CTA per SM, all 512 TMEM columns:
//     S_a | S_b (2 x 128 fp32 cols)   O_a | O_b (2 x 64)   P_a | P_b (2 x 64 cols holding 128 bf16 each)
//   warps 0-3 / 4-7   softmax warpgroups a / b, one query row per thread: ONE pass over S (the 128 scores of the
//                     row live in registers), exp2 with a running reference max, P written back to TENSOR MEMORY
//                     as packed bf16 (tcgen05.st) -- it never touches shared memory.
//   warp 8            TMA producer: both Q tiles once, K/V tiles through a 3-stage ring, all straight out of the
//                     packed qkv activation [nb, L, 3D] via one 3-D tensor map (no head-major repack);
//                     both Q tiles share every K/V tile.
//   warp 9            one thread issues S = Q.K^T (UMMA 128x128x16 x4) and O += P.V (UMMA 128x64x16 x8, A = P from
//                     TMEM, B = V consumed MN-major as it lies in memory), ping-ponging between the two tiles so
//                     one tile's MMAs run under the other tile's softmax.
// O accumulates in TMEM across key tiles.  The online-softmax rescale is LAZY: a warp rescales its O rows (TMEM
// load / multiply / store) only when some row's max grew by more than 2^8 since the reference was taken; otherwise
// probabilities are simply expressed against the older reference (exact after the final 1/l normalisation).
#include <cstdlib>

#include "common.cuh"
#include "ptx.cuh"

namespace pdm {

CUtensorMap make_tmap_bf16_3d(const void* ptr, long long K, long long rows, long long nbatch, long long bs,
                              int box_rows, int box_k);
void attention_tc3(const bf16* qkv, bf16* out, int nb, int L, int H, cudaStream_t s);

namespace {

constexpr int QT = 128, KT = 128, HD = 64;
constexpr int TILE_BYTES = 128 * HD * 2;  // 16 KB
constexpr int KV_STAGES = 3;
constexpr int SMEM_BYTES = 2 * TILE_BYTES + KV_STAGES * 2 * TILE_BYTES + 256;
constexpr int THREADS = 320;
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t S_COL = 0, O_COL = 256, P_COL = 384;  // + tile * {128, 64, 64}
constexpr float RESCALE_LOG2 = 8.f;

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
// packed fp32x2 arithmetic (sm_100 FFMA2 / FADD2): two lanes per FMA-pipe instruction
__device__ __forceinline__ uint64_t pack_f2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack_f2(uint64_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}

__global__ void __launch_bounds__(THREADS, 1)
attention_tc2_kernel(const __grid_constant__ CUtensorMap tmQKV, bf16* __restrict__ out, int L, int H) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sQ = smem;                        // [2] tiles
    uint8_t* sKV = smem + 2 * TILE_BYTES;      // [KV_STAGES] x (K tile | V tile)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * TILE_BYTES + KV_STAGES * 2 * TILE_BYTES);
    uint64_t* q_full = bars;                   // 1
    uint64_t* kv_full = bars + 1;              // [3]
    uint64_t* kv_empty = bars + 4;             // [3]
    uint64_t* s_full = bars + 7;               // [2] per tile
    uint64_t* p_full = bars + 9;               // [2]
    uint64_t* o_full = bars + 11;              // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int D = H * HD;
    const int q0 = blockIdx.x * 2 * QT, h = blockIdx.y, b = blockIdx.z;
    const int nkv = (L + KT - 1) / KT;
    const bool two = q0 + QT < L;              // does the second query tile exist?

    if (threadIdx.x == 0) {
        if (ptx::smem_u32(smem) & 1023) __trap();  // SWIZZLE_128B tiles need 1024-byte alignment
        ptx::prefetch_tmap(&tmQKV);
        ptx::mbar_init(q_full, 1);
        for (int i = 0; i < KV_STAGES; ++i) {
            ptx::mbar_init(&kv_full[i], 1);
            ptx::mbar_init(&kv_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&s_full[i], 1);
            ptx::mbar_init(&p_full[i], 128);
            ptx::mbar_init(&o_full[i], 1);
        }
        ptx::fence_mbar_init();
    }
    if (warp == 9) {
        ptx::tmem_alloc(tmem_slot, TMEM_COLS);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 8) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            ptx::mbar_expect_tx(q_full, (two ? 2 : 1) * TILE_BYTES);
            ptx::tma_load_3d(&tmQKV, q_full, sQ, h * HD, q0, b);
            if (two) ptx::tma_load_3d(&tmQKV, q_full, sQ + TILE_BYTES, h * HD, q0 + QT, b);
            int st = 0;
            uint32_t ph = 0;
            for (int j = 0; j < nkv; ++j) {
                ptx::mbar_wait(&kv_empty[st], ph ^ 1);
                uint8_t* sk = sKV + st * 2 * TILE_BYTES;
                ptx::mbar_expect_tx(&kv_full[st], 2 * TILE_BYTES);
                ptx::tma_load_3d(&tmQKV, &kv_full[st], sk, D + h * HD, j * KT, b);
                ptx::tma_load_3d(&tmQKV, &kv_full[st], sk + TILE_BYTES, 2 * D + h * HD, j * KT, b);
                if (++st == KV_STAGES) {
                    st = 0;
                    ph ^= 1;
                }
            }
        }
    } else if (warp == 9) {
        // ===================== MMA issuer (one thread) =====================
        if (lane == 0) {
            constexpr uint32_t idesc_qk = ptx::make_idesc_bf16(QT, KT, 0, 0);  // A = Q K-major, B = K K-major
            constexpr uint32_t idesc_pv = ptx::make_idesc_bf16(QT, HD, 0, 1);  // A = P (TMEM), B = V MN-major
            const int nt = two ? 2 : 1;
            auto issue_qk = [&](int t, uint32_t kaddr) {
                const uint64_t qdesc = ptx::make_smem_desc_sw128(ptx::smem_u32(sQ + t * TILE_BYTES), 1024);
                const uint64_t kdesc = ptx::make_smem_desc_sw128(kaddr, 1024);
#pragma unroll
                for (int k = 0; k < HD / 16; ++k)
                    ptx::mma_bf16_ss(tmem_base + S_COL + t * 128, qdesc + 2 * k, kdesc + 2 * k, idesc_qk, k != 0);
                ptx::mma_commit(&s_full[t]);
            };
            auto issue_pv = [&](int t, uint32_t vaddr, bool first) {
                // V tile: rows = keys (K dim), 128 bytes of head-dim per row (N contiguous) -> MN-major; 8-key groups are
                // 1024 bytes apart; one UMMA_K step (16 keys) = 2048 bytes.  P: 16 bf16 = 8 TMEM columns per step.
                const uint64_t vdesc = ptx::make_smem_desc_sw128(vaddr, 1024, 1024);
#pragma unroll
                for (int kk = 0; kk < KT / 16; ++kk)
                    ptx::mma_bf16_ts(tmem_base + O_COL + t * 64, tmem_base + P_COL + t * 64 + kk * 8,
                                     vdesc + kk * (2048 >> 4), idesc_pv, !(first && kk == 0));
                ptx::mma_commit(&o_full[t]);
            };
            ptx::mbar_wait(q_full, 0);
            int st = 0;
            uint32_t ph = 0;
            ptx::mbar_wait(&kv_full[0], 0);
            ptx::tc_fence_after();
            for (int t = 0; t < nt; ++t) issue_qk(t, ptx::smem_u32(sKV));
            for (int j = 0; j < nkv; ++j) {
                const uint32_t kv_cur = ptx::smem_u32(sKV + st * 2 * TILE_BYTES);
                int st_n = st + 1;
                uint32_t ph_n = ph;
                if (st_n == KV_STAGES) {
                    st_n = 0;
                    ph_n ^= 1;
                }
                const uint32_t kv_nxt = ptx::smem_u32(sKV + st_n * 2 * TILE_BYTES);
                const bool more = j + 1 < nkv;
                for (int t = 0; t < nt; ++t) {
                    ptx::mbar_wait(&p_full[t], j & 1);
                    ptx::tc_fence_after();
                    issue_pv(t, kv_cur + TILE_BYTES, j == 0);
                    if (t == nt - 1) ptx::mma_commit(&kv_empty[st]);  // K_j / V_j no longer needed by any tile
                    if (more) {
                        if (t == 0) {
                            ptx::mbar_wait(&kv_full[st_n], ph_n);
                            ptx::tc_fence_after();
                        }
                        issue_qk(t, kv_nxt);
                    }
                }
                st = st_n;
                ph = ph_n;
            }
        }
    } else {
        // ===================== softmax warpgroups: thread <-> query row =====================
        const int t = warp >> 2;                 // tile handled by this warpgroup
        if (t == 0 || two) {
            const int r = threadIdx.x & 127;
            const uint32_t lane_base = uint32_t((warp & 3) * 32) << 16;
            const uint32_t s_addr = tmem_base + lane_base + S_COL + t * 128;
            const uint32_t o_addr = tmem_base + lane_base + O_COL + t * 64;
            const uint32_t p_addr = tmem_base + lane_base + P_COL + t * 64;
            const float cs = 0.125f * 1.4426950408889634f;  // softmax scale * log2(e)
            float m_ref = -INFINITY, l = 0.f;
            // one 32-column chunk: p = exp2(s * cs - mb) against the current reference, packed bf16 -> P in TMEM;
            // returns the chunk's row max (raw scores) and adds the chunk's row sum to `lsum` (4-way ILP)
            auto do_chunk = [&](const uint32_t (&v)[32], int c, int nvalid, float mb, float& lsum) -> float {
                float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
                float l4[4] = {0.f, 0.f, 0.f, 0.f};
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int c0 = c * 32 + 2 * i;
                    const float s0 = __uint_as_float(v[2 * i]), s1 = __uint_as_float(v[2 * i + 1]);
                    const bool ok0 = c0 < nvalid, ok1 = c0 + 1 < nvalid;
                    if (ok0) mx4[i & 3] = fmaxf(mx4[i & 3], s0);
                    if (ok1) mx4[i & 3] = fmaxf(mx4[i & 3], s1);
                    const float p0 = ok0 ? ex2(fmaf(s0, cs, -mb)) : 0.f;
                    const float p1 = ok1 ? ex2(fmaf(s1, cs, -mb)) : 0.f;
                    l4[i & 3] += p0 + p1;
                    pk[i] = pack_bf16(p0, p1);
                }
                ptx::tmem_st_32x32_x16(p_addr + c * 16, pk);
                lsum += (l4[0] + l4[1]) + (l4[2] + l4[3]);
                return fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
            };
            // full (unmasked) chunk: 16 FMNMX3 + 16 FFMA2 + 32 MUFU + 16 FADD2 + 16 F2FP = 3 instructions / element
            auto do_chunk_full = [&](const uint32_t (&v)[32], int c, float mb, float& lsum) -> float {
                const uint64_t cs2 = pack_f2(cs, cs), nmb2 = pack_f2(-mb, -mb);
                float mxa = -INFINITY, mxb = -INFINITY;
                uint64_t la = pack_f2(0.f, 0.f), lb = la;
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float s0 = __uint_as_float(v[2 * i]), s1 = __uint_as_float(v[2 * i + 1]);
                    if (i & 1) mxb = max3(mxb, s0, s1); else mxa = max3(mxa, s0, s1);
                    float a0, a1;
                    unpack_f2(fma2(pack_f2(s0, s1), cs2, nmb2), a0, a1);
                    const float p0 = ex2(a0), p1 = ex2(a1);
                    if (i & 1) lb = add2(lb, pack_f2(p0, p1)); else la = add2(la, pack_f2(p0, p1));
                    pk[i] = pack_bf16(p0, p1);
                }
                ptx::tmem_st_32x32_x16(p_addr + c * 16, pk);
                float x0, x1;
                unpack_f2(add2(la, lb), x0, x1);
                lsum += x0 + x1;
                return fmaxf(mxa, mxb);
            };
            // ping-pong token between the two warpgroups: the exp2-heavy phase of one tile runs while the other
            // tile's MMAs (P.V, next Q.K^T) execute, instead of both warpgroups computing and then both waiting
            const int bar_mine = 1 + t, bar_other = 2 - t;
            if (two && t == 1) bar_arrive(1, 256);
            for (int j = 0; j < nkv; ++j) {
                const int nvalid = min(KT, L - j * KT);
                const bool full_tile = nvalid == KT;
                ptx::mbar_wait(&s_full[t], j & 1);
                ptx::tc_fence_after();
                if (two) bar_sync(bar_mine, 256);
                uint32_t va[32], vb[32];
                ptx::tmem_ld_32x32(s_addr, va);
                ptx::tmem_ld_wait();
                if (j == 0) {
                    // no reference yet: take the max of the first chunk (the redo path below covers the rest)
                    float m0 = -INFINITY;
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (i < nvalid) m0 = fmaxf(m0, __uint_as_float(va[i]));
                    m_ref = m0;
                }
                // software pipeline: the TMEM load of chunk c+1 is in flight while chunk c goes through exp2
                const float mb = m_ref * cs;
                float lsum = 0.f, mx;
                ptx::tmem_ld_32x32(s_addr + 32, vb);
                if (full_tile) {
                    mx = do_chunk_full(va, 0, mb, lsum);
                    ptx::tmem_ld_wait();
                    ptx::tmem_ld_32x32(s_addr + 64, va);
                    mx = fmaxf(mx, do_chunk_full(vb, 1, mb, lsum));
                    ptx::tmem_ld_wait();
                    ptx::tmem_ld_32x32(s_addr + 96, vb);
                    mx = fmaxf(mx, do_chunk_full(va, 2, mb, lsum));
                    ptx::tmem_ld_wait();
                    mx = fmaxf(mx, do_chunk_full(vb, 3, mb, lsum));
                } else {
                    mx = do_chunk(va, 0, nvalid, mb, lsum);
                    ptx::tmem_ld_wait();
                    ptx::tmem_ld_32x32(s_addr + 64, va);
                    mx = fmaxf(mx, do_chunk(vb, 1, nvalid, mb, lsum));
                    ptx::tmem_ld_wait();
                    ptx::tmem_ld_32x32(s_addr + 96, vb);
                    mx = fmaxf(mx, do_chunk(va, 2, nvalid, mb, lsum));
                    ptx::tmem_ld_wait();
                    mx = fmaxf(mx, do_chunk(vb, 3, nvalid, mb, lsum));
                }
                const bool need = (mx - m_ref) * cs > RESCALE_LOG2;
                if (__any_sync(0xffffffffu, need)) {
                    // rare: some row outgrew its reference by more than 2^8.  Move this warp's rows to a new reference:
                    // rescale l and the O rows accumulated so far, then redo this tile's P (S is still intact in TMEM).
                    const float m_new = fmaxf(m_ref, mx);
                    const float corr = ex2((m_ref - m_new) * cs);  // j == 0: O is still empty, l == 0
                    if (j > 0) {
                        ptx::mbar_wait(&o_full[t], (j - 1) & 1);  // PV_{j-1} has landed in O
                        ptx::tc_fence_after();
#pragma unroll
                        for (int c = 0; c < 2; ++c) {
                            ptx::tmem_ld_32x32(o_addr + c * 32, va);
                            ptx::tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 32; ++i) va[i] = __float_as_uint(__uint_as_float(va[i]) * corr);
                            ptx::tmem_st_32x32(o_addr + c * 32, va);
                        }
                    }
                    l *= corr;
                    m_ref = m_new;
                    const float mb2 = m_ref * cs;
                    lsum = 0.f;
#pragma unroll 1
                    for (int c = 0; c < 4; ++c) {
                        ptx::tmem_ld_32x32(s_addr + c * 32, va);
                        ptx::tmem_ld_wait();
                        do_chunk(va, c, nvalid, mb2, lsum);
                    }
                }
                l += lsum;
                if (two) bar_arrive(bar_other, 256);
                ptx::tmem_st_wait();
                ptx::tc_fence_before();
                ptx::mbar_arrive(&p_full[t]);
            }
            if (two && t == 0) bar_sync(1, 256);  // consume warpgroup b's last hand-over
            // epilogue: O / l -> bf16 rows
            ptx::mbar_wait(&o_full[t], (nkv - 1) & 1);
            ptx::tc_fence_after();
            const int qi = q0 + t * QT + r;
            const float inv = 1.f / l;
            bf16* dst = out + ((long long)b * L + qi) * D + h * HD;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t v[32];
                ptx::tmem_ld_32x32(o_addr + c * 32, v);
                ptx::tmem_ld_wait();
                if (qi < L) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint4 pk;
                        pk.x = pack_bf16(__uint_as_float(v[8 * q + 0]) * inv, __uint_as_float(v[8 * q + 1]) * inv);
                        pk.y = pack_bf16(__uint_as_float(v[8 * q + 2]) * inv, __uint_as_float(v[8 * q + 3]) * inv);
                        pk.z = pack_bf16(__uint_as_float(v[8 * q + 4]) * inv, __uint_as_float(v[8 * q + 5]) * inv);
                        pk.w = pack_bf16(__uint_as_float(v[8 * q + 6]) * inv, __uint_as_float(v[8 * q + 7]) * inv);
                        reinterpret_cast<uint4*>(dst + c * 32)[q] = pk;
                    }
                }
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 9) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

}  // namespace

void attention_tc_bf16(const bf16* qkv, bf16* out, int nb, int L, int H, cudaStream_t s) {
    // production kernel: attention_tc3.cu; this file's non-persistent predecessor stays selectable for A/B timing
    static const bool v2 = getenv("PDM_ATTN_V2") != nullptr;
    if (!v2) return attention_tc3(qkv, out, nb, L, H, s);
    const int D = H * HD;
    const CUtensorMap tm = make_tmap_bf16_3d(qkv, 3LL * D, L, nb, L, 128, HD);
    static bool attr_set = false;
    if (!attr_set) {
        PDM_CHECK_CUDA(cudaFuncSetAttribute(attention_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        attr_set = true;
    }
    dim3 grid(ceil_div(L, 2 * QT), H, nb);
    attention_tc2_kernel<<<grid, THREADS, SMEM_BYTES, s>>>(tm, out, L, H);
    check_launch("attention_tc2");
}

}  // namespace pdm
