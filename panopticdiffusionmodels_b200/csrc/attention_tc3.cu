// tcgen05 flash attention v3 (bf16 operands, fp32 softmax / accumulate), head_dim 64, non-causal.
// Replaces F.scaled_dot_product_attention in libs/uvit_t2i.py:70-74.
//
// PERSISTENT: one CTA per SM walks a static list of work items; an item is a PAIR of 128-query tiles:
//     "same"  items: two consecutive query tiles of one (row, head) -- both tiles share every K/V tile;
//     "split" items: the ragged last query tile of head 2i paired with the ragged last tile of head 2i+1
//                    (own K/V tiles each), so an odd tile count (L = 590 -> 5, L = 334 -> 3) costs no idle slot.
// Tensor memory (512 columns): S0 | S1 | S2 (3 x 128 fp32 columns, ROTATING over the sequence of tile-steps)
//                              O_a | O_b    (2 x 64).   P (bf16) overwrites the first 64 columns of its S buffer.
// With three S buffers for two tiles, Q.K^T of a tile-step is issued (and finished) long before its softmax
// warpgroup gets there: the exp2 phases of the two warpgroups run back to back on the MUFU pipe, which is the
// bound for head_dim 64 (16 ex2/clk/SM vs 32 scores/clk/SM of MMA).
//   warps 0-3 / 4-7   softmax warpgroups a / b, one query row per thread: whole S row -> registers, row max, lazy
//                     rescale (only when a row max outgrew its reference by > 2^8), exp2, P -> TMEM (tcgen05.st).
//                     Chunks of 32 keys beyond L are skipped, warps whose 32 query rows are all >= L do nothing.
//   warp 8            TMA producer: Q tiles (double-buffered across items) and a 5-slot ring of (K | V) tiles, all
//                     straight out of the packed qkv activation [nb, L, 3D] via one 3-D tensor map.
//   warp 9            one thread issues S = Q.K^T (UMMA 128 x N x 16, N = 128 or the ragged tail rounded to 16) and
//                     O += P.V (UMMA 128 x 64 x 16, A = P from TMEM, B = V MN-major as it lies in memory), keeping
//                     Q.K^T up to three tile-steps ahead of P.V -- across item boundaries too, so the prologue and
//                     the O epilogue of an item are hidden under its neighbours.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "ptx.cuh"

namespace pdm {

CUtensorMap make_tmap_bf16_3d(const void* ptr, long long K, long long rows, long long nbatch, long long bs,
                              int box_rows, int box_k);
namespace {

int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        PDM_CHECK_CUDA(cudaGetDevice(&dev));
        PDM_CHECK_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    }
    return n;
}

constexpr int QT = 128, KT = 128, HD = 64;
constexpr int TILE_BYTES = 128 * HD * 2;  // 16 KB
constexpr int NSLOT = 5;                  // K/V ring slots, (K tile | V tile) each
constexpr int NSBUF = 3;                  // rotating S buffers
constexpr int Q_BYTES = 4 * TILE_BYTES;   // 2 item slots x 2 tiles
constexpr int BAR_BYTES = 256;
constexpr int SMEM_BYTES = 1024 + Q_BYTES + NSLOT * 2 * TILE_BYTES + BAR_BYTES;
constexpr int THREADS = 384;  // warps 0-7 softmax, 8 TMA, 9 P.V, 10 Q.K^T, 11 idle (setmaxnreg works on whole warpgroups)
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t O_COL = 384;  // + t * 64;  S buffer i at column i * 128
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
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}

// ---- work items -------------------------------------------------------------------------------------------------
struct Shape {
    int L, H, H2, nq, fp, odd, ipu, nkv, n_items;
    int last_valid;  // valid keys of the last K/V tile (1..128)
    int last_n16;    // ... rounded up to the UMMA N / K granularity
};
struct Item {
    int b, hA, hB, qA, qB, nt;  // nt = number of live tiles (0: nothing to do)
    bool same;
};
__device__ __forceinline__ Item decode_item(const Shape& sh, int it) {
    Item I;
    const int unit = it / sh.ipu, r = it - unit * sh.ipu;
    I.b = unit / sh.H2;
    const int hp = unit - I.b * sh.H2;
    const int h0 = 2 * hp, h1 = 2 * hp + 1;
    const bool h1ok = h1 < sh.H;
    if (r < sh.fp) {
        I.hA = I.hB = h0; I.qA = 2 * r; I.qB = 2 * r + 1; I.nt = 2; I.same = true;
    } else if (sh.odd && r == sh.fp) {
        I.hA = h0; I.hB = h1; I.qA = I.qB = sh.nq - 1; I.nt = h1ok ? 2 : 1; I.same = false;
    } else {
        const int r2 = r - sh.fp - sh.odd;
        I.hA = I.hB = h1; I.qA = 2 * r2; I.qB = 2 * r2 + 1; I.nt = h1ok ? 2 : 0; I.same = true;
    }
    return I;
}

struct Ring {  // position in the K/V slot ring
    int slot = NSLOT - 1;
    uint32_t ph = 1;
    __device__ __forceinline__ void next() {
        if (++slot == NSLOT) {
            slot = 0;
            ph ^= 1;
        }
    }
};

#ifdef PDM_ATTN_TRACE
__device__ unsigned long long* g_attn_trace = nullptr;  // [warp 0..11][4096] x (event<<56 | n<<40 | clock)
__device__ __forceinline__ void trace_ev(int ev, uint32_t n, int& cnt) {
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && g_attn_trace && cnt < 4096) {
        const unsigned long long c = clock64() & 0xffffffffffull;
        g_attn_trace[(threadIdx.x >> 5) * 4096 + cnt++] = ((unsigned long long)ev << 56) | ((unsigned long long)(n & 0xffff) << 40) | c;
    }
}
#define TRACE(ev, n) trace_ev(ev, n, trace_cnt)
#else
#define TRACE(ev, n)
#endif

// One key tile of one query row: S row (NCH chunks of 32 fp32 columns at s_addr) -> registers, row max, lazy rescale
// of the O row, p = exp2((s - m_ref) * scale * log2 e) -> bf16 P written over the first NCH*16 columns of the S buffer.
// MASK: keys >= nvalid of the last chunk count as -inf.
template <int NCH, bool MASK>
__device__ __forceinline__ void softmax_tile(uint32_t s_addr, uint32_t o_addr, int nvalid, bool first, float& m_ref,
                                             float& l, uint64_t* o_full_bar, uint32_t o_full_parity) {
    const float cs = 0.125f * 1.4426950408889634f;  // softmax scale * log2(e)
    uint32_t s[NCH][32];
#pragma unroll
    for (int c = 0; c < NCH; ++c) ptx::tmem_ld_32x32(s_addr + c * 32, s[c]);
    ptx::tmem_ld_wait();
    if (MASK) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if ((NCH - 1) * 32 + i >= nvalid) s[NCH - 1][i] = 0xff800000u;
    }
    float mxa = -INFINITY, mxb = -INFINITY;
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
            mxa = max3(mxa, __uint_as_float(s[c][i]), __uint_as_float(s[c][i + 1]));
            mxb = max3(mxb, __uint_as_float(s[c][i + 2]), __uint_as_float(s[c][i + 3]));
        }
    const float mx = fmaxf(mxa, mxb);
    if (first) {
        m_ref = mx;
    } else {
        // P.V of the previous key tile has landed in O (long ago, normally).  Every warp consumes EVERY phase of
        // o_full in order: a parity wait is only meaningful while the waiter is at most one phase away.
        ptx::mbar_wait(o_full_bar, o_full_parity);
    }
    if (!first && __any_sync(0xffffffffu, (mx - m_ref) * cs > RESCALE_LOG2)) {
        // rare: some row outgrew its reference by more than 2^8: rescale l and the O rows accumulated so far
        const float m_new = fmaxf(m_ref, mx);
        const float corr = ex2((m_ref - m_new) * cs);
        ptx::tc_fence_after();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            uint32_t o[32];
            ptx::tmem_ld_32x32(o_addr + c * 32, o);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * corr);
            ptx::tmem_st_32x32(o_addr + c * 32, o);
        }
        l *= corr;
        m_ref = m_new;
    }
    const float nmb = -m_ref * cs;
    const uint64_t cs2 = pack_f2(cs, cs), nmb2 = pack_f2(nmb, nmb);
    uint64_t la = pack_f2(0.f, 0.f), lb = la;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            float a0, a1;
            unpack_f2(fma2(pack_f2(__uint_as_float(s[c][2 * i]), __uint_as_float(s[c][2 * i + 1])), cs2, nmb2), a0, a1);
            const float p0 = ex2(a0), p1 = ex2(a1);
            if (i & 1) lb = add2(lb, pack_f2(p0, p1)); else la = add2(la, pack_f2(p0, p1));
            pk[i] = pack_bf16(p0, p1);
        }
        ptx::tmem_st_32x32_x16(s_addr + c * 16, pk);  // P overwrites S (the whole S row is in registers)
    }
    float x0, x1;
    unpack_f2(add2(la, lb), x0, x1);
    l += x0 + x1;
}


__global__ void __launch_bounds__(THREADS, 1)
attention_tc3_kernel(const __grid_constant__ CUtensorMap tmQKV, bf16* __restrict__ out, const Shape sh) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);  // SWIZZLE_128B tiles: 1024 B
    uint8_t* sQ = smem;               // [2 item slots][2 tiles]
    uint8_t* sKV = smem + Q_BYTES;    // [NSLOT] x (K tile | V tile)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Q_BYTES + NSLOT * 2 * TILE_BYTES);
    uint64_t* q_full = bars;                  // [2]
    uint64_t* q_empty = bars + 2;             // [2]
    uint64_t* kv_full = bars + 4;             // [NSLOT]
    uint64_t* kv_empty = bars + 4 + NSLOT;    // [NSLOT]
    uint64_t* s_full = bars + 4 + 2 * NSLOT;  // [NSBUF]
    uint64_t* p_full = s_full + NSBUF;        // [NSBUF]
    uint64_t* o_full = p_full + NSBUF;        // [2]
    uint64_t* o_empty = o_full + 2;           // [2]
    uint64_t* s_free = o_empty + 2;           // [NSBUF]  P.V done with the P in S buffer i
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_free + NSBUF);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int D = sh.H * HD;
    const int nkv = sh.nkv;
#ifdef PDM_ATTN_TRACE
    int trace_cnt = 0;
#endif

    if (threadIdx.x == 0) {
        ptx::prefetch_tmap(&tmQKV);
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&q_full[i], 1);
            ptx::mbar_init(&q_empty[i], 1);
            ptx::mbar_init(&o_full[i], 1);
            ptx::mbar_init(&o_empty[i], 4);
        }
        for (int i = 0; i < NSLOT; ++i) {
            ptx::mbar_init(&kv_full[i], 1);
            ptx::mbar_init(&kv_empty[i], 1);
        }
        for (int i = 0; i < NSBUF; ++i) {
            ptx::mbar_init(&s_full[i], 1);
            ptx::mbar_init(&p_full[i], 4);
            ptx::mbar_init(&s_free[i], 1);
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

    if (warp >= 8) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        if (warp == 8) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            Ring ring;
            int qi = 0;
            for (int it = blockIdx.x; it < sh.n_items; it += gridDim.x) {
                const Item I = decode_item(sh, it);
                if (I.nt == 0) continue;
                const int qs = qi & 1;
                ptx::mbar_wait_relaxed(&q_empty[qs], ((qi >> 1) & 1) ^ 1);
                ptx::mbar_expect_tx(&q_full[qs], I.nt * TILE_BYTES);
                ptx::tma_load_3d(&tmQKV, &q_full[qs], sQ + qs * 2 * TILE_BYTES, I.hA * HD, I.qA * QT, I.b);
                if (I.nt == 2)
                    ptx::tma_load_3d(&tmQKV, &q_full[qs], sQ + qs * 2 * TILE_BYTES + TILE_BYTES, I.hB * HD, I.qB * QT, I.b);
                const int nsrc = (I.nt == 2 && !I.same) ? 2 : 1;
                for (int j = 0; j < nkv; ++j) {
                    for (int u = 0; u < nsrc; ++u) {
                        const int h = u ? I.hB : I.hA;
                        ring.next();
                        ptx::mbar_wait_relaxed(&kv_empty[ring.slot], ring.ph ^ 1);
                        uint8_t* sk = sKV + ring.slot * 2 * TILE_BYTES;
                        ptx::mbar_expect_tx(&kv_full[ring.slot], 2 * TILE_BYTES);
                        ptx::tma_load_3d(&tmQKV, &kv_full[ring.slot], sk, D + h * HD, j * KT, I.b);
                        ptx::tma_load_3d(&tmQKV, &kv_full[ring.slot], sk + TILE_BYTES, 2 * D + h * HD, j * KT, I.b);
                    }
                }
                ++qi;
            }
        }
        } else if (warp == 9 || warp == 10) {
            // ===================== MMA issuers: warp 9 issues O += P.V, warp 10 issues S = Q.K^T =====================
            // The WHOLE warp walks the schedule (warp-uniform control flow and operands, so the descriptors live in
            // uniform registers); only the tcgen05.mma / tcgen05.commit instructions themselves are issued by one
            // elected lane.  (A single-lane region makes the compiler wrap every MMA in a uniformisation loop of
            // ~17 dependent instructions: ~100 clk per MMA, which starved the tensor pipe.)  Two warps because
            // tcgen05.mma issue blocks while the pipe's queue is full: one warp doing both could not keep up with
            // the softmax warpgroups.  Q.K^T (n + 3) reuses the S buffer whose P feeds P.V (n): the Q.K^T warp waits
            // for the COMPLETION of P.V (n) (s_free), since MMAs of different threads are not ordered.
            const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
            const uint32_t sQ_u = __shfl_sync(0xffffffffu, ptx::smem_u32(sQ), 0);
            const uint32_t sKV_u = __shfl_sync(0xffffffffu, ptx::smem_u32(sKV), 0);

            // cursor over the CTA's sequence of tile-steps (item, tile t, key tile j)
            int it = blockIdx.x, k = 0, nsteps = 0, qi = 0;
            uint32_t n = 0;
            Item I;
            Ring ring;
            auto seek = [&]() {  // position on the first live item at or after `it`
                while (it < sh.n_items) {
                    I = decode_item(sh, it);
                    if (I.nt) break;
                    it += gridDim.x;
                }
                k = 0;
                nsteps = it < sh.n_items ? I.nt * nkv : 0;
            };
            auto advance = [&]() {
                ++n;
                if (++k == nsteps) {
                    it += gridDim.x;
                    ++qi;
                    seek();
                }
            };
            seek();

            if (warp == 10) {
                const uint32_t idesc_qk_full = ptx::make_idesc_bf16(QT, KT, 0, 0);
                const uint32_t idesc_qk_last = ptx::make_idesc_bf16(QT, sh.last_n16, 0, 0);
                while (it < sh.n_items) {
                    const int t = I.nt == 2 ? (k & 1) : 0, j = I.nt == 2 ? (k >> 1) : k;
                    const int qs = qi & 1;
                    const uint32_t buf = n % NSBUF, use = n / NSBUF;
                    if (k == 0) ptx::mbar_wait(&q_full[qs], (qi >> 1) & 1);
                    if (!(I.same && t == 1)) {  // first user of a K/V slot
                        ring.next();
                        ptx::mbar_wait(&kv_full[ring.slot], ring.ph);
                    }
                    const uint64_t qdesc = ptx::make_smem_desc_sw128(sQ_u + (qs * 2 + t) * TILE_BYTES, 1024);
                    const uint64_t kdesc = ptx::make_smem_desc_sw128(sKV_u + ring.slot * 2 * TILE_BYTES, 1024);
                    const uint32_t d = tb + buf * 128;
                    const uint32_t idesc = j == nkv - 1 ? idesc_qk_last : idesc_qk_full;
                    const bool item_done = k == nsteps - 1;
                    TRACE(14, n);
                    if (use > 0) ptx::mbar_wait(&s_free[buf], (use - 1) & 1);  // P.V (n - 3) has consumed the buffer
                    ptx::tc_fence_after();
                    TRACE(15, n);
                    if (ptx::elect_one()) {
#pragma unroll
                        for (int kk = 0; kk < HD / 16; ++kk)
                            ptx::mma_bf16_ss(d, qdesc + 2 * kk, kdesc + 2 * kk, idesc, kk != 0);
                        ptx::mma_commit(&s_full[buf]);
                        if (item_done) ptx::mma_commit(&q_empty[qs]);  // all Q.K^T of the item issued: its Q tiles are free
                    }
                    __syncwarp();
                    TRACE(10, n);
                    advance();
                }
            } else {
                constexpr uint32_t idesc_pv = ptx::make_idesc_bf16(QT, HD, 0, 1);  // A = P (TMEM), B = V MN-major
                const int ksteps_last = sh.last_n16 / 16;
                uint32_t o_uses[2] = {0, 0};  // items that have used O_t so far
                while (it < sh.n_items) {
                    const int t = I.nt == 2 ? (k & 1) : 0, j = I.nt == 2 ? (k >> 1) : k;
                    if (!(I.same && t == 1)) ring.next();
                    const uint32_t buf = n % NSBUF;
                    // V tile: rows = keys (K dim), 128 bytes of head-dim per row (N contiguous) -> MN-major; 8-key groups
                    // are 1024 bytes apart; one UMMA_K step (16 keys) = 2048 bytes.  P: 16 bf16 = 8 TMEM columns per step.
                    const uint64_t vdesc =
                        ptx::make_smem_desc_sw128(sKV_u + ring.slot * 2 * TILE_BYTES + TILE_BYTES, 1024, 1024);
                    const bool last_tile = j == nkv - 1;
                    uint64_t* kv_bar = !(I.same && t == 0) ? &kv_empty[ring.slot] : nullptr;  // last user of the slot
                    const uint32_t d_o = tb + O_COL + t * 64, a_p = tb + buf * 128;
                    const uint32_t acc0 = j != 0;
                    if (j == 0) {  // O_t is rewritten: the previous item's epilogue must have drained it
                        ptx::mbar_wait(&o_empty[t], (o_uses[t] & 1) ^ 1);
                        ++o_uses[t];
                    }
                    TRACE(11, n);
                    ptx::mbar_wait(&p_full[buf], (n / NSBUF) & 1);
                    ptx::tc_fence_after();
                    TRACE(12, n);
                    if (ptx::elect_one()) {
                        if (!last_tile) {
                            ptx::mma_bf16_ts(d_o, a_p, vdesc, idesc_pv, acc0);
#pragma unroll
                            for (int kk = 1; kk < KT / 16; ++kk)
                                ptx::mma_bf16_ts(d_o, a_p + kk * 8, vdesc + kk * (2048 >> 4), idesc_pv, 1);
                        } else {
                            ptx::mma_bf16_ts(d_o, a_p, vdesc, idesc_pv, acc0);
                            for (int kk = 1; kk < ksteps_last; ++kk)
                                ptx::mma_bf16_ts(d_o, a_p + kk * 8, vdesc + kk * (2048 >> 4), idesc_pv, 1);
                        }
                        ptx::mma_commit(&s_free[buf]);
                        ptx::mma_commit(&o_full[t]);
                        if (kv_bar) ptx::mma_commit(kv_bar);
                    }
                    __syncwarp();
                    TRACE(13, n);
                    advance();
                }
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
        // ===================== softmax warpgroups: thread <-> query row =====================
        const int t = warp >> 2;  // tile slot handled by this warpgroup
        const int wq = warp & 3;
        const uint32_t lane_base = uint32_t(wq * 32) << 16;
        const uint32_t o_addr = tmem_base + lane_base + O_COL + t * 64;
        const int nch_last = (sh.last_n16 + 31) >> 5;  // 32-key chunks of the ragged last key tile that P.V reads
        uint32_t n_base = 0;   // tile-steps issued before the current item
        uint32_t steps = 0;    // tile-steps of slot t completed so far (o_full phase counter)

        for (int it = blockIdx.x; it < sh.n_items; it += gridDim.x) {
            int nt, q0, h, b;
            {
                const Item I = decode_item(sh, it);
                nt = I.nt; q0 = (t ? I.qB : I.qA) * QT; h = t ? I.hB : I.hA; b = I.b;
            }
            if (nt == 0) continue;
            if (t < nt) {
                const bool live = q0 + wq * 32 < sh.L;  // does this warp own any real query row?
                float m_ref = 0.f, l = 0.f;
                const uint32_t n0 = n_base + (nt == 2 ? t : 0), dn = nt == 2 ? 2 : 1;
                for (int j = 0; j < nkv; ++j) {
                    const uint32_t n = n0 + j * dn;
                    const uint32_t buf = n % NSBUF;
                    const uint32_t s_addr = tmem_base + lane_base + buf * 128;
                    TRACE(0, n);
                    ptx::mbar_wait(&s_full[buf], (n / NSBUF) & 1);
                    ptx::tc_fence_after();
                    TRACE(1, n);
                    if (live) {
                        uint64_t* of = &o_full[t];
                        const uint32_t ofp = (steps - 1) & 1;
                        if (j < nkv - 1) {
                            softmax_tile<4, false>(s_addr, o_addr, KT, j == 0, m_ref, l, of, ofp);
                        } else {
                            switch (nch_last) {
                                case 1: softmax_tile<1, true>(s_addr, o_addr, sh.last_valid, j == 0, m_ref, l, of, ofp); break;
                                case 2: softmax_tile<2, true>(s_addr, o_addr, sh.last_valid, j == 0, m_ref, l, of, ofp); break;
                                case 3: softmax_tile<3, true>(s_addr, o_addr, sh.last_valid, j == 0, m_ref, l, of, ofp); break;
                                default: softmax_tile<4, true>(s_addr, o_addr, sh.last_valid, j == 0, m_ref, l, of, ofp); break;
                            }
                        }
                        TRACE(2, n);
                        ptx::tmem_st_wait();
                    } else if (j > 0) {
                        ptx::mbar_wait(&o_full[t], (steps - 1) & 1);  // keep in step with o_full (see softmax_tile)
                    }
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&p_full[buf]);
                    TRACE(3, n);
                    ++steps;
                }
                // epilogue: O / l -> bf16 rows
                ptx::mbar_wait(&o_full[t], (steps - 1) & 1);
                ptx::tc_fence_after();
                TRACE(4, n0);
                if (live) {
                    const int qi = q0 + wq * 32 + lane;
                    const float inv = 1.f / l;
                    bf16* dst = out + ((long long)b * sh.L + qi) * D + h * HD;
                    uint32_t v[2][32];
                    ptx::tmem_ld_32x32(o_addr, v[0]);
                    ptx::tmem_ld_32x32(o_addr + 32, v[1]);
                    ptx::tmem_ld_wait();
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&o_empty[t]);  // O_t may be overwritten by the next item
                    if (qi < sh.L) {
#pragma unroll
                        for (int c = 0; c < 2; ++c)
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                uint4 pk;
                                pk.x = pack_bf16(__uint_as_float(v[c][8 * q + 0]) * inv, __uint_as_float(v[c][8 * q + 1]) * inv);
                                pk.y = pack_bf16(__uint_as_float(v[c][8 * q + 2]) * inv, __uint_as_float(v[c][8 * q + 3]) * inv);
                                pk.z = pack_bf16(__uint_as_float(v[c][8 * q + 4]) * inv, __uint_as_float(v[c][8 * q + 5]) * inv);
                                pk.w = pack_bf16(__uint_as_float(v[c][8 * q + 6]) * inv, __uint_as_float(v[c][8 * q + 7]) * inv);
                                reinterpret_cast<uint4*>(dst + c * 32)[q] = pk;
                            }
                    }
                } else {
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&o_empty[t]);
                }
                TRACE(5, n0);
            }
            n_base += nt * nkv;
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

void attention_tc3(const bf16* qkv, bf16* out, int nb, int L, int H, cudaStream_t s) {
    const int D = H * HD;
    const CUtensorMap tm = make_tmap_bf16_3d(qkv, 3LL * D, L, nb, L, 128, HD);
    static bool attr_set = false;
    if (!attr_set) {
        PDM_CHECK_CUDA(cudaFuncSetAttribute(attention_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        attr_set = true;
    }
    Shape sh;
    sh.L = L;
    sh.H = H;
    sh.H2 = (H + 1) / 2;
    sh.nq = ceil_div(L, QT);
    sh.fp = sh.nq / 2;
    sh.odd = sh.nq & 1;
    sh.ipu = 2 * sh.fp + sh.odd;
    sh.nkv = ceil_div(L, KT);
    sh.n_items = nb * sh.H2 * sh.ipu;
    sh.last_valid = L - (sh.nkv - 1) * KT;
    sh.last_n16 = (sh.last_valid + 15) & ~15;
    const int grid = std::max(1, std::min(sh.n_items, sm_count()));
#ifdef PDM_ATTN_TRACE
    // development build only: per-warp event trace of CTA 0, dumped to $PDM_ATTN_TRACE_FILE after every launch
    static unsigned long long* trace = nullptr;
    if (!trace) {
        PDM_CHECK_CUDA(cudaMalloc(&trace, 12 * 4096 * 8));
        PDM_CHECK_CUDA(cudaMemcpyToSymbol(g_attn_trace, &trace, sizeof(trace)));
    }
    PDM_CHECK_CUDA(cudaMemsetAsync(trace, 0, 12 * 4096 * 8, s));
#endif
    attention_tc3_kernel<<<grid, THREADS, SMEM_BYTES, s>>>(tm, out, sh);
    check_launch("attention_tc3");
#ifdef PDM_ATTN_TRACE
    if (const char* f = getenv("PDM_ATTN_TRACE_FILE")) {
        PDM_CHECK_CUDA(cudaStreamSynchronize(s));
        std::vector<unsigned long long> h(12 * 4096);
        PDM_CHECK_CUDA(cudaMemcpy(h.data(), trace, h.size() * 8, cudaMemcpyDeviceToHost));
        if (FILE* fp = fopen(f, "wb")) {
            fwrite(h.data(), 8, h.size(), fp);
            fclose(fp);
        }
    }
#endif
}

}  // namespace pdm
