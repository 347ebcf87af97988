// tcgen05 flash attention v4 ("column split"): the tc3 kernel (attention_tc3.cu: persistent items, S / P / O in tensor
// memory, lazy rescale, polynomial exp2 share) with SIXTEEN softmax warps instead of eight.  A tile slot t is served by TWO
// warpgroups: warpgroup (t, h) owns key columns [64 h, 64 h + 64) of every score tile -- thread = (query row, half row).  Four
// softmax warps per scheduler instead of two: the softmax loop is bound by the issue cadence of packed-fp32 code with too few
// warps to interleave (profiles/r02_attention_trace.md), not by a pipe.  Costs: the two halves of a row exchange their partial
// row max through shared memory once per tile-step (64-thread named barrier) and their partial row sums once per item; every
// warp pays the per-step barrier traffic for half the work.
//   warps 0-3 / 4-7     slot a, halves 0 / 1        warps 8-11 / 12-15   slot b, halves 0 / 1
//   warp 16 TMA producer, 17 P.V issuer, 18 Q.K^T issuer, 19 idle (setmaxnreg works on whole warpgroups)
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../common.cuh"
#include "../ptx.cuh"

namespace pdm {

CUtensorMap make_tmap_bf16_3d(const void* ptr, long long K, long long rows, long long nbatch, long long bs,
                              int box_rows, int box_k);
namespace {

constexpr int QT = 128, KT = 128, HD = 64;
constexpr int TILE_BYTES = 128 * HD * 2;  // 16 KB
constexpr int NSLOT = 4;                  // K/V ring slots, (K tile | V tile) each (one less than tc3: room for the exchange area)
constexpr int Q_BYTES = 4 * TILE_BYTES;   // 2 item slots x 2 tiles
constexpr int BAR_BYTES = 256;
constexpr int XCH_BYTES = 2 * 2 * 2 * 128 * 4 + 2 * 2 * 128 * 4;  // row-max exchange [parity][slot][half][row] + row-sum exchange [slot][half][row]
constexpr int SMEM_BYTES = 1024 + Q_BYTES + NSLOT * 2 * TILE_BYTES + BAR_BYTES + XCH_BYTES;
constexpr int THREADS = 640;  // warps 0-15 softmax, 16 TMA, 17 P.V, 18 Q.K^T, 19 idle
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t S_COL = 0, P_COL = 256, O_COL = 384;  // + t * {128, 64, 64}
constexpr float RESCALE_LOG2 = 8.f;
#ifndef PDM_ATTN_LOCKSTEP
#define PDM_ATTN_LOCKSTEP 0  // development switch (measured 0.79x: DESIGN.md, "tried and dropped")
#endif
constexpr bool LOCKSTEP = PDM_ATTN_LOCKSTEP != 0;

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
// Walks the items blockIdx.x, blockIdx.x + gridDim.x, ... ; the (row, head pair, index in unit) coordinates advance
// incrementally (three divisions once, none per item).
struct ItemWalk {
    int it, b, hp, r;
    int d_r, d_hp, d_b;
    __device__ __forceinline__ void init(const Shape& sh) {
        it = blockIdx.x;
        const int unit = it / sh.ipu;
        r = it - unit * sh.ipu;
        b = unit / sh.H2;
        hp = unit - b * sh.H2;
        const int g = gridDim.x, du = g / sh.ipu;
        d_r = g - du * sh.ipu;
        d_b = du / sh.H2;
        d_hp = du - d_b * sh.H2;
    }
    __device__ __forceinline__ void next(const Shape& sh) {
        it += gridDim.x;
        r += d_r;
        int c = 0;
        if (r >= sh.ipu) { r -= sh.ipu; c = 1; }
        hp += d_hp + c;
        c = 0;
        if (hp >= sh.H2) { hp -= sh.H2; c = 1; }
        b += d_b + c;
    }
    __device__ __forceinline__ bool done(const Shape& sh) const { return it >= sh.n_items; }
    __device__ __forceinline__ Item get(const Shape& sh) const {
        Item I;
        I.b = b;
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
};

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
#define TRACE_PARAM , int& trace_cnt
#define TRACE_ARG , trace_cnt
#else
#define TRACE(ev, n)
#define TRACE_PARAM
#define TRACE_ARG
#endif

// exp2 on the FMA pipe: Cody-Waite split x = n + f with the round-to-nearest magic-number add, 2^f on [-0.5, 0.5] by a
// degree-3 minimax polynomial (max relative error 7.5e-5, 25x below the bf16 rounding of P), 2^n by an integer add
// into the exponent field.  Packed f32x2 arithmetic.
#ifndef PDM_ATTN_POLY_PAIRS
#define PDM_ATTN_POLY_PAIRS 0x8888  // which of the 16 pairs of a 32-key chunk take the FMA-pipe path (4 of 16)
#endif
constexpr uint32_t POLY_PAIRS = PDM_ATTN_POLY_PAIRS;
__device__ __forceinline__ void exp2_poly2(float a0, float a1, float& p0, float& p1) {
    a0 = fmaxf(a0, -126.f);  // below: exponent underflow of the integer add
    a1 = fmaxf(a1, -126.f);
    const uint64_t a2 = pack_f2(a0, a1);
    const uint64_t t2 = add2(a2, pack_f2(12582912.f, 12582912.f));      // 1.5 * 2^23: low mantissa bits = round(a)
    const uint64_t n2 = add2(t2, pack_f2(-12582912.f, -12582912.f));    // round(a) as float
    const uint64_t f2 = fma2(n2, pack_f2(-1.f, -1.f), a2);              // f = a - round(a)
    uint64_t q2 = fma2(pack_f2(0.055171646f, 0.055171646f), f2, pack_f2(0.24261113f, 0.24261113f));
    q2 = fma2(q2, f2, pack_f2(0.69326097f, 0.69326097f));
    q2 = fma2(q2, f2, pack_f2(0.99992806f, 0.99992806f));
    float t0, t1, q0, q1;
    unpack_f2(t2, t0, t1);
    unpack_f2(q2, q0, q1);
    p0 = __uint_as_float(__float_as_uint(q0) + (__float_as_uint(t0) << 23));
    p1 = __uint_as_float(__float_as_uint(q1) + (__float_as_uint(t1) << 23));
}

// One key tile of one HALF query row: NCH (0..2) chunks of 32 fp32 score columns at s_addr -> registers (then this warp's
// claim on S is released: s_free), partial row max -> exchanged with the warp that owns the other half, lazy rescale of this
// warp's 32 O columns, p = exp2((s - m_ref) * scale * log2 e) -> bf16 P at p_addr (NCH x 16 columns), partial row sum.
// MASK: keys >= nvalid (index inside this half) of the last chunk count as -inf.
template <int NCH, bool MASK>
__device__ __forceinline__ void softmax_half(uint32_t s_addr, uint32_t p_addr, uint32_t o_addr, int nvalid, bool first,
                                             bool wait_prev, float& m_ref, float& l, uint32_t s_free_bar, uint32_t o_full_bar,
                                             uint32_t o_full_parity, int lane, float* xch_mine, const float* xch_other,
                                             int pair_bar TRACE_PARAM) {
    const float cs = 0.125f * 1.4426950408889634f;  // softmax scale * log2(e)
    uint32_t s[NCH > 0 ? NCH : 1][32];
#pragma unroll
    for (int c = 0; c < NCH; ++c) ptx::tmem_ld_32x32(s_addr + c * 32, s[c]);
    if (NCH > 0) ptx::tmem_ld_wait();
    ptx::tc_fence_before();
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(s_free_bar);  // this warp's share of S_t is in registers
    TRACE(6, 0);
    if (MASK && NCH > 0) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if ((NCH - 1) * 32 + i >= nvalid) s[NCH - 1][i] = 0xff800000u;
    }
    float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
            mx4[0] = max3(mx4[0], __uint_as_float(s[c][i]), __uint_as_float(s[c][i + 1]));
            mx4[1] = max3(mx4[1], __uint_as_float(s[c][i + 2]), __uint_as_float(s[c][i + 3]));
            mx4[2] = max3(mx4[2], __uint_as_float(s[c][i + 4]), __uint_as_float(s[c][i + 5]));
            mx4[3] = max3(mx4[3], __uint_as_float(s[c][i + 6]), __uint_as_float(s[c][i + 7]));
        }
    float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
    // the row max over BOTH halves: exchange through shared memory (64-thread named barrier of the two warps of this row quarter)
    xch_mine[lane] = mx;
    asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
    mx = fmaxf(mx, xch_other[lane]);
    TRACE(7, 0);
    if (wait_prev) {
        ptx::mbar_wait(o_full_bar, o_full_parity);  // the previous P.V of this slot read P_t and wrote O_t
        ptx::tc_fence_after();
    }
    TRACE(8, 0);
    if (first) {
        m_ref = mx;
    } else if (__any_sync(0xffffffffu, (mx - m_ref) * cs > RESCALE_LOG2)) {
        // rare: some row outgrew its reference by more than 2^8 (both halves take the same decision: same mx, same m_ref):
        // rescale this warp's partial l and its 32 columns of the O rows
        const float m_new = fmaxf(m_ref, mx);
        const float corr = ex2((m_ref - m_new) * cs);
        uint32_t o[32];
        ptx::tmem_ld_32x32(o_addr, o);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * corr);
        ptx::tmem_st_32x32(o_addr, o);
        l *= corr;
        m_ref = m_new;
    }
    if (NCH > 0) {
        const float nmb = -m_ref * cs;
        const uint64_t cs2 = pack_f2(cs, cs), nmb2 = pack_f2(nmb, nmb);
        uint64_t la = pack_f2(0.f, 0.f), lb = la;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                float a0, a1, p0, p1;
                unpack_f2(fma2(pack_f2(__uint_as_float(s[c][2 * i]), __uint_as_float(s[c][2 * i + 1])), cs2, nmb2), a0, a1);
                if (!MASK && ((POLY_PAIRS >> i) & 1)) {
                    exp2_poly2(a0, a1, p0, p1);
                } else {
                    p0 = ex2(a0);
                    p1 = ex2(a1);
                }
                if (i & 1) lb = add2(lb, pack_f2(p0, p1)); else la = add2(la, pack_f2(p0, p1));
                pk[i] = pack_bf16(p0, p1);
            }
            ptx::tmem_st_32x32_x16(p_addr + c * 16, pk);
        }
        float x0, x1;
        unpack_f2(add2(la, lb), x0, x1);
        l += x0 + x1;
    }
}

__global__ void __launch_bounds__(THREADS, 1)
attention_tc4_kernel(const __grid_constant__ CUtensorMap tmQKV, bf16* __restrict__ out, const Shape sh) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);  // SWIZZLE_128B tiles: 1024 B
    uint8_t* sQ = smem;               // [2 item slots][2 tiles]
    uint8_t* sKV = smem + Q_BYTES;    // [NSLOT] x (K tile | V tile)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Q_BYTES + NSLOT * 2 * TILE_BYTES);
    uint64_t* q_full = bars;                  // [2]      TMA -> Q.K^T warp
    uint64_t* q_empty = bars + 2;             // [2]      Q.K^T commit -> TMA
    uint64_t* kv_full = bars + 4;             // [NSLOT]  TMA -> Q.K^T warp (P.V follows in program order of the data flow)
    uint64_t* kv_empty = bars + 4 + NSLOT;    // [NSLOT]  P.V commit -> TMA
    uint64_t* s_full = bars + 4 + 2 * NSLOT;  // [2]      Q.K^T commit -> softmax warpgroup t
    uint64_t* s_free = s_full + 2;            // [2]      softmax warpgroup t (4 warps) -> Q.K^T warp: S_t is in registers
    uint64_t* p_full = s_free + 2;            // [2]      softmax warpgroup t (4 warps) -> P.V warp
    uint64_t* o_full = p_full + 2;            // [2]      P.V commit -> softmax warpgroup t (P_t consumed, O_t updated)
    uint64_t* o_empty = o_full + 2;           // [2]      softmax warpgroup t (4 warps) -> P.V warp: O_t drained
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + 2);
    float* xch = reinterpret_cast<float*>(smem + Q_BYTES + NSLOT * 2 * TILE_BYTES + BAR_BYTES);  // [2][2][2][128] max | [2][2][128] sum

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
            ptx::mbar_init(&s_full[i], 1);
            ptx::mbar_init(&s_free[i], 8);
            ptx::mbar_init(&p_full[i], 8);
            ptx::mbar_init(&o_full[i], 1);
            ptx::mbar_init(&o_empty[i], 8);
        }
        for (int i = 0; i < NSLOT; ++i) {
            ptx::mbar_init(&kv_full[i], 1);
            ptx::mbar_init(&kv_empty[i], 1);
        }
        ptx::fence_mbar_init();
    }
    if (warp == 17) {
        ptx::tmem_alloc(tmem_slot, TMEM_COLS);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp >= 16) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        // Producer and MMA warps: the WHOLE warp walks the schedule (warp-uniform control flow and operands, so the
        // descriptors live in uniform registers); only the TMA / tcgen05.mma / tcgen05.commit instructions are issued
        // by one elected lane.  (A single-lane region makes the compiler wrap every such instruction in a
        // uniformisation loop of ~17 dependent instructions: ~100 clk per MMA, which starved the tensor pipe.)
        const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint32_t sQ_u = __shfl_sync(0xffffffffu, ptx::smem_u32(sQ), 0);
        const uint32_t sKV_u = __shfl_sync(0xffffffffu, ptx::smem_u32(sKV), 0);
        ItemWalk w;
        w.init(sh);
        if (warp == 16) {
            // ===================== TMA producer =====================
            Ring ring;
            int qi = 0;
            for (; !w.done(sh); w.next(sh)) {
                const Item I = w.get(sh);
                if (I.nt == 0) continue;
                const int qs = qi & 1;
                ptx::mbar_wait_relaxed(&q_empty[qs], ((qi >> 1) & 1) ^ 1);
                if (ptx::elect_one()) {
                    ptx::mbar_expect_tx(&q_full[qs], I.nt * TILE_BYTES);
                    ptx::tma_load_3d(&tmQKV, &q_full[qs], sQ_u + qs * 2 * TILE_BYTES, I.hA * HD, I.qA * QT, I.b);
                    if (I.nt == 2)
                        ptx::tma_load_3d(&tmQKV, &q_full[qs], sQ_u + qs * 2 * TILE_BYTES + TILE_BYTES, I.hB * HD, I.qB * QT, I.b);
                }
                __syncwarp();
                const int nsrc = (I.nt == 2 && !I.same) ? 2 : 1;
                for (int j = 0; j < nkv; ++j) {
                    for (int u = 0; u < nsrc; ++u) {
                        const int h = u ? I.hB : I.hA;
                        ring.next();
                        ptx::mbar_wait_relaxed(&kv_empty[ring.slot], ring.ph ^ 1);
                        const uint32_t sk = sKV_u + ring.slot * 2 * TILE_BYTES;
                        if (ptx::elect_one()) {
                            ptx::mbar_expect_tx(&kv_full[ring.slot], 2 * TILE_BYTES);
                            ptx::tma_load_3d(&tmQKV, &kv_full[ring.slot], sk, D + h * HD, j * KT, I.b);
                            ptx::tma_load_3d(&tmQKV, &kv_full[ring.slot], sk + TILE_BYTES, 2 * D + h * HD, j * KT, I.b);
                        }
                        __syncwarp();
                    }
                }
                ++qi;
            }
        } else if (warp == 18) {
            // ===================== S_t = Q.K^T issuer =====================
            const uint32_t idesc_qk_full = ptx::make_idesc_bf16(QT, KT, 0, 0);
            const uint32_t idesc_qk_last = ptx::make_idesc_bf16(QT, sh.last_n16, 0, 0);
            Ring ring;
            int qi = 0;
            uint32_t cnt[2] = {0, 0};  // tile-steps issued so far per tile slot
            for (; !w.done(sh); w.next(sh)) {
                const Item I = w.get(sh);
                if (I.nt == 0) continue;
                const int qs = qi & 1;
                ptx::mbar_wait(&q_full[qs], (qi >> 1) & 1);
                for (int j = 0; j < nkv; ++j) {
                    for (int t = 0; t < I.nt; ++t) {
                        if (!(I.same && t == 1)) {  // first user of a K/V slot
                            ring.next();
                            ptx::mbar_wait(&kv_full[ring.slot], ring.ph);
                        }
                        const uint64_t qdesc = ptx::make_smem_desc_sw128(sQ_u + (qs * 2 + t) * TILE_BYTES, 1024);
                        const uint64_t kdesc = ptx::make_smem_desc_sw128(sKV_u + ring.slot * 2 * TILE_BYTES, 1024);
                        const uint32_t d = tb + S_COL + t * 128;
                        const uint32_t idesc = j == nkv - 1 ? idesc_qk_last : idesc_qk_full;
                        const bool item_done = j == nkv - 1 && t == I.nt - 1;
                        const uint32_t c = cnt[t]++;
                        TRACE(14, c);
                        if (c > 0) ptx::mbar_wait(&s_free[t], (c - 1) & 1);  // the previous S_t sits in registers
                        ptx::tc_fence_after();
                        TRACE(15, c);
                        if (ptx::elect_one()) {
#pragma unroll
                            for (int kk = 0; kk < HD / 16; ++kk)
                                ptx::mma_bf16_ss(d, qdesc + 2 * kk, kdesc + 2 * kk, idesc, kk != 0);
                            ptx::mma_commit(&s_full[t]);
                            if (item_done) ptx::mma_commit(&q_empty[qs]);  // all Q.K^T of the item issued: Q tiles are free
                        }
                        __syncwarp();
                        TRACE(10, c);
                    }
                }
                ++qi;
            }
        } else if (warp == 17) {
            // ===================== O_t += P_t.V issuer =====================
            constexpr uint32_t idesc_pv = ptx::make_idesc_bf16(QT, HD, 0, 1);  // A = P (TMEM), B = V MN-major
            const int ksteps_last = sh.last_n16 / 16;
            Ring ring;
            uint32_t cnt[2] = {0, 0};     // tile-steps issued so far per tile slot
            uint32_t o_uses[2] = {0, 0};  // items that have used O_t so far
            for (; !w.done(sh); w.next(sh)) {
                const Item I = w.get(sh);
                if (I.nt == 0) continue;
                for (int j = 0; j < nkv; ++j) {
                    for (int t = 0; t < I.nt; ++t) {
                        if (!(I.same && t == 1)) ring.next();
                        // V tile: rows = keys (K dim), 128 bytes of head-dim per row (N contiguous) -> MN-major; 8-key
                        // groups are 1024 bytes apart; one UMMA_K step (16 keys) = 2048 bytes.  P: 16 bf16 = 8 TMEM columns.
                        const uint64_t vdesc =
                            ptx::make_smem_desc_sw128(sKV_u + ring.slot * 2 * TILE_BYTES + TILE_BYTES, 1024, 1024);
                        const bool last_tile = j == nkv - 1;
                        uint64_t* kv_bar = !(I.same && t == 0) ? &kv_empty[ring.slot] : nullptr;  // last user of the slot
                        const uint32_t d_o = tb + O_COL + t * 64, a_p = tb + P_COL + t * 64;
                        const uint32_t acc0 = j != 0;
                        if (j == 0) {  // O_t is rewritten: the previous item's epilogue must have drained it
                            ptx::mbar_wait(&o_empty[t], (o_uses[t] & 1) ^ 1);
                            ++o_uses[t];
                        }
                        const uint32_t c = cnt[t]++;
                        TRACE(11, c);
                        ptx::mbar_wait(&p_full[t], c & 1);
                        ptx::tc_fence_after();
                        TRACE(12, c);
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
                            ptx::mma_commit(&o_full[t]);
                            if (kv_bar) ptx::mma_commit(kv_bar);
                        }
                        __syncwarp();
                        TRACE(13, c);
                    }
                }
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");  // pool = 640 x 96 launch registers: 512 x 104 + 128 x 56 fits, 112 would block forever
        // ===================== softmax warpgroups: thread <-> (query row, half of the key columns) =====================
        const int t = warp >> 3;         // tile slot
        const int hf = (warp >> 2) & 1;  // which 64 key columns of every score tile
        const int wq = warp & 3;         // TMEM lane quarter = 32 query rows
        const uint32_t lane_base = uint32_t(wq * 32) << 16;
        const uint32_t s_addr = tmem_base + lane_base + S_COL + t * 128 + hf * 64;
        const uint32_t p_addr = tmem_base + lane_base + P_COL + t * 64 + hf * 32;
        const uint32_t o_addr = tmem_base + lane_base + O_COL + t * 64 + hf * 32;
        // 32-key chunks of this half in the ragged last key tile, and the valid keys inside the half
        const int nch_all = (sh.last_n16 + 31) >> 5;
        const int nch_last = min(2, max(0, nch_all - 2 * hf));
        const int nv_half = sh.last_valid - 64 * hf;
        uint32_t steps = 0;  // tile-steps of slot t completed so far (phase counter of s_full / o_full)
        const uint32_t b_s_full = ptx::smem_u32(&s_full[t]), b_s_free = ptx::smem_u32(&s_free[t]);
        const uint32_t b_p_full = ptx::smem_u32(&p_full[t]), b_o_full = ptx::smem_u32(&o_full[t]);
        const uint32_t b_o_empty = ptx::smem_u32(&o_empty[t]);
        const int pair_bar = 1 + t * 4 + wq;  // named barrier of the two warps (halves 0 / 1) of this slot and row quarter
        float* xmax = xch;                    // [parity][slot][half][128]
        float* xsum = xch + 2 * 2 * 2 * 128;  // [slot][half][128]
        float* xsum_mine = xsum + (t * 2 + hf) * 128 + wq * 32;
        const float* xsum_other = xsum + (t * 2 + (hf ^ 1)) * 128 + wq * 32;

        struct Pending {
            bool any = false, live = false, row_ok = false;
            uint32_t parity = 0;
            float l = 0.f;
            bf16* dst = nullptr;
        } pend;
        auto flush_epilogue = [&]() {  // this warp's 32 columns of O / l -> bf16 rows of the pending item
            ptx::mbar_wait(b_o_full, pend.parity);  // the item's last P.V has landed
            ptx::tc_fence_after();
            TRACE(4, 0);
            // total row sum = both halves
            xsum_mine[lane] = pend.l;
            asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
            const float inv = 1.f / (pend.l + xsum_other[lane]);
            if (pend.live) {
                uint32_t v[32];
                ptx::tmem_ld_32x32(o_addr, v);
                ptx::tmem_ld_wait();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(b_o_empty);  // O_t may be overwritten by the next item
                if (pend.row_ok) {
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        uint32_t pk[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e)
                            pk[e] = pack_bf16(__uint_as_float(v[16 * q + 2 * e]) * inv, __uint_as_float(v[16 * q + 2 * e + 1]) * inv);
                        asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                                     ::"l"(pend.dst + q * 16), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]), "r"(pk[4]), "r"(pk[5]),
                                     "r"(pk[6]), "r"(pk[7])
                                     : "memory");
                    }
                }
            } else {
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(b_o_empty);
            }
            // the exchange slot is reused by the next item's epilogue: both warps must have read it
            asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
            pend.any = false;
            TRACE(5, 0);
        };

        ItemWalk w;
        for (w.init(sh); !w.done(sh); w.next(sh)) {
            int nt, q0, h, b;
            {
                const Item I = w.get(sh);
                nt = I.nt; q0 = (t ? I.qB : I.qA) * QT; h = t ? I.hB : I.hA; b = I.b;
            }
            if (nt == 0) continue;
            if (t < nt) {
                const bool live = q0 + wq * 32 < sh.L;  // does this warp own any real query row?
                float m_ref = 0.f, l = 0.f;
                for (int j = 0; j < nkv; ++j) {
                    TRACE(0, steps);
                    ptx::mbar_wait(b_s_full, steps & 1);
                    ptx::tc_fence_after();
                    TRACE(1, steps);
                    const bool wait_prev = j > 0 || pend.any;
                    const uint32_t ofp = (steps - 1) & 1;
                    float* xm_mine = xmax + (((steps & 1) * 2 + t) * 2 + hf) * 128 + wq * 32;
                    const float* xm_other = xmax + (((steps & 1) * 2 + t) * 2 + (hf ^ 1)) * 128 + wq * 32;
                    if (live) {
                        if (j < nkv - 1) {
                            softmax_half<2, false>(s_addr, p_addr, o_addr, 64, j == 0, wait_prev, m_ref, l, b_s_free, b_o_full, ofp, lane,
                                                   xm_mine, xm_other, pair_bar TRACE_ARG);
                        } else {
                            switch (nch_last) {
                                case 0: softmax_half<0, true>(s_addr, p_addr, o_addr, nv_half, j == 0, wait_prev, m_ref, l, b_s_free, b_o_full, ofp, lane, xm_mine, xm_other, pair_bar TRACE_ARG); break;
                                case 1: softmax_half<1, true>(s_addr, p_addr, o_addr, nv_half, j == 0, wait_prev, m_ref, l, b_s_free, b_o_full, ofp, lane, xm_mine, xm_other, pair_bar TRACE_ARG); break;
                                default: softmax_half<2, true>(s_addr, p_addr, o_addr, nv_half, j == 0, wait_prev, m_ref, l, b_s_free, b_o_full, ofp, lane, xm_mine, xm_other, pair_bar TRACE_ARG); break;
                            }
                        }
                        TRACE(2, steps);
                        ptx::tmem_st_wait();
                    } else {
                        if (lane == 0) ptx::mbar_arrive(b_s_free);
                        if (wait_prev) ptx::mbar_wait(b_o_full, ofp);  // keep in step with o_full
                    }
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(b_p_full);
                    TRACE(3, steps);
                    ++steps;
                    if (j == 0 && pend.any) flush_epilogue();
                }
                const int qi = q0 + wq * 32 + lane;
                pend.any = true;
                pend.live = live;
                pend.row_ok = qi < sh.L;
                pend.parity = (steps - 1) & 1;
                pend.l = l;
                pend.dst = out + ((long long)b * sh.L + qi) * D + h * HD + hf * 32;
            } else if (pend.any) {
                flush_epilogue();  // this tile slot sits the item out (odd head count)
            }
        }
        if (pend.any) flush_epilogue();
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 17) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

}  // namespace

void attention_tc4_bf16(const bf16* qkv, bf16* out, int nb, int L, int H, cudaStream_t s) {
    const int D = H * HD;
    const CUtensorMap tm = make_tmap_bf16_3d(qkv, 3LL * D, L, nb, L, 128, HD);
    static std::atomic<bool> attr_set[MAX_DEVICES];
    ensure_dyn_smem(attention_tc4_kernel, SMEM_BYTES, attr_set);
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
    attention_tc4_kernel<<<grid, THREADS, SMEM_BYTES, s>>>(tm, out, sh);
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
