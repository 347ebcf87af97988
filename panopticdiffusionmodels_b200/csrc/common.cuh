// Shared host/device helpers for libpdm.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <stdexcept>
#include <string>

namespace pdm {

typedef __nv_bfloat16 bf16;

struct Error : public std::runtime_error {
    explicit Error(const std::string& s) : std::runtime_error(s) {}
};

#define PDM_CHECK_CUDA(expr)                                                                              \
    do {                                                                                                  \
        cudaError_t _e = (expr);                                                                          \
        if (_e != cudaSuccess)                                                                            \
            throw ::pdm::Error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " (" __FILE__ ":" + \
                               std::to_string(__LINE__) + ")");                                           \
    } while (0)

#define PDM_REQUIRE(cond, msg)                                                  \
    do {                                                                        \
        if (!(cond)) throw ::pdm::Error(std::string("pdm: ") + (msg));          \
    } while (0)

void set_last_error(const std::string& msg);  // thread-local message behind pdm_last_error() (engine.cu)
extern std::atomic<long long> g_launch_count;
inline void count_launch() { g_launch_count.fetch_add(1, std::memory_order_relaxed); }
inline void check_launch(const char* what) {
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) throw Error(std::string(what) + " launch failed: " + cudaGetErrorString(e));
}

// Per-device state: function attributes (the opt-in for > 48 KB of dynamic shared memory) and device properties belong to
// a device / context, not to the process -- a module moved to a second GPU in the same process needs them again.
constexpr int MAX_DEVICES = 64;
inline int current_device() {
    int dev = 0;
    PDM_CHECK_CUDA(cudaGetDevice(&dev));
    PDM_REQUIRE(dev >= 0 && dev < MAX_DEVICES, "device ordinal out of range");
    return dev;
}
inline int sm_count() {
    static std::atomic<int> n[MAX_DEVICES];
    const int dev = current_device();
    int v = n[dev].load(std::memory_order_relaxed);
    if (v == 0) {
        PDM_CHECK_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
        n[dev].store(v, std::memory_order_relaxed);
    }
    return v;
}
// Opt a kernel in to `bytes` of dynamic shared memory once per device.  `done` is a per-kernel flag array.
template <typename K>
inline void ensure_dyn_smem(K kernel, int bytes, std::atomic<bool>* done) {
    const int dev = current_device();
    if (!done[dev].load(std::memory_order_acquire)) {
        PDM_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        done[dev].store(true, std::memory_order_release);
    }
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

// Row addressing of a logically [nb * Lr, width] matrix living inside a [nb, bs, width] buffer:
// logical row m = b * Lr + t  ->  physical row b * bs + t.
struct RowMap {
    int Lr;  // logical rows per batch
    int bs;  // physical rows per batch (>= Lr)
};

// ---- GEMM interface shared by the SIMT fp32 kernel and the tcgen05 bf16 kernel ----
// out = [gelu]( A1[:, :K1] . W[:, :K1]^T + A2[:, :K2] . W[:, K1:]^T + bias ) [+ resid]
//   A*: activation dtype (fp32 in PDM_PREC_FP32, bf16 in PDM_PREC_BF16), row-major, width K*
//   W : [N, K1+K2] row-major (nn.Linear layout), fp32 master + bf16 copy
//   out32 (fp32, may alias resid) and/or out2 (activation dtype)
struct GemmProblem {
    const void* A1 = nullptr;
    const void* A2 = nullptr;
    int K1 = 0, K2 = 0;
    int a1_bs = 0, a2_bs = 0;  // physical rows per batch of A1 / A2
    const float* W32 = nullptr;
    const bf16* W16 = nullptr;
    const float* bias = nullptr;
    int N = 0;
    int nb = 1, Lr = 0;  // logical rows = nb * Lr
    const float* resid = nullptr;
    int resid_bs = 0;
    float* out32 = nullptr;
    int out32_bs = 0;
    float* out32b = nullptr;  // optional second copy of out32 (tcgen05 kernel only)
    int out32b_bs = 0;
    void* out2 = nullptr;
    int out2_bs = 0;
    bool gelu = false;
    // ---- deferred LayerNorm (tcgen05 kernel only; see gemm_tc.cu) ----
    // producer side (fp32-output form): a second activation-dtype copy for rows >= out2b_row0 of every batch, and the
    // per-row partial sums (sum, sum of squares) over each 128-column slice: stats[row][ceil(N/128)][2]
    void* out2b = nullptr;
    int out2b_bs = 0, out2b_row0 = 0, out2b_mod = 0;  // rows with (row % out2b_mod if out2b_mod else row) >= out2b_row0
    // fp32 store filter: out32 is written only for rows with (row % out32_mod if out32_mod else row) >= out32_row0 (0 = all
    // rows).  For a residual GEMM whose fp32 result nobody reads (the next block starts with a long-skip GEMM fed by the bf16
    // copies, or the head follows): the value is still formed (residual read, bf16 copies, row sums), only the dead store goes.
    int out32_row0 = 0, out32_mod = 0;
    float* stats = nullptr;
    int stats_bs = 0;
    float* statsb = nullptr;  // same values, second destination (rows of the concatenated mask stream)
    int statsb_bs = 0;
    // consumer side (bf16-output form): A1 holds the RAW rows bf16(x); W16 holds the LN-folded weight of fold_ln_weight
    // (gamma * W, centred along K) and `bias` holds bias + W.beta, so that out = rstd * acc + bias == LN(x).W^T + b
    const float* ln_rstd = nullptr;  // [rows] 1 / sqrt(var + eps) of x, precomputed by the caller
    int ln_rstd_bs = 0;
    // ... or the producer's partial row sums themselves: [rows][ceil(ln_D / 128)][2]; the epilogue then forms
    // 1 / sqrt(E[x^2] - E[x]^2 + 1e-5) itself (no separate kernel between the GEMMs)
    const float* ln_stats = nullptr;
    int ln_stats_bs = 0, ln_D = 0;
    // plain fp32-output form only (no residual, no copies): ln_rstd is honoured there too (out = rstd * acc + bias), and
    // rowbias [Lr, N] (fp32) is added per ROW OF THE BATCH after the bias (positional embedding of the patch-embed GEMM)
    const float* rowbias = nullptr;
    // ---- implicit-GEMM 3x3 convolution, stride 1, zero padding 1 (tcgen05 kernel only; VAE decoder, vae.cu) ----
    // conv_H > 0: A1 is an NHWC activation [conv_N, conv_H, conv_W, conv_C] (bf16), rows = output pixels in (n, h, w) order
    // (Lr = conv_N * conv_H * conv_W, nb = 1), K1 = 9 * conv_C with k = (ky * 3 + kx) * conv_C + c, W16 = [N, 9 * conv_C].
    // The A tile of tap (ky, kx) is ONE shifted TMA box of the activation; the border comes from the out-of-bounds zero fill.
    int conv_N = 0, conv_H = 0, conv_W = 0, conv_C = 0;
    // conv_up = 1 + 2 a + b (0 = off): phase (a, b) of "nearest 2x upsample -> 3x3 convolution" as a 2x2 convolution over the
    // LOW-resolution activation: K1 = 4 * conv_C with k = (dy * 2 + dx) * conv_C + c reading source pixel (y + dy + a - 1,
    // x + dx + b - 1); row (n, y, x) of the result is written to row (n, 2 y + a, 2 x + b) of out32 [conv_N, 2 H, 2 W, N].
    int conv_up = 0;
    // GroupNorm statistics of the OUTPUT (fp32 plain / += / upsample-phase forms, nb = 1, rows = pixels of whole images):
    // per (image, 32-row slab, group) partial (sum, sum of squares) in a fixed order -> gn_part[img][blk][32 groups][2],
    // blk = slab_in_image * gn_stride + gn_slot0, gn_nblk slabs-times-stride per image.  N = 32 * channels-per-group.
    float* gn_part = nullptr;
    int gn_hw = 0, gn_stride = 1, gn_slot0 = 0;
};

void gemm_simt_f32(const GemmProblem& p, cudaStream_t s);
void gemm_tc_bf16(const GemmProblem& p, cudaStream_t s);

// ---- attention: qkv [nb, L, 3D] (q | k | v, each H heads x 64) -> out [nb, L, D] ----
void attention_simt(const void* qkv, void* out, int nb, int L, int H, bool is_bf16, cudaStream_t s);
void attention_tc_bf16(const bf16* qkv, bf16* out, int nb, int L, int H, cudaStream_t s);
#ifdef PDM_ATTN_EXPERIMENTS
void attention_tc4_bf16(const bf16* qkv, bf16* out, int nb, int L, int H, cudaStream_t s);  // experiments/attention_tc4.cu (measured slower; not in the product build)
#endif

// ---- bandwidth-bound kernels (elementwise.cu) ----
void layernorm(const float* x, const float* w, const float* b, void* out, bool out_bf16, long long rows, int D,
               cudaStream_t s);
void convert_f32_bf16(const float* in, bf16* out, long long n, cudaStream_t s);
// deferred LayerNorm helpers: raw bf16 copy + per-row partial sums of an fp32 stream; weight folding
void rowstats_convert(const float* x, bf16* out, float* stats, long long rows, int D, cudaStream_t s);
void fold_ln_weight(const float* W, const float* bias, const float* gamma, const float* beta, bf16* Wf, float* d, int N,
                    int K, bool for_gelu, cudaStream_t s);
constexpr int LN_PART = 128;  // columns per partial-sum slice (= accumulator columns per GEMM epilogue warp)
void copy_rows(void* dst, int dst_bs, const void* src, int src_bs, int Lr, int nb, int row_bytes, cudaStream_t s);

struct EmbedArgs {
    const float* img;      // [Bx, C, S, S]
    const float* mask;     // [Bx, Cm, S, S] or null
    int Bx;                // input batch (rows b use input b % Bx)
    int nb;                // output batch
    const float* t_dev;    // [Bx] or null
    float t_scalar;
    const float* freqs;    // [D/2]
    const float* ctxtok;   // [nb, T, D] (bias already added)
    const float* wT_img;   // [C*p*p, D]  patch_embed.proj.weight transposed (engine builds it at finalize)
    const float* wT_msk;   // [Cm*p*p, D]
    const float* w_img;    // [D, C*p*p]
    const float* b_img;
    const float* w_msk;    // [D, Cm*p*p]
    const float* b_msk;
    const float* pos;      // [ext + P (+P), D]
    const float* pos_m;    // positional table of the mask tokens, indexed by patch
    float* out_x;          // [nb, Lx, D]
    int Lx;
    float* out_m;          // [nb, Lm, D] mask tokens written at token offset m_off
    int Lm, m_off;
    int C, Cm, S, p, D, T;
    // bf16 engine path (embed_extras): the time / context rows also leave their bf16 copy and LayerNorm row sums, and (two-stream)
    // the same three things in the mask stream's buffers -- the concat of libs/uvit_t2i.py:427 and the first row-statistics pass
    bf16* xb = nullptr;        // [nb, Lx, D]
    float* stats = nullptr;    // [nb, Lx, ceil(D / 128), 2]
    float* out_x2 = nullptr;   // [nb, L2rows, D] (null: no mirror)
    bf16* xb2 = nullptr;
    float* stats2 = nullptr;
    int L2rows = 0;
};
void embed_tokens(const EmbedArgs& a, cudaStream_t s);
void transpose_f32(const float* in, float* out, int R, int C, cudaStream_t s);  // [R, C] -> [C, R]

struct HeadArgs {
    const float* x;       // [nb, Lx, D] image tokens at offset x_off
    int Lx, x_off;
    const float* m;       // [nb, Lm, D] mask tokens at offset m_off (null: image only)
    int Lm, m_off;
    bool ln_m;            // apply the final LayerNorm to the mask tokens too (single-stream)
    bool gt;              // use_ground_truth (libs/uvit_t2i.py:486-496): decode image + mask features, no mask decoder
    const float* ln_w;
    const float* ln_b;
    const float* w_dec;   // [p*p*C, D]
    const float* b_dec;
    const float* w_decm;  // [p*p*Cm, D]
    const float* b_decm;
    const float* w_fin;   // [C, C, 3, 3]
    const float* b_fin;
    const float* w_finm;  // [Cm, Cm, 3, 3]
    const float* b_finm;
    float* tmp_img;       // [nb, C, S, S]
    float* tmp_msk;       // [nb, Cm, S, S]
    float* out_img;       // [nb, C, S, S]
    float* out_msk;       // [nb, Cm, S, S]
    int nb, C, Cm, S, p, D;
};
void head_decode(const HeadArgs& a, cudaStream_t s);
// bf16 engine path: patches -> bf16 GEMM operand rows, and the 3x3 heads on token-major decoder outputs
void im2col_patches(const float* img, bf16* out, int Bx, int nb, int C, int S, int p, cudaStream_t s);
void embed_extras(const EmbedArgs& a, cudaStream_t s);
void conv3x3_tokens(const float* tok, const float* w, const float* bias, float* out, int nb, int C, int S, int p, int do_tanh,
                    cudaStream_t s);

struct UpdateArgs {
    const float* eps_c;
    const float* eps_u;
    const float* pm_c;
    const float* pm_u;
    const float* x_in;
    const float* x_base;
    float* X0;
    float* x_out;
    const float* m_base;
    float* P0;
    float* m_out;
    float alpha, sigma, A, B_img, C_img, B_msk, C_msk, scale;
    float A_msk = 0.f;   // mask_plain: m_out = A_msk * m_base + B_msk * P0 (enable_mask_opt=False pass-through)
    int stage, has_c;
    int mask_plain = 0;
    long long n_img, n_mask;
};
void cfg_solver_update(const UpdateArgs& a, cudaStream_t s);

struct MultistepArgs {
    const float* eps_c;
    const float* eps_u;
    const float* pm_c;
    const float* pm_u;
    const float* x;       // state the network was evaluated at (time t_0)
    const float* X1;      // data prediction at t_{-1}
    const float* X2;      // data prediction at t_{-2}
    float* X0;            // out: data prediction at t_0
    float* x_out;
    const float* m;
    const float* P1;
    const float* P2;
    float* P0;
    float* m_out;
    float alpha, sigma, A, B, C1, C2, inv_r0, inv_r1, q, inv_r01, halfB, scale;
    int order;
    long long n_img, n_mask;
};
void multistep_update(const MultistepArgs& a, cudaStream_t s);

// host planner (plan.cu)
int solver_plan(const float* betas, int n_betas, int steps, int order, int method, int skip_type, float eps, float T,
                int mask_opt, float n_time, float* out, int cap, int* n_evals);
const char* plan_last_error();

void bits2int(const float* pm, int32_t* labels, int B, int nbits, int hw, cudaStream_t s);
void int2bits(const int32_t* ids, float* bits, int B, int nbits, int hw, cudaStream_t s);

}  // namespace pdm
