/*
 * pdm.h -- C ABI of libpdm.so: the B200 (sm_100a) joint image+mask sampling hot path of
 * Panoptic Diffusion Models (U-ViT t2i forward wrapped in the DPM-Solver++ data-prediction loop).
 *
 * The reference is pure Python (no FFI of its own); each entry point below names the reference
 * function it replaces (paths relative to the reference tree).  INTEGRATION.md shows the ctypes
 * binding a maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; pdm_last_error() gives the message
 *     (thread-local).
 *   - all tensor pointers are DEVICE pointers to contiguous float32 unless stated otherwise and are
 *     owned by the caller; the engine owns its parameters and a private workspace.
 *   - all device work is enqueued on the cudaStream_t passed as `stream` (void*; NULL = legacy
 *     default stream); no function synchronises the stream except where noted.
 *   - a handle is thread-compatible (one thread at a time), one handle per GPU/process.
 *   - there is NO CPU path: every entry point that computes needs a CUDA device.
 */
#ifndef PDM_H_
#define PDM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PDM_ABI_VERSION 2

/* precision modes of the network evaluation */
#define PDM_PREC_BF16 0 /* bf16 operands on tcgen05 tensor cores, fp32 accumulate, fp32 residual stream */
#define PDM_PREC_FP32 1 /* fp32 everywhere (parity anchor: <=1e-3 max-rel vs the reference forward) */

/* number of floats per model evaluation in a solver plan (see pdm_sample) */
#define PDM_PLAN_STRIDE 16
/* plan record layout (float32 each):
 *  [0] t_model   time fed to the network (= 1000 * t_continuous, train_t2i_discrete.py:508)
 *  [1] alpha     alpha(t_eval)   [2] sigma   sigma(t_eval)      (dpm_solver_pp.py:313-316)
 *  [3] A         sigma(t_next)/sigma(s)                coefficient of the step-start state
 *  [4] B_img     signed coefficient of X_0 (image)     [5] C_img  signed coefficient of (X_j - X_0)
 *  [6] B_msk     signed coefficient of P_0 (mask)      [7] C_msk  signed coefficient of (P_j - P_0)
 *  [8] stage     0,1,2 : index of this evaluation inside its solver step
 *  [9] has_c     1 if the (X_j - X_0) term is present (stage > 0)
 *  [10] last     1 if this evaluation closes its solver step (output becomes the next step-start state)
 *  [11] A_msk    [12] split: when split == 1 the mask stream is the reference's enable_mask_opt=False pass-through,
 *                m_out = A_msk * m_base + B_msk * P_0 (no difference term): the intermediate evaluations see the step-start
 *                mask, the next step starts from the prediction itself (dpm_solver_pp.py:441-457, 536-557, 730-766)
 *  [13..14] reserved (0)
 *  [15] kind     0: singlestep record (above); 1: multistep record (layout: see pdm_multistep_update)
 */

/* pdm_solver_plan: method / skip_type codes (dpm_solver_pp.py:927-930 `method`, `skip_type`) */
#define PDM_METHOD_FAST 0        /* 'fast': singlestep orders 3,...,3,2 / 3,...,3,1 / 3,...,3,2,1 (dpm_solver_pp.py:386-395) */
#define PDM_METHOD_SINGLESTEP 1  /* 'singlestep': steps // order updates of one order */
#define PDM_METHOD_MULTISTEP 2   /* 'multistep': 2M / 3M with lower-order warm-up */
#define PDM_SKIP_TIME_UNIFORM 0
#define PDM_SKIP_LOGSNR 1
#define PDM_SKIP_T2 2 /* 't2' (dpm_solver_pp.py:351-354); the reference's 10^7-point 'time_quadratic' search is not offered */

typedef struct pdm_engine* pdm_handle;

/* Mirrors the keyword arguments of libs/uvit_t2i.py:259-261 (UViT.__init__) that shape the graph. */
typedef struct pdm_config {
    int32_t img_size;           /* latent H = W (32 @256px, 64 @512px) */
    int32_t patch_size;         /* 2 */
    int32_t in_chans;           /* 4 */
    int32_t embed_dim;          /* D, multiple of 64 */
    int32_t depth;              /* even */
    int32_t num_heads;          /* D / 64 (head dim must be 64) */
    int32_t mlp_ratio;          /* 4 */
    int32_t clip_dim;           /* 768 */
    int32_t num_clip_token;     /* 77 */
    int32_t num_panoptic_class; /* 8 analog bits */
    int32_t enable_panoptic;    /* 1: mask_embed / decoder_pred_mask exist */
    int32_t separate;           /* 1: two-stream topology (mask blocks + zero-conv bridges) */
} pdm_config;

/* replaces: utils.py:291-299 get_nnet / UViT.__init__ (libs/uvit_t2i.py:259-353) */
int pdm_create(const pdm_config* cfg, pdm_handle* out);
int pdm_destroy(pdm_handle h);

/* replaces: nn.Module.load_state_dict (eval_t2i_discrete.py:51).  `key` is the reference state_dict
 * key, `dev_f32` a device float32 tensor of the given shape; the engine copies and
 * repacks it (fp32 master + bf16 GEMM operand).  Unknown key -> error.  Dead-but-present keys
 * (mask_embed_0.*, even-indexed zero_convs.*) are accepted and ignored. */
int pdm_set_param(pdm_handle h, const char* key, const void* dev_f32, const int64_t* shape, int32_t ndim,
                  void* stream);
/* checks that every parameter the graph needs was set; builds the derived operands (bf16 copies, the LayerNorm-folded
 * qkv / fc1 weights of the deferred-LayerNorm path, transposed patch-embed weights).  Must be called again after any
 * pdm_set_param: evaluations are refused in between.  Synchronises `stream`. */
int pdm_finalize_params(pdm_handle h, void* stream);

/* bytes of private workspace the engine will hold for `n` network rows (n = samples in one forward) */
int pdm_workspace_bytes(pdm_handle h, int32_t n, int32_t precision, size_t* bytes);

/* replaces: UViT.forward (libs/uvit_t2i.py:378-525).
 *   x [n,in_chans,H,W]; t [n] (model time, e.g. 0..1000); ctx [n,num_clip_token,clip_dim];
 *   mask [n,num_panoptic_class,H,W] or NULL (image-only path, uvit_t2i.py:407-410);
 *   out_noise [n,in_chans,H,W]; out_mask [n,num_panoptic_class,H,W] (required iff mask != NULL). */
int pdm_nnet_forward(pdm_handle h, const float* x, const float* t, const float* ctx, const float* mask,
                     float* out_noise, float* out_mask, int32_t n, int32_t precision, void* stream);

/* pdm_nnet_forward with option flags.  PDM_FWD_GROUND_TRUTH = the reference's `use_ground_truth=True` evaluation
 * (libs/uvit_t2i.py:380, 486-496): the noise is decoded from image feature + mask feature (mask tokens normalised only in
 * the single-stream topology), the mask decoder is skipped and out_mask receives the mask that was passed in.  This is
 * the evaluation the second phase of DPM_Solver.sample(use_twophases=True) runs (dpm_solver_pp.py:1071-1075). */
#define PDM_FWD_GROUND_TRUTH 1
int pdm_nnet_forward_ex(pdm_handle h, const float* x, const float* t, const float* ctx, const float* mask, float* out_noise,
                        float* out_mask, int32_t n, int32_t precision, int32_t flags, void* stream);

/* replaces: cfg_nnet's guidance combine (train_t2i_discrete.py:429-431) + DPM_Solver.model_fn's
 * eps->x0 (dpm_solver_pp.py:316) + one singlestep linear update (dpm_solver_pp.py:444-456, 529-555,
 * 724-764), fused in ONE elementwise kernel.  `coef` is a HOST pointer to one plan record.
 *   eps_c/eps_u [B,4hw] (eps_u NULL -> no guidance), pm_c/pm_u [B,8hw] (NULL -> no mask stream),
 *   x_in  state the network was evaluated at; x_base step-start state (== x_in at stage 0);
 *   X0/P0 data-prediction history of stage 0 (written at stage 0, read later);
 *   x_out/m_out next state.  m_in is not needed (the mask "x0" is the prediction itself).
 *   coef[12] != 0 selects the pass-through mask update m_out = coef[11] * m_base + coef[6] * P0 (see the record layout). */
int pdm_cfg_update(const float* eps_c, const float* eps_u, const float* pm_c, const float* pm_u,
                   const float* x_in, const float* x_base, float* X0, float* x_out,
                   const float* m_base, float* P0, float* m_out,
                   const float* coef, float cfg_scale, int64_t n_img, int64_t n_mask, void* stream);

/* replaces: dpm_multistep_update / dpm_multistep_second_update / dpm_multistep_third_update (dpm_solver_pp.py:602-677,
 * 852-871; data prediction, solver_type='dpm_solver') fused with the guidance combine and eps->x0, ONE kernel per step.
 * `coef` (HOST, PDM_PLAN_STRIDE floats): [0] t_model [1] alpha(t_0) [2] sigma(t_0) [3] A=sigma_t/sigma_0 [4] B
 * [5] C1 [6] C2 [7] 1/r0 [8] 1/r1 [9] r0/(r0+r1) [10] 1/(r0+r1) [11] order (1,2,3) [12] 0.5*B.
 *   x: state the network was evaluated at (t_0); X1/X2: cached data predictions at t_-1/t_-2 (NULL if order is lower);
 *   X0 (out): data prediction at t_0; x_out: state at t.  The mask stream (m, P1, P2 -> P0, m_out) gets the same
 *   update with the mask prediction as its data prediction (no reference behaviour exists: parity unpinned). */
int pdm_multistep_update(const float* eps_c, const float* eps_u, const float* pm_c, const float* pm_u, const float* x,
                         const float* X1, const float* X2, float* X0, float* x_out, const float* m, const float* P1,
                         const float* P2, float* P0, float* m_out, const float* coef, float cfg_scale, int64_t n_img,
                         int64_t n_mask, void* stream);

/* replaces: the host-side scalar work of DPM_Solver.sample -- NoiseScheduleVP('discrete', betas) (dpm_solver_pp.py:55-169,
 * interpolate_fn :9-52), get_time_steps / get_orders_and_timesteps_for_singlestep_solver (:330-405) and the per-step
 * coefficients of the 1S / 2S / 3S (:432-457, :511-557, :700-766) and 1 / 2M / 3M (:602-677, driver :995-1017) updates
 * (data prediction, solver_type 'dpm_solver').  HOST function, no device work, no handle: fills `out_plan`
 * [n_evals][PDM_PLAN_STRIDE] (one record per network evaluation) for pdm_sample.
 *   betas [n_betas] HOST float32 (the discrete schedule, train_t2i_discrete.py:40-44, 504); steps = number of function
 *   evaluations; order 1..3; eps = end time t_0 (1/N in the t2i path), T = start time (1.0);
 *   mask_opt != 0: the mask stream follows the solver update (enable_mask_opt=True, the live path); 0: pass-through;
 *   n_time: factor from continuous to model time (1000, train_t2i_discrete.py:508).
 *   out_plan may be NULL to query *n_evals; cap_evals = capacity of out_plan in records.
 * Agrees with the Python host planner (float32 torch ops = what the reference computes) to a few float32 ulp per
 * coefficient (libm vs SLEEF transcendental rounding), see csrc/plan.cu. */
int pdm_solver_plan(const float* betas, int32_t n_betas, int32_t steps, int32_t order, int32_t method, int32_t skip_type,
                    float eps, float T, int32_t mask_opt, float n_time, float* out_plan, int32_t cap_evals,
                    int32_t* n_evals);

/* replaces: DPM_Solver.sample(method='fast') (dpm_solver_pp.py:1018-1044) driven by cfg_nnet
 * (train_t2i_discrete.py:387-439, 506-516): the whole denoising loop on the device.
 *   plan: HOST array [n_evals][PDM_PLAN_STRIDE] built by the host planner (solver scalars are data
 *   independent); z_init [B,4,H,W]; mask_init [B,8,H,W] (NULL = image-only);
 *   ctx [B,T,clip]; empty_ctx [T,clip] (NULL -> no guidance, cfg_scale ignored);
 *   out_z [B,4,H,W]; out_pred_mask [B,8,H,W] = CFG'd mask prediction of the first evaluation of the
 *   last solver step (dpm_solver_pp.py:827,1044).
 * A plan of multistep records (kind 1, pdm_solver_plan with PDM_METHOD_MULTISTEP) runs DPM-Solver++ 2M / 3M the same way:
 * one network evaluation + ONE fused update kernel per step, the data-prediction history kept in the engine's workspace;
 * out_pred_mask is then the mask prediction of the last evaluation.
 * use_graph != 0 captures the loop into a CUDA graph (cached per (B, n_evals, precision)). */
int pdm_sample(pdm_handle h, const float* plan, int32_t n_evals, const float* z_init, const float* mask_init,
               const float* ctx, const float* empty_ctx, float cfg_scale, float* out_z, float* out_pred_mask,
               int32_t B, int32_t precision, int32_t use_graph, void* stream);

/* replaces: utils.py:490-518 bits2int applied to (pred_mask > 0) (utils.py:596): sign-threshold the 8
 * analog bits, MSB first.  pred_mask [B,8,H,W] float32 -> labels [B,H,W] int32 (device). */
int pdm_bits2int(const float* pred_mask, int32_t* labels, int32_t B, int32_t nbits, int32_t hw, void* stream);
/* replaces: utils.py:475-488 int2bits followed by *2-1 (train_t2i_discrete.py:489-490):
 * ids [B,H,W] int32 -> analog bits [B,8,H,W] float32 in {-1,+1}. */
int pdm_int2bits(const int32_t* ids, float* bits, int32_t B, int32_t nbits, int32_t hw, void* stream);

/* ---- VAE decoder (the step after the loop: latents -> images) ---------------------------------------------------------
 * replaces: libs/autoencoder.py:303-410 (Decoder) + :446-450 (FrozenAutoencoderKL.decode), called per mini-batch by
 * eval_t2i_discrete.py:74-84 (decode_large_batch) / utils.py:627-637.  Mirrors `ddconfig` of libs/autoencoder.py:471-485. */
typedef struct pdm_vae* pdm_vae_handle;
typedef struct pdm_vae_config {
    int32_t ch;             /* 128 */
    int32_t num_levels;     /* len(ch_mult) = 4 */
    int32_t ch_mult[8];     /* 1, 2, 4, 4 */
    int32_t num_res_blocks; /* 2 */
    int32_t z_channels;     /* 4 */
    int32_t embed_dim;      /* 4 */
    int32_t out_ch;         /* 3 */
    float scale_factor;     /* 0.18215 (configs: 0.23010) */
} pdm_vae_config;
int pdm_vae_create(const pdm_vae_config* cfg, pdm_vae_handle* out);
int pdm_vae_destroy(pdm_vae_handle h);
/* `key` = key of the reference autoencoder state_dict (`decoder.*`, `post_quant_conv.*`); `encoder.*` / `quant_conv.*` /
 * `loss.*` are accepted and ignored (the sampling path only decodes).  dev_f32: device float32 tensor of the given shape. */
int pdm_vae_set_param(pdm_vae_handle h, const char* key, const void* dev_f32, const int64_t* shape, int32_t ndim, void* stream);
int pdm_vae_finalize_params(pdm_vae_handle h, void* stream);
/* z [n, 4, s, s] float32 (scaled latents, as the sampler returns them) -> out [n, out_ch, 8 s, 8 s] float32 in [-1, 1]-ish
 * (the caller applies unpreprocess: 0.5 (x + 1) clamp, train_t2i_discrete.py:584).  latent_size s in {16, 32, 64, ...}. */
int pdm_vae_decode(pdm_vae_handle h, const float* z, float* out, int32_t n, int32_t latent_size, void* stream);
int pdm_vae_workspace_bytes(pdm_vae_handle h, int32_t n, int32_t latent_size, size_t* bytes);

/* diagnostics */
const char* pdm_last_error(void);
int pdm_abi_version(void);
/* number of kernels launched by this library since load (all handles); lets a caller count launches
 * inside a timed region. */
int64_t pdm_launch_count(void);
/* per-forward timing hooks: when enabled the engine records CUDA events around named phases. */
int pdm_set_profiling(pdm_handle h, int32_t enabled);
/* after a forward with profiling enabled (synchronises): milliseconds spent in up to `cap` phases;
 * names are written as a '\n'-separated list into name_buf. */
int pdm_get_profile(pdm_handle h, float* ms, int32_t cap, char* name_buf, int32_t name_cap, int32_t* count);

/* ---- kernel-level diagnostics (unit tests, per-kernel roofline timing in bench.py) ---------------
 * One Linear layer through the production GEMM kernel of the given precision:
 *   out[M,N] = [gelu](A[M,K] . W[N,K]^T + bias) [+ resid]   (all float32 device tensors; for
 *   PDM_PREC_BF16 A and W are rounded to bf16 first, exactly as the engine feeds the tcgen05 kernel).
 * If A2/K2 are given the K loop streams [A | A2] (the long-skip GEMM).  iters > 0 and ms != NULL:
 * the GEMM kernel alone is launched `iters` extra times between CUDA events and the average duration
 * (milliseconds) is written to *ms (synchronises). */
int pdm_debug_linear(const float* A, const float* A2, const float* W, const float* bias, const float* resid,
                     float* out, int32_t M, int32_t N, int32_t K, int32_t K2, int32_t precision, int32_t gelu,
                     int32_t iters, float* ms, void* stream);
/* Self-attention core through the production kernel: qkv [nb,L,3*H*64] -> out [nb,L,H*64] (float32
 * device tensors, rounded to bf16 for PDM_PREC_BF16); same timing convention. */
int pdm_debug_attention(const float* qkv, float* out, int32_t nb, int32_t L, int32_t H, int32_t precision,
                        int32_t iters, float* ms, void* stream);
/* LayerNorm kernel: x [rows,D] -> out (float32 result; bf16 path rounds through bf16); timing as above. */
int pdm_debug_layernorm(const float* x, const float* w, const float* b, float* out, int64_t rows, int32_t D,
                        int32_t precision, int32_t iters, float* ms, void* stream);

/* Deferred-LayerNorm chain exactly as the bf16 engine runs a block:
 *   x[M,D]   = resid + A[M,K1] . W1[D,K1]^T + b1      (A == NULL: x = resid; the row sums then come from the
 *                                                      rowstats pass instead of the producing GEMM's epilogue)
 *   out[M,N] = [gelu]( LayerNorm(x; gamma, beta, eps 1e-5) . W2[N,D]^T + b2 )     (b2 may be NULL)
 * with the LayerNorm folded into W2 and applied per row in the consuming GEMM's epilogue.  All float32 device
 * tensors; x_out (nullable) receives the fp32 x.  Replaces nn.LayerNorm + nn.Linear of libs/uvit_t2i.py:189-190,
 * 209-210 (norm1 -> attn.qkv, norm2 -> mlp.fc1).  Timing of the consuming GEMM as in pdm_debug_linear. */
int pdm_debug_ln_chain(const float* A, const float* W1, const float* b1, const float* resid, const float* gamma,
                       const float* beta, const float* W2, const float* b2, float* out, float* x_out, int32_t M,
                       int32_t N, int32_t D, int32_t K1, int32_t gelu, int32_t iters, float* ms, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PDM_H_ */
