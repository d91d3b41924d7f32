/*
 * pno_b200.h -- C ABI of libpno_b200.so: the B200 (sm_100a) implementation of pno-physics-bench's probabilistic
 * spectral-convolution hot path.
 *
 * The reference (danieleschmidt/pno-physics-bench) is 100 % Python and exposes no FFI for this path:
 * the drop-in boundary is the Python class API of src/pno_physics_bench/models.py.  This header is therefore the
 * builder-defined C ABI that the Python mirror of that API (pno_physics_bench_b200/models.py) binds through ctypes;
 * each entry point cites the reference lines whose work it replaces (paths relative to src/pno_physics_bench/).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (torch's caching allocator); the library never frees
 *     caller memory and allocates only the small twiddle tables a plan owns (plus, for the 3xTF32 engine, one grow-only scratch
 *     per stream holding the tf32 hi / lo parts of the two 1x1 weight matrices of pno_local_fwd: 16 bytes per weight);
 *   - `stream` is a cudaStream_t passed as void*; every launch goes on it; no hidden synchronisation;
 *   - every function returns 0 on success, a negative PNO_E_* code otherwise; pno_last_error() gives the message
 *     (thread-local);
 *   - activations are fp32 NCHW contiguous; complex numbers are interleaved (re, im) fp32 pairs;
 *   - spectral parameters use the kernel-native layout  mu[kx][ky][i][o] (complex) / log_var[kx][ky][i][o] (fp32),
 *     one array per block (weights1 = +kx rows, weights2 = -kx rows).  The Python modules store the reference's
 *     [Ci,Co,m1,m2] parameters as strided views of exactly this storage, so no repacking happens per call;
 *   - truncated spectra use the layout  xhat[mode][b][c], mode = kx2*m2 + ky, kx2 in [0, 2*m1): rows [0,m1) are
 *     kx = kx2 (weights1), rows [m1,2*m1) are kx = H - 2*m1 + kx2 (weights2); real and imaginary planes are
 *     separate arrays;
 *   - noise is a counter-based Philox4x32-10 stream keyed by (seed, sample_idx, site, element) -- see
 *     oracle/philox.py for the normative definition; nothing depends on launch geometry or GPU count.
 *
 * Tuning / diagnostic environment variables (read by the host side of the library; none is needed in normal use):
 *   PNO_PDL=1            launch the tensor-core kernels with programmatic stream serialization (measured: no gain)
 *   PNO_LOCAL_TN=256     256-pixel tiles in the single-pass local-branch kernel (measured slower than the default 128)
 *   PNO_LOCAL_EW=16      16 epilogue warps in the local-branch kernel instead of 8 (measured slower)
 *   PNO_LOCAL_PREFETCH=0 no L2 prefetch of the next tile's x boxes
 *   PNO_INV_NB1=2, PNO_INV_NZT=3, PNO_INV_ND2=n, PNO_INV_ADD=0|2|3|4 buffer counts of the inverse-transform kernel
 *   PNO_INV_TS=0         single pass: transposer path instead of stage 2 from tensor memory
 *   PNO_INV_TMAOUT=0     global store instructions instead of the in-place TMA tile store
 *   PNO_INV_G=4|8, PNO_INV_EPI=8, PNO_INV_CAP=kb                     channel-group size, epilogue warps, shared-memory budget
 *   PNO_MIX_RAW=0        single pass: register-prefetched mu / log_var instead of the TMA-staged tiles
 *   PNO_FWD_STAGES=n, PNO_FWD_KPS=1|2                                ring depth / stage size of the forward-transform kernel
 *   PNO_MIX_STAGES=2..4                                              ring depth of the channel-mix kernel
 *   PNO_DFT_TWOPASS=1    use the two-pass transforms (pno_tc_dft2.cu) wherever they tile the shape, not only where the fused kernels do not
 *   PNO_DEBUG_MMA=1      (PNO_TRACE builds) serialise the inverse kernel's MMA batches to time them
 * A library built with `make EXTRA=-DPNO_TRACE` records per-role event timelines of CTA 0 into the pno_debug_set_buffer() buffer
 * (tools/trace_*.py); production builds compile all of that away.
 */
#ifndef PNO_B200_H
#define PNO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PNO_VERSION 100

enum {
    PNO_OK = 0,
    PNO_E_INVALID = -1,   /* bad shape / null pointer / misalignment                          */
    PNO_E_MODES = -2,     /* m1 > H or m2 > W/2+1 (reference: RuntimeError from einsum, models.py:93-94) */
    PNO_E_CUDA = -3,      /* a CUDA runtime call failed                                        */
    PNO_E_ARCH = -4,      /* device is not sm_100 (no fallback paths exist)                    */
    PNO_E_UNSUPPORTED = -5 /* the requested engine does not support this shape                 */
};

enum { PNO_ACT_NONE = 0, PNO_ACT_GELU = 1, PNO_ACT_RELU = 2 };   /* models.py:224-229 (exact-erf GELU) */

/* arithmetic engines.  The tensor-core engines tile the shapes documented in each kernel file (the BASELINE configs).  A stage
 * whose shape a tensor-core engine does not tile FAILS with PNO_E_UNSUPPORTED -- unless the caller ORs
 * PNO_ENGINE_ALLOW_FALLBACK into the engine argument, in which case that stage runs on the CUDA-core fp32 engine instead (same
 * results to fp32 accuracy).  pno_stage_counts() tells which engine actually ran each stage.  1x1 convs with <= 8 input or <= 4
 * output channels (the lift / projection heads) are streaming CUDA-core kernels under every engine: there is no GEMM in them. */
enum {
    PNO_ENGINE_FP32 = 0,      /* fp32 FMA on CUDA cores: any shape, reference-grade fp32 (rel-L2 <= 1e-5)      */
    PNO_ENGINE_TC_TF32X3 = 1, /* tcgen05 tensor cores, error-compensated 3xTF32 (rel-L2 <= 1e-5)                */
    PNO_ENGINE_TC_TF32 = 2,   /* tcgen05 tensor cores, single-pass TF32 operands / fp32 accumulate: the reduced-
                                 precision mode (rel-L2 ~1e-3, inside the 2e-2 budget of the bf16-compute mode)   */
    PNO_ENGINE_ALLOW_FALLBACK = 0x100   /* flag: opt into the CUDA-core fallback for shapes the engine does not tile */
};

/* stages whose engine is dispatched (index into pno_stage_counts) */
enum {
    PNO_STAGE_DFT_FWD = 0, PNO_STAGE_DFT_INV = 1, PNO_STAGE_MIX_FWD = 2, PNO_STAGE_MIX_BWD_DX = 3, PNO_STAGE_MIX_BWD_DW = 4,
    PNO_STAGE_LOCAL_FWD = 5, PNO_STAGE_LOCAL_BWD = 6,
    PNO_STAGE_LOCAL_THIN = 7,   /* lift- / projection-shaped 1x1 convs, forward or backward: streaming CUDA-core kernels by design */
    PNO_N_STAGES = 8
};

typedef struct pno_plan pno_plan_t;

int pno_version(void);
const char* pno_last_error(void);
/* 0 if the current device can run this library (compute capability 10.x), PNO_E_ARCH otherwise. */
int pno_check_device(void);
/* number of CUDA kernels this library has launched in this process (monotonic; bench.py reports deltas) */
uint64_t pno_launch_count(void);
/* The persistent tensor-core kernels size their grids to (SM count - n): n SMs stay free for a kernel of another stream, e.g. the
 * NCCL all-reduce of data-parallel training launched while backward is still running (reference: DDP bucket overlap,
 * distributed_training.py:96-101).  Process-wide, 0 by default; takes effect at the next launch (captured graphs keep theirs). */
int pno_set_sm_reserve(int n);
/* out[2*stage + 0] / out[2*stage + 1] = successful calls of `stage` that ran on CUDA cores / on the tensor cores since the process
 * started (monotonic; PNO_N_STAGES pairs).  Tests assert on deltas that the tensor-core kernel really ran. */
int pno_stage_counts(uint64_t* out);
/* Host-only query (no CUDA call, works without a GPU): 1 if `engine` (PNO_ENGINE_TC_*) runs `stage` on its tensor-core kernel
 * for a [B, Ci|Co, H, W] problem with (m1, m2) retained modes -- every variant of the stage the block path uses (with / without
 * addend and pre-activation output, mean / sampled weights) -- else 0.  Answered by the tilings' own validation code. */
int pno_engine_supports(int stage, int engine, int H, int W, int m1, int m2, int B, int Ci, int Co);
/* Device-resident noise key for CUDA-graph replay (reference models.py:361-364 is a host loop over samples; here one captured
 * forward is replayed S times).  While device_key != NULL (per host thread) every stochastic entry point records that pointer in
 * its launches, and the kernels read (seed low word, seed high word, sample_idx) = device_key[0..2] at RUN time instead of using
 * their by-value seed / sample_idx arguments; `site` and batch_offset stay by-value.  NULL restores the by-value behaviour. */
int pno_set_dynamic_noise(const uint32_t* device_key);
/* device_key[0..3] = (seed low, seed high, sample_idx, 0) written by a one-thread kernel on `stream` (no host buffer to keep
 * alive); pno_noise_key_advance adds delta to device_key[2] (captured at the end of a graphed forward: next sample). */
int pno_noise_key_write(uint32_t* device_key, uint64_t seed, uint32_t sample_idx, void* stream);
int pno_noise_key_advance(uint32_t* device_key, uint32_t delta, void* stream);

/* Plan = twiddle tables for one (H, W, m1, m2) on the current device.  Immutable after creation. */
int pno_plan_create(int H, int W, int m1, int m2, pno_plan_t** plan);
int pno_plan_destroy(pno_plan_t* plan);
/* bytes of scratch pno_spectral_fwd / pno_spectral_bwd need for batch B and channel counts Ci, Co */
size_t pno_spectral_workspace_bytes(const pno_plan_t* plan, int B, int Ci, int Co);

/* ---- noise stream (parity / debugging; the hot path draws the same numbers in-kernel) -------------------------- */
/* eps[first_elem .. first_elem+n) of an activation-shaped site (replaces torch.randn_like, models.py:272). */
int pno_philox_normal_activation(uint64_t seed, uint32_t sample_idx, uint32_t site, uint64_t first_elem, uint64_t n,
                                 float* out, void* stream);
/* eps_re, eps_im, each [2][Ci][Co][m1][m2] in the reference's logical layout (models.py:67-68, 74, 76). */
int pno_philox_normal_spectral(uint64_t seed, uint32_t sample_idx, uint32_t site, int Ci, int Co, int m1, int m2,
                               float* eps_re, float* eps_im, void* stream);
/* raw Philox4x32-10 output words for quads [q0, q0+nq): out[nq][4] */
int pno_philox_bits(uint64_t seed, uint32_t sample_idx, uint32_t site, uint64_t q0, uint64_t nq, uint32_t* out,
                    void* stream);

/* ---- stages of the spectral layer ------------------------------------------------------------------------------ */
/* Pruned forward transform (replaces torch.fft.rfft2 + the two slices, models.py:89, 93-94):
 *   xhat[mode][b][c] = sum_{h,w} x[b][c][h][w] exp(-2 pi i (kx h/H + ky w/W)).
 * grad_scale != 0 multiplies by alive(kx2) * c_ky / (H W): the adjoint of the inverse transform (backward of irfft2). */
int pno_dft_fwd(const pno_plan_t* plan, int engine, const float* x, int B, int C, int grad_scale, float* xhat_re,
                float* xhat_im, void* stream);
/* Pruned inverse transform + fused epilogue (replaces zeros + scatter + torch.fft.irfft2, models.py:92-97, and the
 * add + activation of the block, models.py:306):
 *   s = sum_mode coef * Re(yhat[mode][b][c] exp(+2 pi i (...)))          coef = alive c_ky/(HW) (adjoint != 0: coef = alive)
 *   v = s + (addend ? addend[b][c][h][w] : 0);  if (pre_act) pre_act = v;  y = act(v)                               */
int pno_dft_inv(const pno_plan_t* plan, int engine, const float* yhat_re, const float* yhat_im, int B, int C, int adjoint,
                const float* addend, int act, float* y, float* pre_act, void* stream);
/* Per-mode complex channel mix with in-kernel weight sampling (replaces sample_weights + compl_mul2d,
 * models.py:63-78, 33-35):  yhat[mode][b][o] = sum_i xhat[mode][b][i] * (mu + exp(lv/2) * eps)[mode][i][o].
 * lv1 == lv2 == NULL selects the mean weights (sample=False / eval gate, models.py:79-80, 86). */
int pno_mix_fwd(const pno_plan_t* plan, int engine, const float* xhat_re, const float* xhat_im, int B, int Ci, int Co,
                const float* mu1, const float* mu2, const float* lv1, const float* lv2, uint64_t seed,
                uint32_t sample_idx, uint32_t site, float* yhat_re, float* yhat_im, void* stream);
/* dxhat[mode][b][i] = sum_o ghat[mode][b][o] * conj(W[mode][i][o])  (W re-sampled from the same counters) */
int pno_mix_bwd_dx(const pno_plan_t* plan, int engine, const float* ghat_re, const float* ghat_im, int B, int Ci, int Co,
                   const float* mu1, const float* mu2, const float* lv1, const float* lv2, uint64_t seed,
                   uint32_t sample_idx, uint32_t site, float* dxhat_re, float* dxhat_im, void* stream);
/* dmu[mode][i][o] = sum_b conj(xhat[mode][b][i]) * ghat[mode][b][o];
 * dlv = 0.5 exp(lv/2) (eps_re Re dmu + eps_im Im dmu)   (written, not accumulated; dlv* may be NULL) */
int pno_mix_bwd_dw(const pno_plan_t* plan, int engine, const float* xhat_re, const float* xhat_im, const float* ghat_re,
                   const float* ghat_im, int B, int Ci, int Co, const float* lv1, const float* lv2, uint64_t seed,
                   uint32_t sample_idx, uint32_t site, float* dmu1, float* dmu2, float* dlv1, float* dlv2, void* stream);

/* Mode windows: the same three kernels on global modes [mode_first, mode_first + mode_count) only (a range inside ONE weight
 * block, i.e. not crossing mode m1*m2) for MODE-PARALLEL sharding of the spectral weights across GPUs: each
 * rank keeps the weights / log-variances / gradients / optimiser state of its own modes.  The spectra are [mode_count][B][C]
 * (local mode index), mu / lv / dmu / dlv are [mode_count][Ci][Co] shards, the Philox counters use the GLOBAL mode index, so the
 * sharded model draws exactly the weights of the unsharded one.  Tensor-core engines only (PNO_E_UNSUPPORTED otherwise). */
int pno_mix_fwd_window(const pno_plan_t* plan, int engine, const float* xhat_re, const float* xhat_im, int B, int Ci, int Co,
                       const float* mu, const float* lv, uint64_t seed, uint32_t sample_idx, uint32_t site, int mode_first,
                       int mode_count, float* yhat_re, float* yhat_im, void* stream);
int pno_mix_bwd_dx_window(const pno_plan_t* plan, int engine, const float* ghat_re, const float* ghat_im, int B, int Ci, int Co,
                          const float* mu, const float* lv, uint64_t seed, uint32_t sample_idx, uint32_t site, int mode_first,
                          int mode_count, float* dxhat_re, float* dxhat_im, void* stream);
int pno_mix_bwd_dw_window(const pno_plan_t* plan, int engine, const float* xhat_re, const float* xhat_im, const float* ghat_re,
                          const float* ghat_im, int B, int Ci, int Co, const float* lv, uint64_t seed, uint32_t sample_idx,
                          uint32_t site, int mode_first, int mode_count, float* dmu, float* dlv, void* stream);

/* ---- whole layer / block --------------------------------------------------------------------------------------- */
/* SpectralConv2d(_Probabilistic).forward (models.py:37-50, 82-98) with the PNO-block epilogue fused
 * (models.py:306): y = act(spectral(x) + addend).  addend may be NULL and act PNO_ACT_NONE for the bare layer.
 * keep_xhat != 0 leaves xhat in the first 2*M*B*Ci floats of `workspace` for pno_spectral_bwd. */
int pno_spectral_fwd(const pno_plan_t* plan, int engine, const float* x, int B, int Ci, int Co, const float* mu1,
                     const float* mu2, const float* lv1, const float* lv2, uint64_t seed, uint32_t sample_idx,
                     uint32_t site, const float* addend, int act, float* y, float* pre_act, void* workspace,
                     void* stream);
/* Backward of the spectral branch for upstream gradient g = dL/d(spectral output) [B][Co][H][W].
 * xhat_saved: the xhat planes kept by pno_spectral_fwd (NULL: recomputed from x).  Any of dx / dmu* / dlv* may be
 * NULL to skip it. */
int pno_spectral_bwd(const pno_plan_t* plan, int engine, const float* x, const float* xhat_saved, const float* g, int B,
                     int Ci, int Co, const float* mu1, const float* mu2, const float* lv1, const float* lv2,
                     uint64_t seed, uint32_t sample_idx, uint32_t site, float* dx, float* dmu1, float* dmu2, float* dlv1,
                     float* dlv2, void* workspace, void* stream);

/* Local branch: 1x1 conv mean (+ reparameterised activation-space noise), models.py:288-290, 298-303, 309-314, 264-275:
 *   m = Wm x + bm;  if (Wl) { lv = clamp(Wl x + bl, -10, 10); s = max(exp(lv/2), 1e-6); out = m + eps s } else out = m
 * x [B][Ci][HW], out [B][Co][HW]; eps element index uses the GLOBAL batch index batch_offset + b.
 * dout_dlv (may be NULL) receives d out / d(Wl x + bl) = 0.5 eps s [|lv|<10][s>1e-6] for the backward pass. */
int pno_local_fwd(int engine, const float* x, int B, int Ci, int Co, int HW, const float* Wm, const float* bm,
                  const float* Wl, const float* bl, uint64_t seed, uint32_t sample_idx, uint32_t site,
                  uint64_t batch_offset, float* out, float* dout_dlv, void* stream);
/* Backward of pno_local_fwd on the tensor-core engines (autograd of models.py:298-303): with dYm = g and dYl = g * dout_dlv,
 *   dx = Wm^T dYm + Wl^T dYl ([B,Ci,HW]);  dw = [sum_b dYm x^T ; sum_b dYl x^T] ([2Co][Ci] stacked, [Co][Ci] when Wl == NULL);
 *   db = [sum dYm ; sum dYl] per output channel.  dx / dw / db may each be NULL.  Shapes: Co % 128 == 0, Ci in {128, 256},
 * HW % 128 == 0; or a lift-shaped layer (Ci <= 4, dx == NULL, HW % 4 == 0: streaming dot products); or a projection-shaped
 * layer (Co <= 4, HW % 4 == 0: one pass over x); else PNO_E_UNSUPPORTED (as for PNO_ENGINE_FP32): the caller then uses plain library GEMMs.
 * workspace: pno_local_bwd_workspace_bytes(...) bytes (the stacked gradient [B][2Co][HW] and the stacked transposed weights;
 * 1 KB for lift-shaped layers).  A lift-shaped call with dx != NULL is PNO_E_UNSUPPORTED. */
size_t pno_local_bwd_workspace_bytes(int B, int Ci, int Co, int HW, int noisy);
int pno_local_bwd(int engine, const float* x, const float* g, const float* dout_dlv, int B, int Ci, int Co, int HW,
                  const float* Wm, const float* Wl, float* dx, float* dw, float* db, void* workspace, void* stream);

/* gz = g * act'(pre_act)  (backward of models.py:306) */
int pno_act_bwd(const float* pre_act, const float* g, uint64_t n, int act, float* gz, void* stream);

/* ---- KL (models.py:100-104) ------------------------------------------------------------------------------------ */
/* *kl_accum (double, device) += -0.5 sum(1 + lv - re^2 - im^2 - exp(lv)) over n elements of one block */
int pno_kl_fwd(const float* mu, const float* lv, uint64_t n, double* kl_accum, void* stream);
/* dmu (+)= gscale * mu ; dlv (+)= gscale * (-0.5 (1 - exp(lv))) ; gscale read from device memory */
int pno_kl_bwd(const float* mu, const float* lv, uint64_t n, const float* gscale, int accumulate, float* dmu, float* dlv,
               void* stream);

/* ---- MC predictive moments (models.py:361-368) ----------------------------------------------------------------- */
int pno_mc_accumulate(const float* y, uint64_t n, double* sum, double* sumsq, void* stream);
/* One forward pass of UncertaintyDecomposer.decompose (uncertainty.py:57-73): p = mean + sqrt(exp(log_var) + 1e-8) * eps with
 * eps = Philox normal of (seed, sample_idx, site, flat element index); sum += p, sumsq += p^2, sum_avar += exp(log_var);
 * pred_out (optional) receives p.  log_var == NULL: p = mean, aleatoric variance 0 (uncertainty.py:75-77). */
int pno_mc_accumulate_decompose(const float* mean, const float* log_var, uint64_t n, uint64_t seed, uint32_t sample_idx,
                                uint32_t site, double* sum, double* sumsq, double* sum_avar, float* pred_out, void* stream);
/* mean = sum/S; std = sqrt((sumsq - S mean^2)/(S-1)) (unbiased; S == 1 -> NaN like torch.std) */
int pno_mc_finalize(const double* sum, const double* sumsq, uint64_t n, uint32_t S, float* mean, float* std,
                    void* stream);

/* ---- loss / calibration epilogue on device ----------------------------------------------------------
 * One pass over (mean, spread, target), n elements each; spread_kind 0: spread is the predictive std, 1: spread is log_var
 * (std = exp(0.5 log_var), training/losses.py:60).  ADDS into acc (double, device), so batches accumulate on device and
 * PNOTrainer.evaluate needs no per-batch host copies (training/trainer.py:540-552).  Layout of acc (L = n_levels,
 * A = n_alphas, NB = n_bins):
 *   [0] n   [1] sum |mean-target|   [2] sum (mean-target)^2   [3] sum Gaussian NLL, std clamped at 1e-6 (losses.py:97-117,
 *   metrics.py:295-324)   [4] sum std   [5] sum std^2   [6] sum std*|mean-target|
 *   [7 .. 7+L)        count(|mean-target| <= z_levels[l] * std)        (metrics.py:99-128, losses.py:119-149)
 *   [7+L .. 7+L+A)    sum interval score for (alphas[a], z_alphas[a])  (metrics.py:130-159)
 *   [7+L+A .. +3 NB)  per ECE bin b = (bin_edges[b], bin_edges[b+1]]: count, sum confidence, count(|e|/(std+1e-8) <= 1)
 *                     with confidence = clamp(1 - 2 pdf(|e|/(std+1e-8)), 0, 1)                  (metrics.py:32-97)
 * z_levels / alphas / z_alphas are HOST arrays (copied by value), bin_edges is a DEVICE array of NB+1 floats. */
#define PNO_UQ_MAX_LEVELS 8
#define PNO_UQ_MAX_ALPHAS 4
#define PNO_UQ_MAX_BINS 64
int pno_uq_stats(const float* mean, const float* spread, const float* target, uint64_t n, int spread_kind,
                 const float* z_levels, int n_levels, const float* alphas, const float* z_alphas, int n_alphas,
                 const float* bin_edges, int n_bins, double* acc, void* stream);
/* Gradient of  g[0] * sum (mean-target)^2 + g[1] * sum nll  (g: 2 doubles on device) w.r.t. mean and log_var
 * (log_var / dlog_var may be NULL: MSE only).  The coverage term of the loss is piecewise constant (losses.py:140-143). */
int pno_loss_data_bwd(const float* mean, const float* log_var, const float* target, uint64_t n, const double* g,
                      float* dmean, float* dlog_var, void* stream);

/* ---- optimiser step of the trainer (training/trainer.py:283-291: clip_grad_norm_, AdamW.step) without host syncs ---- */
/* kl_kind (both calls): 0 = none; 1 / 2 = the tensor is a spectral mean (re, im parts) / log-variance whose KL gradient
 * (models.py:100-104) is ANALYTIC in the parameter itself -- *kl_gscale * p, resp. *kl_gscale * (-0.5 (1 - exp(p))) -- and is
 * added to the data gradient on the fly instead of being materialised and accumulated by a separate pass (kl_gscale: one float
 * on the device = d loss / d KL). */
/* *acc (double, device) += sum of squares of the n (total) gradient values */
int pno_sumsq_accumulate(const float* grad, uint64_t n, double* acc, const float* param, int kl_kind, const float* kl_gscale,
                         void* stream);
/* torch.optim.AdamW semantics (decoupled weight decay, bias correction by step >= 1, denom = sqrt(v_hat) + eps) on one
 * flat tensor; the gradient is first scaled by min(1, max_norm / (sqrt(*grad_sumsq) + 1e-6)) (torch clip_grad_norm_) when
 * grad_sumsq != NULL and max_norm > 0, and written back clipped when write_clipped_grad != 0. */
int pno_adamw_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, uint64_t n, float lr, float beta1, float beta2,
                   float eps, float weight_decay, uint32_t step, const double* grad_sumsq, float max_norm,
                   int write_clipped_grad, int kl_kind, const float* kl_gscale, void* stream);

/* Multi-tensor forms of the two calls above: ONE launch (per 40 tensors) over HOST arrays of `count` device pointers / sizes /
 * kl_kinds (kl_kinds may be NULL = all 0; params may be NULL in the norm call when every kind is 0).  All tensors of one call share
 * the hyper-parameters and the step count (one torch param_group whose states are in step).  Same arithmetic per element as the
 * per-tensor calls; the norm accumulates in a different order (fp64), so results agree to rounding. */
int pno_sumsq_accumulate_multi(int count, float* const* grads, const uint64_t* ns, float* const* params, const int* kl_kinds,
                               const float* kl_gscale, double* acc, void* stream);
int pno_adamw_step_multi(int count, float* const* params, float* const* grads, float* const* exp_avgs, float* const* exp_avg_sqs,
                         const uint64_t* ns, const int* kl_kinds, float lr, float beta1, float beta2, float eps, float weight_decay,
                         uint32_t step, const double* grad_sumsq, float max_norm, int write_clipped_grad, const float* kl_gscale,
                         void* stream);

/* Diagnostics: device buffer (>= 148*16 int64) into which the tensor-core kernels dump per-CTA barrier-wait cycle counters
 * (NULL switches it off, the default).  Used by tools/phase_times.py to find the stage a pipeline is waiting on. */
int pno_debug_set_buffer(void* device_ptr);

/* ---- hardware self-test of the tcgen05 / TMA building blocks ---------------------------------------------------- */
/* One CTA computes D[128][N] = (+-A)[128][K] * (+-B)[K][N] on the tensor cores (kind::tf32).
 * flags: bit0 A MN-major, bit1 B MN-major, bit2 A staged by TMA, bit3 B staged by TMA, bit4 negate A, bit5 negate B,
 * bit6 both operands K-major in SWIZZLE_32B slab form (one 32-byte slab per K=8 step; K%8 == 0, N%16 == 0).
 * A [128][K], B [K][N] row-major; the TMA sources are the arrays whose contiguous axis is the operand's major axis
 * (K-major A: A itself; MN-major A: A^T [K][128]; K-major B: B^T [N][K]; MN-major B: B itself). */
int pno_selftest_umma(int flags, int N, int K, const float* A, const float* B, const float* A_tma_src,
                      const float* B_tma_src, float* D, void* stream);
/* The fp32-parity product of the tensor-core engines on RAW fp32 operands: one kind::tf32 MMA on the fp32 tiles plus two
 * kind::f16 MMAs on bf16 copies (bf16(A) x bf16(A-lo part of B), bf16(lo part of A) x bf16(B)) into one accumulator.
 * flags: bit0 B MN-major, bit1 the bf16 tiles of A staged by TMA (scratch: 512*K bytes), bit2 main term only (what the
 * tensor core does with the 13 low mantissa bits of a raw fp32 operand).  K%32 == 0, K <= 128, N%16 == 0 (N%64 MN-major). */
int pno_selftest_umma_mixed(int flags, int N, int K, const float* A, const float* B, float* D, void* scratch, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PNO_B200_H */
