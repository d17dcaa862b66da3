/* libshotvae C ABI -- the drop-in boundary for the SHOT-VAE training-step hot path on B200 (sm_100a).
 *
 * The reference (FengHZ/SHOT-VAE) has no FFI of its own: its hot path sits behind a Python module
 * API (shot_vae_model/vae.py:140-151, lib/criterion.py:32-57,97-108, lib/utils/mixup.py:5-41,
 * main_shot_vae.py:281-366) and every FLOP is an ATen/cuDNN call.  This header is the boundary the
 * Python shim (shot-vae_b200/shotvae_b200/_abi.py, ctypes) binds; each entry point names the
 * reference call site(s) whose arithmetic it replaces.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer owned by the caller (torch allocates; the library never
 *    allocates, frees or retains pointers past the call, except cached TMA descriptors);
 *  - every call is asynchronous on `stream` (a cudaStream_t passed as void*), performs no host
 *    synchronisation and is CUDA-graph capturable;
 *  - activations are NHWC bf16 with the channel count padded to a multiple of 16; `NB = G * B`
 *    images where G independent "pass groups" are batched and BatchNorm statistics are per group;
 *  - return value 0 on success, negative on error (sv_last_error() gives the text, thread local).
 *  - there is no CPU fallback anywhere.
 */
#ifndef SHOTVAE_H_
#define SHOTVAE_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SV_ABI_VERSION 1
#define SV_MAX_TAPS 16

int sv_abi_version(void);
const char* sv_last_error(void);
/* 1 if the loaded library was built with the tcgen05/TMA implicit-GEMM path */
int sv_has_tcgen05(void);
/* number of kernel launches issued by this library since process start (bench `gpu_launches`) */
long long sv_launch_count(void);
/* Persistent-grid kernels (the tcgen05 convolutions and weight gradients: one CTA per SM, statically strided tiles) launch
 * at most n CTAs from now on (0 = one per SM).  A data-parallel step sets it around the part of the backward that overlaps the
 * NCCL all-reduces (replacement of the reference's nn.DataParallel reduce, wideresnet.py:78-94): a collective that holds k SMs
 * while a 148-CTA grid is launched makes that grid's last k CTAs a second wave, i.e. doubles the kernel's time.  Returns the
 * previous limit.  The value is read at launch (and at CUDA-graph capture) time. */
int sv_set_cta_limit(int32_t n);

/* ---- implicit GEMM (replaces every nn.Conv2d / nn.ConvTranspose2d fprop + dgrad call site:
 *      wideresnet.py:12-14,29-35,41-43; preactresnet.py:31-36,56-59; decoder.py:13-58) ------------
 * out[m, n] = sum_t sum_c A[gather(m, t), c] * Wt[t][n][c]   (+bias[n]) (+residual[m, n])
 * rows m enumerate (nb, oh, ow) over NB x OH x OW; tap t reads input pixel
 * (oh*in_stride + dy[t], ow*in_stride + dx[t]) (zero outside the image); the result is stored at
 * output pixel (oh*out_stride + out_off_y, ow*out_stride + out_off_x) of an NB x OHf x OWf image.
 * Per-channel sum / sum-of-squares of the stored values are atomically accumulated into
 * stats[g][0][n], stats[g][1][n] (g = nb / group_images) for the following BatchNorm. */
typedef struct {
  const void* A;        /* bf16 [NB, H, W, C] */
  const void* Wt;       /* bf16 [T][N][C] (packed by sv_pack_weight) */
  void* out_bf16;       /* bf16 [NB, OHf, OWf, N] or NULL */
  float* out_f32;       /* fp32 [NB, OHf, OWf, n_valid] or NULL */
  const void* residual; /* bf16, layout of out_bf16, or NULL */
  const float* bias;    /* fp32 [N] or NULL */
  float* stats;         /* fp32 [G][2][N] (accumulated) or NULL */
  int32_t NB, H, W, C;
  int32_t OH, OW, N, T;
  int32_t in_stride, out_stride, out_off_y, out_off_x;
  int32_t OHf, OWf, n_valid, group_images;
  int8_t dy[SV_MAX_TAPS];
  int8_t dx[SV_MAX_TAPS];
  int32_t impl;         /* 0 = auto, 1 = mma.sync kernel, 2 = tcgen05 + per-tap TMA kernel,
                           3 = tcgen05 halo-tile kernel (input read once, taps = shifted descriptors),
                           4 = FP32 kernel of the parity-grade mode (see "FP32 mode" below) */
  int32_t w_layout;     /* layout of Wt: 0 = [T][N][C]; 1 = [T][C/8][N][8] (required by, and selects, kernel 3);
                           2 = fp32 [T][N][C] (required by, and selects, kernel 4) */
  /* Fused BatchNorm-backward statistics (input-gradient launches): when bn_y != NULL the output IS the gradient
   * w.r.t. the activated tensor a = act(scale*y + shift) of the BatchNorm whose input y (bf16, layout of out_bf16)
   * and per-group coefficients [G][N] are given here, and instead of sum / sum-of-squares the epilogue accumulates
   *   stats[0][g][n] += sum dz            (d loss / d beta),   dz = out * (scale*y + shift > 0 ? 1 : bn_slope)
   *   stats[1][g][n] += sum dz * x_hat    (d loss / d gamma),  x_hat = (y - mean) * rsqrt(var + bn_eps)
   * (note the [2][G][N] layout).  Kernels 2 and 3 only; residual must be NULL. */
  const void* bn_y;
  const float* bn_scale;
  const float* bn_shift;
  const float* bn_mean;
  const float* bn_var;
  float bn_slope, bn_eps;
} sv_igemm_args;
int sv_igemm_fprop(const sv_igemm_args* a, void* stream);
/* FP32 mode (parity-grade precision, selected per network: plan.Net(precision="fp32")).  The reference's CPU path is
 * FP32 end to end (main_shot_vae.py:281-366 on ATen); the production kernels round conv operands to bf16.  With
 * impl = 4 / w_layout = 2 the SAME problem description is evaluated on fp32 tensors: A, residual and Wt point to
 * float data, out_bf16 must be NULL, out_f32 is [NB, OHf, OWf, n_valid > 0 ? n_valid : N], products and sums are fp32
 * FMAs on the CUDA cores.  sv_igemm_wgrad with impl = 4 takes float A / Gr.  Every activation-tensor entry point below
 * has an `_f32` twin with float tensors in place of bf16 ones (same arguments otherwise).  About 20x slower than the
 * bf16 tensor-core path; used by the parity tests (gradients within 2e-2 of the oracle end to end), never by bench.py. */
/* 1 if kernel `impl` (1, 2, 3, 4) can run this problem; impl = 0 returns the kernel auto mode selects */
int sv_igemm_fprop_supports(const sv_igemm_args* a, int32_t impl);
/* n independent problems (the output-parity phases of one transposed convolution, reference decoder.py:19-63 ->
 * nn.ConvTranspose2d(k=4, s=2, p=1)).  Up to four problems of identical geometry that run on the per-tap tcgen05
 * kernel share ONE grid; anything else is launched one after the other.  Same result as n sv_igemm_fprop calls. */
int sv_igemm_fprop_batch(const sv_igemm_args* args, int32_t n, void* stream);

/* weight-gradient GEMM (replaces the cuDNN wgrad behind every conv / convT backward):
 * part[s][n][t*C + c] = sum over the s-th slice of rows m of  Gr[m, n] * A[gather(m, t), c]
 * Gr is dense bf16 [NB*OH*OW, N]; geometry fields as for sv_igemm_fprop.  `splits` slices of the
 * row range are written to `partial` (fp32 [splits][N][T*C]); sv_wgrad_reduce sums them. */
typedef struct {
  const void* A;   /* bf16 [NB, H, W, C] */
  const void* Gr;  /* bf16 [NB*OH*OW, N] */
  float* partial;  /* fp32 [splits][N][T*C] */
  int32_t NB, H, W, C;
  int32_t OH, OW, N, T;
  int32_t in_stride, splits;
  int8_t dy[SV_MAX_TAPS];
  int8_t dx[SV_MAX_TAPS];
  int32_t impl;    /* 0 = auto, 1 = mma.sync kernel, 2 = tcgen05 (halo-tile kernel for C, N <= 128; TMA-fed kernel for wider layers),
                      4 = FP32 mode (A and Gr are float tensors; any splits >= 1) */
} sv_wgrad_args;
int sv_igemm_wgrad(const sv_wgrad_args* a, void* stream);
/* The tcgen05 kernel writes one partial slice per persistent CTA: returns the `splits` value the caller
 * must use (and size `partial` for) if that kernel will run this problem, or 0 if the mma.sync kernel
 * will (any splits >= 1). */
int sv_igemm_wgrad_splits(const sv_wgrad_args* a);
/* grad[n*sn + c*sc + tap_index[t]*st] += sum_s partial[s][n][t*C + c]  for n < n_real, c < c_real */
int sv_wgrad_reduce(const float* partial, float* grad, int32_t splits, int32_t N, int32_t C, int32_t T,
                    int32_t n_real, int32_t c_real, int64_t sn, int64_t sc, int64_t st,
                    const int8_t* tap_index /* host, T entries */, void* stream);

/* The same reduction for several weight tensors in ONE launch (launch latency, not bytes, is what 30-odd small reductions
 * per step cost): descs is a HOST array, every record has the meaning of the sv_wgrad_reduce arguments.  The partial buffers
 * must all be live until the launch has run (one workspace per weight tensor, not a shared one). */
typedef struct {
  const float* partial;
  float* grad;
  int64_t sn, sc, st;
  int32_t splits, N, C, T, n_real, c_real;
  int8_t tap_index[SV_MAX_TAPS];
} sv_wgrad_reduce_desc;
int sv_wgrad_reduce_batched(const sv_wgrad_reduce_desc* descs /* host */, int32_t n, void* stream);
int sv_sizeof_wgrad_reduce_desc(void);

/* dst(t, n, c) (bf16) = src[n*sn + c*sc + tap_index[t]*st] for n < n_real, c < c_real else 0;
 * layout 0: dst[t][n][c]; layout 1: dst[t][c/8][n][c%8] (8-channel planes, for the halo-tile kernel);
 * layout 2: dst is FLOAT [t][n][c] (FP32 mode) */
int sv_pack_weight(const float* src, void* dst, int32_t N, int32_t C, int32_t T, int32_t n_real, int32_t c_real,
                   int64_t sn, int64_t sc, int64_t st, const int8_t* tap_index /* host */, int32_t layout, void* stream);
/* All packs of a network in ONE launch.  `table_dev` is a device array of n_packs records
 * { const float* src; void* dst; int64 sn, sc, st; int32 N, C, T, n_real, c_real, layout; int8 tap[16]; }
 * (sv_sizeof_pack_desc() bytes each, same meaning as the sv_pack_weight arguments). */
int sv_pack_weights_batched(const void* table_dev, int32_t n_packs, int32_t blocks_per_pack, void* stream);
int sv_sizeof_pack_desc(void);
/* fp32 NCHW [NB, c_real, H, W] -> bf16 NHWC [NB, H, W, C] (zero padded channels); _f32: fp32 NHWC */
int sv_pack_image(const float* src, void* dst, int32_t NB, int32_t c_real, int32_t HW, int32_t C, void* stream);
int sv_pack_image_f32(const float* src, void* dst, int32_t NB, int32_t c_real, int32_t HW, int32_t C, void* stream);
/* fp32 [rows, c_real] -> fp32 [rows, C] with zero padded channels (FP32 mode: ELBO gradient -> decoder backward) */
int sv_pad_channels_f32(const float* src, float* dst, int64_t rows, int32_t c_real, int32_t C, void* stream);
/* fp32 NHWC [NB, HW, c_real] -> fp32 NCHW [NB, c_real, HW] */
int sv_nhwc_to_nchw_f32(const float* src, float* dst, int32_t NB, int32_t c_real, int32_t HW, void* stream);

/* ---- BatchNorm2d (train mode) + LeakyReLU/ReLU (wideresnet.py:27-28,32-33,39-40,90-91;
 *      preactresnet.py:29-30,33-34,54; decoder.py:19-20,...) -------------------------------------- */
/* stats[G][2][C] (sum, sumsq over `count` values) -> mean, var(biased), scale=gamma*rstd,
 * shift=beta-mean*scale, all [G][C] */
int sv_bn_finalize(const float* stats, const float* gamma, const float* beta, float count, float eps, int32_t G,
                   int32_t C, int32_t c_real, float* mean, float* var, float* scale, float* shift, void* stream);
/* a = act(y*scale + shift), act = x>0 ? x : slope*x ; y, a bf16 [NB*HW, C] */
int sv_bn_act_fwd(const void* y, void* a, const float* scale, const float* shift, float slope, int64_t rows_per_group,
                  int32_t G, int32_t C, void* stream);
/* sv_bn_finalize + sv_bn_act_fwd in one launch (scale/shift derived from the raw statistics in the kernel
 * prologue; mean/var/scale/shift are still written for the backward pass) */
int sv_bn_finalize_act_fwd(const void* y, void* a, const float* stats, const float* gamma, const float* beta, float count,
                           float eps, float slope, int64_t rows_per_group, int32_t G, int32_t C, float* mean, float* var,
                           float* scale, float* shift, void* stream);
/* feat[nb][c] = mean_hw act(y*scale+shift)   (transition BN + AdaptiveAvgPool2d, vae.py:107,142) */
int sv_bn_act_gap_fwd(const void* y, float* feat, const float* scale, const float* shift, float slope, int32_t NB,
                      int32_t HW, int32_t C, int32_t group_images, void* stream);
/* dbeta[g][c] += sum g', dgamma[g][c] += sum g' * xhat, with g' = g_a * act'(y*scale+shift).
 * g_a is bf16 [rows, C], or (g_feat != NULL) fp32 [NB][C] broadcast over HW and scaled by 1/HW. */
int sv_bn_bwd_reduce(const void* g_a, const float* g_feat, const void* y, const float* scale, const float* shift,
                     const float* mean, const float* var, float eps, float slope, int64_t rows_per_group, int32_t HW,
                     int32_t G, int32_t C, float* dgamma, float* dbeta, void* stream);
/* one term of the input gradient of BN+act */
typedef struct {
  const void* g_a;      /* bf16 [rows, C] or NULL when g_feat is used */
  const float* g_feat;  /* fp32 [NB][C] or NULL */
  const float* scale; const float* shift; const float* mean; const float* var;
  const float* dgamma; const float* dbeta;    /* [G][C], from sv_bn_bwd_reduce */
  float* grad_gamma; float* grad_beta;        /* parameter grads [c_real], accumulated (+=) */
  float slope;
  int32_t c_real;
} sv_bn_bwd_term;
/* g_y = [addend] + sum_terms scale*(g' - dbeta/M - xhat*dgamma/M)   (bf16 out) */
int sv_bn_bwd_apply(const sv_bn_bwd_term* terms, int32_t nterms, const void* y, const void* addend, void* g_y,
                    float eps, int64_t rows_per_group, int32_t HW, int32_t G, int32_t C, void* stream);
/* running_mean/var update in reference order for `npass` passes (momentum 0.1, unbiased var);
 * mean/var are [npass][C] slices selected by the caller */
int sv_bn_running_update(const float* const* mean_ptrs, const float* const* var_ptrs, int32_t npass, float count,
                         float momentum, int32_t c_real, float* running_mean, float* running_var,
                         int64_t* num_batches_tracked, void* stream);
/* every BatchNorm of the step in one launch: `table_dev` = n_bn records { const float* mean[4], var[4]; float*
 * running_mean, running_var; int64* nbt; float count; int32 npass, C } (sv_sizeof_run_desc() bytes each) */
int sv_bn_running_update_batched(const void* table_dev, int32_t n_bn, int32_t max_c, float momentum, void* stream);
int sv_sizeof_run_desc(void);
/* out[c] += sum_rows x[row][c]  (conv0 bias gradient) */
int sv_colsum_bf16(const void* x, float* out, int64_t rows, int32_t C, int32_t c_real, void* stream);
int sv_colsum_f32(const void* x, float* out, int64_t rows, int32_t C, int32_t c_real, void* stream);
/* FP32-mode twins of the BatchNorm entry points above: y / a / g_a / addend / g_y are float tensors */
int sv_bn_act_fwd_f32(const void* y, void* a, const float* scale, const float* shift, float slope, int64_t rows_per_group,
                      int32_t G, int32_t C, void* stream);
int sv_bn_finalize_act_fwd_f32(const void* y, void* a, const float* stats, const float* gamma, const float* beta, float count,
                               float eps, float slope, int64_t rows_per_group, int32_t G, int32_t C, float* mean, float* var,
                               float* scale, float* shift, void* stream);
int sv_bn_act_gap_fwd_f32(const void* y, float* feat, const float* scale, const float* shift, float slope, int32_t NB,
                          int32_t HW, int32_t C, int32_t group_images, void* stream);
int sv_bn_bwd_reduce_f32(const void* g_a, const float* g_feat, const void* y, const float* scale, const float* shift,
                         const float* mean, const float* var, float eps, float slope, int64_t rows_per_group, int32_t HW,
                         int32_t G, int32_t C, float* dgamma, float* dbeta, void* stream);
int sv_bn_bwd_apply_f32(const sv_bn_bwd_term* terms, int32_t nterms, const void* y, const void* addend, void* g_y,
                        float eps, int64_t rows_per_group, int32_t HW, int32_t G, int32_t C, void* stream);

/* ---- small FP32 linears: the three inference heads (vae.py:10-15,143-145) and the k=1
 *      ConvTranspose2d decoder stem (decoder.py:13-18) ------------------------------------------- */
/* out[b][n] = sum_k x[b][k] * W(n,k) + bias[n];  W(n,k) = w_kn ? W[k*ldw + n] : W[n*ldw + k].
 * x fp32 [B][ldx]; writes out_f32 [B][ldo] and/or out_bf16 [B][ldo]; optional stats[G][2][ldo]. */
int sv_linear_fwd(const float* x, int32_t ldx, const float* W, int32_t ldw, int32_t w_kn, const float* bias,
                  float* out_f32, void* out_bf16, int32_t ldo, float* stats, int32_t group_rows, int32_t B, int32_t N,
                  int32_t K, void* stream);
/* gx[b][k] (+)= sum_n g[b][n] * W(n,k); g fp32 or bf16 */
int sv_linear_bwd_input(const float* g_f32, const void* g_bf16, int32_t ldg, const float* W, int32_t ldw, int32_t w_kn,
                        float* gx, int32_t ldgx, int32_t accumulate, int32_t B, int32_t N, int32_t K, void* stream);
/* dW(n,k) += sum_b g[b][n] * x[b][k]; dbias[n] += sum_b g[b][n] */
int sv_linear_bwd_weight(const float* g_f32, const void* g_bf16, int32_t ldg, const float* x, int32_t ldx, float* dW,
                         int32_t ldw, int32_t w_kn, float* dbias, int32_t B, int32_t N, int32_t K, void* stream);
/* The inference heads of the VAE (vae.py:113-129: continuous mean, continuous log sigma, discrete logits) share their input:
 * one launch for all of them in each direction.  W[i] is [N[i]][K] row-major (nn.Linear), out[i] / g[i] are [B][N[i]] fp32.
 *   sv_heads_fwd:        out[i][b][n] = sum_k x[b][k] W[i][n][k] + bias[i][n]        (same arithmetic as sv_linear_fwd)
 *   sv_heads_bwd_input:  gx[b][k]     = sum_i sum_n g[i][b][n] W[i][n][k]
 *   sv_heads_bwd_weight: dW[i][n][k] += sum_b g[i][b][n] x[b][k];  dbias[i][n] += sum_b g[i][b][n]   (dbias[i] may be NULL) */
#define SV_MAX_HEADS 4
typedef struct {
  const float* W[SV_MAX_HEADS];
  const float* bias[SV_MAX_HEADS];
  float* out[SV_MAX_HEADS];
  const float* g[SV_MAX_HEADS];
  float* dW[SV_MAX_HEADS];
  float* dbias[SV_MAX_HEADS];
  int32_t N[SV_MAX_HEADS];
  int32_t n;
} sv_heads;
int sv_sizeof_heads(void);
int sv_heads_fwd(const float* x, int32_t ldx, const sv_heads* heads, int32_t B, int32_t K, void* stream);
int sv_heads_bwd_input(const sv_heads* heads, float* gx, int32_t ldgx, int32_t B, int32_t K, void* stream);
int sv_heads_bwd_weight(const sv_heads* heads, const float* x, int32_t ldx, int32_t B, int32_t K, void* stream);
int sv_log_softmax_fwd(const float* logits, float* out, int32_t B, int32_t N, void* stream);
/* g_logits = g_out - exp(out) * sum_n g_out */
int sv_log_softmax_bwd(const float* g_out, const float* out, float* g_logits, int32_t B, int32_t N, void* stream);

/* ---- Sample (vae.py:18-86): latent[b] = [ mu + exp(ls)*eps | y ] , ld = padded row length ---------
 * mode 0: y = onehot(label)               (vae.py:47-49)
 * mode 1: y = lam*onehot(label) + (1-lam)*onehot(label_mix), lam = *lam_dev, (1-lam) = lam_dev[1]
 * mode 2: y = softmax((la + gumbel(u)) / temperature)  (vae.py:58-73) */
int sv_sample_fwd(const float* mu, const float* ls, const float* la, const float* eps, const float* unif,
                  const int64_t* label, const int64_t* label_mix, const float* lam_dev, int32_t mode,
                  float temperature, int32_t B, int32_t D, int32_t nd, float* latent, int32_t ld, void* stream);
/* g_mu (+)= g_z ; g_ls (+)= g_z*eps*exp(ls) ; mode 2: g_la (+)= softmax-backward(g_y)/temperature */
int sv_sample_bwd(const float* g_latent, int32_t ld, const float* ls, const float* eps, const float* latent,
                  int32_t mode, float temperature, int32_t B, int32_t D, int32_t nd, float* g_mu, float* g_ls,
                  float* g_la, int32_t accumulate, void* stream);

/* ---- VAECriterion (criterion.py:32-57), one fused loss + gradient pass --------------------------
 * terms[0] = rec (BCE-with-logits sum / B, or MSE(sigmoid) sum / (2 B sigma^2)), terms[1] = KLc,
 * terms[2] = KLd (all accumulated with atomics into a zeroed `terms`).
 * x fp32 NCHW [B, ch, HW]; xhat fp32, NHWC [B, HW, ch] if xhat_nhwc else NCHW.
 * g_xhat (optional): d rec / d xhat * (*g_scale or 1), written as bf16 NHWC [B, HW, g_ld] (zero
 * padded) when g_bf16 != NULL and/or fp32 in xhat's layout when g_f32 != NULL. */
int sv_elbo_rec_fwd_bwd(const float* x, const float* xhat, int32_t xhat_nhwc, int32_t B, int32_t ch, int32_t HW,
                        int32_t bce, float x_sigma, const float* g_scale, float* terms, void* g_bf16, int32_t g_ld,
                        float* g_f32, void* stream);
int sv_elbo_kl_fwd(const float* mu, const float* ls, const float* la, int32_t B, int32_t D, int32_t nd, float* terms,
                   void* stream);
/* latent gradients of  w * ( kbc*|KLc - cmi| + kbd*|KLd - dmi| ) using terms[1], terms[2] from the
 * forward; coef (device) = {w, kbc, cmi, kbd, dmi}.  unit != 0: instead write the plain
 * d KLc/d mu, d KLc/d ls, d KLd/d la (the autograd path scales them itself). */
int sv_elbo_kl_bwd(const float* mu, const float* ls, const float* la, const float* terms, const float* coef,
                   int32_t unit, int32_t B, int32_t D, int32_t nd, float* g_mu, float* g_ls, float* g_la,
                   int32_t accumulate, void* stream);

/* ---- posterior-matching terms (main_shot_vae.py:316-321,358-361; ClsCriterion criterion.py:97-108)
 * terms[0] += -(1/B) sum_b w_b sum_c la[b,c]*target[b,c];  terms[1] += (|mu-mu_t|^2 + |exp(ls)-sig_t|^2)/B
 * target: dense fp32 [B][nd] (target != NULL) or lam*onehot(label_a) + (1-lam)*onehot(label_b).
 * grads (optional): g_la (+)= -c_disc*target/B ; g_mu (+)= c_cont*2(mu-mu_t)/B ;
 * g_ls (+)= c_cont*2(exp(ls)-sig_t)exp(ls)/B with c_disc = coef[0], c_cont = coef[1] (device). */
int sv_posterior_fwd_bwd(const float* la, const float* target, const int64_t* label_a, const int64_t* label_b,
                         const float* lam_dev, const float* mu, const float* ls, const float* mu_t, const float* sig_t,
                         const float* coef, int32_t B, int32_t D, int32_t nd, float* terms, float* g_la, float* g_mu,
                         float* g_ls, int32_t accumulate, void* stream);
/* inference-KL monitor (main_shot_vae.py:331-339): *out += sum alpha*(la - log smooth_onehot(label))/B */
int sv_inference_kl(const float* la, const int64_t* label, int32_t B, int32_t nd, float* out, void* stream);

/* ---- device-side noise and accumulator clears of the fused step -----------------------------------
 * normal[0..n_normal) ~ N(0,1) (the eps of Sample.forward, vae.py:82-84) and uniform[0..n_uniform) ~ U[0,1) (the Gumbel
 * uniforms, vae.py:69) from Philox4x32-10 at counter `offset`; `state` is a device struct {uint64 seed, uint64 offset,
 * uint32 ticket, pad} (sv_sizeof_noise_state() = 32 bytes) whose offset the launch advances, so CUDA-graph replays draw
 * fresh numbers.  Parity runs feed host-drawn noise (torch.randn / torch.rand in the reference's order) instead. */
int sv_noise_fill(float* normal, int64_t n_normal, float* uniform, int64_t n_uniform, void* state, void* stream);
int sv_sizeof_noise_state(void);
/* graph-capturable memset (cudaMemsetAsync: a memset node, no kernel) for the per-step accumulators */
int sv_fill_zero(void* dst, int64_t nbytes, void* stream);

/* ---- mixup / label smoothing (mixup.py:5-41) ------------------------------------------------------
 * lam_dev = {lam, 1-lam} (fp32, device).  image fp32 NCHW [B, img_elems]; writes the mixed image as
 * fp32 NCHW (mixed_f32, optional) and as bf16 NHWC with `img_ld` channels (mixed_bf16, optional). */
int sv_mixup_lerp(const float* image, const float* mu, const float* ls, const float* la, const int64_t* index,
                  const float* lam_dev, int32_t B, int32_t ch, int32_t HW, int32_t D, int32_t nd, float* mixed_f32,
                  void* mixed_bf16, int32_t img_ld, float* mixed_mu, float* mixed_sigma, float* mixed_alpha,
                  void* stream);
/* --om pairing (mixup.py:11-18, 93-99): index[i] = argmin_{2nd} KL(N_i || N_j), ties -> lower j.
 * kl_out (optional) fp32 [B][B]. */
int sv_pairwise_kl_second_nearest(const float* mu, const float* ls, int32_t B, int32_t D, int64_t* index,
                                  float* kl_out, void* stream);

/* ---- torch.optim.SGD step (main_shot_vae.py:198,365-366) over a flat FP32 arena -----------------
 * g = grad*grad_scale + wd*p ; m = first ? g : mom*m + g ; p -= lr*m ; grad = 0.  hyper (device) =
 * {lr, momentum, wd, grad_scale, first_step_flag}. */
int sv_sgd_step(float* param, float* grad, float* momentum_buf, const float* hyper, int64_t n, void* stream);

/* ---- entry points the reference defines but never calls on the training path -----------
 * Pairwise distances of lib/utils/calculate_dist.py:94-160: out[i][j] fp32 [n1][n2] between rows of (u1, ls1)
 * [n1][D] and (u2, ls2) [n2][D]; ls operands may be NULL for SV_DIST_SQ_EUCLID / SV_DIST_COSINE. */
#define SV_DIST_GAUSSIAN_KL 0   /* pairwise_norm_kl_dist_gpu (:94-107): KL(N(u1_i, e^ls1_i) || N(u2_j, e^ls2_j)) */
#define SV_DIST_SQ_EUCLID 1     /* pairwise_square_euclidean_gpu (:110-117) */
#define SV_DIST_WASSERSTEIN 2   /* pairwise_norm_wasserstein_dist_gpu (:120-130) */
#define SV_DIST_COSINE 3        /* calculate_mean_dist_pairwise(distance="cosine") (:146-149), norms squared as written */
int sv_pairwise_dist(const float* u1, const float* ls1, const float* u2, const float* ls2, int32_t n1, int32_t n2, int32_t D,
                     int32_t mode, float* out, void* stream);
/* Two-distribution forms of KLNormCriterion / KLDiscCriterion (lib/criterion.py:151-157,172-176): *loss += sum/batch
 * (loss must be zeroed by the caller), g0..g3 (optional) receive d loss / d x0..x3.
 * SV_KLPAIR_NORM: x0 = mean_pre, x1 = log_sigma_pre, x2 = mean_gt, x3 = sigma_gt;
 * SV_KLPAIR_DISC_QP / _PQ: x0 = log q (prediction), x1 = p (target probabilities), x2 = x3 = NULL. */
#define SV_KLPAIR_NORM 0
#define SV_KLPAIR_DISC_QP 1
#define SV_KLPAIR_DISC_PQ 2
int sv_kl_pair_fwd_bwd(int32_t mode, const float* x0, const float* x1, const float* x2, const float* x3, int64_t n, int32_t batch,
                       float* loss, float* g0, float* g1, float* g2, float* g3, void* stream);

/* ---- device input pipeline (lib/dataloader.py:42-70: Pad(4, reflect) -> RandomHorizontalFlip -> RandomCrop(32)
 *      -> ToTensor) over a device-resident uint8 dataset ------------------------------------------------------
 * data: uint8 [n_images][src_h][src_w][ch] (src_hwc = 1: CIFAR / MNIST storage) or [n_images][ch][src_h][src_w]
 * (src_hwc = 0: SVHN storage); index (optional) int64 [B] selects the images; params (optional) int32 [B][3] =
 * {crop row i, crop column j, flip} with i, j in [0, src + 2*pad - out] (NULL: no augmentation, i = j = flip = 0).
 * out: fp32 NCHW [B][ch][out_h][out_w] in [0, 1], bit-exact with torchvision (uint8 / 255 in fp32). */
int sv_augment_batch(const uint8_t* data, const int64_t* index, const int32_t* params, int32_t B, int32_t ch, int32_t src_h,
                     int32_t src_w, int32_t pad, int32_t out_h, int32_t out_w, int32_t src_hwc, float* out, void* stream);

/* debugging aid: per-tile clock64 stamps of CTA 0 of the last halo-kernel launch run with
 * SHOTVAE_HALO_TRACE=1 ([role: producer, mma, mma-acc-wait, epilogue][tile < 64][begin, end]) */
int sv_debug_halo_trace(long long* host_out);

/* struct sizes, so the ctypes mirror can be checked at load time */
int sv_sizeof_igemm_args(void);
int sv_sizeof_wgrad_args(void);
int sv_sizeof_bn_bwd_term(void);

#ifdef __cplusplus
}
#endif
#endif /* SHOTVAE_H_ */
