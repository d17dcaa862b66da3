// C-ABI plumbing: error state, launch accounting, argument validation, kernel selection.
#include <stdarg.h>
#include <stdlib.h>
#include <atomic>
#include "common.cuh"
#include "igemm.h"

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};
static std::atomic<int> g_cta_limit{0};

int sv_cta_limit() { return g_cta_limit.load(std::memory_order_relaxed); }

void sv_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

bool sv_pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("SHOTVAE_PDL");
    on = (e && e[0] == '1') ? 1 : 0;     // MEASURED: 6.33 vs 6.35 ms/step inside the CUDA graph -> off by default
  }
  return on != 0;
}

int sv_check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    sv_set_error("%s: %s", what, cudaGetErrorString(e));
    cudaGetLastError();
    return SV_ERR_CUDA;
  }
  return SV_OK;
}

extern "C" {

int sv_abi_version(void) { return SV_ABI_VERSION; }
const char* sv_last_error(void) { return g_err; }
long long sv_launch_count(void) { return g_launches.load(); }
int sv_set_cta_limit(int32_t n) {
  const int old = g_cta_limit.load();
  g_cta_limit.store(n > 0 ? n : 0);
  return old;
}
int sv_sizeof_igemm_args(void) { return (int)sizeof(sv_igemm_args); }
int sv_sizeof_wgrad_args(void) { return (int)sizeof(sv_wgrad_args); }
int sv_sizeof_bn_bwd_term(void) { return (int)sizeof(sv_bn_bwd_term); }

#ifdef SV_NO_TCGEN05
int sv_has_tcgen05(void) { return 0; }
#else
int sv_has_tcgen05(void) { return 1; }
#endif

static int fill_params(const sv_igemm_args* a, IgemmParams& p) {
  SV_REQUIRE(a && a->A && a->Wt, "sv_igemm_fprop: null operand");
  SV_REQUIRE(a->C % 16 == 0 && a->N % 16 == 0, "sv_igemm_fprop: C (%d) and N (%d) must be multiples of 16", a->C, a->N);
  SV_REQUIRE(a->T >= 1 && a->T <= SV_MAX_TAPS, "sv_igemm_fprop: T=%d out of range", a->T);
  SV_REQUIRE(a->NB > 0 && a->OH > 0 && a->OW > 0 && a->group_images > 0, "sv_igemm_fprop: bad geometry");
  SV_REQUIRE(a->out_bf16 || a->out_f32, "sv_igemm_fprop: no output");
  SV_REQUIRE((long long)a->NB * a->OH * a->OW < (1ll << 31), "sv_igemm_fprop: too many rows");
  p.A = (const bf16*)a->A; p.Wt = (const bf16*)a->Wt; p.out = (bf16*)a->out_bf16; p.outf = a->out_f32;
  p.res = (const bf16*)a->residual; p.bias = a->bias; p.stats = a->stats;
  p.NB = a->NB; p.H = a->H; p.W = a->W; p.C = a->C; p.OH = a->OH; p.OW = a->OW; p.N = a->N; p.T = a->T;
  p.in_stride = a->in_stride; p.out_stride = a->out_stride; p.out_off_y = a->out_off_y; p.out_off_x = a->out_off_x;
  p.OHf = a->OHf; p.OWf = a->OWf; p.n_valid = a->n_valid; p.group_images = a->group_images;
  p.M = a->NB * a->OH * a->OW;
  p.rows_per_group = a->group_images * a->OH * a->OW;
  p.w_layout = a->w_layout;
  memcpy(p.dy, a->dy, SV_MAX_TAPS); memcpy(p.dx, a->dx, SV_MAX_TAPS);
  p.bn_y = (const bf16*)a->bn_y; p.bn_scale = a->bn_scale; p.bn_shift = a->bn_shift; p.bn_mean = a->bn_mean; p.bn_var = a->bn_var;
  p.bn_slope = a->bn_slope; p.bn_eps = a->bn_eps;
  if (p.bn_y != nullptr) {
    SV_REQUIRE(p.stats && p.bn_scale && p.bn_shift && p.bn_mean && p.bn_var, "sv_igemm_fprop: fused BatchNorm-backward statistics need stats + coefficients");
    SV_REQUIRE(p.res == nullptr && p.out != nullptr && p.outf == nullptr, "sv_igemm_fprop: fused BatchNorm-backward statistics: bf16 output, no residual");
  }
  return SV_OK;
}

// SHOTVAE_IGEMM=mma forces the mma.sync kernels in auto mode (debug switch, not a dispatch layer)
static bool auto_tc_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SHOTVAE_IGEMM");
    v = (e && strcmp(e, "mma") == 0) ? 0 : 1;
  }
  return v == 1;
}

static int select_impl(const IgemmParams& p) {
  if (p.w_layout == 2) return 4;
#ifndef SV_NO_TCGEN05
  if (p.w_layout == 1) return 3;
  if (auto_tc_enabled() && igemm_fprop_tc_supported(p)) return 2;
#endif
  return 1;
}

int sv_igemm_fprop_supports(const sv_igemm_args* a, int32_t impl) {
  IgemmParams p;
  if (fill_params(a, p) != SV_OK) return 0;
  if (impl == 0) {
    const int sel = select_impl(p);
    if (sel == 4) return igemm_fprop_f32_supported(p) ? 4 : 0;
    return (p.bn_y != nullptr && sel == 1) ? 0 : sel;      // the mma.sync kernel has no fused BatchNorm-backward epilogue
  }
  if (impl == 1) return (p.w_layout == 0 && p.bn_y == nullptr) ? 1 : 0;
  if (impl == 4) return igemm_fprop_f32_supported(p) ? 1 : 0;
#ifndef SV_NO_TCGEN05
  if (impl == 2) return igemm_fprop_tc_supported(p) ? 1 : 0;
  if (impl == 3) return igemm_fprop_halo_supported(p) ? 1 : 0;
#endif
  return 0;
}

int sv_igemm_fprop(const sv_igemm_args* a, void* stream) {
  IgemmParams p;
  int rc = fill_params(a, p);
  if (rc != SV_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  int impl = a->impl;
  if (impl == 0) impl = select_impl(p);
#ifndef SV_NO_TCGEN05
  if (impl == 2) {
    SV_REQUIRE(igemm_fprop_tc_supported(p), "sv_igemm_fprop: shape not supported by the tcgen05 per-tap kernel");
    return igemm_fprop_tc(p, st);
  }
  if (impl == 3) {
    SV_REQUIRE(igemm_fprop_halo_supported(p), "sv_igemm_fprop: shape not supported by the tcgen05 halo kernel");
    return igemm_fprop_halo(p, st);
  }
#endif
  if (impl == 4) {
    SV_REQUIRE(igemm_fprop_f32_supported(p), "sv_igemm_fprop: the FP32 kernel needs fp32 weights (w_layout 2), an fp32 output and no fused statistics");
    return igemm_fprop_f32(p, st);
  }
  SV_REQUIRE(impl == 1, "sv_igemm_fprop: unknown impl %d", impl);
  SV_REQUIRE(p.bn_y == nullptr, "sv_igemm_fprop: fused BatchNorm-backward statistics are not available on the mma.sync kernel");
  SV_REQUIRE(p.w_layout == 0, "sv_igemm_fprop: plane-interleaved weights are only consumed by the halo kernel");
  return igemm_fprop_mma(p, st);
}

int sv_igemm_fprop_batch(const sv_igemm_args* args, int32_t n, void* stream) {
  SV_REQUIRE(args && n >= 1, "sv_igemm_fprop_batch: no problems");
#ifndef SV_NO_TCGEN05
  if (n <= 4) {
    // one grid when every problem runs on the per-tap tcgen05 kernel with the same geometry
    IgemmParams ps[4];
    bool same = true;
    for (int i = 0; i < n && same; ++i) {
      if (fill_params(&args[i], ps[i]) != SV_OK) return SV_ERR_ARG;
      int impl = args[i].impl;
      if (impl == 0) impl = select_impl(ps[i]);
      same = impl == 2 && igemm_fprop_tc_supported(ps[i]) && ps[i].NB == ps[0].NB && ps[i].H == ps[0].H && ps[i].W == ps[0].W &&
             ps[i].C == ps[0].C && ps[i].N == ps[0].N && ps[i].OH == ps[0].OH && ps[i].OW == ps[0].OW;
    }
    if (same && n > 1) return igemm_fprop_tc_batch(ps, n, (cudaStream_t)stream);
  }
#endif
  if (n > 1 && n <= 4) {
    // ... or when every problem lands on the mma.sync kernel with the same geometry (last decoder layer: 64 -> 3 channels)
    IgemmParams ps[4];
    bool same = true;
    for (int i = 0; i < n && same; ++i) {
      if (fill_params(&args[i], ps[i]) != SV_OK) return SV_ERR_ARG;
      int impl = args[i].impl;
      if (impl == 0) impl = select_impl(ps[i]);
      same = impl == 1 && ps[i].bn_y == nullptr && ps[i].w_layout == 0 && ps[i].NB == ps[0].NB && ps[i].H == ps[0].H && ps[i].W == ps[0].W && ps[i].C == ps[0].C &&
             ps[i].N == ps[0].N && ps[i].OH == ps[0].OH && ps[i].OW == ps[0].OW && ps[i].T == ps[0].T && ps[i].w_layout == ps[0].w_layout;
    }
    if (same) return igemm_fprop_mma_batch(ps, n, (cudaStream_t)stream);
  }
  for (int i = 0; i < n; ++i) {
    const int rc = sv_igemm_fprop(&args[i], stream);
    if (rc != SV_OK) return rc;
  }
  return SV_OK;
}

static int fill_wgrad(const sv_wgrad_args* a, WgradParams& p) {
  SV_REQUIRE(a && a->A && a->Gr && a->partial, "sv_igemm_wgrad: null operand");
  SV_REQUIRE(a->C % 8 == 0 && a->N % 16 == 0, "sv_igemm_wgrad: C (%d) %% 8, N (%d) %% 16", a->C, a->N);
  SV_REQUIRE(a->T >= 1 && a->T <= SV_MAX_TAPS && a->splits >= 1, "sv_igemm_wgrad: bad T/splits");
  p.A = (const bf16*)a->A; p.Gr = (const bf16*)a->Gr; p.partial = a->partial;
  p.NB = a->NB; p.H = a->H; p.W = a->W; p.C = a->C; p.OH = a->OH; p.OW = a->OW; p.N = a->N; p.T = a->T;
  p.in_stride = a->in_stride; p.splits = a->splits;
  p.M = a->NB * a->OH * a->OW;
  p.rows_per_split = ((p.M + a->splits - 1) / a->splits + 31) / 32 * 32;
  memcpy(p.dy, a->dy, SV_MAX_TAPS); memcpy(p.dx, a->dx, SV_MAX_TAPS);
  return SV_OK;
}

// 0 = mma.sync kernel, 2 = tcgen05 halo-tile kernel (narrow layers), 3 = tcgen05 + TMA kernel (wide layers)
static int wgrad_kernel(const sv_wgrad_args* a, const WgradParams& p) {
  if (a->impl == 4) return 4;
#ifndef SV_NO_TCGEN05
  if (a->impl == 1) return 0;
  if (a->impl == 0 && !auto_tc_enabled()) return 0;
  if (wgrad_halo_supported(p)) return 2;
  if (wgrad_tc_supported(p)) return 3;
#endif
  return 0;
}

int sv_igemm_wgrad_splits(const sv_wgrad_args* a) {
  WgradParams p;
  sv_wgrad_args tmp = *a;
  if (tmp.splits < 1) tmp.splits = 1;
  if (fill_wgrad(&tmp, p) != SV_OK) return 0;
#ifndef SV_NO_TCGEN05
  const int k = wgrad_kernel(a, p);
  if (k == 2) return wgrad_halo_splits(p);
  if (k == 3) return wgrad_tc_splits(p);
#endif
  return 0;
}

int sv_igemm_wgrad(const sv_wgrad_args* a, void* stream) {
  WgradParams p;
  int rc = fill_wgrad(a, p);
  if (rc != SV_OK) return rc;
  if (a->impl == 4) return igemm_wgrad_f32(p, (cudaStream_t)stream);      // FP32 mode: A and Gr are float tensors
#ifndef SV_NO_TCGEN05
  const int k = wgrad_kernel(a, p);
  if (k == 2) return wgrad_halo(p, (cudaStream_t)stream);
  if (k == 3) return wgrad_tc(p, (cudaStream_t)stream);
  SV_REQUIRE(a->impl != 2, "sv_igemm_wgrad: shape not supported by the tcgen05 kernels");
#endif
  return igemm_wgrad_mma(p, (cudaStream_t)stream);
}

}  // extern "C"
