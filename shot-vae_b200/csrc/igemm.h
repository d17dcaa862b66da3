// Internal parameter blocks shared by the implicit-GEMM kernels (mma.sync and tcgen05 paths).
#pragma once
#include "common.cuh"
#include "../../include/shotvae.h"

struct IgemmParams {
  const bf16* A;
  const bf16* Wt;
  bf16* out;
  float* outf;
  const bf16* res;
  const float* bias;
  float* stats;
  int NB, H, W, C;
  int OH, OW, N, T;
  int in_stride, out_stride, out_off_y, out_off_x;
  int OHf, OWf, n_valid, group_images;
  int M;               // NB*OH*OW
  int rows_per_group;  // group_images*OH*OW
  int w_layout;        // 0 = [T][N][C], 1 = [T][C/8][N][8], 2 = fp32 [T][N][C] (FP32 mode: A / res / Wt are float tensors)
  int8_t dy[SV_MAX_TAPS];
  int8_t dx[SV_MAX_TAPS];
  // fused BatchNorm-backward statistics (see sv_igemm_args)
  const bf16* bn_y;
  const float* bn_scale;
  const float* bn_shift;
  const float* bn_mean;
  const float* bn_var;
  float bn_slope, bn_eps;
};

// Epilogue helper shared by the tcgen05 kernels: coef = {scale, shift, rstd, -mean*rstd} of 4 consecutive channels.
// v: output-gradient values (already rounded to bf16) -> dz in v, dz * x_hat in w.
__device__ __forceinline__ void bn_bwd_terms4(const float* yv, const float4 sc, const float4 sh, const float4 rs, const float4 mr,
                                              float slope, float* v, float* w) {
  const float s[4] = {sc.x, sc.y, sc.z, sc.w}, h[4] = {sh.x, sh.y, sh.z, sh.w}, r[4] = {rs.x, rs.y, rs.z, rs.w},
              m[4] = {mr.x, mr.y, mr.z, mr.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float pre = fmaf(yv[j], s[j], h[j]);
    const float dz = pre > 0.f ? v[j] : slope * v[j];
    v[j] = dz;
    w[j] = dz * fmaf(yv[j], r[j], m[j]);
  }
}

struct WgradParams {
  const bf16* A;
  const bf16* Gr;
  float* partial;
  int NB, H, W, C;
  int OH, OW, N, T;
  int in_stride, splits;
  int M, rows_per_split;
  int8_t dy[SV_MAX_TAPS];
  int8_t dx[SV_MAX_TAPS];
};

int igemm_fprop_mma(const IgemmParams& p, cudaStream_t st);
int igemm_fprop_mma_batch(const IgemmParams* batch, int nb, cudaStream_t st);   // same-geometry problems, one grid
int igemm_wgrad_mma(const WgradParams& p, cudaStream_t st);
// tcgen05 + TMA path (igemm_tc.cu). Returns SV_ERR_UNSUPPORTED when the shape is not covered.
bool igemm_fprop_tc_supported(const IgemmParams& p);
int igemm_fprop_tc(const IgemmParams& p, cudaStream_t st);
int igemm_fprop_tc_batch(const IgemmParams* ps, int n, cudaStream_t st);   // <= 4 problems of identical tile shape, one grid
// tcgen05 halo-tile path (igemm_halo.cu)
bool igemm_fprop_halo_supported(const IgemmParams& p);
int igemm_fprop_halo(const IgemmParams& p, cudaStream_t st);
// tcgen05 halo-tile weight gradient (wgrad_halo.cu)
bool wgrad_halo_supported(const WgradParams& p);
int wgrad_halo_splits(const WgradParams& p);
int wgrad_halo(const WgradParams& p, cudaStream_t st);
// tcgen05 + TMA weight gradient for wide layers (wgrad_tc.cu)
bool wgrad_tc_supported(const WgradParams& p);
int wgrad_tc_splits(const WgradParams& p);
int wgrad_tc(const WgradParams& p, cudaStream_t st);
// parity-grade FP32 mode (igemm_f32.cu): fp32 activations and weights, CUDA-core FMA
bool igemm_fprop_f32_supported(const IgemmParams& p);
int igemm_fprop_f32(const IgemmParams& p, cudaStream_t st);
int igemm_wgrad_f32(const WgradParams& p, cudaStream_t st);
