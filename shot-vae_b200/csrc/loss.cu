// Bandwidth-bound kernels of the SHOT-VAE step: reparameterisation / gumbel-softmax sampling, the
// smooth-ELBO loss with its gradient in one pass, posterior-matching terms, optimal-interpolation
// mixup (gather + lerp), the --om pairwise-KL pairing and the SGD update.
#include <math.h>
#include "common.cuh"
#include "../../include/shotvae.h"

namespace {

// ---------------------------------------------------------------------------------------- sample
__global__ void __launch_bounds__(128) sample_fwd_kernel(const float* __restrict__ mu, const float* __restrict__ ls,
                                                         const float* __restrict__ la, const float* __restrict__ eps,
                                                         const float* __restrict__ unif, const long long* __restrict__ label,
                                                         const long long* __restrict__ label_mix, const float* __restrict__ lam_dev,
                                                         int mode, float temperature, int D, int nd, float* __restrict__ latent,
                                                         int ld) {
  const int b = blockIdx.x;
  float* out = latent + (size_t)b * ld;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const size_t k = (size_t)b * D + d;
    out[d] = mu[k] + expf(ls[k]) * eps[k];
  }
  for (int d = D + nd + threadIdx.x; d < ld; d += blockDim.x) out[d] = 0.f;
  if (threadIdx.x >= 32) return;
  const int lane = threadIdx.x;
  if (mode == 0) {
    const int y = (int)label[b];
    for (int c = lane; c < nd; c += 32) out[D + c] = (c == y) ? 1.f : 0.f;
  } else if (mode == 1) {
    const int ya = (int)label[b], yb = (int)label_mix[b];
    const float l = lam_dev[0], oml = lam_dev[1];
    for (int c = lane; c < nd; c += 32)
      out[D + c] = __fadd_rn(__fmul_rn(l, (c == ya) ? 1.f : 0.f), __fmul_rn(oml, (c == yb) ? 1.f : 0.f));
  } else {
    const float EPS = 1e-12f;
    float mx = -INFINITY;
    for (int c = lane; c < nd; c += 32) {
      const size_t k = (size_t)b * nd + c;
      const float gmb = -logf(-logf(unif[k] + EPS) + EPS);
      const float logit = (la[k] + gmb) / temperature;
      out[D + c] = logit;
      mx = fmaxf(mx, logit);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float s = 0.f;
    for (int c = lane; c < nd; c += 32) {
      const float e = expf(out[D + c] - mx);
      out[D + c] = e;
      s += e;
    }
    s = warp_sum(s);
    for (int c = lane; c < nd; c += 32) out[D + c] = out[D + c] / s;
  }
}

__global__ void __launch_bounds__(128) sample_bwd_kernel(const float* __restrict__ g_latent, int ld, const float* __restrict__ ls,
                                                         const float* __restrict__ eps, const float* __restrict__ latent, int mode,
                                                         float temperature, int D, int nd, float* g_mu, float* g_ls, float* g_la,
                                                         int accumulate) {
  const int b = blockIdx.x;
  const float* g = g_latent + (size_t)b * ld;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const size_t k = (size_t)b * D + d;
    const float gz = g[d];
    const float gl = gz * eps[k] * expf(ls[k]);
    if (accumulate) { g_mu[k] += gz; g_ls[k] += gl; } else { g_mu[k] = gz; g_ls[k] = gl; }
  }
  if (mode != 2 || threadIdx.x >= 32 || g_la == nullptr) return;
  const int lane = threadIdx.x;
  const float* y = latent + (size_t)b * ld + D;
  float dot = 0.f;
  for (int c = lane; c < nd; c += 32) dot += y[c] * g[D + c];
  dot = warp_sum(dot);
  for (int c = lane; c < nd; c += 32) {
    const float v = y[c] * (g[D + c] - dot) / temperature;
    const size_t k = (size_t)b * nd + c;
    if (accumulate) g_la[k] += v; else g_la[k] = v;
  }
}

// ------------------------------------------------------------------------------------------ ELBO
template <bool BCE>
__global__ void __launch_bounds__(256) elbo_rec_kernel(const float* __restrict__ x, const float* __restrict__ xhat, int xhat_nhwc,
                                                       long long npix, int ch, int HW, float inv_b, float mse_scale,
                                                       const float* __restrict__ g_scale, float* terms, bf16* g_bf16, int g_ld,
                                                       float* g_f32) {
  __shared__ float red[8];
  const float gs = (g_scale ? *g_scale : 1.f) * inv_b;
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < npix; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / HW, p = i - b * HW;
    float gv[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) gv[c] = 0.f;
    for (int c = 0; c < ch; ++c) {
      const size_t xi = ((size_t)b * ch + c) * HW + p;
      const size_t hi = xhat_nhwc ? (size_t)i * ch + c : xi;
      const float t = x[xi], z = xhat[hi];
      float g;
      if (BCE) {
        // binary_cross_entropy_with_logits: (1-t)*z + max(-z,0) + log1p(exp(-|z|))
        acc += (1.f - t) * z + fmaxf(-z, 0.f) + log1pf(expf(-fabsf(z)));
        g = 1.f / (1.f + expf(-z)) - t;
      } else {
        const float s = 1.f / (1.f + expf(-z));
        const float d = s - t;
        acc += d * d * mse_scale;
        g = 2.f * mse_scale * d * s * (1.f - s);
      }
      g *= gs;
      if (c < 16) gv[c] = g;
      if (g_f32) g_f32[hi] = g;
    }
    if (g_bf16) {
      for (int c0 = 0; c0 < g_ld; c0 += 8) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (c0 + j < 16) ? gv[c0 + j] : 0.f;
        *reinterpret_cast<bf16x8*>(g_bf16 + (size_t)i * g_ld + c0) = pack8(v);
      }
    }
  }
  const float s = block_sum<256>(acc, red);
  if (threadIdx.x == 0) atomicAdd(&terms[0], s * inv_b);
}

// Vectorised variant (HW % 4 == 0, CH = 1 or 3 channels, 16-byte aligned operands): one thread owns 4 consecutive pixels of
// one image, so every global access is a 16-byte vector (x planes and an NCHW x_hat / gradient as float4 per channel, an
// NHWC x_hat / gradient as CH consecutive float4, the bf16 NHWC gradient as 4 x 32 bytes) -- the kernel streams
// x + x_hat in and the gradient out at HBM rate instead of issuing 4-byte accesses.
template <bool BCE, int CH>
__global__ void __launch_bounds__(256) elbo_rec_vec_kernel(const float* __restrict__ x, const float* __restrict__ xhat, int xhat_nhwc,
                                                           long long nquad, int HW, float inv_b, float mse_scale,
                                                           const float* __restrict__ g_scale, float* terms, bf16* g_bf16, int g_ld,
                                                           float* g_f32) {
  __shared__ float red[8];
  const float gs = (g_scale ? *g_scale : 1.f) * inv_b;
  const int qpi = HW >> 2;      // quads per image
  float acc = 0.f;
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < nquad; q += (long long)gridDim.x * blockDim.x) {
    const long long b = q / qpi;
    const int p = (int)(q - b * qpi) << 2;
    float t[CH][4], z[CH][4], g[CH][4];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const float4 v = *reinterpret_cast<const float4*>(x + ((size_t)b * CH + c) * HW + p);
      t[c][0] = v.x; t[c][1] = v.y; t[c][2] = v.z; t[c][3] = v.w;
    }
    if (xhat_nhwc) {
      float f[4 * CH];
#pragma unroll
      for (int k = 0; k < CH; ++k) {
        const float4 v = *reinterpret_cast<const float4*>(xhat + ((size_t)b * HW + p) * CH + 4 * k);
        f[4 * k] = v.x; f[4 * k + 1] = v.y; f[4 * k + 2] = v.z; f[4 * k + 3] = v.w;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int c = 0; c < CH; ++c) z[c][j] = f[j * CH + c];
    } else {
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const float4 v = *reinterpret_cast<const float4*>(xhat + ((size_t)b * CH + c) * HW + p);
        z[c][0] = v.x; z[c][1] = v.y; z[c][2] = v.z; z[c][3] = v.w;
      }
    }
#pragma unroll
    for (int c = 0; c < CH; ++c)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        // one exponential serves the softplus and the sigmoid: e = exp(-|z|) in (0, 1], sigmoid(z) = (z >= 0 ? 1 : e) / (1 + e).
        // Fast-math intrinsics (ex2.approx / lg2.approx / rcp.approx, ~1e-6 relative) keep the kernel bandwidth bound: with
        // expf x 2 + log1pf + a division per element it was issue bound at 63 % of the HBM rate (MEASURED, B = 16384).
        const float zz = z[c][j], tt = t[c][j];
        const float e = __expf(-fabsf(zz));
        const float r = __fdividef(1.f, 1.f + e);
        const float sg = zz >= 0.f ? r : e * r;
        float gg;
        if (BCE) {
          acc += (1.f - tt) * zz + fmaxf(-zz, 0.f) + __logf(1.f + e);
          gg = sg - tt;
        } else {
          const float d = sg - tt;
          acc += d * d * mse_scale;
          gg = 2.f * mse_scale * d * sg * (1.f - sg);
        }
        g[c][j] = gg * gs;
      }
    if (g_f32) {
      if (xhat_nhwc) {
        float f[4 * CH];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int c = 0; c < CH; ++c) f[j * CH + c] = g[c][j];
#pragma unroll
        for (int k = 0; k < CH; ++k)
          *reinterpret_cast<float4*>(g_f32 + ((size_t)b * HW + p) * CH + 4 * k) = make_float4(f[4 * k], f[4 * k + 1], f[4 * k + 2], f[4 * k + 3]);
      } else {
#pragma unroll
        for (int c = 0; c < CH; ++c)
          *reinterpret_cast<float4*>(g_f32 + ((size_t)b * CH + c) * HW + p) = make_float4(g[c][0], g[c][1], g[c][2], g[c][3]);
      }
    }
    if (g_bf16) {
      const float zero[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      const bf16x8 z8 = pack8(zero);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        bf16* dst = g_bf16 + ((size_t)b * HW + p + j) * g_ld;
        float v[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) v[c] = c < CH ? g[c < CH ? c : 0][j] : 0.f;
        const bf16x8 v8 = pack8(v);
        if (g_ld == 16 && (reinterpret_cast<uintptr_t>(dst) & 31) == 0) {
          st_global_32B(dst, v8, z8);          // one full 32-byte sector per pixel
        } else {
          *reinterpret_cast<bf16x8*>(dst) = v8;
          for (int c0 = 8; c0 < g_ld; c0 += 8) *reinterpret_cast<bf16x8*>(dst + c0) = z8;
        }
      }
    }
  }
  const float s = block_sum<256>(acc, red);
  if (threadIdx.x == 0) atomicAdd(&terms[0], s * inv_b);
}

__global__ void __launch_bounds__(256) elbo_kl_fwd_kernel(const float* __restrict__ mu, const float* __restrict__ ls,
                                                          const float* __restrict__ la, int nc, int nda, float inv_b, float log_prior,
                                                          float* terms) {
  __shared__ float red[8];
  float kc = 0.f, kd = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nc; i += gridDim.x * blockDim.x) {
    const float m = mu[i], l2 = 2.f * ls[i];
    kc += m * m + expf(l2) - l2 - 1.f;
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nda; i += gridDim.x * blockDim.x) {
    const float a = la[i];
    kd += expf(a) * (a - log_prior);
  }
  const float skc = block_sum<256>(kc, red);
  const float skd = block_sum<256>(kd, red);
  if (threadIdx.x == 0) {
    atomicAdd(&terms[1], 0.5f * skc * inv_b);
    atomicAdd(&terms[2], skd * inv_b);
  }
}

__device__ __forceinline__ float sgn(float v) { return (v > 0.f) ? 1.f : ((v < 0.f) ? -1.f : 0.f); }

__global__ void elbo_kl_bwd_kernel(const float* __restrict__ mu, const float* __restrict__ ls, const float* __restrict__ la,
                                   const float* __restrict__ terms, const float* __restrict__ coef, int unit, int nc, int nda,
                                   float inv_b, float log_prior, float* g_mu, float* g_ls, float* g_la, int accumulate) {
  float cc = inv_b, cd = inv_b;
  if (!unit) {
    // coef = {w, kbc, cmi, kbd, dmi}
    cc = coef[0] * coef[1] * sgn(terms[1] - coef[2]) * inv_b;
    cd = coef[0] * coef[3] * sgn(terms[2] - coef[4]) * inv_b;
  }
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nc) {
    const float gm = cc * mu[i];
    const float gl = cc * (expf(2.f * ls[i]) - 1.f);
    if (accumulate) { g_mu[i] += gm; g_ls[i] += gl; } else { g_mu[i] = gm; g_ls[i] = gl; }
  }
  if (i < nda) {
    const float a = la[i];
    const float ga = cd * expf(a) * (a - log_prior + 1.f);
    if (accumulate) g_la[i] += ga; else g_la[i] = ga;
  }
}

// ------------------------------------------------------------------------- posterior matching
__global__ void __launch_bounds__(256) posterior_kernel(const float* __restrict__ la, const float* __restrict__ target,
                                                        const long long* __restrict__ label_a, const long long* __restrict__ label_b,
                                                        const float* __restrict__ lam_dev, const float* __restrict__ mu,
                                                        const float* __restrict__ ls, const float* __restrict__ mu_t,
                                                        const float* __restrict__ sig_t, const float* __restrict__ coef, int B, int D,
                                                        int nd, float inv_b, float* terms, float* g_la, float* g_mu, float* g_ls,
                                                        int accumulate) {
  __shared__ float red[8];
  const float c_disc = coef ? coef[0] : 1.f, c_cont = coef ? coef[1] : 1.f;
  float ce = 0.f, mse = 0.f;
  const int nda = B * nd, nc = B * D;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nda; i += gridDim.x * blockDim.x) {
    float t;
    if (target) {
      t = target[i];
    } else {
      const int b = i / nd, c = i - b * nd;
      const float l = lam_dev ? lam_dev[0] : 1.f, oml = lam_dev ? lam_dev[1] : 0.f;
      t = l * ((c == (int)label_a[b]) ? 1.f : 0.f);
      if (label_b) t += oml * ((c == (int)label_b[b]) ? 1.f : 0.f);
    }
    ce += la[i] * t;
    if (g_la) {
      const float g = -c_disc * t * inv_b;
      if (accumulate) g_la[i] += g; else g_la[i] = g;
    }
  }
  if (mu != nullptr) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nc; i += gridDim.x * blockDim.x) {
      const float dm = mu[i] - mu_t[i];
      const float e = expf(ls[i]);
      const float dsg = e - sig_t[i];
      mse += dm * dm + dsg * dsg;
      if (g_mu) {
        const float gm = c_cont * 2.f * dm * inv_b, gl = c_cont * 2.f * dsg * e * inv_b;
        if (accumulate) { g_mu[i] += gm; g_ls[i] += gl; } else { g_mu[i] = gm; g_ls[i] = gl; }
      }
    }
  }
  const float sce = block_sum<256>(ce, red);
  const float smse = block_sum<256>(mse, red);
  if (threadIdx.x == 0) {
    atomicAdd(&terms[0], -sce * inv_b);
    if (mu != nullptr) atomicAdd(&terms[1], smse * inv_b);
  }
}

__global__ void __launch_bounds__(256) inference_kl_kernel(const float* __restrict__ la, const long long* __restrict__ label, int B,
                                                           int nd, float inv_b, float* out) {
  __shared__ float red[8];
  const float off = 0.001f / (float)(nd - 1);
  const float on = 1.f - 0.001f - 0.001f / (float)(nd - 1) + off;
  float acc = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B * nd; i += gridDim.x * blockDim.x) {
    const int b = i / nd, c = i - b * nd;
    const float a = la[i], al = expf(a);
    const float sm = (c == (int)label[b]) ? on : off;
    acc += al * a - al * logf(sm);
  }
  const float s = block_sum<256>(acc, red);
  if (threadIdx.x == 0) atomicAdd(out, s * inv_b);
}

// ----------------------------------------------------------------------------------------- mixup
__global__ void __launch_bounds__(256) mixup_image_kernel(const float* __restrict__ image, const long long* __restrict__ index,
                                                          const float* __restrict__ lam_dev, long long npix, int ch, int HW,
                                                          float* mixed_f32, bf16* mixed_bf16, int img_ld) {
  const float l = lam_dev[0], oml = lam_dev[1];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < npix; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / HW, p = i - b * HW;
    const long long b2 = index[b];
    float v[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) v[c] = 0.f;
    for (int c = 0; c < ch; ++c) {
      const size_t k = ((size_t)b * ch + c) * HW + p;
      const float m = __fadd_rn(__fmul_rn(l, image[k]), __fmul_rn(oml, image[((size_t)b2 * ch + c) * HW + p]));
      if (mixed_f32) mixed_f32[k] = m;
      if (c < 16) v[c] = m;
    }
    if (mixed_bf16) {
      for (int c0 = 0; c0 < img_ld; c0 += 8) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (c0 + j < 16) ? v[c0 + j] : 0.f;
        *reinterpret_cast<bf16x8*>(mixed_bf16 + (size_t)i * img_ld + c0) = pack8(o);
      }
    }
  }
}

// 4 consecutive pixels per thread, 16-byte accesses (HW % 4 == 0, CH = 1 or 3)
template <int CH>
__global__ void __launch_bounds__(256) mixup_image_vec_kernel(const float* __restrict__ image, const long long* __restrict__ index,
                                                              const float* __restrict__ lam_dev, long long nquad, int HW,
                                                              float* mixed_f32, bf16* mixed_bf16, int img_ld) {
  const float l = lam_dev[0], oml = lam_dev[1];
  const int qpi = HW >> 2;
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < nquad; q += (long long)gridDim.x * blockDim.x) {
    const long long b = q / qpi;
    const int p = (int)(q - b * qpi) << 2;
    const long long b2 = index[b];
    float m[CH][4];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const float4 a = *reinterpret_cast<const float4*>(image + ((size_t)b * CH + c) * HW + p);
      const float4 o = *reinterpret_cast<const float4*>(image + ((size_t)b2 * CH + c) * HW + p);
      m[c][0] = __fadd_rn(__fmul_rn(l, a.x), __fmul_rn(oml, o.x));
      m[c][1] = __fadd_rn(__fmul_rn(l, a.y), __fmul_rn(oml, o.y));
      m[c][2] = __fadd_rn(__fmul_rn(l, a.z), __fmul_rn(oml, o.z));
      m[c][3] = __fadd_rn(__fmul_rn(l, a.w), __fmul_rn(oml, o.w));
      if (mixed_f32) *reinterpret_cast<float4*>(mixed_f32 + ((size_t)b * CH + c) * HW + p) = make_float4(m[c][0], m[c][1], m[c][2], m[c][3]);
    }
    if (mixed_bf16) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        bf16* dst = mixed_bf16 + ((size_t)b * HW + p + j) * img_ld;
        float v[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) v[c] = c < CH ? m[c < CH ? c : 0][j] : 0.f;
        *reinterpret_cast<bf16x8*>(dst) = pack8(v);
        const float zero[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int c0 = 8; c0 < img_ld; c0 += 8) *reinterpret_cast<bf16x8*>(dst + c0) = pack8(zero);
      }
    }
  }
}

__global__ void mixup_latent_kernel(const float* __restrict__ mu, const float* __restrict__ ls, const float* __restrict__ la,
                                    const long long* __restrict__ index, const float* __restrict__ lam_dev, int B, int D, int nd,
                                    float* mixed_mu, float* mixed_sigma, float* mixed_alpha) {
  const float l = lam_dev[0], oml = lam_dev[1];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B * D) {
    const int b = i / D, d = i - b * D;
    const size_t j = (size_t)index[b] * D + d;
    mixed_mu[i] = __fadd_rn(__fmul_rn(l, mu[i]), __fmul_rn(oml, mu[j]));
    mixed_sigma[i] = __fadd_rn(__fmul_rn(l, expf(ls[i])), __fmul_rn(oml, expf(ls[j])));
  }
  if (i < B * nd) {
    const int b = i / nd, c = i - b * nd;
    const size_t j = (size_t)index[b] * nd + c;
    mixed_alpha[i] = __fadd_rn(__fmul_rn(l, expf(la[i])), __fmul_rn(oml, expf(la[j])));
  }
}

__device__ __forceinline__ unsigned long long kl_key(float v, int idx) {
  unsigned int u = __float_as_uint(v);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return ((unsigned long long)u << 32) | (unsigned int)idx;
}

__device__ __forceinline__ unsigned long long block_min_u64(unsigned long long k, unsigned long long* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, k, o);
    k = other < k ? other : k;
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = k;
  __syncthreads();
  unsigned long long r = red[0];
  for (int i = 1; i < (int)(blockDim.x >> 5); ++i) r = red[i] < r ? red[i] : r;
  return r;
}

// one block per row i; thread j strides over the candidates
__global__ void __launch_bounds__(128) pairwise_kl_kernel(const float* __restrict__ mu, const float* __restrict__ ls, int B, int D,
                                                          long long* __restrict__ index, float* __restrict__ kl_out) {
  extern __shared__ float sm[];  // mu_i[D], ls_i[D], s1sq[D]
  __shared__ unsigned long long red[4];
  float* mu_i = sm;
  float* ls_i = sm + D;
  float* s1sq = sm + 2 * D;
  const int i = blockIdx.x;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    mu_i[d] = mu[(size_t)i * D + d];
    const float l = ls[(size_t)i * D + d];
    ls_i[d] = l;
    const float s = expf(l);
    s1sq[d] = __fmul_rn(s, s);
  }
  __syncthreads();
  unsigned long long best1 = ~0ull, best2 = ~0ull;
  for (int j = threadIdx.x; j < B; j += blockDim.x) {
    // gaussian_kl_divergence_calculation (mixup.py:93-99), FP32 terms, wide accumulators
    double a = 0.0, b = 0.0, c = 0.0;
    for (int d = 0; d < D; ++d) {
      const float m2 = mu[(size_t)j * D + d], l2 = ls[(size_t)j * D + d];
      const float s2 = expf(l2);
      const float s2sq = __fmul_rn(s2, s2);
      a += (double)__fsub_rn(l2, ls_i[d]);
      b += (double)__fdiv_rn(s1sq[d], s2sq);
      const float dm = __fsub_rn(mu_i[d], m2);
      c += (double)__fdiv_rn(__fmul_rn(dm, dm), s2sq);
    }
    const float kl = (float)a + 0.5f * (float)b + 0.5f * (float)c - 0.5f * (float)D;
    if (kl_out) kl_out[(size_t)i * B + j] = kl;
    const unsigned long long k = kl_key(kl, j);
    if (k < best1) { best2 = best1; best1 = k; } else if (k < best2) { best2 = k; }
  }
  const unsigned long long m1 = block_min_u64(best1, red);
  const unsigned long long cand = (best1 == m1) ? best2 : best1;
  const unsigned long long m2 = block_min_u64(cand, red);
  if (threadIdx.x == 0) index[i] = (long long)(m2 & 0xffffffffull);
}

// ------------------------------------------------------------------------------------------- SGD
__global__ void __launch_bounds__(256) sgd_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                  const float* __restrict__ hyper, long long n) {
  const float lr = hyper[0], mom = hyper[1], wd = hyper[2], gscale = hyper[3];
  const bool first = hyper[4] != 0.f;
  const long long n4 = n >> 2;
  float4* p4 = reinterpret_cast<float4*>(p);
  float4* g4 = reinterpret_cast<float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pv = p4[i], gv = g4[i], mv = m4[i];
    float* pp = &pv.x; float* gg = &gv.x; float* mm = &mv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float d = gg[k] * gscale + wd * pp[k];
      mm[k] = first ? d : mom * mm[k] + d;
      pp[k] -= lr * mm[k];
    }
    p4[i] = pv; m4[i] = mv;
    g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (long long i = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float d = g[i] * gscale + wd * p[i];
    const float mv = first ? d : mom * m[i] + d;
    m[i] = mv;
    p[i] -= lr * mv;
    g[i] = 0.f;
  }
}

static inline int grid_for(long long n, int threads, int cap = 148 * 8) {
  long long b = (n + threads - 1) / threads;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace

extern "C" {

int sv_sample_fwd(const float* mu, const float* ls, const float* la, const float* eps, const float* unif, const int64_t* label,
                  const int64_t* label_mix, const float* lam_dev, int32_t mode, float temperature, int32_t B, int32_t D,
                  int32_t nd, float* latent, int32_t ld, void* stream) {
  SV_REQUIRE(mu && ls && eps && latent && ld >= D + nd, "sv_sample_fwd: bad arguments");
  SV_REQUIRE(mode == 2 ? (la && unif) : (label != nullptr), "sv_sample_fwd: mode %d operands missing", mode);
  SV_REQUIRE(mode != 1 || (label_mix && lam_dev), "sv_sample_fwd: mixup operands missing");
  sample_fwd_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(mu, ls, la, eps, unif, (const long long*)label,
                                                         (const long long*)label_mix, lam_dev, mode, temperature, D, nd, latent, ld);
  return sv_check_launch("sample_fwd");
}

int sv_sample_bwd(const float* g_latent, int32_t ld, const float* ls, const float* eps, const float* latent, int32_t mode,
                  float temperature, int32_t B, int32_t D, int32_t nd, float* g_mu, float* g_ls, float* g_la, int32_t accumulate,
                  void* stream) {
  SV_REQUIRE(g_latent && ls && eps && g_mu && g_ls, "sv_sample_bwd: null pointer");
  sample_bwd_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(g_latent, ld, ls, eps, latent, mode, temperature, D, nd, g_mu, g_ls, g_la,
                                                         accumulate);
  return sv_check_launch("sample_bwd");
}

int sv_elbo_rec_fwd_bwd(const float* x, const float* xhat, int32_t xhat_nhwc, int32_t B, int32_t ch, int32_t HW, int32_t bce,
                        float x_sigma, const float* g_scale, float* terms, void* g_bf16, int32_t g_ld, float* g_f32, void* stream) {
  SV_REQUIRE(x && xhat && terms && ch <= 16, "sv_elbo_rec_fwd_bwd: bad arguments");
  SV_REQUIRE(!g_bf16 || (g_ld % 8 == 0 && g_ld >= ch), "sv_elbo_rec_fwd_bwd: g_ld");
  const long long npix = (long long)B * HW;
  const float inv_b = 1.f / (float)B;
  const float mse_scale = 1.f / (2.f * x_sigma * x_sigma);
  const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(xhat) | reinterpret_cast<uintptr_t>(g_bf16) |
                         reinterpret_cast<uintptr_t>(g_f32)) & 15) == 0;
  if (HW % 4 == 0 && aligned && (ch == 3 || ch == 1)) {
    const long long nquad = npix / 4;
    const int gridv = grid_for(nquad, 256, 148 * 8);
    cudaStream_t st = (cudaStream_t)stream;
#define SV_ELBO_VEC(BB, CC) elbo_rec_vec_kernel<BB, CC><<<gridv, 256, 0, st>>>(x, xhat, xhat_nhwc, nquad, HW, inv_b, mse_scale, g_scale, terms, \
                                                                             (bf16*)g_bf16, g_ld, g_f32)
    if (bce) { if (ch == 3) SV_ELBO_VEC(true, 3); else SV_ELBO_VEC(true, 1); }
    else { if (ch == 3) SV_ELBO_VEC(false, 3); else SV_ELBO_VEC(false, 1); }
#undef SV_ELBO_VEC
    return sv_check_launch("elbo_rec");
  }
  const int grid = grid_for(npix, 256);
  if (bce)
    elbo_rec_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(x, xhat, xhat_nhwc, npix, ch, HW, inv_b, mse_scale, g_scale, terms,
                                                                  (bf16*)g_bf16, g_ld, g_f32);
  else
    elbo_rec_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(x, xhat, xhat_nhwc, npix, ch, HW, inv_b, mse_scale, g_scale, terms,
                                                                   (bf16*)g_bf16, g_ld, g_f32);
  return sv_check_launch("elbo_rec");
}

static float host_log_prior(int nd) { return logf((float)(1.0 / (double)nd)); }

int sv_elbo_kl_fwd(const float* mu, const float* ls, const float* la, int32_t B, int32_t D, int32_t nd, float* terms, void* stream) {
  SV_REQUIRE(mu && ls && la && terms, "sv_elbo_kl_fwd: null pointer");
  const int n = B * (D > nd ? D : nd);
  elbo_kl_fwd_kernel<<<grid_for(n, 256, 64), 256, 0, (cudaStream_t)stream>>>(mu, ls, la, B * D, B * nd, 1.f / (float)B,
                                                                            host_log_prior(nd), terms);
  return sv_check_launch("elbo_kl_fwd");
}

int sv_elbo_kl_bwd(const float* mu, const float* ls, const float* la, const float* terms, const float* coef, int32_t unit, int32_t B,
                   int32_t D, int32_t nd, float* g_mu, float* g_ls, float* g_la, int32_t accumulate, void* stream) {
  SV_REQUIRE(mu && ls && la && g_mu && g_ls && g_la, "sv_elbo_kl_bwd: null pointer");
  SV_REQUIRE(unit || (terms && coef), "sv_elbo_kl_bwd: terms/coef required");
  const int n = B * (D > nd ? D : nd);
  elbo_kl_bwd_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(mu, ls, la, terms, coef, unit, B * D, B * nd, 1.f / (float)B,
                                                                        host_log_prior(nd), g_mu, g_ls, g_la, accumulate);
  return sv_check_launch("elbo_kl_bwd");
}

int sv_posterior_fwd_bwd(const float* la, const float* target, const int64_t* label_a, const int64_t* label_b, const float* lam_dev,
                         const float* mu, const float* ls, const float* mu_t, const float* sig_t, const float* coef, int32_t B,
                         int32_t D, int32_t nd, float* terms, float* g_la, float* g_mu, float* g_ls, int32_t accumulate,
                         void* stream) {
  SV_REQUIRE(la && terms && (target || label_a), "sv_posterior_fwd_bwd: null pointer");
  SV_REQUIRE(!mu || (ls && mu_t && sig_t), "sv_posterior_fwd_bwd: continuous operands");
  const int n = B * (D > nd ? D : nd);
  posterior_kernel<<<grid_for(n, 256, 64), 256, 0, (cudaStream_t)stream>>>(la, target, (const long long*)label_a,
                                                                          (const long long*)label_b, lam_dev, mu, ls, mu_t, sig_t,
                                                                          coef, B, D, nd, 1.f / (float)B, terms, g_la, g_mu, g_ls,
                                                                          accumulate);
  return sv_check_launch("posterior");
}

int sv_inference_kl(const float* la, const int64_t* label, int32_t B, int32_t nd, float* out, void* stream) {
  inference_kl_kernel<<<grid_for((long long)B * nd, 256, 16), 256, 0, (cudaStream_t)stream>>>(la, (const long long*)label, B, nd,
                                                                                            1.f / (float)B, out);
  return sv_check_launch("inference_kl");
}

int sv_mixup_lerp(const float* image, const float* mu, const float* ls, const float* la, const int64_t* index, const float* lam_dev,
                  int32_t B, int32_t ch, int32_t HW, int32_t D, int32_t nd, float* mixed_f32, void* mixed_bf16, int32_t img_ld,
                  float* mixed_mu, float* mixed_sigma, float* mixed_alpha, void* stream) {
  SV_REQUIRE(index && lam_dev, "sv_mixup_lerp: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (image) {
    SV_REQUIRE(ch <= 16 && (mixed_f32 || mixed_bf16), "sv_mixup_lerp: image operands");
    SV_REQUIRE(!mixed_bf16 || img_ld % 8 == 0, "sv_mixup_lerp: img_ld");
    const long long npix = (long long)B * HW;
    const bool aligned = ((reinterpret_cast<uintptr_t>(image) | reinterpret_cast<uintptr_t>(mixed_f32) | reinterpret_cast<uintptr_t>(mixed_bf16)) & 15) == 0;
    if (HW % 4 == 0 && aligned && ch == 3)
      mixup_image_vec_kernel<3><<<grid_for(npix / 4, 256), 256, 0, st>>>(image, (const long long*)index, lam_dev, npix / 4, HW, mixed_f32,
                                                                         (bf16*)mixed_bf16, img_ld);
    else if (HW % 4 == 0 && aligned && ch == 1)
      mixup_image_vec_kernel<1><<<grid_for(npix / 4, 256), 256, 0, st>>>(image, (const long long*)index, lam_dev, npix / 4, HW, mixed_f32,
                                                                         (bf16*)mixed_bf16, img_ld);
    else
      mixup_image_kernel<<<grid_for(npix, 256), 256, 0, st>>>(image, (const long long*)index, lam_dev, npix, ch, HW, mixed_f32,
                                                              (bf16*)mixed_bf16, img_ld);
    int rc = sv_check_launch("mixup_image");
    if (rc) return rc;
  }
  if (mu) {
    SV_REQUIRE(ls && la && mixed_mu && mixed_sigma && mixed_alpha, "sv_mixup_lerp: latent operands");
    const int n = B * (D > nd ? D : nd);
    mixup_latent_kernel<<<ceil_div(n, 256), 256, 0, st>>>(mu, ls, la, (const long long*)index, lam_dev, B, D, nd, mixed_mu,
                                                          mixed_sigma, mixed_alpha);
    return sv_check_launch("mixup_latent");
  }
  return SV_OK;
}

int sv_pairwise_kl_second_nearest(const float* mu, const float* ls, int32_t B, int32_t D, int64_t* index, float* kl_out,
                                  void* stream) {
  SV_REQUIRE(mu && ls && index && B >= 2, "sv_pairwise_kl_second_nearest: bad arguments");
  pairwise_kl_kernel<<<B, 128, 3 * (size_t)D * sizeof(float), (cudaStream_t)stream>>>(mu, ls, B, D, (long long*)index, kl_out);
  return sv_check_launch("pairwise_kl");
}

int sv_sgd_step(float* param, float* grad, float* momentum_buf, const float* hyper, int64_t n, void* stream) {
  SV_REQUIRE(param && grad && momentum_buf && hyper, "sv_sgd_step: null pointer");
  SV_REQUIRE((((uintptr_t)param | (uintptr_t)grad | (uintptr_t)momentum_buf) & 15) == 0, "sv_sgd_step: arenas must be 16-byte aligned");
  sgd_kernel<<<grid_for(n / 4 + 1, 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(param, grad, momentum_buf, hyper, n);
  return sv_check_launch("sgd");
}

}  // extern "C"
