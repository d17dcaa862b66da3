// BatchNorm (train mode) + LeakyReLU/ReLU forward/backward on NHWC bf16 tensors, per pass-group
// statistics.  All kernels are HBM/L2-bandwidth bound: 16-byte vector accesses, a fixed channel
// chunk per thread so the per-channel coefficients live in registers, shuffle-free register
// accumulation with one shared-memory + one global atomic per (block, channel).
#include "common.cuh"
#include "../../include/shotvae.h"

namespace {

struct ColShape {
  int cpr;      // 16-byte chunks per row (C/8)
  int rl;       // row lanes per block
  int threads;  // cpr*rl
  int slabs;    // blocks per group
  long long slab_rows;
};

static ColShape col_shape(long long rows_per_group, int G, int C, int blocks_per_sm = 8) {
  ColShape s;
  s.cpr = C / 8;
  s.rl = 256 / s.cpr;
  if (s.rl < 1) s.rl = 1;
  s.threads = s.cpr * s.rl;
  long long want = (148ll * blocks_per_sm) / (G > 0 ? G : 1);
  if (want < 1) want = 1;
  long long by_rows = ceil_div_ll(rows_per_group, (long long)s.rl * 8);
  s.slabs = (int)(by_rows < want ? by_rows : want);
  if (s.slabs < 1) s.slabs = 1;
  s.slab_rows = ceil_div_ll(rows_per_group, s.slabs);
  return s;
}

__global__ void bn_finalize_kernel(const float* __restrict__ stats, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float count, float eps, int G, int C, int c_real,
                                   float* mean, float* var, float* scale, float* shift) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= G * C) return;
  const int g = i / C, c = i - g * C;
  float m = 0.f, v = 0.f, sc = 0.f, sh = 0.f;
  if (c < c_real) {
    const double s1 = stats[(size_t)(g * 2 + 0) * C + c], s2 = stats[(size_t)(g * 2 + 1) * C + c];
    const double dm = s1 / count;
    double dv = s2 / count - dm * dm;
    if (dv < 0.0) dv = 0.0;
    m = (float)dm;
    v = (float)dv;
    const float rstd = rsqrtf(v + eps);
    sc = gamma[c] * rstd;
    sh = beta[c] - m * sc;
  }
  mean[i] = m; var[i] = v; scale[i] = sc; shift[i] = sh;
}

template <typename TA>
__global__ void __launch_bounds__(256) bn_act_fwd_kernel(const TA* __restrict__ y, TA* __restrict__ a,
                                                         const float* __restrict__ scale, const float* __restrict__ shift,
                                                         float slope, long long rows_per_group, long long slab_rows, int C) {
  pdl_trigger();
  pdl_wait();
  const int cpr = C / 8;
  const int chunk = threadIdx.x % cpr, rl = threadIdx.x / cpr, nrl = blockDim.x / cpr;
  const int g = blockIdx.y;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = scale[(size_t)g * C + chunk * 8 + j];
    sh[j] = shift[(size_t)g * C + chunk * 8 + j];
  }
  const long long r0 = (long long)blockIdx.x * slab_rows;
  const long long r1 = min(r0 + slab_rows, rows_per_group);
  constexpr int U = 4;
  for (long long rb = r0 + rl; rb < r1; rb += (long long)nrl * U) {
    typename V8<TA>::raw q[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = rb + (long long)u * nrl;
      if (r < r1) q[u] = V8<TA>::load(y + ((size_t)g * rows_per_group + r) * C + chunk * 8);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = rb + (long long)u * nrl;
      if (r < r1) {
        float v[8];
        V8<TA>::unpack(q[u], v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float t = fmaf(v[j], sc[j], sh[j]);
          v[j] = t > 0.f ? t : slope * t;
        }
        V8<TA>::store(a + ((size_t)g * rows_per_group + r) * C + chunk * 8, v);
      }
    }
  }
}

// BatchNorm finalize fused into the apply: the block derives scale / shift of its C channels ONCE from the raw (sum, sum^2)
// statistics (FP64 mean / variance, two divisions per channel) into shared memory -- round 2 had every THREAD do that for its
// 8 channels: 16 double divisions x 256 threads x ~600 blocks per launch, about 4 us of FP64 pipe time in an 11 us kernel --
// while the first rows of y are already in flight; block (0, g) also publishes mean / var / scale / shift for the backward pass
// and the running-statistics update.  Saves one launch per BatchNorm.
template <typename TA>
__global__ void __launch_bounds__(256) bn_finalize_act_fwd_kernel(const TA* __restrict__ y, TA* __restrict__ a,
                                                                  const float* __restrict__ stats, const float* __restrict__ gamma,
                                                                  const float* __restrict__ beta, float count, float eps, float slope,
                                                                  long long rows_per_group, long long slab_rows, int C, float* mean,
                                                                  float* var, float* scale, float* shift) {
  __shared__ float s_sc[2048], s_sh[2048];      // C <= 2048 (C / 8 <= 256 chunks per row)
  pdl_trigger();
  pdl_wait();
  const int cpr = C / 8;
  const int chunk = threadIdx.x % cpr, rl = threadIdx.x / cpr, nrl = blockDim.x / cpr;
  const int g = blockIdx.y;
  const long long r0 = (long long)blockIdx.x * slab_rows;
  const long long r1 = min(r0 + slab_rows, rows_per_group);
  constexpr int U = 4;
  typename V8<TA>::raw q[U];
  // first rows of this thread: in flight while the coefficients are computed
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const long long r = r0 + rl + (long long)u * nrl;
    if (r < r1) q[u] = V8<TA>::load(y + ((size_t)g * rows_per_group + r) * C + chunk * 8);
  }
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const double s1 = stats[(size_t)(g * 2 + 0) * C + c], s2 = stats[(size_t)(g * 2 + 1) * C + c];
    const double dm = s1 / count;
    double dv = s2 / count - dm * dm;
    if (dv < 0.0) dv = 0.0;
    const float m = (float)dm, v = (float)dv;
    const float sc = gamma[c] * rsqrtf(v + eps);
    const float sh = beta[c] - m * sc;
    s_sc[c] = sc;
    s_sh[c] = sh;
    if (blockIdx.x == 0) {
      const size_t k = (size_t)g * C + c;
      mean[k] = m; var[k] = v; scale[k] = sc; shift[k] = sh;
    }
  }
  __syncthreads();
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = s_sc[chunk * 8 + j];
    sh[j] = s_sh[chunk * 8 + j];
  }
  for (long long rb = r0 + rl; rb < r1; rb += (long long)nrl * U) {
    typename V8<TA>::raw cur[U];
#pragma unroll
    for (int u = 0; u < U; ++u) cur[u] = q[u];
    // next batch of rows in flight while this one is transformed and stored
    const long long rn = rb + (long long)nrl * U;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = rn + (long long)u * nrl;
      if (r < r1) q[u] = V8<TA>::load(y + ((size_t)g * rows_per_group + r) * C + chunk * 8);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = rb + (long long)u * nrl;
      if (r < r1) {
        float v[8];
        V8<TA>::unpack(cur[u], v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float t = fmaf(v[j], sc[j], sh[j]);
          v[j] = t > 0.f ? t : slope * t;
        }
        V8<TA>::store(a + ((size_t)g * rows_per_group + r) * C + chunk * 8, v);
      }
    }
  }
}

struct RunDesc {
  const float* mean[4];
  const float* var[4];
  float* running_mean;
  float* running_var;
  long long* nbt;
  float count;
  int npass, C;
};

// every BatchNorm's running-statistics update of the step in one launch (blockIdx.y = BatchNorm)
__global__ void bn_running_update_batched_kernel(const RunDesc* __restrict__ table, float momentum) {
  const RunDesc d = table[blockIdx.y];
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && d.nbt != nullptr) *d.nbt += d.npass;
  if (c >= d.C) return;
  float rm = d.running_mean[c], rv = d.running_var[c];
  const float unbias = d.count > 1.f ? d.count / (d.count - 1.f) : 1.f;
  for (int p = 0; p < d.npass; ++p) {
    rm = (1.f - momentum) * rm + momentum * d.mean[p][c];
    rv = (1.f - momentum) * rv + momentum * (d.var[p][c] * unbias);
  }
  d.running_mean[c] = rm;
  d.running_var[c] = rv;
}

template <typename TA>
__global__ void __launch_bounds__(256) bn_act_gap_kernel(const TA* __restrict__ y, float* __restrict__ feat,
                                                         const float* __restrict__ scale, const float* __restrict__ shift, float slope,
                                                         int NB, int HW, int C, int group_images) {
  // one CTA per image: threads = (8-channel chunk) x (pixel slice); the slices are folded through shared memory
  // (the one-thread-per-(image, chunk) version walked the HW pixels serially on 32 CTAs: 44 us for a 4 MB tensor)
  extern __shared__ float s_gap[];     // [slices][C]
  pdl_trigger();
  pdl_wait();
  const int cpr = C / 8;
  const int slices = blockDim.x / cpr;
  const int nb = blockIdx.x;
  const int chunk = threadIdx.x % cpr, slice = threadIdx.x / cpr;
  const int g = nb / group_images;
  if (slice < slices) {
    float sc[8], sh[8], acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sc[j] = scale[(size_t)g * C + chunk * 8 + j];
      sh[j] = shift[(size_t)g * C + chunk * 8 + j];
      acc[j] = 0.f;
    }
    const TA* base = y + (size_t)nb * HW * C + chunk * 8;
#pragma unroll 4
    for (int p = slice; p < HW; p += slices) {
      float v[8];
      V8<TA>::unpack(V8<TA>::load(base + (size_t)p * C), v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float t = fmaf(v[j], sc[j], sh[j]);
        acc[j] += t > 0.f ? t : slope * t;
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) s_gap[slice * C + chunk * 8 + j] = acc[j];
  }
  __syncthreads();
  const float inv = 1.f / (float)HW;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < slices; ++k) s += s_gap[k * C + c];
    feat[(size_t)nb * C + c] = s * inv;
  }
}

// dgamma/dbeta reduction
template <bool FEAT, typename TA>
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const TA* __restrict__ g_a, const float* __restrict__ g_feat,
                                                            const TA* __restrict__ y, const float* __restrict__ scale,
                                                            const float* __restrict__ shift, const float* __restrict__ mean,
                                                            const float* __restrict__ var, float eps, float slope,
                                                            long long rows_per_group, long long slab_rows, int HW, int C,
                                                            float* dgamma, float* dbeta) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float s_acc[];  // [2][C]
  const int cpr = C / 8;
  const int chunk = threadIdx.x % cpr, rl = threadIdx.x / cpr, nrl = blockDim.x / cpr;
  const int g = blockIdx.y;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) s_acc[i] = 0.f;
  float sc[8], sh[8], mu[8], rs[8], ab[8], ag[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const size_t k = (size_t)g * C + chunk * 8 + j;
    sc[j] = scale[k]; sh[j] = shift[k]; mu[j] = mean[k]; rs[j] = rsqrtf(var[k] + eps);
    ab[j] = 0.f; ag[j] = 0.f;
  }
  const float inv_hw = 1.f / (float)HW;
  const long long r0 = (long long)blockIdx.x * slab_rows;
  const long long r1 = min(r0 + slab_rows, rows_per_group);
  constexpr int U = 4;      // rows in flight per thread (memory-level parallelism)
  for (long long rb = r0 + rl; rb < r1; rb += (long long)nrl * U) {
    typename V8<TA>::raw gq[U], yq[U];
    float gf[U][8];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = rb + (long long)u * nrl;
      if (r < r1) {
        const size_t row = (size_t)g * rows_per_group + r;
        const size_t off = row * C + chunk * 8;
        if (FEAT) {
          const size_t nb = row / HW;
#pragma unroll
          for (int j = 0; j < 8; ++j) gf[u][j] = g_feat[nb * C + chunk * 8 + j] * inv_hw;
        } else {
          gq[u] = V8<TA>::load(g_a + off);
        }
        yq[u] = V8<TA>::load(y + off);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = rb + (long long)u * nrl;
      if (r < r1) {
        float gv[8], yv[8];
        if (FEAT) {
#pragma unroll
          for (int j = 0; j < 8; ++j) gv[j] = gf[u][j];
        } else {
          V8<TA>::unpack(gq[u], gv);
        }
        V8<TA>::unpack(yq[u], yv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float pre = fmaf(yv[j], sc[j], sh[j]);
          const float gp = pre > 0.f ? gv[j] : slope * gv[j];
          ab[j] += gp;
          ag[j] += gp * (yv[j] - mu[j]) * rs[j];
        }
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    atomicAdd(&s_acc[chunk * 8 + j], ag[j]);
    atomicAdd(&s_acc[C + chunk * 8 + j], ab[j]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    atomicAdd(&dgamma[(size_t)g * C + i], s_acc[i]);
    atomicAdd(&dbeta[(size_t)g * C + i], s_acc[C + i]);
  }
}

struct BwdTerms {
  sv_bn_bwd_term t[2];
  int n;
};

// NT = number of BatchNorm branches fed by y (2 only at a residual unit with a projection shortcut).  The
// single-branch instantiation keeps half the per-channel coefficients, fits 2 CTAs per SM and streams faster.
template <int NT, typename TA>
__global__ void __launch_bounds__(256, (NT == 1 && sizeof(TA) == 2) ? 2 : 1) bn_bwd_apply_kernel(const BwdTerms T, const TA* __restrict__ y,
                                                           const TA* __restrict__ addend, TA* __restrict__ g_y, float eps,
                                                           long long rows_per_group, long long slab_rows, int HW, int G, int C) {
  pdl_trigger();
  pdl_wait();
  const int cpr = C / 8;
  const int chunk = threadIdx.x % cpr, rl = threadIdx.x / cpr, nrl = blockDim.x / cpr;
  const int g = blockIdx.y;
  const float invM = 1.f / (float)rows_per_group;
  float sc[NT][8], sh[NT][8], k1[NT][8], k0[NT][8];
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    if (t < T.n) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const size_t k = (size_t)g * C + chunk * 8 + j;
        const float s = T.t[t].scale[k], rstd = rsqrtf(T.t[t].var[k] + eps);
        sc[t][j] = s;
        sh[t][j] = T.t[t].shift[k];
        k1[t][j] = -s * rstd * T.t[t].dgamma[k] * invM;
        k0[t][j] = -s * T.t[t].dbeta[k] * invM - k1[t][j] * T.t[t].mean[k];
      }
      // parameter gradients: one block adds sum_g dgamma / dbeta
      if (blockIdx.x == 0 && blockIdx.y == 0 && rl == 0 && T.t[t].grad_gamma != nullptr) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = chunk * 8 + j;
          if (c < T.t[t].c_real) {
            float sg = 0.f, sb = 0.f;
            for (int gg = 0; gg < G; ++gg) {
              sg += T.t[t].dgamma[(size_t)gg * C + c];
              sb += T.t[t].dbeta[(size_t)gg * C + c];
            }
            T.t[t].grad_gamma[c] += sg;
            T.t[t].grad_beta[c] += sb;
          }
        }
      }
    }
  }
  const float inv_hw = 1.f / (float)HW;
  const long long r0 = (long long)blockIdx.x * slab_rows;
  const long long r1 = min(r0 + slab_rows, rows_per_group);
  constexpr int U = 4;      // rows in flight per thread
  // (MEASURED, not kept: requesting the first rows before the per-channel coefficients are fetched -- C4 2.06 -> 2.83 ms of
  // bn_bwd_apply per step: the 12 extra live vectors cost more than the ~1 us of exposed latency)
  for (long long rb = r0 + rl; rb < r1; rb += (long long)nrl * U) {
    typename V8<TA>::raw yq[U], aq[U], gq[NT][U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = rb + (long long)u * nrl;
      if (r < r1) {
        const size_t off = ((size_t)g * rows_per_group + r) * C + chunk * 8;
        yq[u] = V8<TA>::load(y + off);
        if (addend != nullptr) aq[u] = V8<TA>::load(addend + off);
#pragma unroll
        for (int t = 0; t < NT; ++t)
          if (t < T.n && T.t[t].g_feat == nullptr) gq[t][u] = V8<TA>::load((const TA*)T.t[t].g_a + off);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = rb + (long long)u * nrl;
      if (r < r1) {
        const size_t row = (size_t)g * rows_per_group + r;
        const size_t off = row * C + chunk * 8;
        float yv[8], o[8];
        V8<TA>::unpack(yq[u], yv);
        if (addend != nullptr) {
          V8<TA>::unpack(aq[u], o);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = 0.f;
        }
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          if (t < T.n) {
            float gv[8];
            if (T.t[t].g_feat != nullptr) {
              const size_t nb = row / HW;
#pragma unroll
              for (int j = 0; j < 8; ++j) gv[j] = T.t[t].g_feat[nb * C + chunk * 8 + j] * inv_hw;
            } else {
              V8<TA>::unpack(gq[t][u], gv);
            }
            const float slope = T.t[t].slope;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float pre = fmaf(yv[j], sc[t][j], sh[t][j]);
              const float gp = pre > 0.f ? gv[j] : slope * gv[j];
              o[j] += sc[t][j] * gp + k1[t][j] * yv[j] + k0[t][j];
            }
          }
        }
        V8<TA>::store(g_y + off, o);
      }
    }
  }
}

struct PassPtrs {
  const float* mean[8];
  const float* var[8];
};

__global__ void bn_running_update_kernel(PassPtrs P, int npass, float count, float momentum, int c_real,
                                         float* running_mean, float* running_var, long long* nbt) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && nbt != nullptr) *nbt += npass;
  if (c >= c_real) return;
  float rm = running_mean[c], rv = running_var[c];
  const float unbias = count > 1.f ? count / (count - 1.f) : 1.f;
  for (int p = 0; p < npass; ++p) {
    rm = (1.f - momentum) * rm + momentum * P.mean[p][c];
    rv = (1.f - momentum) * rv + momentum * (P.var[p][c] * unbias);
  }
  running_mean[c] = rm;
  running_var[c] = rv;
}

template <typename TA>
__global__ void __launch_bounds__(256) colsum_kernel(const TA* __restrict__ x, float* __restrict__ out, long long rows,
                                                     long long slab_rows, int C, int c_real) {
  extern __shared__ float s_acc[];  // [C]
  const int cpr = C / 8;
  const int chunk = threadIdx.x % cpr, rl = threadIdx.x / cpr, nrl = blockDim.x / cpr;
  for (int i = threadIdx.x; i < C; i += blockDim.x) s_acc[i] = 0.f;
  float a[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = 0.f;
  const long long r0 = (long long)blockIdx.x * slab_rows;
  const long long r1 = min(r0 + slab_rows, rows);
  for (long long r = r0 + rl; r < r1; r += nrl) {
    float v[8];
    V8<TA>::unpack(V8<TA>::load(x + (size_t)r * C + chunk * 8), v);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] += v[j];
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 8; ++j) atomicAdd(&s_acc[chunk * 8 + j], a[j]);
  __syncthreads();
  for (int i = threadIdx.x; i < c_real; i += blockDim.x) atomicAdd(&out[i], s_acc[i]);
}

template <typename TA>
__global__ void pack_image_kernel(const float* __restrict__ src, TA* __restrict__ dst, long long npix, int c_real, int HW,
                                  int C) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= npix) return;
  const long long nb = i / HW, p = i - nb * HW;
  for (int c0 = 0; c0 < C; c0 += 8) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c0 + j;
      v[j] = c < c_real ? src[((size_t)nb * c_real + c) * HW + p] : 0.f;
    }
    V8<TA>::store(dst + (size_t)i * C + c0, v);
  }
}

// fp32 [rows][c_real] -> fp32 [rows][C], zero padded channels (FP32 mode: the ELBO gradient enters the decoder backward)
__global__ void pad_channels_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, long long total, int c_real, int C) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long r = i / C;
  const int c = (int)(i - r * C);
  dst[i] = c < c_real ? src[r * c_real + c] : 0.f;
}

__global__ void nhwc_to_nchw_kernel(const float* __restrict__ src, float* __restrict__ dst, long long total, int c_real,
                                    int HW) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long p = i % HW;
  const long long r = i / HW;
  const int c = (int)(r % c_real);
  const long long nb = r / c_real;
  dst[i] = src[((size_t)nb * HW + p) * c_real + c];
}


// ---- host launchers, templated on the activation element type (bf16: production; float: parity-grade FP32 mode) ----
template <typename TA>
int bn_act_fwd_t(const void* y, void* a, const float* scale, const float* shift, float slope, int64_t rows_per_group, int32_t G,
                 int32_t C, void* stream) {
  SV_REQUIRE(C % 8 == 0 && C / 8 <= 256, "sv_bn_act_fwd: unsupported C=%d", C);
  const ColShape s = col_shape(rows_per_group, G, C, 4);
  sv_launch_pdl(bn_act_fwd_kernel<TA>, dim3(s.slabs, G), dim3(s.threads), 0, (cudaStream_t)stream, (const TA*)y, (TA*)a, scale, shift, slope,
                (long long)rows_per_group, s.slab_rows, C);
  return sv_check_launch("bn_act_fwd");
}

template <typename TA>
int bn_finalize_act_fwd_t(const void* y, void* a, const float* stats, const float* gamma, const float* beta, float count, float eps,
                          float slope, int64_t rows_per_group, int32_t G, int32_t C, float* mean, float* var, float* scale, float* shift,
                          void* stream) {
  SV_REQUIRE(C % 8 == 0 && C / 8 <= 256, "sv_bn_finalize_act_fwd: unsupported C=%d", C);
  SV_REQUIRE(y && a && stats && gamma && beta && mean && var && scale && shift, "sv_bn_finalize_act_fwd: null pointer");
  const ColShape s = col_shape(rows_per_group, G, C, 4);
  sv_launch_pdl(bn_finalize_act_fwd_kernel<TA>, dim3(s.slabs, G), dim3(s.threads), 0, (cudaStream_t)stream, (const TA*)y, (TA*)a, stats, gamma,
                beta, count, eps, slope, (long long)rows_per_group, s.slab_rows, C, mean, var, scale, shift);
  return sv_check_launch("bn_finalize_act_fwd");
}

template <typename TA>
int bn_act_gap_fwd_t(const void* y, float* feat, const float* scale, const float* shift, float slope, int32_t NB, int32_t HW, int32_t C,
                     int32_t group_images, void* stream) {
  SV_REQUIRE(C % 8 == 0 && C / 8 <= 256, "sv_bn_act_gap_fwd: unsupported C=%d", C);
  const int cpr = C / 8, slices = 256 / cpr;
  const size_t smem = (size_t)slices * C * sizeof(float);
  sv_launch_pdl(bn_act_gap_kernel<TA>, dim3(NB), dim3(256), smem, (cudaStream_t)stream, (const TA*)y, feat, scale, shift, slope, NB, HW, C,
                group_images);
  return sv_check_launch("bn_act_gap");
}

template <typename TA>
int bn_bwd_reduce_t(const void* g_a, const float* g_feat, const void* y, const float* scale, const float* shift, const float* mean,
                    const float* var, float eps, float slope, int64_t rows_per_group, int32_t HW, int32_t G, int32_t C, float* dgamma,
                    float* dbeta, void* stream) {
  SV_REQUIRE(C % 8 == 0 && C / 8 <= 256, "sv_bn_bwd_reduce: unsupported C=%d", C);
  SV_REQUIRE((g_a != nullptr) != (g_feat != nullptr), "sv_bn_bwd_reduce: exactly one of g_a / g_feat");
  const ColShape s = col_shape(rows_per_group, G, C, 2);   // few blocks: each ends with 2*C global atomics
  const size_t smem = 2 * (size_t)C * sizeof(float);
  if (g_feat)
    sv_launch_pdl(bn_bwd_reduce_kernel<true, TA>, dim3(s.slabs, G), dim3(s.threads), smem, (cudaStream_t)stream, (const TA*)nullptr, g_feat,
                  (const TA*)y, scale, shift, mean, var, eps, slope, (long long)rows_per_group, s.slab_rows, HW, C, dgamma, dbeta);
  else
    sv_launch_pdl(bn_bwd_reduce_kernel<false, TA>, dim3(s.slabs, G), dim3(s.threads), smem, (cudaStream_t)stream, (const TA*)g_a,
                  (const float*)nullptr, (const TA*)y, scale, shift, mean, var, eps, slope, (long long)rows_per_group, s.slab_rows, HW, C,
                  dgamma, dbeta);
  return sv_check_launch("bn_bwd_reduce");
}

template <typename TA>
int bn_bwd_apply_t(const sv_bn_bwd_term* terms, int32_t nterms, const void* y, const void* addend, void* g_y, float eps,
                   int64_t rows_per_group, int32_t HW, int32_t G, int32_t C, void* stream) {
  SV_REQUIRE(nterms >= 1 && nterms <= 2, "sv_bn_bwd_apply: nterms must be 1 or 2");
  SV_REQUIRE(C % 8 == 0 && C / 8 <= 256, "sv_bn_bwd_apply: unsupported C=%d", C);
  BwdTerms T;
  memset(&T, 0, sizeof(T));
  T.n = nterms;
  for (int i = 0; i < nterms; ++i) T.t[i] = terms[i];
  const ColShape s = col_shape(rows_per_group, G, C, 4);
  if (nterms == 1)
    sv_launch_pdl(bn_bwd_apply_kernel<1, TA>, dim3(s.slabs, G), dim3(s.threads), 0, (cudaStream_t)stream, T, (const TA*)y, (const TA*)addend,
                  (TA*)g_y, eps, (long long)rows_per_group, s.slab_rows, HW, G, C);
  else
    sv_launch_pdl(bn_bwd_apply_kernel<2, TA>, dim3(s.slabs, G), dim3(s.threads), 0, (cudaStream_t)stream, T, (const TA*)y, (const TA*)addend,
                  (TA*)g_y, eps, (long long)rows_per_group, s.slab_rows, HW, G, C);
  return sv_check_launch("bn_bwd_apply");
}

template <typename TA>
int colsum_t(const void* x, float* out, int64_t rows, int32_t C, int32_t c_real, void* stream) {
  SV_REQUIRE(C % 8 == 0 && C / 8 <= 256, "sv_colsum: unsupported C=%d", C);
  const ColShape s = col_shape(rows, 1, C);
  colsum_kernel<TA><<<s.slabs, s.threads, (size_t)C * sizeof(float), (cudaStream_t)stream>>>((const TA*)x, out, (long long)rows, s.slab_rows, C,
                                                                                             c_real);
  return sv_check_launch("colsum");
}

template <typename TA>
int pack_image_t(const float* src, void* dst, int32_t NB, int32_t c_real, int32_t HW, int32_t C, void* stream) {
  SV_REQUIRE(C % 8 == 0, "sv_pack_image: C %% 8");
  const long long npix = (long long)NB * HW;
  pack_image_kernel<TA><<<(int)ceil_div_ll(npix, 256), 256, 0, (cudaStream_t)stream>>>(src, (TA*)dst, npix, c_real, HW, C);
  return sv_check_launch("pack_image");
}

}  // namespace

extern "C" {

int sv_bn_finalize(const float* stats, const float* gamma, const float* beta, float count, float eps, int32_t G, int32_t C,
                   int32_t c_real, float* mean, float* var, float* scale, float* shift, void* stream) {
  SV_REQUIRE(stats && gamma && beta && mean && var && scale && shift, "sv_bn_finalize: null pointer");
  const int n = G * C;
  bn_finalize_kernel<<<ceil_div(n, 128), 128, 0, (cudaStream_t)stream>>>(stats, gamma, beta, count, eps, G, C, c_real, mean,
                                                                        var, scale, shift);
  return sv_check_launch("bn_finalize");
}

// every activation-tensor entry point exists twice: bf16 tensors (production) and fp32 tensors (`_f32`, the parity-grade mode)
#define SV_BN_PAIR(name, impl, params, args)                         \
  int name params { return impl<bf16> args; }                       \
  int name##_f32 params { return impl<float> args; }

SV_BN_PAIR(sv_bn_act_fwd, bn_act_fwd_t,
           (const void* y, void* a, const float* scale, const float* shift, float slope, int64_t rows_per_group, int32_t G, int32_t C,
            void* stream),
           (y, a, scale, shift, slope, rows_per_group, G, C, stream))
SV_BN_PAIR(sv_bn_finalize_act_fwd, bn_finalize_act_fwd_t,
           (const void* y, void* a, const float* stats, const float* gamma, const float* beta, float count, float eps, float slope,
            int64_t rows_per_group, int32_t G, int32_t C, float* mean, float* var, float* scale, float* shift, void* stream),
           (y, a, stats, gamma, beta, count, eps, slope, rows_per_group, G, C, mean, var, scale, shift, stream))
SV_BN_PAIR(sv_bn_act_gap_fwd, bn_act_gap_fwd_t,
           (const void* y, float* feat, const float* scale, const float* shift, float slope, int32_t NB, int32_t HW, int32_t C,
            int32_t group_images, void* stream),
           (y, feat, scale, shift, slope, NB, HW, C, group_images, stream))
SV_BN_PAIR(sv_bn_bwd_reduce, bn_bwd_reduce_t,
           (const void* g_a, const float* g_feat, const void* y, const float* scale, const float* shift, const float* mean,
            const float* var, float eps, float slope, int64_t rows_per_group, int32_t HW, int32_t G, int32_t C, float* dgamma,
            float* dbeta, void* stream),
           (g_a, g_feat, y, scale, shift, mean, var, eps, slope, rows_per_group, HW, G, C, dgamma, dbeta, stream))
SV_BN_PAIR(sv_bn_bwd_apply, bn_bwd_apply_t,
           (const sv_bn_bwd_term* terms, int32_t nterms, const void* y, const void* addend, void* g_y, float eps, int64_t rows_per_group,
            int32_t HW, int32_t G, int32_t C, void* stream),
           (terms, nterms, y, addend, g_y, eps, rows_per_group, HW, G, C, stream))
SV_BN_PAIR(sv_pack_image, pack_image_t,
           (const float* src, void* dst, int32_t NB, int32_t c_real, int32_t HW, int32_t C, void* stream),
           (src, dst, NB, c_real, HW, C, stream))
#undef SV_BN_PAIR

int sv_colsum_bf16(const void* x, float* out, int64_t rows, int32_t C, int32_t c_real, void* stream) {
  return colsum_t<bf16>(x, out, rows, C, c_real, stream);
}
int sv_colsum_f32(const void* x, float* out, int64_t rows, int32_t C, int32_t c_real, void* stream) {
  return colsum_t<float>(x, out, rows, C, c_real, stream);
}

int sv_sizeof_run_desc(void) { return (int)sizeof(RunDesc); }

int sv_bn_running_update_batched(const void* table_dev, int32_t n_bn, int32_t max_c, float momentum, void* stream) {
  SV_REQUIRE(table_dev && n_bn > 0 && max_c > 0, "sv_bn_running_update_batched: bad arguments");
  bn_running_update_batched_kernel<<<dim3(ceil_div(max_c, 128), n_bn), 128, 0, (cudaStream_t)stream>>>((const RunDesc*)table_dev, momentum);
  return sv_check_launch("bn_running_update_batched");
}

int sv_bn_running_update(const float* const* mean_ptrs, const float* const* var_ptrs, int32_t npass, float count,
                         float momentum, int32_t c_real, float* running_mean, float* running_var,
                         int64_t* num_batches_tracked, void* stream) {
  SV_REQUIRE(npass >= 1 && npass <= 8, "sv_bn_running_update: npass out of range");
  PassPtrs P;
  for (int i = 0; i < npass; ++i) { P.mean[i] = mean_ptrs[i]; P.var[i] = var_ptrs[i]; }
  bn_running_update_kernel<<<ceil_div(c_real, 128), 128, 0, (cudaStream_t)stream>>>(
      P, npass, count, momentum, c_real, running_mean, running_var, (long long*)num_batches_tracked);
  return sv_check_launch("bn_running_update");
}

int sv_pad_channels_f32(const float* src, float* dst, int64_t rows, int32_t c_real, int32_t C, void* stream) {
  SV_REQUIRE(src && dst && c_real <= C, "sv_pad_channels_f32: bad arguments");
  const long long total = (long long)rows * C;
  pad_channels_f32_kernel<<<(int)ceil_div_ll(total, 256), 256, 0, (cudaStream_t)stream>>>(src, dst, total, c_real, C);
  return sv_check_launch("pad_channels_f32");
}

int sv_nhwc_to_nchw_f32(const float* src, float* dst, int32_t NB, int32_t c_real, int32_t HW, void* stream) {
  const long long total = (long long)NB * c_real * HW;
  nhwc_to_nchw_kernel<<<(int)ceil_div_ll(total, 256), 256, 0, (cudaStream_t)stream>>>(src, dst, total, c_real, HW);
  return sv_check_launch("nhwc_to_nchw");
}

}  // extern "C"
