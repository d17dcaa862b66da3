// Small FP32 linears: the three inference heads and the k=1 ConvTranspose2d decoder stem.
// These are tiny (<= 20 MFLOP) and feed the bit-exact mixup pairing, so they stay in FP32 on the
// CUDA cores ("keep FP32").  32x32 output tile, 256 threads, 4 outputs each.
#include "common.cuh"
#include "../../include/shotvae.h"

namespace {

template <typename T>
__device__ __forceinline__ float ldf(const T* p, size_t i);
template <>
__device__ __forceinline__ float ldf<float>(const float* p, size_t i) { return p[i]; }
template <>
__device__ __forceinline__ float ldf<bf16>(const bf16* p, size_t i) { return __bfloat162float(p[i]); }

// ---- tile bodies -----------------------------------------------------------------------------------------------------------
// acc[i] += sum_{k in [kbeg, kend)} x[b0 + ty + 8 i][k] * W(n0 + tx, k)        (32 x 32 output tile, 256 threads)
template <typename TX>
__device__ __forceinline__ void linear_tile_accum(float (&acc)[4], float (*xs)[33], float (*ws)[33], const TX* __restrict__ x, int ldx,
                                                  const float* __restrict__ W, int ldw, int w_kn, int B, int N, int K, int b0, int n0,
                                                  int kbeg, int kend) {
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int k0 = kbeg; k0 < kend; k0 += 32) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = ty + 8 * i;
      const int b = b0 + r, k = k0 + tx;
      xs[r][tx] = (b < B && k < kend) ? ldf<TX>(x, (size_t)b * ldx + k) : 0.f;
      // ws[n_local][k_local]
      if (w_kn) {
        const int kk = k0 + r, n = n0 + tx;  // coalesced over n
        ws[tx][r] = (kk < kend && n < N) ? W[(size_t)kk * ldw + n] : 0.f;
      } else {
        const int n = n0 + r;
        ws[r][tx] = (n < N && k < kend) ? W[(size_t)n * ldw + k] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < 32; ++kk) {
      const float wv = ws[tx][kk];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = fmaf(xs[ty + 8 * i][kk], wv, acc[i]);
    }
    __syncthreads();
  }
}

// out[b][n] (+)= sum_k x[b][k] * W(n,k) + bias[n].  gridDim.z > 1 splits the K range: every slice adds its partial sum
// with an atomic (the caller has zeroed / prepared out_f32; bias from slice 0; no bf16 output / statistics then).
template <typename TX>
__global__ void __launch_bounds__(256) linear_fwd_kernel(const TX* __restrict__ x, int ldx, const float* __restrict__ W, int ldw,
                                                         int w_kn, const float* __restrict__ bias, float* out_f32, bf16* out_bf16,
                                                         int ldo, float* stats, int group_rows, int accumulate, int B, int N,
                                                         int K) {
  __shared__ float xs[32][33];
  __shared__ float ws[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int b0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  const bool split = gridDim.z > 1;
  int kbeg = 0, kend = K;
  if (split) {
    const int per = ((K + gridDim.z - 1) / gridDim.z + 31) & ~31;
    kbeg = blockIdx.z * per;
    kend = min(K, kbeg + per);
  }
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  linear_tile_accum<TX>(acc, xs, ws, x, ldx, W, ldw, w_kn, B, N, K, b0, n0, kbeg, kend);
  const int n = n0 + tx;
  if (n >= N) return;
  const float bv = (bias && blockIdx.z == 0) ? bias[n] : 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int b = b0 + ty + 8 * i;
    if (b >= B) continue;
    float v = acc[i] + bv;
    const size_t o = (size_t)b * ldo + n;
    if (split) {
      atomicAdd(&out_f32[o], v);
      continue;
    }
    if (out_f32) {
      if (accumulate) v += out_f32[o];
      out_f32[o] = v;
    }
    if (out_bf16) {
      const bf16 h = __float2bfloat16(v);
      out_bf16[o] = h;
      v = __bfloat162float(h);
    }
    if (stats) {
      const int g = b / group_rows;
      atomicAdd(&stats[(size_t)(g * 2 + 0) * ldo + n], v);
      atomicAdd(&stats[(size_t)(g * 2 + 1) * ldo + n], v * v);
    }
  }
}

// ---- the three inference heads in one launch each way (sv_heads_*): same tile bodies, head = a range of blockIdx ------------
struct HeadSet {
  const float* W[SV_MAX_HEADS];      // [N_h][K] row-major
  const float* bias[SV_MAX_HEADS];
  float* out[SV_MAX_HEADS];          // forward outputs [B][N_h]
  const float* g[SV_MAX_HEADS];      // output gradients [B][N_h]
  float* dW[SV_MAX_HEADS];
  float* dbias[SV_MAX_HEADS];
  int N[SV_MAX_HEADS];
  int tile0[SV_MAX_HEADS + 1];       // first 32-column tile of every head
  int n;
};

__global__ void __launch_bounds__(256) heads_fwd_kernel(const float* __restrict__ x, int ldx, const __grid_constant__ HeadSet hs, int B,
                                                        int K) {
  __shared__ float xs[32][33];
  __shared__ float ws[32][33];
  int h = 0;
  while (h + 1 < hs.n && (int)blockIdx.x >= hs.tile0[h + 1]) ++h;
  const int N = hs.N[h], n0 = (blockIdx.x - hs.tile0[h]) * 32, b0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  linear_tile_accum<float>(acc, xs, ws, x, ldx, hs.W[h], K, 0, B, N, K, b0, n0, 0, K);
  const int n = n0 + tx;
  if (n >= N) return;
  const float bv = hs.bias[h] ? hs.bias[h][n] : 0.f;
  float* out = hs.out[h];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int b = b0 + ty + 8 * i;
    if (b < B) out[(size_t)b * N + n] = acc[i] + bv;
  }
}

// gx[b][k] = sum_h sum_n g_h[b][n] * W_h[n][k]
__global__ void __launch_bounds__(256) heads_bwd_input_kernel(const __grid_constant__ HeadSet hs, float* __restrict__ gx, int ldgx, int B,
                                                              int K) {
  __shared__ float xs[32][33];
  __shared__ float ws[32][33];
  const int k0 = blockIdx.x * 32, b0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int h = 0; h < hs.n; ++h)     // roles of N and K exchanged: the reduction runs over the head's outputs
    linear_tile_accum<float>(acc, xs, ws, hs.g[h], hs.N[h], hs.W[h], K, 1, B, K, hs.N[h], b0, k0, 0, hs.N[h]);
  const int k = k0 + tx;
  if (k >= K) return;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int b = b0 + ty + 8 * i;
    if (b < B) gx[(size_t)b * ldgx + k] = acc[i];
  }
}

// acc[i] += sum_{b in [bbeg, bend)} g[b][n0 + ty + 8 i] * x[b][k0 + tx];  bacc += sum_b g[b][n0 + tx] (warp ty == 0)
template <typename TG>
__device__ __forceinline__ void wgrad_tile_accum(float (&acc)[4], float& bacc, float (*gs)[33], float (*xs)[33], const TG* __restrict__ g,
                                                 int ldg, const float* __restrict__ x, int ldx, int N, int K, int n0, int k0, int bbeg,
                                                 int bend, bool want_bias) {
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int b0 = bbeg; b0 < bend; b0 += 32) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = ty + 8 * i, b = b0 + r;
      gs[r][tx] = (b < bend && n0 + tx < N) ? ldf<TG>(g, (size_t)b * ldg + n0 + tx) : 0.f;
      xs[r][tx] = (b < bend && k0 + tx < K) ? x[(size_t)b * ldx + k0 + tx] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int bb = 0; bb < 32; ++bb) {
      const float xv = xs[bb][tx];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = fmaf(gs[bb][ty + 8 * i], xv, acc[i]);
    }
    if (want_bias && ty == 0) {
      for (int bb = 0; bb < 32; ++bb) bacc += gs[bb][tx];
    }
    __syncthreads();
  }
}

// adds the tile to dW: coalesced in both weight layouts (w_kn: through a shared-memory transpose), atomics when the batch
// range is split over gridDim.z
__device__ __forceinline__ void wgrad_tile_store(const float (&acc)[4], float bacc, float (*tr)[33], float* dW, int ldw, int w_kn,
                                                 float* dbias, int N, int K, int n0, int k0, bool atomic, bool want_bias) {
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  if (w_kn) {
#pragma unroll
    for (int i = 0; i < 4; ++i) tr[ty + 8 * i][tx] = acc[i];      // tr[n_local][k_local]
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = k0 + ty + 8 * i, n = n0 + tx;
      if (n < N && k < K) {
        float* d = dW + (size_t)k * ldw + n;
        const float v = tr[tx][ty + 8 * i];
        if (atomic) atomicAdd(d, v); else *d += v;
      }
    }
  } else {
    const int k = k0 + tx;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int n = n0 + ty + 8 * i;
      if (n < N && k < K) {
        float* d = dW + (size_t)n * ldw + k;
        if (atomic) atomicAdd(d, acc[i]); else *d += acc[i];
      }
    }
  }
  if (want_bias && ty == 0 && n0 + tx < N) {
    if (atomic) atomicAdd(&dbias[n0 + tx], bacc); else dbias[n0 + tx] += bacc;
  }
}

// dW(n,k) += sum_b g[b][n] * x[b][k] ; dbias[n] += sum_b g[b][n].  gridDim.z slices the batch.
template <typename TG>
__global__ void __launch_bounds__(256) linear_bwd_weight_kernel(const TG* __restrict__ g, int ldg, const float* __restrict__ x,
                                                                int ldx, float* dW, int ldw, int w_kn, float* dbias, int B, int N,
                                                                int K) {
  __shared__ float gs[32][33];  // [b][n]
  __shared__ float xs[32][33];  // [b][k]
  const int k0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
  const int per = ((B + gridDim.z - 1) / gridDim.z + 31) & ~31;
  const int bbeg = blockIdx.z * per, bend = min(B, bbeg + per);
  const bool want_bias = dbias != nullptr && blockIdx.x == 0;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  float bacc = 0.f;
  wgrad_tile_accum<TG>(acc, bacc, gs, xs, g, ldg, x, ldx, N, K, n0, k0, bbeg, bend, want_bias);
  wgrad_tile_store(acc, bacc, gs, dW, ldw, w_kn, dbias, N, K, n0, k0, gridDim.z > 1, want_bias);
}

__global__ void __launch_bounds__(256) heads_bwd_weight_kernel(const __grid_constant__ HeadSet hs, const float* __restrict__ x, int ldx,
                                                               int B, int K) {
  __shared__ float gs[32][33];
  __shared__ float xs[32][33];
  int h = 0;
  while (h + 1 < hs.n && (int)blockIdx.y >= hs.tile0[h + 1]) ++h;
  const int N = hs.N[h], n0 = (blockIdx.y - hs.tile0[h]) * 32, k0 = blockIdx.x * 32;
  const int per = ((B + gridDim.z - 1) / gridDim.z + 31) & ~31;
  const int bbeg = blockIdx.z * per, bend = min(B, bbeg + per);
  const bool want_bias = hs.dbias[h] != nullptr && blockIdx.x == 0;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  float bacc = 0.f;
  wgrad_tile_accum<float>(acc, bacc, gs, xs, hs.g[h], N, x, ldx, N, K, n0, k0, bbeg, bend, want_bias);
  wgrad_tile_store(acc, bacc, gs, hs.dW[h], K, 0, hs.dbias[h], N, K, n0, k0, gridDim.z > 1, want_bias);
}

__global__ void log_softmax_fwd_kernel(const float* __restrict__ logits, float* __restrict__ out, int B, int N) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const float* r = logits + (size_t)b * N;
  float mx = -INFINITY;
  for (int i = lane; i < N; i += 32) mx = fmaxf(mx, r[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float s = 0.f;
  for (int i = lane; i < N; i += 32) s += expf(r[i] - mx);
  s = warp_sum(s);
  const float lse = mx + logf(s);
  for (int i = lane; i < N; i += 32) out[(size_t)b * N + i] = r[i] - lse;
}

__global__ void log_softmax_bwd_kernel(const float* __restrict__ g_out, const float* __restrict__ out, float* __restrict__ g_logits,
                                       int B, int N) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  float s = 0.f;
  for (int i = lane; i < N; i += 32) s += g_out[(size_t)b * N + i];
  s = warp_sum(s);
  for (int i = lane; i < N; i += 32) {
    const size_t k = (size_t)b * N + i;
    g_logits[k] = g_out[k] - expf(out[k]) * s;
  }
}

}  // namespace

extern "C" {

int sv_linear_fwd(const float* x, int32_t ldx, const float* W, int32_t ldw, int32_t w_kn, const float* bias, float* out_f32,
                  void* out_bf16, int32_t ldo, float* stats, int32_t group_rows, int32_t B, int32_t N, int32_t K,
                  void* stream) {
  SV_REQUIRE(x && W && (out_f32 || out_bf16), "sv_linear_fwd: null pointer");
  SV_REQUIRE(!stats || group_rows > 0, "sv_linear_fwd: group_rows");
  dim3 grid(ceil_div(N, 32), ceil_div(B, 32));
  linear_fwd_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(x, ldx, W, ldw, w_kn, bias, out_f32, (bf16*)out_bf16, ldo, stats,
                                                                   group_rows > 0 ? group_rows : 1, 0, B, N, K);
  return sv_check_launch("linear_fwd");
}

int sv_linear_bwd_input(const float* g_f32, const void* g_bf16, int32_t ldg, const float* W, int32_t ldw, int32_t w_kn, float* gx,
                        int32_t ldgx, int32_t accumulate, int32_t B, int32_t N, int32_t K, void* stream) {
  SV_REQUIRE((g_f32 != nullptr) != (g_bf16 != nullptr), "sv_linear_bwd_input: exactly one gradient input");
  // gx[b][k] = sum_n g[b][n] W(n,k): the forward kernel with the roles of N and K exchanged.  A long reduction with few
  // output tiles (the decoder stem: 1024 outputs -> 138 latents, 40 tiles) is a serial chain of 32 L2 round trips per block:
  // split it over gridDim.z, partial sums added with atomics into the (zeroed unless accumulating) result.
  int z = 1;
  const int tiles = ceil_div(K, 32) * ceil_div(B, 32);
  if (N >= 256 && tiles < 148) z = N / 128 < 8 ? N / 128 : 8;
  dim3 grid(ceil_div(K, 32), ceil_div(B, 32), z);
  if (z > 1 && !accumulate) {
    if (cudaMemset2DAsync(gx, (size_t)ldgx * sizeof(float), 0, (size_t)K * sizeof(float), B, (cudaStream_t)stream) != cudaSuccess)
      return sv_check_launch("linear_bwd_input (memset)");
  }
  if (g_f32)
    linear_fwd_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(g_f32, ldg, W, ldw, !w_kn, nullptr, gx, nullptr, ldgx, nullptr,
                                                                     1, accumulate, B, K, N);
  else
    linear_fwd_kernel<bf16><<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)g_bf16, ldg, W, ldw, !w_kn, nullptr, gx, nullptr,
                                                                    ldgx, nullptr, 1, accumulate, B, K, N);
  return sv_check_launch("linear_bwd_input");
}

static int batch_slices(int B, int tiles) {
  // slices of the batch so that about one wave of blocks runs: every slice adds its tile with atomics
  int z = 1;
  while (z < 8 && tiles * z * 2 <= 296 && B / (z * 2) >= 32) z *= 2;
  return z;
}

int sv_linear_bwd_weight(const float* g_f32, const void* g_bf16, int32_t ldg, const float* x, int32_t ldx, float* dW, int32_t ldw,
                         int32_t w_kn, float* dbias, int32_t B, int32_t N, int32_t K, void* stream) {
  SV_REQUIRE((g_f32 != nullptr) != (g_bf16 != nullptr), "sv_linear_bwd_weight: exactly one gradient input");
  dim3 grid(ceil_div(K, 32), ceil_div(N, 32), batch_slices(B, ceil_div(K, 32) * ceil_div(N, 32)));
  if (g_f32)
    linear_bwd_weight_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(g_f32, ldg, x, ldx, dW, ldw, w_kn, dbias, B, N, K);
  else
    linear_bwd_weight_kernel<bf16><<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)g_bf16, ldg, x, ldx, dW, ldw, w_kn, dbias,
                                                                           B, N, K);
  return sv_check_launch("linear_bwd_weight");
}

static int fill_heads(HeadSet& hs, const sv_heads* h, const char* what) {
  SV_REQUIRE(h != nullptr && h->n >= 1 && h->n <= SV_MAX_HEADS, "%s: 1..%d heads", what, SV_MAX_HEADS);
  int tiles = 0;
  hs.n = h->n;
  for (int i = 0; i < h->n; ++i) {
    SV_REQUIRE(h->W[i] != nullptr && h->N[i] > 0, "%s: head %d has no weight", what, i);
    hs.W[i] = h->W[i]; hs.bias[i] = h->bias[i]; hs.out[i] = h->out[i]; hs.g[i] = h->g[i]; hs.dW[i] = h->dW[i]; hs.dbias[i] = h->dbias[i];
    hs.N[i] = h->N[i];
    hs.tile0[i] = tiles;
    tiles += ceil_div(h->N[i], 32);
  }
  hs.tile0[h->n] = tiles;
  return SV_OK;
}

int sv_sizeof_heads(void) { return (int)sizeof(sv_heads); }

int sv_heads_fwd(const float* x, int32_t ldx, const sv_heads* heads, int32_t B, int32_t K, void* stream) {
  HeadSet hs;
  if (int rc = fill_heads(hs, heads, "sv_heads_fwd")) return rc;
  SV_REQUIRE(x != nullptr, "sv_heads_fwd: null input");
  for (int i = 0; i < hs.n; ++i) SV_REQUIRE(hs.out[i] != nullptr, "sv_heads_fwd: head %d has no output", i);
  heads_fwd_kernel<<<dim3(hs.tile0[hs.n], ceil_div(B, 32)), 256, 0, (cudaStream_t)stream>>>(x, ldx, hs, B, K);
  return sv_check_launch("heads_fwd");
}

int sv_heads_bwd_input(const sv_heads* heads, float* gx, int32_t ldgx, int32_t B, int32_t K, void* stream) {
  HeadSet hs;
  if (int rc = fill_heads(hs, heads, "sv_heads_bwd_input")) return rc;
  SV_REQUIRE(gx != nullptr, "sv_heads_bwd_input: null output");
  for (int i = 0; i < hs.n; ++i) SV_REQUIRE(hs.g[i] != nullptr, "sv_heads_bwd_input: head %d has no gradient", i);
  heads_bwd_input_kernel<<<dim3(ceil_div(K, 32), ceil_div(B, 32)), 256, 0, (cudaStream_t)stream>>>(hs, gx, ldgx, B, K);
  return sv_check_launch("heads_bwd_input");
}

int sv_heads_bwd_weight(const sv_heads* heads, const float* x, int32_t ldx, int32_t B, int32_t K, void* stream) {
  HeadSet hs;
  if (int rc = fill_heads(hs, heads, "sv_heads_bwd_weight")) return rc;
  SV_REQUIRE(x != nullptr, "sv_heads_bwd_weight: null input");
  for (int i = 0; i < hs.n; ++i) SV_REQUIRE(hs.g[i] != nullptr && hs.dW[i] != nullptr, "sv_heads_bwd_weight: head %d incomplete", i);
  dim3 grid(ceil_div(K, 32), hs.tile0[hs.n], batch_slices(B, ceil_div(K, 32) * hs.tile0[hs.n]));
  heads_bwd_weight_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(hs, x, ldx, B, K);
  return sv_check_launch("heads_bwd_weight");
}

int sv_log_softmax_fwd(const float* logits, float* out, int32_t B, int32_t N, void* stream) {
  log_softmax_fwd_kernel<<<ceil_div(B, 4), 128, 0, (cudaStream_t)stream>>>(logits, out, B, N);
  return sv_check_launch("log_softmax_fwd");
}

int sv_log_softmax_bwd(const float* g_out, const float* out, float* g_logits, int32_t B, int32_t N, void* stream) {
  log_softmax_bwd_kernel<<<ceil_div(B, 4), 128, 0, (cudaStream_t)stream>>>(g_out, out, g_logits, B, N);
  return sv_check_launch("log_softmax_bwd");
}

}  // extern "C"
