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

// out[b][n] (+)= sum_k x[b][k] * W(n,k) + bias[n]
template <typename TX>
__global__ void __launch_bounds__(256) linear_fwd_kernel(const TX* __restrict__ x, int ldx, const float* __restrict__ W, int ldw,
                                                         int w_kn, const float* __restrict__ bias, float* out_f32, bf16* out_bf16,
                                                         int ldo, float* stats, int group_rows, int accumulate, int B, int N,
                                                         int K) {
  __shared__ float xs[32][33];
  __shared__ float ws[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int b0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k0 = 0; k0 < K; k0 += 32) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = ty + 8 * i;
      const int b = b0 + r, k = k0 + tx;
      xs[r][tx] = (b < B && k < K) ? ldf<TX>(x, (size_t)b * ldx + k) : 0.f;
      // ws[n_local][k_local]
      if (w_kn) {
        const int kk = k0 + r, n = n0 + tx;  // coalesced over n
        ws[tx][r] = (kk < K && n < N) ? W[(size_t)kk * ldw + n] : 0.f;
      } else {
        const int n = n0 + r;
        ws[r][tx] = (n < N && k < K) ? W[(size_t)n * ldw + k] : 0.f;
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
  const int n = n0 + tx;
  if (n >= N) return;
  const float bv = bias ? bias[n] : 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int b = b0 + ty + 8 * i;
    if (b >= B) continue;
    float v = acc[i] + bv;
    const size_t o = (size_t)b * ldo + n;
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

// dW(n,k) += sum_b g[b][n] * x[b][k] ; dbias[n] += sum_b g[b][n]
template <typename TG>
__global__ void __launch_bounds__(256) linear_bwd_weight_kernel(const TG* __restrict__ g, int ldg, const float* __restrict__ x,
                                                                int ldx, float* dW, int ldw, int w_kn, float* dbias, int B, int N,
                                                                int K) {
  __shared__ float gs[32][33];  // [b][n]
  __shared__ float xs[32][33];  // [b][k]
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int k0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  float bacc = 0.f;
  for (int b0 = 0; b0 < B; b0 += 32) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = ty + 8 * i, b = b0 + r;
      gs[r][tx] = (b < B && n0 + tx < N) ? ldf<TG>(g, (size_t)b * ldg + n0 + tx) : 0.f;
      xs[r][tx] = (b < B && k0 + tx < K) ? x[(size_t)b * ldx + k0 + tx] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int bb = 0; bb < 32; ++bb) {
      const float xv = xs[bb][tx];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = fmaf(gs[bb][ty + 8 * i], xv, acc[i]);
    }
    if (dbias != nullptr && blockIdx.x == 0 && ty == 0) {
      for (int bb = 0; bb < 32; ++bb) bacc += gs[bb][tx];
    }
    __syncthreads();
  }
  const int k = k0 + tx;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ty + 8 * i;
    if (n < N && k < K) {
      const size_t o = w_kn ? (size_t)k * ldw + n : (size_t)n * ldw + k;
      dW[o] += acc[i];
    }
  }
  if (dbias != nullptr && blockIdx.x == 0 && ty == 0 && n0 + tx < N) dbias[n0 + tx] += bacc;
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
  // gx[b][k] = sum_n g[b][n] W(n,k): the forward kernel with the roles of N and K exchanged
  dim3 grid(ceil_div(K, 32), ceil_div(B, 32));
  if (g_f32)
    linear_fwd_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(g_f32, ldg, W, ldw, !w_kn, nullptr, gx, nullptr, ldgx, nullptr,
                                                                     1, accumulate, B, K, N);
  else
    linear_fwd_kernel<bf16><<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)g_bf16, ldg, W, ldw, !w_kn, nullptr, gx, nullptr,
                                                                    ldgx, nullptr, 1, accumulate, B, K, N);
  return sv_check_launch("linear_bwd_input");
}

int sv_linear_bwd_weight(const float* g_f32, const void* g_bf16, int32_t ldg, const float* x, int32_t ldx, float* dW, int32_t ldw,
                         int32_t w_kn, float* dbias, int32_t B, int32_t N, int32_t K, void* stream) {
  SV_REQUIRE((g_f32 != nullptr) != (g_bf16 != nullptr), "sv_linear_bwd_weight: exactly one gradient input");
  dim3 grid(ceil_div(K, 32), ceil_div(N, 32));
  if (g_f32)
    linear_bwd_weight_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(g_f32, ldg, x, ldx, dW, ldw, w_kn, dbias, B, N, K);
  else
    linear_bwd_weight_kernel<bf16><<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)g_bf16, ldg, x, ldx, dW, ldw, w_kn, dbias,
                                                                           B, N, K);
  return sv_check_launch("linear_bwd_weight");
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
