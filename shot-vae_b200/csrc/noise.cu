// Device-side noise for the fused step (benchmark / production mode; parity runs feed host-drawn noise instead):
// the reparameterisation noise eps ~ N(0, 1) (reference vae.py:82-84, torch.randn) and the Gumbel-softmax uniforms
// u ~ U[0, 1) (vae.py:69, torch.rand) of all four passes in ONE launch, counter-based (Philox4x32-10) so that a CUDA-graph
// replay draws fresh numbers: the (seed, offset) pair lives in device memory and the last block to finish advances the offset.
// sv_fill_zero is the graph-capturable memset used for the per-step accumulators (a memset node, no kernel).
#include "common.cuh"
#include "../../include/shotvae.h"

namespace {

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

struct NoiseState {
  unsigned long long seed, offset;
  unsigned int ticket, pad0;
  unsigned long long pad1;
};

__global__ void __launch_bounds__(256) noise_fill_kernel(float* __restrict__ normal, long long n_normal, float* __restrict__ uniform,
                                                         long long n_uniform, NoiseState* state) {
  const long long qn = (n_normal + 3) >> 2, qu = (n_uniform + 3) >> 2;
  const unsigned long long seed = state->seed, offset = state->offset;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < qn + qu; i += (long long)gridDim.x * blockDim.x) {
    const unsigned long long ctr = offset + (unsigned long long)i;
    const uint4 r = philox4x32_10(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    float v[4];
    if (i < qn) {
      // Box-Muller on two pairs; u1 in (0, 1] so the logarithm is finite
      const float u1 = (float)((r.x >> 8) + 1u) * (1.f / 16777216.f), u2 = (float)(r.y >> 8) * (1.f / 16777216.f);
      const float u3 = (float)((r.z >> 8) + 1u) * (1.f / 16777216.f), u4 = (float)(r.w >> 8) * (1.f / 16777216.f);
      const float ra = sqrtf(-2.f * logf(u1)), rb = sqrtf(-2.f * logf(u3));
      float sa, ca, sb, cb;
      sincospif(2.f * u2, &sa, &ca);
      sincospif(2.f * u4, &sb, &cb);
      v[0] = ra * ca; v[1] = ra * sa; v[2] = rb * cb; v[3] = rb * sb;
      const long long e = i << 2;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (e + j < n_normal) normal[e + j] = v[j];
    } else {
      const uint32_t x[4] = {r.x, r.y, r.z, r.w};
      const long long e = (i - qn) << 2;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (e + j < n_uniform) uniform[e + j] = (float)(x[j] >> 8) * (1.f / 16777216.f);      // [0, 1), 24 bits like torch.rand
    }
  }
  // the last block to get here has seen every other block finish reading `offset`
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int t = atomicAdd(&state->ticket, 1u);
    if (t == gridDim.x - 1) {
      state->offset = offset + (unsigned long long)(qn + qu);
      state->ticket = 0u;
    }
  }
}

}  // namespace

extern "C" {

int sv_noise_fill(float* normal, int64_t n_normal, float* uniform, int64_t n_uniform, void* state, void* stream) {
  SV_REQUIRE(state && (normal || n_normal == 0) && (uniform || n_uniform == 0) && n_normal >= 0 && n_uniform >= 0, "sv_noise_fill: bad arguments");
  const long long quads = ((n_normal + 3) >> 2) + ((n_uniform + 3) >> 2);
  if (quads == 0) return SV_OK;
  long long blocks = (quads + 255) / 256;
  if (blocks > 148 * 4) blocks = 148 * 4;
  noise_fill_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(normal, (long long)n_normal, uniform, (long long)n_uniform, (NoiseState*)state);
  return sv_check_launch("noise_fill");
}

int sv_sizeof_noise_state(void) { return (int)sizeof(NoiseState); }

int sv_fill_zero(void* dst, int64_t nbytes, void* stream) {
  SV_REQUIRE(dst && nbytes >= 0, "sv_fill_zero: bad arguments");
  if (nbytes == 0) return SV_OK;
  const cudaError_t e = cudaMemsetAsync(dst, 0, (size_t)nbytes, (cudaStream_t)stream);
  if (e != cudaSuccess) {
    sv_set_error("sv_fill_zero: %s", cudaGetErrorString(e));
    return SV_ERR_CUDA;
  }
  return SV_OK;
}

}  // extern "C"
