// Device input pipeline (reference lib/dataloader.py:42-70): the torchvision train transform
//   Pad(4, reflect) -> RandomHorizontalFlip -> RandomCrop(32) -> ToTensor
// as ONE gather kernel over a device-resident uint8 dataset: out[b, c, y, x] = float(src[index[b]][..]) / 255.
// The random parameters (crop offsets i, j in [0, 2*pad], flip bit) are drawn by the caller; the arithmetic is
// bit-exact with torchvision's (uint8 -> float32, IEEE division by 255).  Bandwidth bound: 3 KB read + 12 KB
// written per image.
#include "common.cuh"
#include "../../include/shotvae.h"

namespace {

__device__ __forceinline__ int reflect_index(int t, int n) {
  // numpy / torch 'reflect' (no edge repeat): -1 -> 1, n -> n-2
  if (t < 0) t = -t;
  if (t >= n) t = 2 * (n - 1) - t;
  return t;
}

// One block per output image: the source image (<= 4 KB of uint8) is staged in shared memory with 16-byte loads, then
// every thread converts quads of 4 consecutive output pixels (one 16-byte coalesced store each).  The per-image index /
// parameter loads happen once per block instead of once per output element, and no thread waits on a dependent chain of
// byte loads from global memory.
constexpr int AUG_MAX_BYTES = 4096;

__global__ void __launch_bounds__(256) augment_kernel(const uint8_t* __restrict__ data, const int64_t* __restrict__ index,
                                                      const int32_t* __restrict__ params, int ch, int sh, int sw, int pad,
                                                      int oh, int ow, int hwc, float* __restrict__ out) {
  __shared__ __align__(16) uint8_t img[AUG_MAX_BYTES];
  const int b = blockIdx.x;
  const int nbytes = ch * sh * sw;
  const uint8_t* src = data + (size_t)(index != nullptr ? index[b] : b) * nbytes;
  if ((nbytes & 15) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    for (int i = threadIdx.x; i < (nbytes >> 4); i += blockDim.x)
      reinterpret_cast<uint4*>(img)[i] = reinterpret_cast<const uint4*>(src)[i];
  } else {
    for (int i = threadIdx.x; i < nbytes; i += blockDim.x) img[i] = src[i];
  }
  int ci = 0, cj = 0, flip = 0;
  if (params != nullptr) { ci = params[3 * b]; cj = params[3 * b + 1]; flip = params[3 * b + 2]; }
  __syncthreads();
  const int owq = ow >> 2, pw = sw + 2 * pad;
  float* dst = out + (size_t)b * ch * oh * ow;
  for (int i = threadIdx.x; i < ch * oh * owq; i += blockDim.x) {
    const int xq = i % owq, r = i / owq;
    const int y = r % oh, c = r / oh;
    const int sy = reflect_index(ci + y - pad, sh);
    float v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      int px = cj + xq * 4 + k;                       // column in the padded (and possibly flipped) image
      if (flip) px = pw - 1 - px;
      const int sx = reflect_index(px - pad, sw);
      const uint8_t u = hwc ? img[(sy * sw + sx) * ch + c] : img[(c * sh + sy) * sw + sx];
      v[k] = __fdiv_rn((float)u, 255.0f);
    }
    *reinterpret_cast<float4*>(dst + ((size_t)c * oh + y) * ow + xq * 4) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

}  // namespace

extern "C" int sv_augment_batch(const uint8_t* data, const int64_t* index, const int32_t* params, int32_t B, int32_t ch, int32_t src_h,
                                int32_t src_w, int32_t pad, int32_t out_h, int32_t out_w, int32_t src_hwc, float* out, void* stream) {
  SV_REQUIRE(data && out, "sv_augment_batch: null pointer");
  SV_REQUIRE(B > 0 && ch > 0 && src_h > 1 && src_w > 1 && pad >= 0 && pad < src_h && pad < src_w, "sv_augment_batch: bad geometry");
  SV_REQUIRE(out_w % 4 == 0 && out_h <= src_h + 2 * pad && out_w <= src_w + 2 * pad, "sv_augment_batch: output %dx%d does not fit", out_h, out_w);
  SV_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "sv_augment_batch: output must be 16-byte aligned");
  SV_REQUIRE(ch * src_h * src_w <= AUG_MAX_BYTES, "sv_augment_batch: source image of %d bytes exceeds the %d-byte staging buffer", ch * src_h * src_w, AUG_MAX_BYTES);
  augment_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(data, index, params, ch, src_h, src_w, pad, out_h, out_w, src_hwc, out);
  return sv_check_launch("sv_augment_batch");
}
