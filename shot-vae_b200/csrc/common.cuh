// Shared device/host helpers for libshotvae (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define SV_OK 0
#define SV_ERR_ARG (-1)
#define SV_ERR_CUDA (-2)
#define SV_ERR_UNSUPPORTED (-3)

void sv_set_error(const char* fmt, ...);
int sv_check_launch(const char* what);
int sv_cta_limit();          // 0 = every SM; see sv_set_cta_limit

#define SV_REQUIRE(cond, ...)                                                                     \
  do {                                                                                            \
    if (!(cond)) {                                                                                \
      sv_set_error(__VA_ARGS__);                                                                  \
      return SV_ERR_ARG;                                                                          \
    }                                                                                             \
  } while (0)

// ---- programmatic dependent launch (PDL) --------------------------------------------------------------
// Kernels launched with sv_launch_pdl may start while their stream predecessor is still draining: everything
// before pdl_wait() (barrier init, TMEM allocation, shared-memory zeroing) overlaps the predecessor's tail;
// pdl_wait() returns once the predecessor grid has completed and its memory is visible, so no global access
// may precede it.  pdl_trigger() lets the successor begin launching.  Opt-in: SHOTVAE_PDL=1 (no measurable gain inside the step graph).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool sv_pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t sv_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = sv_pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

typedef __nv_bfloat16 bf16;
typedef __nv_bfloat162 bf162;

struct __align__(16) bf16x8 {
  bf162 v[4];
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// 16 values per lane (one row each) -> lane j (< 16) of each half-warp pair ends with the column sums.
// Butterfly transpose-reduce over the 32 lanes: 16 columns => 8+4+2+1 exchanges + 1 final fold.
__device__ __forceinline__ float colsum16(const float* v, int lane) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const bool up = lane & 8;
    const float send = up ? v[i] : v[i + 8];
    const float keep = up ? v[i + 8] : v[i];
    a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const bool up = lane & 4;
    const float send = up ? a[i] : a[i + 4];
    const float keep = up ? a[i + 4] : a[i];
    a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const bool up = lane & 2;
    const float send = up ? a[i] : a[i + 2];
    const float keep = up ? a[i + 2] : a[i];
    a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  {
    const bool up = lane & 1;
    const float send = up ? a[0] : a[1];
    const float keep = up ? a[1] : a[0];
    a[0] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
  }
  // lanes L and L^16 hold partial sums of the same column (L & 15): fold the two half warps
  return a[0] + __shfl_xor_sync(0xffffffffu, a[0], 16);
}


// Block-wide sum; result valid in thread 0 (and broadcast through smem to all threads).
template <int NT>
__device__ __forceinline__ float block_sum(float v, float* red /* >= NT/32 floats */) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float r = 0.f;
  if (w == 0) {
    r = (l < NT / 32) ? red[l] : 0.f;
    r = warp_sum(r);
    if (l == 0) red[0] = r;
  }
  __syncthreads();
  return red[0];
}

__device__ __forceinline__ void unpack8(const bf16x8& p, float* f) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(p.v[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ bf16x8 pack8(const float* f) {
  bf16x8 p;
#pragma unroll
  for (int i = 0; i < 4; ++i) p.v[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return p;
}

// 8 consecutive channels of an activation row: bf16 (production, 16 bytes) or fp32 (parity-grade FP32 mode, 32 bytes).
// The elementwise kernels are templated on the element type through this helper.
template <typename T>
struct V8;
template <>
struct V8<bf16> {
  typedef bf16x8 raw;
  static __device__ __forceinline__ raw load(const bf16* p) { return *reinterpret_cast<const bf16x8*>(p); }
  static __device__ __forceinline__ void unpack(const raw& r, float* f) { unpack8(r, f); }
  static __device__ __forceinline__ void store(bf16* p, const float* f) { *reinterpret_cast<bf16x8*>(p) = pack8(f); }
};
struct __align__(16) f32x8 {
  float4 a, b;
};
template <>
struct V8<float> {
  typedef f32x8 raw;
  static __device__ __forceinline__ raw load(const float* p) {
    raw r;
    r.a = *reinterpret_cast<const float4*>(p);
    r.b = *reinterpret_cast<const float4*>(p + 4);
    return r;
  }
  static __device__ __forceinline__ void unpack(const raw& r, float* f) {
    f[0] = r.a.x; f[1] = r.a.y; f[2] = r.a.z; f[3] = r.a.w; f[4] = r.b.x; f[5] = r.b.y; f[6] = r.b.z; f[7] = r.b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float* f) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
  }
};

// 256-bit global accesses (sm_100): one full 32-byte sector per lane and instruction.  MEASURED
// (tools/store_probe.cu): a 128-row tile of 64-byte rows costs ~600 cycles as 2 x 16-byte stores per thread
// and ~300 as one 32-byte store; 128-byte rows 1330 -> 1010.
__device__ __forceinline__ void st_global_32B(void* ptr, const bf16x8& a, const bf16x8& b) {
  const uint32_t* x = reinterpret_cast<const uint32_t*>(&a);
  const uint32_t* y = reinterpret_cast<const uint32_t*>(&b);
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(ptr), "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(y[0]), "r"(y[1]),
               "r"(y[2]), "r"(y[3])
               : "memory");
}
__device__ __forceinline__ void ld_global_32B(const void* ptr, bf16x8& a, bf16x8& b) {
  uint32_t* x = reinterpret_cast<uint32_t*>(&a);
  uint32_t* y = reinterpret_cast<uint32_t*>(&b);
  asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(x[0]), "=r"(x[1]), "=r"(x[2]), "=r"(x[3]), "=r"(y[0]), "=r"(y[1]), "=r"(y[2]), "=r"(y[3])
               : "l"(ptr));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// One lane of a converged warp.  Unlike `lane == 0`, ptxas knows that exactly one thread passes an
// elect.sync predicate, so tcgen05.mma / tcgen05.commit behind it are emitted back to back instead of
// inside a per-active-thread ELECT/BRA.U.ANY serialisation loop (~6 extra instructions per MMA).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(smem)), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }
