// Micro-probe: how fast can ONE thread per SM pull operand tiles into shared memory with TMA?
// Every CTA (one per SM) issues `nloads` tensor-map (or 1-D bulk) loads through a ring of `stages` shared-memory
// buffers and waits for each in order; cycles / bytes per SM answer "what bounds a TMA-fed implicit GEMM":
// the box-row rate, the bytes, or the latency x stages in flight.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/build/tma_probe tools/tma_probe.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!try_wait(bar, parity))
    if (++spins > (1u << 20)) { printf("tma_probe: timeout block %d\n", blockIdx.x); __trap(); }
}

struct Cfg {
  int ndim;            // 0 = 1-D bulk copy, 2..5 = tensor map rank
  int box_bytes;
  int stages, nloads;
  int mode;            // coordinate pattern
  int span;            // number of distinct tiles along the outermost dimension per CTA
  int hq;              // plane-major cases: image height / 4
  int issuers, burst;  // warps issuing their own rings; burst: issue `stages` loads, then wait for all of them
  const uint8_t* gsrc; // bulk mode source
  long long gbytes;
};

__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ CUtensorMap tm, const Cfg c, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t full_all[4][8];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t stage_bytes = (uint32_t)((c.box_bytes + 1023) & ~1023);
  if (threadIdx.x == 0) {
    for (int w = 0; w < 4; ++w)
      for (int s = 0; s < c.stages; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full_all[w][s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int wid = threadIdx.x >> 5;
  uint64_t* full = full_all[wid];
  smem += (size_t)wid * c.stages * stage_bytes;
  if ((threadIdx.x & 31) == 0 && wid < c.issuers) {
    if (c.ndim) asm volatile("prefetch.tensormap [%0];" ::"l"(&tm) : "memory");
    const long long t0 = clock64();
    for (int i = 0; i < c.nloads + c.stages; ++i) {
      const int s = i % c.stages;
      if (c.burst) {
        if (s == 0 && i >= c.stages)
          for (int k = 0; k < c.stages; ++k) wait(&full[k], (uint32_t)(((i / c.stages) - 1) & 1));
      } else if (i >= c.stages) wait(&full[s], (uint32_t)(((i / c.stages) - 1) & 1));
      if (i < c.nloads) {
        const uint32_t dst = smem_u32(smem) + (uint32_t)s * stage_bytes, bar = smem_u32(&full[s]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(c.box_bytes) : "memory");
        const int tile = (blockIdx.x * 4 + wid) * c.span + (i % c.span);
        if (c.ndim == 0) {
          const uint8_t* src = c.gsrc + ((long long)tile * c.box_bytes) % (c.gbytes - c.box_bytes);
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(dst), "l"(src), "r"(c.box_bytes), "r"(bar) : "memory");
        } else if (c.ndim == 2) {
          // mode: rows per box in `mode`
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                       ::"r"(dst), "l"(&tm), "r"(bar), "r"(0), "r"(tile * c.mode) : "memory");
        } else if (c.ndim == 3) {
          // plane-major halo tile: coords (-1 [x, border slot], y0 - 1, plane-row index)
          const int y0 = (i % 3) * c.hq - 1;
          asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                       ::"r"(dst), "l"(&tm), "r"(bar), "r"(-2), "r"(y0), "r"(tile * c.mode) : "memory");
        } else {
          // NHWC per-tap box: coords (c0, dx, dy, img0)
          const int t = i % 9, dy = t / 3 - 1, dx = t % 3 - 1;
          asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                       ::"r"(dst), "l"(&tm), "r"(bar), "r"(0), "r"(dx), "r"(dy), "r"(tile * c.mode) : "memory");
        }
      }
    }
    const long long t1 = clock64();
    if (wid == 0) out[blockIdx.x] = t1 - t0;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
  return reinterpret_cast<EncodeTiledFn>(sym);
}

struct Case {
  const char* name;
  int ndim;
  CUtensorMapDataType dt;
  int esz;
  cuuint64_t dims[5];
  cuuint32_t box[5];
  CUtensorMapSwizzle sw;
  int mode;     // coordinate multiplier of the outermost coordinate per tile
};

int main(int argc, char** argv) {
  const long long GB = 1ll << 28;   // 256 MB source (larger than L2); small-span runs stay L2 resident
  uint8_t* g;
  cudaMalloc(&g, GB);
  cudaMemset(g, 1, GB);
  long long* out;
  cudaMalloc(&out, 148 * sizeof(long long));
  EncodeTiledFn enc = get_encode();
  if (!enc) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  int nsm = 148;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);

  {   // ~0.3 s of work so the SM clock has ramped up before anything is measured
    for (int i = 0; i < 300; ++i) cudaMemset(g, i & 1, GB);
    cudaDeviceSynchronize();
  }
  Case cases[] = {
      {"2d gemm tile 64x128 bf16 SW128 (128 rows x 128 B)", 2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, {64, 1u << 21}, {64, 128}, CU_TENSOR_MAP_SWIZZLE_128B, 128},
      {"2d gemm tile 64x256 bf16 SW128 (256 rows x 128 B)", 2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, {64, 1u << 21}, {64, 256}, CU_TENSOR_MAP_SWIZZLE_128B, 256},
      {"2d 32x128 bf16 SW64 (128 rows x 64 B)", 2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, {32, 1u << 22}, {32, 128}, CU_TENSOR_MAP_SWIZZLE_64B, 128},
      {"2d 256x32 bf16 no swizzle (32 rows x 512 B)", 2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, {256, 1u << 19}, {256, 32}, CU_TENSOR_MAP_SWIZZLE_NONE, 32},
      {"2d 256x128 bf16 no swizzle (128 rows x 512 B = 64 KB)", 2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, {256, 1u << 19}, {256, 128}, CU_TENSOR_MAP_SWIZZLE_NONE, 128},
      {"2d 8x256 bf16 no swizzle (256 rows x 16 B)", 2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, {8, 1u << 24}, {8, 256}, CU_TENSOR_MAP_SWIZZLE_NONE, 256},
      {"4d NHWC C128 8x8: box 64ch x 8 x 8 x 2img SW128 (128 rows x 128 B, per tap)", 4, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, {128, 8, 8, 1u << 14}, {64, 8, 8, 2}, CU_TENSOR_MAP_SWIZZLE_128B, 2},
      {"4d NHWC C64 16x16: box 64ch x 16 x 8 x 1 SW128", 4, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, {64, 16, 16, 1u << 13}, {64, 16, 8, 1}, CU_TENSOR_MAP_SWIZZLE_128B, 1},
      {"4d NHWC C32 32x32: box 32ch x 32 x 4 x 1 SW64", 4, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, {32, 32, 32, 1u << 12}, {32, 32, 4, 1}, CU_TENSOR_MAP_SWIZZLE_64B, 1},
      {"3d plane-major 32x32 C32 halo: box 68 u64 (34 slots) x 7 rows x 4 planes (28 rows x 544 B)", 3, CU_TENSOR_MAP_DATA_TYPE_UINT64, 8, {64, 32, 1u << 14}, {68, 7, 4}, CU_TENSOR_MAP_SWIZZLE_NONE, 4},
      {"3d plane-major 16x16 C64 halo: box 36 u64 x 11 rows x 8 planes (88 rows x 288 B)", 3, CU_TENSOR_MAP_DATA_TYPE_UINT64, 8, {32, 16, 1u << 16}, {36, 11, 8}, CU_TENSOR_MAP_SWIZZLE_NONE, 8},
      {"3d plane-major 8x8 C128 halo: box 20 u64 x 10 rows x 16 planes (160 rows x 160 B)", 3, CU_TENSOR_MAP_DATA_TYPE_UINT64, 8, {16, 8, 1u << 18}, {20, 10, 16}, CU_TENSOR_MAP_SWIZZLE_NONE, 16},
      {"1-D bulk copy 16 KB", 0, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, {16384}, {16384}, CU_TENSOR_MAP_SWIZZLE_NONE, 1},
      {"1-D bulk copy 32 KB", 0, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, {32768}, {32768}, CU_TENSOR_MAP_SWIZZLE_NONE, 1},
  };
  const int stages_list[] = {1, 2, 4, 6};
  printf("%-92s %6s %5s %8s %10s %10s %9s\n", "case", "span", "stg", "boxB", "cyc/load", "B/cyc/SM", "rows/cyc");
  for (const Case& cs : cases) {
    CUtensorMap tm;
    memset(&tm, 0, sizeof(tm));
    long long box_bytes = cs.esz;
    long long rows = 1;
    for (int d = 0; d < (cs.ndim ? cs.ndim : 1); ++d) { box_bytes *= cs.box[d]; if (d > 0) rows *= cs.box[d]; }
    if (cs.ndim) {
      cuuint64_t strides[4];
      cuuint64_t s = cs.dims[0] * cs.esz;
      for (int d = 0; d + 1 < cs.ndim; ++d) { strides[d] = s; s *= cs.dims[d + 1]; }
      if ((long long)s > GB) { printf("%s: tensor larger than the buffer\n", cs.name); continue; }
      cuuint32_t es[5] = {1, 1, 1, 1, 1};
      CUresult r = enc(&tm, cs.dt, cs.ndim, g, cs.dims, strides, cs.box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, cs.sw,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("%s: encode failed %d\n", cs.name, (int)r); continue; }
    }
    for (int variant = 0; variant < 4; ++variant) {       // 0: 1 issuer ring; 1: burst; 2: 2 issuers; 3: 4 issuers
      for (int st : stages_list) {
        const int issuers = variant == 2 ? 2 : (variant == 3 ? 4 : 1);
        if ((long long)st * issuers * ((box_bytes + 1023) & ~1023ll) > 190 * 1024) continue;
        if (variant >= 2 && st != 2 && st != 4) continue;
        Cfg c;
        c.ndim = cs.ndim; c.box_bytes = (int)box_bytes; c.stages = st; c.nloads = 240; c.mode = cs.mode; c.span = 8;
        c.gsrc = g; c.gbytes = GB; c.hq = cs.ndim == 3 ? (int)cs.dims[1] / 4 : 0;
        c.issuers = issuers; c.burst = variant == 1;
        const long long outer = cs.ndim ? (long long)cs.dims[cs.ndim - 1] : GB / box_bytes;
        long long max_span = outer / ((long long)cs.mode * nsm * 4);
        if (c.span > max_span) c.span = (int)max_span;
        if (c.span < 1) { printf("%s: tensor too small\n", cs.name); continue; }
        const size_t smem = (size_t)st * issuers * ((box_bytes + 1023) & ~1023ll) + 1024;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        probe<<<nsm, 128, smem>>>(tm, c, out);
        cudaEventRecord(e0);
        probe<<<nsm, 128, smem>>>(tm, c, out);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: %s\n", cs.name, cudaGetErrorString(e)); return 1; }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        long long h[148];
        cudaMemcpy(h, out, nsm * sizeof(long long), cudaMemcpyDeviceToHost);
        double avg = 0;
        for (int i = 0; i < nsm; ++i) avg += (double)h[i];
        avg /= nsm;
        const double tot = (double)box_bytes * c.nloads * issuers;
        printf("%-84s iss%d %s stg%d box %6lld  cyc/load %7.1f  B/cyc/SM %6.2f  kernel %.1f us -> %.2f GHz, %.0f GB/s chip\n", cs.name, issuers,
               c.burst ? "burst" : "ring ", st, box_bytes, avg / c.nloads, tot / avg, ms * 1e3, avg / (ms * 1e6), tot * nsm / (ms * 1e-3) / 1e9);
      }
    }
  }
  return 0;
}
