// Micro-probe: tcgen05.mma issue/execute rate for different shared-memory operand layouts.
// One CTA per SM-sample, one warp issues NREP MMAs back to back on garbage data and times them.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

__global__ void probe(int N, int mode, int a_off_bytes, int nrep, long long* out, int lbo_a, int stress, const uint4* gsrc) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tbase)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  __shared__ volatile int stop;
  if (threadIdx.x == 0) stop = 0;
  __syncthreads();
  if (warp != 0) {
    uint32_t sink = 0;
    while (!stop) {
      if (stress & 1) {
        uint32_t r[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(tbase + ((uint32_t)((warp & 3) * 32) << 16) + 128u));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        sink += r[0];
      }
      if (stress & 2) {
        for (int k = 0; k < 8; ++k) { uint4 val = make_uint4(sink, 1, 2, 3); asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(smem_u32(smem + 49152 + 16 * ((threadIdx.x + 288 * k) & 1023))), "r"(val.x), "r"(val.y), "r"(val.z), "r"(val.w) : "memory"); }
      }
      if (stress & 4) {
        for (int k = 0; k < 7; ++k) {
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem + 49152 + 16 * ((threadIdx.x + 288 * k) & 1023))), "l"(gsrc + ((blockIdx.x * 4096 + threadIdx.x + 288 * k + sink) & 0xFFFFF)) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 2;" ::: "memory");
        if (stress & 8) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        sink += 7;
      }
      if (!(stress & 7)) __nanosleep(100);
    }
    if (sink == 0x12345) out[3] = sink;
  }
  if (warp == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a_base = smem_u32(smem) + a_off_bytes, b_base = smem_u32(smem) + 32768;
    uint64_t ad, bd;
    if (mode == 0) {            // SW128 K-major, rows 128 B apart
      ad = desc(a_base, 16, 1024, 2); bd = desc(b_base, 16, 1024, 2);
    } else if (mode == 1) {     // SW64 K-major
      ad = desc(a_base, 16, 512, 4); bd = desc(b_base, 16, 512, 4);
    } else {                    // no swizzle, interleaved planes: LBO = plane stride, SBO = 128
      ad = desc(a_base, lbo_a, 128, 0); bd = desc(b_base, 2048, 128, 0);
    }
    const uint32_t d = tbase;
    long long t0 = clock64();
    for (int i = 0; i < nrep; ++i) {
      uint64_t ad2 = ad, bd2 = bd;
      if (stress & 16) ad2 = ad + (uint64_t)(((i % 9) * 35) + (i & 1) * 2 * (lbo_a >> 4));
      if (stress & 32) bd2 = bd + (uint64_t)((i % 18) * 64);
      if (lane == 0)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d), "l"(ad2), "l"(bd2), "r"(idesc), "r"(i ? 1u : 0u) : "memory");
    }
    long long t1 = clock64();
    if (lane == 0) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    }
    long long t2 = clock64();
    if (lane == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    stop = 1;
    asm volatile("tcgen05.fence::before_thread_sync;");
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(256));
}

int main() {
  long long* out;
  uint4* gsrc;
  cudaMalloc(&out, 64);
  cudaMalloc(&gsrc, (size_t)(1 << 20) * 16 + 65536);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const char* names[3] = {"SW128", "SW64", "NOSWZ"};
  struct Cfg { int mode, N, off, lbo, stress, grid; };
  Cfg cfgs[] = {{2, 32, 16, 3808, 16, 1}, {2, 32, 16, 3808, 48, 1}, {2, 32, 16, 3808, 48, 148}, {2, 32, 16, 4096, 16, 1}, {2, 32, 0, 4096, 16, 1}, {2, 64, 16, 3168, 48, 148}, {0, 32, 0, 4096, 32, 1}, {2, 32, 16, 3808, 0, 1}, {2, 32, 16, 3808, 0, 148}, {2, 32, 16, 3808, 1, 148}, {2, 32, 16, 3808, 2, 1}, {2, 32, 16, 3808, 2, 148},
                {2, 32, 16, 3808, 4, 1}, {2, 32, 16, 3808, 4, 148}, {2, 32, 16, 3808, 12, 1}, {2, 32, 16, 3808, 12, 148}, {2, 32, 16, 3808, 15, 148},
                {0, 128, 0, 4096, 0, 148}, {0, 256, 0, 4096, 0, 148}};
  for (auto c : cfgs) {
    long long h[2];
    for (int rep = 0; rep < 2; ++rep) {
      probe<<<c.grid, 288, 80 * 1024>>>(c.N, c.mode, c.off, 256, out, c.lbo, c.stress, gsrc);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s N=%d: %s\n", names[c.mode], c.N, cudaGetErrorString(e)); return 1; }
    }
    cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("%-6s N=%3d lbo=%4d stress=%2d grid=%3d : issue %6.1f cyc/MMA, complete %6.1f cyc/MMA\n", names[c.mode], c.N, c.lbo, c.stress, c.grid,
           h[0] / 256.0, h[1] / 256.0);
  }
  return 0;
}
