// Micro-probe: cost of the halo kernel's per-tile tcgen05 issue pattern (NB MMAs + NC commits per batch,
// alternating accumulators) with and without concurrent global stores / cp.async traffic from other warps.
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

template <int NMMA, int NCOMMIT>
__global__ void probe(int N, int nbatch, long long* out, int stress, uint4* gdst, const uint4* gsrc, int waitmode) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[4];
  __shared__ uint32_t tbase;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;");
    stop = 0;
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
  if (warp != 0) {
    uint32_t sink = 0;
    uint32_t k = 0;
    while (!stop) {
      if (stress & 1) {   // scattered 16-byte global stores, 64-byte stride between lanes (the direct epilogue's pattern)
        uint4 v = make_uint4(sink, 1, 2, 3);
        size_t o = ((size_t)blockIdx.x * 65536 + ((k * 288 + threadIdx.x) & 16383)) * 4;
        gdst[o] = v; gdst[o + 1] = v;
      }
      if (stress & 2) {   // coalesced 16-byte global stores
        uint4 v = make_uint4(sink, 1, 2, 3);
        size_t o = ((size_t)blockIdx.x * 65536 + ((k * 288 + threadIdx.x) & 16383)) * 4;
        gdst[o / 4 * 2] = v; gdst[o / 4 * 2 + 1] = v;   // 32 B per lane, contiguous across the warp
      }
      if (stress & 4) {   // cp.async traffic into shared memory
        for (int j = 0; j < 7; ++j)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem + 49152 + 16 * ((threadIdx.x + 288 * j) & 1023))), "l"(gsrc + ((blockIdx.x * 4096 + threadIdx.x + 288 * (j + k)) & 0xFFFFF)) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 2;" ::: "memory");
      }
      if (stress & 8) {   // tcgen05.ld traffic
        uint32_t r[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(tbase + ((uint32_t)((warp & 3) * 32) << 16) + 128u));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        sink += r[0];
      }
      if (!(stress & 15)) __nanosleep(100);
      ++k; sink += k;
      if (stress & 16) __nanosleep(200);    // paced (roughly the real epilogue's duty cycle)
    }
    if (sink == 0x12345) out[3] = sink;
  }
  if (warp == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a_base = smem_u32(smem) + 16, b_base = smem_u32(smem) + 32768;
    const uint64_t ad = desc(a_base, 3808, 128, 0), bd = desc(b_base, 2048, 128, 0);
    long long t0 = clock64();
    uint32_t ph[4] = {0, 0, 0, 0};
    for (int b = 0; b < nbatch; ++b) {
      const uint32_t d = tbase + (uint32_t)((b & 1) * N);
      if (waitmode == 1 && b >= 2) {      // like the real kernel: wait until the batch two back has completed (accumulator reuse)
        const int w = 2 + (b & 1);
        uint32_t ok = 0;
        while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar[w])), "r"(ph[w]) : "memory");
        ph[w] ^= 1;
      }
      if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NMMA; ++i) {
          const uint64_t ad2 = ad + (uint64_t)(((i >> 1) % 9) * 35 + (i & 1) * 2 * (3808 >> 4));
          const uint64_t bd2 = bd + (uint64_t)((i % 18) * 64);
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(d), "l"(ad2), "l"(bd2), "r"(idesc), "r"(i ? 1u : 0u) : "memory");
        }
        if (NCOMMIT >= 2) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[b & 1])) : "memory");
        if (NCOMMIT >= 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[2 + (b & 1)])) : "memory");
      }
      __syncwarp();
    }
    long long t1 = clock64();
    if (lane == 0 && blockIdx.x == 0) { out[0] = t1 - t0; }
    // drain
    if (lane == 0) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[0])) : "memory");
    __nanosleep(20000);
    long long t2 = clock64();
    if (lane == 0 && blockIdx.x == 0) out[1] = t2 - t0;
    stop = 1;
    asm volatile("tcgen05.fence::before_thread_sync;");
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(256));
}

template <int NMMA, int NCOMMIT>
void run(const char* tag, int N, int stress, int grid, int waitmode, long long* out, uint4* gdst, uint4* gsrc) {
  cudaFuncSetAttribute(probe<NMMA, NCOMMIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  long long h[2];
  const int nb = 64;
  for (int rep = 0; rep < 2; ++rep) {
    probe<NMMA, NCOMMIT><<<grid, 288, 80 * 1024>>>(N, nb, out, stress, gdst, gsrc, waitmode);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", tag, cudaGetErrorString(e)); return; }
  }
  cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
  printf("%-28s N=%3d mma/batch=%2d commits=%d stress=%2d grid=%3d wait=%d : %7.1f cycles/batch (%5.1f per MMA)\n", tag, N, NMMA, NCOMMIT, stress, grid,
         waitmode, h[0] / (double)nb, h[0] / (double)nb / (NMMA ? NMMA : 1));
}

int main() {
  long long* out; uint4 *gdst, *gsrc;
  cudaMalloc(&out, 64);
  cudaMalloc(&gdst, (size_t)148 * 65536 * 4 * 16 + 65536);
  cudaMalloc(&gsrc, (size_t)(1 << 20) * 16 + 65536);
  for (int wm = 0; wm < 2; ++wm) {
    run<18, 2>("halo pattern", 32, 0, 1, wm, out, gdst, gsrc);
    run<18, 1>("one commit", 32, 0, 1, wm, out, gdst, gsrc);
    run<18, 0>("no commit", 32, 0, 1, wm == 1 ? 0 : 0, out, gdst, gsrc);
    run<36, 2>("36 mma", 32, 0, 1, wm, out, gdst, gsrc);
    run<0, 2>("commits only", 32, 0, 1, wm, out, gdst, gsrc);
    run<18, 2>("N=64", 64, 0, 1, wm, out, gdst, gsrc);
    run<18, 2>("+scattered stores", 32, 1, 148, wm, out, gdst, gsrc);
    run<18, 2>("+scattered stores paced", 32, 17, 148, wm, out, gdst, gsrc);
    run<18, 2>("+coalesced stores", 32, 2, 148, wm, out, gdst, gsrc);
    run<18, 2>("+coalesced stores paced", 32, 18, 148, wm, out, gdst, gsrc);
    run<18, 2>("+cp.async", 32, 4, 148, wm, out, gdst, gsrc);
    run<18, 2>("+tcgen05.ld", 32, 8, 148, wm, out, gdst, gsrc);
    run<18, 2>("+tcgen05.ld paced", 32, 24, 148, wm, out, gdst, gsrc);
    run<18, 2>("+all paced", 32, 16 + 1 + 4 + 8, 148, wm, out, gdst, gsrc);
  }
  return 0;
}
