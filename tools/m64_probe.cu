// Probe: tcgen05.mma with M = 64 (cta_group::1): (1) where do the 64 accumulator rows land in TMEM,
// (2) cycles per MMA against M = 128 for the N the weight-gradient kernel uses.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// K-major no-swizzle canonical layout: core matrix = 8 rows x 16 bytes contiguous (128 B); SBO = stride between
// 8-row groups, LBO = stride between the two 8-element K chunks of one K=16 MMA.
__global__ void probe(int M, int N, int nrep, float* dump, long long* cyc) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __nv_bfloat16* A = (__nv_bfloat16*)smem;              // 128 rows x 16 K: [kchunk(2)][row/8][row%8][8]
  __nv_bfloat16* B = (__nv_bfloat16*)(smem + 16384);    // 256 rows(N) x 16 K
  for (int i = threadIdx.x; i < 128 * 16; i += blockDim.x) {
    const int r = i / 16, k = i % 16;
    A[(k / 8) * 1024 + (r / 8) * 64 + (r % 8) * 8 + (k % 8)] = __float2bfloat16(k == 0 ? (float)(r + 1) : 0.f);
  }
  for (int i = threadIdx.x; i < 256 * 16; i += blockDim.x) {
    const int r = i / 16, k = i % 16;
    B[(k / 8) * 2048 + (r / 8) * 64 + (r % 8) * 8 + (k % 8)] = __float2bfloat16(k == 0 ? 1.f : 0.f);
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tbase)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  // zero the accumulator region first (so untouched lanes read back 0): tcgen05.st of zeros
  {
    const uint32_t taddr = tbase + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < 16; c += 8)
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr + c), "r"(0) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (warp == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint64_t ad = desc(smem_u32(A), 2048, 128), bd = desc(smem_u32(B), 4096, 128);
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    long long t0 = clock64();
    if (pred) {
      for (int i = 0; i < nrep; ++i)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tbase), "l"(ad), "l"(bd), "r"(idesc), "r"(i ? 1u : 0u) : "memory");
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    __syncwarp();
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    long long t1 = clock64();
    if (lane == 0) cyc[0] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  {
    uint32_t r[8];
    const uint32_t taddr = tbase + ((uint32_t)(warp * 32) << 16);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int c = 0; c < 8; ++c) dump[(warp * 32 + lane) * 8 + c] = __uint_as_float(r[c]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512));
}

int main() {
  float* dump; long long* cyc;
  cudaMalloc(&dump, 128 * 8 * 4); cudaMalloc(&cyc, 64);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int M : {128, 64}) {
    probe<<<1, 128, 32 * 1024>>>(M, 32, 1, dump, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("M=%d: %s\n", M, cudaGetErrorString(e)); return 1; }
    float h[128 * 8];
    cudaMemcpy(h, dump, sizeof(h), cudaMemcpyDeviceToHost);
    printf("M=%d: accumulator column 0 by TMEM lane (value = row+1, nrep=1):\n", M);
    for (int l = 0; l < 128; ++l) printf("%s%3.0f", (l % 32 == 0) ? "\n  " : " ", h[l * 8] / 1.f);
    printf("\n  column 1 lanes 0..7:");
    for (int l = 0; l < 8; ++l) printf(" %3.0f", h[l * 8 + 1]);
    printf("\n");
  }
  for (int M : {128, 64})
    for (int N : {32, 64, 128, 256}) {
      long long h = 0;
      for (int rep = 0; rep < 2; ++rep) { probe<<<1, 128, 32 * 1024>>>(M, N, 512, dump, cyc); cudaDeviceSynchronize(); }
      cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      printf("M=%3d N=%3d : %.1f cycles/MMA\n", M, N, h / 512.0);
    }
  return 0;
}
