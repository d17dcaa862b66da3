// Micro-probe: per-SM cost of the epilogue's global store patterns.  256 threads per CTA (8 warps), one CTA
// per SM, each "tile" = 128 rows x ROWB bytes written to a fresh region (streams through L2 like the conv output).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void st16(void* p, uint32_t v) {
  asm volatile("st.global.v4.b32 [%0], {%1,%1,%1,%1};" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st32(void* p, uint32_t v) {
  asm volatile("st.global.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p), "r"(v) : "memory");
}

// mode 0: thread = (row, half): two 16-byte stores (current epilogue, N=32: 32 B per thread at 64-byte row pitch)
// mode 1: same bytes, one 32-byte store
// mode 2: 4 warps active, each thread a whole 64-byte row: four 16-byte stores
// mode 3: 4 warps active, whole row: two 32-byte stores
// mode 4: fully coalesced: every warp instruction writes 512 contiguous bytes (16 B per lane)
// mode 5: fully coalesced 32 B per lane (1 KB per instruction)
__global__ void probe(uint8_t* out, int mode, int tiles, int rowb, long long* cyc) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const size_t tile_bytes = (size_t)128 * rowb;
  uint8_t* base = out + (size_t)blockIdx.x * tiles * tile_bytes;
  __syncthreads();
  long long t0 = clock64();
  for (int t = 0; t < tiles; ++t) {
    uint8_t* tb = base + (size_t)t * tile_bytes;
    const int q = warp & 3, half = warp >> 2;
    const int row = q * 32 + lane;
    if (mode == 0) {
      uint8_t* p = tb + (size_t)row * rowb + half * (rowb / 2);
      for (int k = 0; k < rowb / 2; k += 16) st16(p + k, t);
    } else if (mode == 1) {
      uint8_t* p = tb + (size_t)row * rowb + half * (rowb / 2);
      for (int k = 0; k < rowb / 2; k += 32) st32(p + k, t);
    } else if (mode == 2) {
      if (half == 0) { uint8_t* p = tb + (size_t)row * rowb; for (int k = 0; k < rowb; k += 16) st16(p + k, t); }
    } else if (mode == 3) {
      if (half == 0) { uint8_t* p = tb + (size_t)row * rowb; for (int k = 0; k < rowb; k += 32) st32(p + k, t); }
    } else if (mode == 4) {
      for (size_t o = (size_t)tid * 16; o < tile_bytes; o += 256 * 16) st16(tb + o, t);
    } else {
      for (size_t o = (size_t)tid * 32; o < tile_bytes; o += 256 * 32) st32(tb + o, t);
    }
  }
  long long t1 = clock64();
  if (tid == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
  const int tiles = 64;
  uint8_t* out; long long* cyc;
  cudaMalloc(&out, (size_t)148 * tiles * 128 * 256 + 4096);
  cudaMalloc(&cyc, 64);
  const char* names[] = {"2x16B per thread, column split", "1x32B per thread, column split", "whole row 16B stores (4 warps)", "whole row 32B stores (4 warps)",
                         "coalesced 16B/lane", "coalesced 32B/lane"};
  for (int rowb : {64, 128, 256}) {
    for (int mode = 0; mode < 6; ++mode) {
      long long h = 0;
      for (int rep = 0; rep < 3; ++rep) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        probe<<<148, 256>>>(out, mode, tiles, rowb, cyc);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        if (rep == 2)
          printf("row %3d B  %-34s : %7.1f cycles/tile (issue side), kernel %.1f us -> %.0f GB/s\n", rowb, names[mode], h / (double)tiles, ms * 1e3,
                 148.0 * tiles * 128 * rowb / (ms * 1e-3) / 1e9);
      }
    }
  }
  return 0;
}
