// Weight-gradient implicit GEMM on tcgen05 for stride-1 convolutions (3x3 / 1x1 / 2x2 phases):
//
//   dW[t][n][c] = sum over output pixels s of  G[s][n] * A[s + (dy_t, dx_t)][c]
//
// Same "slot space" as igemm_halo.cu: both the activation halo tile A (R rows) and the output
// gradient tile G (Rg rows) are staged ONCE per 128-slot tile in the no-swizzle interleaved layout
// (planes of 8 channels, consecutive slots 16 bytes apart, the two border slots of every row zero).
// In that layout a tile is directly an MN-major UMMA operand whose K dimension runs over slots:
//   A operand (M side) = G      : M = output channel n (planes at stride SBO), K = 16 consecutive slots
//   B operand (N side) = A halo : N = input channel c,  K = the same slots shifted by the tap offset
// so every tap is again only a descriptor start-address shift.  Each CTA owns one group of taps and a
// slice of the tiles; it accumulates over ALL of its tiles in TMEM (taps x C columns, FP32) and writes
// its partial sums once at the end -- there is no per-tile epilogue and no atomics.  sv_wgrad_reduce
// adds the per-CTA partials into the FP32 gradient arena.
//
// M is fixed at 128: with fewer than 128 output channels the upper accumulator rows read whatever
// follows the G planes in shared memory (zero-initialised, so finite) and are never stored.
//
// warp roles: 0..3 = final epilogue, 4..7 = producers (zero-filling 16-byte cp.async, completion
// signalled through cp.async.mbarrier.arrive), 8 = MMA issuer + TMEM allocator.
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"
#include "igemm.h"

namespace {

constexpr int WG_THREADS = 288;
constexpr int PRODUCERS = 128;
constexpr int BM = 128;
constexpr int MAX_STAGES = 6;
constexpr uint32_t SPIN_LIMIT = 1u << 28;
constexpr int PAD_SLOTS = 8;

struct WgParams {
  const bf16* A;
  const bf16* G;
  float* partial;            // [ctas][N][T*C]  (ctas = splits of sv_wgrad_reduce)
  int NB, H, W, C, N, T;     // N = output channels of THIS launch (a power-of-two slice of the layer's Ntot)
  int Ntot, n_off;           // row pitch of G / of the partial slices, first channel of the slice
  int P, R, Rg;
  int tiles_per_img, items;
  int groups, taps_per_group, ctas_per_group;
  int stages;
  int a_planes, a_planes_log2, a_chunks_per_row;
  int g_planes, g_planes_log2, g_chunks_per_row;
  uint32_t inv_P;
  uint32_t stage_bytes, a_off, a_plane_bytes, g_plane_bytes;
  uint32_t tmem_cols;
  int mma_m;                // 64 when N <= 64 (operand fetch 2 KB instead of 4 KB per MMA: ~1.6x faster), else 128
  int8_t dy[SV_MAX_TAPS];
  int8_t dx[SV_MAX_TAPS];
};

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > SPIN_LIMIT) {
      printf("wgrad_halo: mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// MN-major no-swizzle operand: 8 elements (16 B) contiguous along M/N, consecutive K (slots) 16 B apart,
// `lbo` between groups of 8 K-rows, `sbo` between 8-element chunks along M/N (= plane stride)
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

template <int TG>   // taps handled by one CTA
__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_halo_kernel(const WgParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t full_bar[MAX_STAGES], empty_bar[MAX_STAGES], done_bar;
  __shared__ uint32_t tmem_base_s;

  pdl_trigger();
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int group = blockIdx.x % p.groups;          // tap group of this CTA
  const int cta = blockIdx.x / p.groups;            // slice of the tiles
  const int t0 = group * p.taps_per_group;
  const int ntaps = min(p.taps_per_group, p.T - t0);

  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], PRODUCERS); mbar_init(&empty_bar[s], 1); }
    mbar_init(&done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // zero every staging byte once: border slots, slack and the phantom planes read by M = 128 must be finite
  {
    const uint32_t total16 = (p.stages * p.stage_bytes + (p.mma_m / 8 - p.g_planes) * p.g_plane_bytes + 4096) >> 4;
    for (uint32_t i = tid; i < total16; i += WG_THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async();
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  pdl_wait();      // the prologue above (196 KB of shared-memory zeroing, TMEM allocation) overlapped the previous kernel

  if (warp >= 4 && warp < 8) {
    // ===================================== producers ========================================
    const int ptid = tid - 128;
    const size_t a_row = (size_t)p.W * p.C, g_row = (size_t)p.W * p.Ntot;
    int stage = 0;
    uint32_t phase = 0;
    int img = cta / p.tiles_per_img, j = cta - img * p.tiles_per_img;
    const int img_step = p.ctas_per_group / p.tiles_per_img, j_step = p.ctas_per_group - img_step * p.tiles_per_img;
    for (int item = cta; item < p.items; item += p.ctas_per_group) {
      const int s0 = p.P + BM * j;
      const int yg0 = (int)(((uint32_t)s0 * p.inv_P) >> 16);     // first padded row of the G tile
      const int y0 = yg0 - 1;                                     // first padded row of the A tile
      mbar_wait(&empty_bar[stage], phase ^ 1);
      const uint32_t sbase = smem_u32(smem) + (uint32_t)stage * p.stage_bytes + PAD_SLOTS * 16;
      {   // output-gradient tile: rows yg0 .. yg0+Rg-1 (padded), i.e. image rows yg0-1 ..
        const bf16* img_base = p.G + (size_t)img * p.H * g_row + p.n_off;
        for (int ch = ptid; ch < p.g_chunks_per_row; ch += PRODUCERS) {
          const int slot = ch >> p.g_planes_log2, pl = ch & (p.g_planes - 1);
          uint32_t dst = sbase + (uint32_t)pl * p.g_plane_bytes + (uint32_t)(1 + slot) * 16;
          int yu = yg0 - 1;
          const bf16* src = img_base + (size_t)slot * p.Ntot + (size_t)pl * 8 + (ptrdiff_t)yu * (ptrdiff_t)g_row;
#pragma unroll 4
          for (int r = 0; r < p.Rg; ++r, ++yu, dst += (uint32_t)p.P * 16, src += g_row) {
            const bool ok = yu >= 0 && yu < p.H;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(ok ? src : p.G), "r"(ok ? 16 : 0));
          }
        }
      }
      {   // activation halo tile: rows y0 .. y0+R-1 (padded)
        const bf16* img_base = p.A + (size_t)img * p.H * a_row;
        for (int ch = ptid; ch < p.a_chunks_per_row; ch += PRODUCERS) {
          const int slot = ch >> p.a_planes_log2, pl = ch & (p.a_planes - 1);
          uint32_t dst = sbase + p.a_off + (uint32_t)pl * p.a_plane_bytes + (uint32_t)(1 + slot) * 16;
          int yu = y0 - 1;
          const bf16* src = img_base + (size_t)ch * 8 + (ptrdiff_t)yu * (ptrdiff_t)a_row;
#pragma unroll 4
          for (int r = 0; r < p.R; ++r, ++yu, dst += (uint32_t)p.P * 16, src += a_row) {
            const bool ok = yu >= 0 && yu < p.H;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(ok ? src : p.A), "r"(ok ? 16 : 0));
          }
        }
      }
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&full_bar[stage])) : "memory");
      img += img_step; j += j_step;
      if (j >= p.tiles_per_img) { j -= p.tiles_per_img; ++img; }
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
    }
    cp_async_wait<0>();
  } else if (warp == 8) {
    // ===================================== MMA issuer =======================================
    // idesc: FP32 accumulate, BF16 x BF16, A and B both MN-major, N = C, M = 64 / 128
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(p.C >> 3) << 17) |
                           ((uint32_t)(p.mma_m >> 4) << 24);
    const bool issuer = elect_one();            // one lane issues every MMA and commit of this CTA
    int stage = 0;
    uint32_t phase = 0;
    int j = cta % p.tiles_per_img;
    const int j_step = p.ctas_per_group % p.tiles_per_img;
    uint32_t first = 0;
    for (int item = cta; item < p.items; item += p.ctas_per_group) {
      const int s0 = p.P + BM * j;
      const int yg0 = (int)(((uint32_t)s0 * p.inv_P) >> 16);
      const int rel_g = s0 - yg0 * p.P;          // tile's first slot inside the G buffer
      const int rel_a = rel_g + p.P;             // ... inside the A buffer (starts one row earlier)
      j += j_step;
      if (j >= p.tiles_per_img) j -= p.tiles_per_img;
      mbar_wait(&full_bar[stage], phase);
      fence_proxy_async();
      tc_fence_after();
      const uint32_t sbase = smem_u32(smem) + (uint32_t)stage * p.stage_bytes + PAD_SLOTS * 16;
      const uint64_t g_desc0 = make_desc_mn(sbase + (uint32_t)(rel_g * 16), 128, p.g_plane_bytes);
      const uint64_t a_desc0 = make_desc_mn(sbase + p.a_off + (uint32_t)(rel_a * 16), 128, p.a_plane_bytes);
      if (issuer) {
#pragma unroll
        for (int t = 0; t < TG; ++t) {
          if (t < ntaps) {
            const uint64_t a_tap = a_desc0 + (uint64_t)(int64_t)((int)p.dy[t0 + t] * p.P + (int)p.dx[t0 + t]);
            const uint32_t d_tmem = tmem_base + (uint32_t)(t * p.C);
#pragma unroll
            for (int ks = 0; ks < BM / 16; ++ks)     // 16 slots (K) per MMA: +256 bytes per step
              tc_mma_bf16(d_tmem, g_desc0 + (uint64_t)(ks * 16), a_tap + (uint64_t)(ks * 16), idesc, first | (uint32_t)ks);
          }
        }
        tc_commit(&empty_bar[stage]);
      }
      __syncwarp();
      first = 1;
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
    }
    if (issuer) tc_commit(&done_bar);
    __syncwarp();
  } else if (warp < 4) {
    // ===================================== final epilogue ====================================
    mbar_wait(&done_bar, 0);
    tc_fence_after();
    // accumulator row = output channel.  M = 128: row r lives in TMEM lane r; M = 64: rows 16q .. 16q+15 live in lanes
    // 32q .. 32q+15 (MEASURED, tools/m64_probe.cu), i.e. the low half of every warp's lane quarter.
    const int n = (p.mma_m == 64) ? (lane < 16 ? warp * 16 + lane : p.N) : warp * 32 + lane;
    const int TC = p.T * p.C;
    float* dst = p.partial + ((size_t)cta * p.Ntot + p.n_off + n) * TC + (size_t)t0 * p.C;
    const int ncols = ntaps * p.C;
    const bool has_items = cta < p.items;
    for (int c0 = 0; c0 < ncols; c0 += 16) {
      uint32_t raw[16];
      tc_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, raw);
      tc_ld_wait();
      if (n < p.N) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float4 v = has_items ? make_float4(__uint_as_float(raw[4 * q]), __uint_as_float(raw[4 * q + 1]), __uint_as_float(raw[4 * q + 2]),
                                             __uint_as_float(raw[4 * q + 3]))
                               : make_float4(0.f, 0.f, 0.f, 0.f);
          *reinterpret_cast<float4*>(dst + c0 + 4 * q) = v;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  const int lim = sv_cta_limit();      // sv_set_cta_limit: SMs left to a concurrent collective (data parallel backward)
  return (lim > 0 && lim < n) ? lim : n;
}

int ilog2(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}

}  // namespace

bool wgrad_halo_supported(const WgradParams& p) {
  if (p.in_stride != 1 || p.H != p.OH || p.W != p.OW) return false;
  if (!((p.W == 32 || p.W == 16 || p.W == 8) && p.H >= 8 && p.H <= 32)) return false;
  if (!(p.C == 16 || p.C == 32 || p.C == 64 || p.C == 128)) return false;
  // output channels: a power of two up to 128, or a sum of such slices (WRN-28-10's 16 -> 160 stem convs: 128 + 32),
  // one launch per slice over the same activation tiles
  if (p.N % 16 != 0 || p.N > 256) return false;
  for (int t = 0; t < p.T; ++t)
    if (p.dy[t] < -1 || p.dy[t] > 1 || p.dx[t] < -1 || p.dx[t] > 1) return false;
  if ((reinterpret_cast<uintptr_t>(p.A) & 15) || (reinterpret_cast<uintptr_t>(p.Gr) & 15) || (reinterpret_cast<uintptr_t>(p.partial) & 15))
    return false;
  return true;
}

// number of per-CTA partial slices the kernel writes (the `splits` sv_wgrad_reduce must sum)
int wgrad_halo_splits(const WgradParams& p) {
  const int groups = ceil_div(p.T * p.C, 512);
  int ctas = sm_count() / groups;
  const int P = p.W + 2;
  const int items = p.NB * ceil_div(p.H * P, BM);
  if (ctas > items) ctas = items;
  if (ctas < 1) ctas = 1;
  return ctas;
}

static int wgrad_halo_slice(const WgradParams& p, int n_off, int n_slice, cudaStream_t st);

int wgrad_halo(const WgradParams& p, cudaStream_t st) {
  int n_off = 0;
  while (n_off < p.N) {
    int n_slice = 128;
    while (n_slice > p.N - n_off) n_slice >>= 1;
    const int rc = wgrad_halo_slice(p, n_off, n_slice, st);
    if (rc != SV_OK) return rc;
    n_off += n_slice;
  }
  return SV_OK;
}

static int wgrad_halo_slice(const WgradParams& p0, int n_off, int n_slice, cudaStream_t st) {
  WgradParams p = p0;
  p.N = n_slice;
  WgParams q;
  memset(&q, 0, sizeof(q));
  q.A = p.A; q.G = p.Gr; q.partial = p.partial;
  q.NB = p.NB; q.H = p.H; q.W = p.W; q.C = p.C; q.N = p.N; q.T = p.T;
  q.Ntot = p0.N; q.n_off = n_off;
  q.P = p.W + 2;
  int R = (4 * q.P + BM) / q.P;
  if (R > p.H + 2) R = p.H + 2;
  q.R = R;
  int Rg = (q.P - 1 + BM - 1) / q.P + 1;
  if (Rg > p.H + 2) Rg = p.H + 2;
  q.Rg = Rg;
  q.tiles_per_img = ceil_div(p.H * q.P, BM);
  q.items = p.NB * q.tiles_per_img;
  q.groups = ceil_div(p.T * p.C, 512);
  q.taps_per_group = ceil_div(p.T, q.groups);
  q.ctas_per_group = wgrad_halo_splits(p);
  if (q.ctas_per_group != p.splits) { sv_set_error("wgrad_halo: splits must be %d (got %d)", q.ctas_per_group, p.splits); return SV_ERR_ARG; }
  q.a_planes = p.C / 8; q.a_planes_log2 = ilog2(q.a_planes); q.a_chunks_per_row = p.W * q.a_planes;
  q.g_planes = p.N / 8; q.g_planes_log2 = ilog2(q.g_planes); q.g_chunks_per_row = p.W * q.g_planes;
  q.inv_P = (65536u + q.P - 1) / q.P;
  const uint32_t slack = (uint32_t)(3 * q.P + BM + 2 + PAD_SLOTS);
  q.g_plane_bytes = (uint32_t)((Rg * q.P > q.P + BM ? Rg * q.P : q.P + BM) + PAD_SLOTS) * 16;
  q.a_plane_bytes = (uint32_t)((R * q.P > (int)slack ? R * q.P : (int)slack) + PAD_SLOTS) * 16;
  q.a_off = (uint32_t)q.g_planes * q.g_plane_bytes;
  q.stage_bytes = (q.a_off + (uint32_t)q.a_planes * q.a_plane_bytes + PAD_SLOTS * 16 + 1023) & ~1023u;
  q.mma_m = p.N <= 64 ? 64 : 128;
  const size_t tail = (size_t)(q.mma_m / 8 - q.g_planes) * q.g_plane_bytes + 4096;   // phantom G planes of the last stage stay inside the allocation
  int stages = (int)((200 * 1024 - tail) / q.stage_bytes);
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages < 2) { sv_set_error("wgrad_halo: tile does not fit"); return SV_ERR_UNSUPPORTED; }
  q.stages = stages;
  const int cols = q.taps_per_group * p.C;
  q.tmem_cols = cols <= 32 ? 32 : cols <= 64 ? 64 : cols <= 128 ? 128 : cols <= 256 ? 256 : 512;
  memcpy(q.dy, p.dy, SV_MAX_TAPS);
  memcpy(q.dx, p.dx, SV_MAX_TAPS);
  const size_t smem = (size_t)stages * q.stage_bytes + tail + 1024;
  const int grid = q.ctas_per_group * q.groups;
#define SV_WG_CASE(TG)                                                                                     \
  if (q.taps_per_group == TG) {                                                                            \
    static bool configured = false;                                                                        \
    if (!configured) {                                                                                     \
      cudaFuncSetAttribute(wgrad_halo_kernel<TG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024); \
      configured = true;                                                                                   \
    }                                                                                                      \
    sv_launch_pdl(wgrad_halo_kernel<TG>, dim3(grid), dim3(WG_THREADS), smem, st, q);                                              \
    return sv_check_launch("wgrad_halo");                                                                  \
  }
  SV_WG_CASE(1) SV_WG_CASE(2) SV_WG_CASE(3) SV_WG_CASE(4) SV_WG_CASE(5) SV_WG_CASE(9)
#undef SV_WG_CASE
  sv_set_error("wgrad_halo: no instantiation for %d taps per CTA", q.taps_per_group);
  return SV_ERR_UNSUPPORTED;
}
