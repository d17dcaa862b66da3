// Halo-tile implicit-GEMM convolution for tcgen05: every input element is read from L2/HBM ONCE per
// CTA tile, not once per filter tap.
//
// The 3x3 (or 1x1 / 2x2-phase) convolution over an H x W image is evaluated in "slot space": the
// image is viewed with a one-pixel zero border, pitch P = W + 2, slot = y*P + x.  A tile is 128
// consecutive output slots.  Producer warps bring the R padded rows that the tile and its halo touch
// into shared memory in the NO-SWIZZLE K-major canonical layout ("interleaved": planes of 8 channels,
// consecutive slots 16 bytes apart).  In that layout the operand for tap (dy, dx) is the SAME buffer
// read from a start address shifted by (dy*P + dx) slots, so each tap is just a different UMMA shared
// memory descriptor: 9 taps x C/16 MMAs per tile, zero data movement between taps.  Out-of-image
// rows are zero-filled by the copies (that is the padding); the two border slots per row produce
// garbage accumulator rows that the epilogue simply does not store.
//
// The weights of all taps for this CTA's output-channel slice stay resident in shared memory (same
// interleaved layout) for the lifetime of the persistent CTA.
//
// Measured on B200: the TMA unit retires roughly one innermost box row per ~5 cycles whatever its
// length, so a tensor-map load with 16-byte rows (what the interleaved layout needs) is ~10x too slow.
// The A tile is therefore staged by three producer warps with zero-filling 16-byte cp.async (global
// reads fully coalesced: one image row = W*C*2 contiguous bytes), published to the tensor core with
// cp.async.mbarrier.arrive.noinc + a consumer-side fence.proxy.async; the resident weights arrive by 1-D bulk
// copies (cp.async.bulk).  (A shared-memory staged, bulk-store epilogue was measured ~35 % slower than the direct 32-byte
// stores in round 1 and has been removed.)
//
// warp roles (12 warps = 3 per scheduler, so every thread may use 168 registers): 0..3 and 8..11 = epilogue
// (TMEM -> registers -> +bias/+residual -> bf16 NHWC 32-byte stores, BatchNorm sum / sum^2), 4..6 = A-tile
// producers, 7 = MMA issuer + TMEM allocator + weight loader.
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"
#include "igemm.h"

// optional timeline instrumentation (SHOTVAE_HALO_TRACE=1): CTA 0 stamps clock64 per tile and role
__device__ long long g_halo_trace[4][64][2];
__device__ long long g_halo_mma_trace[3][80];
__device__ long long g_halo_marks[8];

namespace {

constexpr int HL_THREADS = 384;   // 12 warps (3 per scheduler -> 168 registers): 0-3 + 8-11 epilogue, 4-6 producers, 7 MMA
constexpr int MMA_WARP = 7;
constexpr int PRODUCERS = 96;
constexpr int LAG = 4;          // cp.async groups in flight per producer thread before a stage is published
constexpr int BM = 128;
constexpr int MAX_GROUPS = 4;
constexpr int MAX_STAGES = 8;
constexpr int MAX_ACC = 8;      // TMEM accumulator ring (the commit -> epilogue signal lags the MMAs by 1-4k cycles under store traffic)
constexpr uint32_t SPIN_LIMIT = 1u << 28;
#ifdef SV_HALO_TRACE
constexpr bool kTrace = true;    // debug build: clock64 timeline of CTA 0 + role ablation (tools/halo_trace.py, tools/halo_lag.py)
#else
constexpr bool kTrace = false;   // production build: the instrumentation is compiled out of every role loop
#endif
constexpr int PAD_SLOTS = 8;   // slack slots (128 B) in front of the A tile: tap (-1,-1) of slot 0 reads 1 slot before it

struct HaloParams {
  bf16* out;
  const bf16* res;
  const float* bias;
  float* stats;
  int NB, H, W, C, N, T;
  int P, R;                 // slot pitch (W+2), rows per A tile
  int tiles_per_img, items; // items = NB * tiles_per_img
  int BN, n_tiles, ctas_per_nt;
  int OHf, OWf, out_stride, out_off_y, out_off_x;
  int group_images, groups;
  int stages;
  int n_acc;                // accumulators in the TMEM ring (n_acc * BN columns, power of two)
  int one_commit;           // the epilogue releases the A stage (one tcgen05.commit per tile instead of two)
  int planes, planes_log2, chunks_per_row;
  int trace, ablate, lag;
  uint32_t inv_P;           // ceil(65536 / P): s / P == (s * inv_P) >> 16 for the slot range used here
  const bf16* A;
  const bf16* Wp;           // weights packed [T][C/8][N][8]
  // fused BatchNorm-backward statistics (input-gradient launches, see sv_igemm_args): y + coefficients [G][N]
  const bf16* bn_y;
  const float* bn_scale;
  const float* bn_shift;
  const float* bn_mean;
  const float* bn_var;
  float bn_slope, bn_eps;
  uint32_t a_stage_bytes, a_plane_bytes, b_bytes, b_tap_bytes, b_plane_bytes;
  int8_t dy[SV_MAX_TAPS];
  int8_t dx[SV_MAX_TAPS];
};

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
      printf("igemm_halo: mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
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

// no-swizzle K-major canonical layout: 8-row core matrices of 128 contiguous bytes; `sbo` between
// consecutive 8-row groups (M/N direction), `lbo` between the two 8-channel planes of one K=16 step
__device__ __forceinline__ uint64_t make_desc_nosw(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (sm_100); layout_type (bits 61-63) = 0: no swizzle
  return d;
}

// BNB_: the epilogue accumulates BatchNorm-backward statistics (input-gradient launches) instead of sum / sum of squares;
// a template parameter so that the forward instantiations keep their register allocation
template <int T_, int KC_, bool BNB_>
__global__ void __launch_bounds__(HL_THREADS, 1)
igemm_halo_kernel(const HaloParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t full_bar[MAX_STAGES], empty_bar[MAX_STAGES], tmem_full[MAX_ACC], tmem_empty[MAX_ACC], b_bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_stat[MAX_GROUPS][2][128];
  __shared__ __align__(16) float s_coef[BNB_ ? MAX_GROUPS : 1][4][128];     // {scale, shift, rstd, -mean*rstd} of this CTA's channel slice

  const long long t_entry = clock64();
  pdl_trigger();
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // warp-uniform by construction
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_b = smem;                         // resident weights [T][C/8][BN][8]
  uint8_t* smem_a = smem + ((p.b_bytes + 1023) & ~1023u);
  const int nt = blockIdx.x % p.n_tiles;
  const int cta = blockIdx.x / p.n_tiles;
  const uint32_t tmem_cols = (uint32_t)max(32, p.n_acc * p.BN);      // host guarantees a power of two <= 512

  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], PRODUCERS); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < p.n_acc; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], 8); }
    mbar_init(&b_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < MAX_GROUPS * 2 * 128; i += HL_THREADS) (&s_stat[0][0][0])[i] = 0.f;
  // the two border slots of every buffered row are image padding: zero them once, producers never touch them
  {
    const int per_stage = p.planes * p.R * 2;
    for (int i = tid; i < p.stages * per_stage; i += HL_THREADS) {
      const int st = i / per_stage, rem = i - st * per_stage;
      const int pl = rem / (p.R * 2), rr = rem - pl * (p.R * 2);
      const int r = rr >> 1, x = (rr & 1) ? (p.P - 1) : 0;
      uint8_t* dst = smem_a + (size_t)st * p.a_stage_bytes + PAD_SLOTS * 16 + (size_t)pl * p.a_plane_bytes + (size_t)(r * p.P + x) * 16;
      *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
    }
    fence_proxy_async();
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  // everything above touched only shared / tensor memory and overlapped the previous kernel's tail
  pdl_wait();
  if (kTrace && p.trace && blockIdx.x == 0 && tid == 0) { g_halo_marks[0] = t_entry; g_halo_marks[1] = clock64(); }

  if (warp >= 4 && warp < MMA_WARP) {
    // ===================================== A-tile producers (4 warps) =======================
    const int ptid = tid - 128;
    const bf16* __restrict__ Ag = p.A;
    const size_t row_elems = (size_t)p.W * p.C;
    int stage = 0, it = 0;
    uint32_t phase = 0;
    int img = cta / p.tiles_per_img, j = cta - img * p.tiles_per_img;
    const int img_step = p.ctas_per_nt / p.tiles_per_img, j_step = p.ctas_per_nt - img_step * p.tiles_per_img;
    for (int item = cta; item < p.items; item += p.ctas_per_nt, ++it) {
      const int s0 = p.P + BM * j;              // first output slot of the tile (padded-grid slot index)
      const int y0 = (int)(((uint32_t)s0 * p.inv_P) >> 16) - 1;   // first padded row held in shared memory
      mbar_wait(&empty_bar[stage], phase ^ 1);
      if (kTrace && p.trace && blockIdx.x == 0 && ptid == 0 && it < 64) g_halo_trace[0][it][0] = clock64();
      const uint32_t abase = smem_u32(smem_a) + (uint32_t)stage * p.a_stage_bytes + PAD_SLOTS * 16;
      const bf16* img_base = Ag + (size_t)img * p.H * row_elems;
      for (int ch = ptid; ch < p.chunks_per_row; ch += PRODUCERS) {
        const int slot = ch >> p.planes_log2, pl = ch & (p.planes - 1);
        uint32_t dst = abase + (uint32_t)pl * p.a_plane_bytes + (uint32_t)(1 + slot) * 16;
        const bf16* src = img_base + (size_t)ch * 8 + (ptrdiff_t)(y0 - 1) * (ptrdiff_t)row_elems;
        int yu = y0 - 1;
#pragma unroll 4
        for (int r = 0; r < p.R; ++r, ++yu, dst += (uint32_t)p.P * 16, src += row_elems) {
          const bool ok = yu >= 0 && yu < p.H;
          const int sz = (ok && !(kTrace && (p.ablate & 1))) ? 16 : 0;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(ok ? src : Ag), "r"(sz));
        }
      }
      // completion of this thread's copies arrives on the stage barrier asynchronously: the producer
      // never waits on memory (a wait_group + fence.proxy.async here compiles to MEMBAR.ALL.CTA, which
      // drains every copy in flight and serialises the pipeline on DRAM latency)
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&full_bar[stage])) : "memory");
      if (kTrace && p.trace && blockIdx.x == 0 && ptid == 0 && it < 64) g_halo_trace[0][it][1] = clock64();
      img += img_step; j += j_step;
      if (j >= p.tiles_per_img) { j -= p.tiles_per_img; ++img; }
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
    }
    cp_async_wait<0>();      // do not exit with copies in flight
  } else if (warp == MMA_WARP) {
    // ===================================== MMA issuer =======================================
    // The whole warp runs this loop so that every address / descriptor computation is warp-uniform
    // (uniform datapath, no per-thread register -> uniform register moves in front of each
    // tcgen05.mma); only the issue itself is predicated on one lane.
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    constexpr int kc_n = KC_;
    const bool leader = lane == 0;
    // resident weights of this CTA's channel slice: one 1-D bulk copy per (tap, 8-channel plane)
    if (leader) mbar_expect_tx(&b_bar, p.b_bytes);
    __syncwarp();
    for (int i = lane; i < p.T * p.planes; i += 32)     // all lanes issue: the copies are latency bound
      bulk_load(smem_b + (size_t)i * p.b_plane_bytes, p.Wp + ((size_t)i * p.N + (size_t)nt * p.BN) * 8, p.b_plane_bytes, &b_bar);
    __syncwarp();
    const uint64_t b_desc0 = make_desc_nosw(smem_u32(smem_b), p.b_plane_bytes, 128);
    const uint32_t a_plane16 = p.a_plane_bytes >> 4, b_plane16 = p.b_plane_bytes >> 4, b_tap16 = p.b_tap_bytes >> 4;
    const uint32_t a_first = smem_u32(smem_a) + PAD_SLOTS * 16;
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    int j = cta % p.tiles_per_img;
    const int j_step = p.ctas_per_nt % p.tiles_per_img;
    mbar_wait(&b_bar, 0);
    if (kTrace && p.trace && blockIdx.x == 0 && leader) g_halo_marks[2] = clock64();
    int ti = 0;
    for (int item = cta; item < p.items; item += p.ctas_per_nt, ++ti) {
      const int s0 = p.P + BM * j;
      const int y0 = (int)(((uint32_t)s0 * p.inv_P) >> 16) - 1;
      const int rel0 = s0 - y0 * p.P;           // tile's first output slot relative to the buffer
      j += j_step;
      if (j >= p.tiles_per_img) j -= p.tiles_per_img;
      if (kTrace && p.trace && blockIdx.x == 0 && leader && ti < 64) g_halo_trace[2][ti][0] = clock64();
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      if (kTrace && p.trace && blockIdx.x == 0 && leader && ti < 64) g_halo_trace[2][ti][1] = clock64();
      mbar_wait(&full_bar[stage], phase);
      fence_proxy_async();     // generic-proxy (cp.async) writes -> visible to the tensor core's async-proxy reads
      tc_fence_after();
      if (kTrace && p.trace && blockIdx.x == 0 && leader && ti < 64) g_halo_trace[1][ti][0] = clock64();
      const uint64_t a_desc0 = make_desc_nosw(a_first + (uint32_t)stage * p.a_stage_bytes + (uint32_t)(rel0 * 16), p.a_plane_bytes, 128);
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.BN);
      const bool issuer = elect_one();
      if (issuer && !(kTrace && (p.ablate & 4))) {
#pragma unroll
        for (int t = 0; t < T_; ++t) {
          // tap (dy, dx) of the A operand == start address shifted by (dy*P + dx) slots of 16 bytes
          const uint64_t a_tap = a_desc0 + (uint64_t)(int64_t)((int)p.dy[t] * p.P + (int)p.dx[t]);
          const uint64_t b_tap = b_desc0 + (uint64_t)((uint32_t)t * b_tap16);
#pragma unroll
          for (int kc = 0; kc < kc_n; ++kc)
            tc_mma_bf16(d_tmem, a_tap + (uint64_t)((uint32_t)(2 * kc) * a_plane16), b_tap + (uint64_t)((uint32_t)(2 * kc) * b_plane16), idesc,
                        (t | kc) ? 1u : 0u);
        }
      }
      if (issuer) {
        if (!p.one_commit) tc_commit(&empty_bar[stage]);
        if (kTrace && p.trace && blockIdx.x == 0 && ti >= 2 && ti < 5) g_halo_mma_trace[ti - 2][78] = clock64();
        tc_commit(&tmem_full[acc]);
        if (kTrace && p.trace && blockIdx.x == 0 && ti >= 2 && ti < 5) g_halo_mma_trace[ti - 2][79] = clock64();
      }
      __syncwarp();
      if (kTrace && p.trace && blockIdx.x == 0 && leader && ti < 64) g_halo_trace[1][ti][1] = clock64();
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
      if (++acc == p.n_acc) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ===================================== epilogue (8 warps) ================================
    // warps 0-3 take the low half of the tile's channels, warps 8-11 the high half; warp w reads the
    // TMEM lane quarter w % 4 (hardware restriction), i.e. 32 of the tile's 128 output slots.
    const int q = warp & 3, half = warp >> 3;
    constexpr bool bnb = BNB_;
    if (bnb) {
      const int et0 = (warp < 4) ? tid : tid - 128;      // 0..255 over the eight epilogue warps
      for (int i = et0; i < p.groups * p.BN; i += 256) {
        const int g = i / p.BN, c = i - g * p.BN;
        const size_t k = (size_t)g * p.N + nt * p.BN + c;
        const float rs = rsqrtf(p.bn_var[k] + p.bn_eps);
        s_coef[g][0][c] = p.bn_scale[k]; s_coef[g][1][c] = p.bn_shift[k]; s_coef[g][2][c] = rs; s_coef[g][3][c] = -p.bn_mean[k] * rs;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    const int nchunk = (p.BN >= 32) ? (p.BN >> 5) : (half == 0 ? 1 : 0);   // 16-column chunks owned by this warp
    const int cbase = half * (p.BN >> 1);                                  // first column of this warp (BN >= 32)
    const bool reg_stats = p.BN <= 64;
    int acc = 0;
    uint32_t acc_phase = 0;
    int img = cta / p.tiles_per_img, j = cta - img * p.tiles_per_img;
    const int img_step = p.ctas_per_nt / p.tiles_per_img, j_step = p.ctas_per_nt - img_step * p.tiles_per_img;
    // BatchNorm statistics.  BN <= 64: every thread keeps per-channel partial sums of ITS rows in
    // registers across all tiles (2 FMAs per element, no shuffles in the loop); the cross-row reduction
    // runs once per pass group.  BN = 128: per-tile shuffle transpose-reduce.
    float a1[2][16], a2[2][16];
#pragma unroll
    for (int k = 0; k < 2; ++k)
#pragma unroll
      for (int jj = 0; jj < 16; ++jj) { a1[k][jj] = 0.f; a2[k][jj] = 0.f; }
    int cur_g = -1;
    auto flush_stats = [&](int g) {
      if (g < 0 || !reg_stats) return;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        if (k < nchunk) {
          const float s1 = colsum16(a1[k], lane), s2 = colsum16(a2[k], lane);
          if (lane < 16) {
            atomicAdd(&s_stat[g][0][cbase + k * 16 + lane], s1);
            atomicAdd(&s_stat[g][1][cbase + k * 16 + lane], s2);
          }
        }
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) { a1[k][jj] = 0.f; a2[k][jj] = 0.f; }
      }
    };
    int ti = 0, estage = 0;
    // The side operand of the epilogue (residual, or the BatchNorm input of the fused backward statistics) is loaded into
    // registers before the accumulator wait.  MEASURED alternatives that did not pay (profiles/r02_kernel_findings.md):
    // prefetch.global.L1 three tiles ahead (no change) and staging it in the tile's shared-memory stage with one bulk copy
    // per tile (64 -> 70 us for the fused-statistics input gradient): the epilogue is bound by LSU work, not by this latency.
    const bf16* side = bnb ? p.bn_y : p.res;
    for (int item = cta; item < p.items; item += p.ctas_per_nt, ++ti) {
      const int s = p.P + BM * j + q * 32 + lane;      // this thread's output slot
      const int y = (int)(((uint32_t)s * p.inv_P) >> 16), x = s - y * p.P;   // padded coordinates
      const bool row_ok = (x >= 1) && (x <= p.W) && (y <= p.H);
      const int oh = y - 1, ow = x - 1;
      const uint32_t pix = (uint32_t)((img * p.OHf + oh * p.out_stride + p.out_off_y) * p.OWf + ow * p.out_stride + p.out_off_x);
      const uint32_t obase = pix * (uint32_t)p.N + (uint32_t)(nt * p.BN + cbase);   // element offset (< 2^31 by construction)
      const int g = img / p.group_images;
      img += img_step; j += j_step;
      if (j >= p.tiles_per_img) { j -= p.tiles_per_img; ++img; }
      if (p.stats != nullptr && g != cur_g) { flush_stats(cur_g); cur_g = g; }
      // residual rows are fetched BEFORE waiting for the accumulator, so their latency is hidden
      bf16x8 rres[8];
      const bool use_res = !bnb && p.res != nullptr && row_ok;
      const bool use_y = bnb && row_ok;                // BatchNorm input at the same pixels (same layout as the output)
      if ((use_res || use_y) && !(kTrace && (p.ablate & 16))) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (k < nchunk) ld_global_32B(side + obase + k * 16, rres[2 * k], rres[2 * k + 1]);
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      if (p.one_commit) {
        // the tile's MMAs are complete: its A stage can be refilled (saves the MMA warp a second tcgen05.commit)
        if (warp == 0 && lane == 0) mbar_arrive(&empty_bar[estage]);
        if (++estage == p.stages) estage = 0;
      }
      if (kTrace && p.trace && blockIdx.x == 0 && tid == 0 && ti < 64) g_halo_trace[3][ti][0] = clock64();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.BN + cbase);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (k < nchunk) {
          const int c0 = k * 16;
          uint32_t raw[16];
          tc_ld16(taddr + c0, raw);
          tc_ld_wait();
          float v[16];
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) v[jj] = __uint_as_float(raw[jj]);
          if (p.bias != nullptr) {
            const int n0 = nt * p.BN + cbase + c0;
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) v[jj] += p.bias[n0 + jj];
          }
          if (row_ok && !(kTrace && (p.ablate & 2))) {
            if (use_res) {
              float rr[16];
              unpack8(rres[2 * k], rr);
              unpack8(rres[2 * k + 1], rr + 8);
#pragma unroll
              for (int jj = 0; jj < 16; ++jj) v[jj] += rr[jj];
            }
            const bf16x8 o0 = pack8(v), o1 = pack8(v + 8);
            st_global_32B(p.out + obase + c0, o0, o1);
            if (p.stats != nullptr) { unpack8(o0, v); unpack8(o1, v + 8); }
          } else {
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) v[jj] = 0.f;
          }
          if (p.stats != nullptr) {
            float sq[16];
            if (bnb) {
              // v = stored output gradient w.r.t. the activated tensor -> dz (through the activation) and dz * x_hat
              if (use_y && !(kTrace && (p.ablate & 8))) {
                float yv[16];
                unpack8(rres[2 * k], yv);
                unpack8(rres[2 * k + 1], yv + 8);
                const float4* cf = reinterpret_cast<const float4*>(&s_coef[g][0][cbase + c0]);
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4)
                  bn_bwd_terms4(yv + 4 * j4, cf[j4], cf[32 + j4], cf[64 + j4], cf[96 + j4], p.bn_slope, v + 4 * j4, sq + 4 * j4);
              } else {
#pragma unroll
                for (int jj = 0; jj < 16; ++jj) sq[jj] = 0.f;
              }
            }
            if (reg_stats) {
              if (k < 2) {
#pragma unroll
                for (int jj = 0; jj < 16; ++jj) {
                  a1[k][jj] += v[jj];
                  a2[k][jj] = bnb ? a2[k][jj] + sq[jj] : fmaf(v[jj], v[jj], a2[k][jj]);
                }
              }
            } else {
              if (!bnb) {
#pragma unroll
                for (int jj = 0; jj < 16; ++jj) sq[jj] = v[jj] * v[jj];
              }
              const float s1 = colsum16(v, lane), s2 = colsum16(sq, lane);
              if (lane < 16) {
                atomicAdd(&s_stat[g][0][cbase + c0 + lane], s1);
                atomicAdd(&s_stat[g][1][cbase + c0 + lane], s2);
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (kTrace && p.trace && blockIdx.x == 0 && tid == 0 && ti < 64) g_halo_trace[3][ti][1] = clock64();
      if (++acc == p.n_acc) { acc = 0; acc_phase ^= 1; }
    }
    if (p.stats != nullptr) {
      flush_stats(cur_g);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const int et = (warp < 4) ? tid : tid - 128;       // 0..255 over the eight epilogue warps
      for (int i = et; i < p.groups * 2 * p.BN; i += 256) {
        const int g = i / (2 * p.BN), rem = i - g * 2 * p.BN;
        const int st = rem / p.BN, c = rem - st * p.BN;
        const float val = s_stat[g][st][c];
        // forward statistics are [G][2][N]; the BatchNorm-backward pair is [2][G][N] (dbeta block, then dgamma block)
        const size_t slot = BNB_ ? (size_t)(st * p.groups + g) : (size_t)(g * 2 + st);
        if (val != 0.f) atomicAdd(&p.stats[slot * p.N + nt * p.BN + c], val);
      }
    }
  }
  tc_fence_before();
  if (kTrace && p.trace && blockIdx.x == 0 && tid == 0) g_halo_marks[3] = clock64();
  __syncthreads();
  if (kTrace && p.trace && blockIdx.x == 0 && tid == 0) g_halo_marks[4] = clock64();
  if (warp == MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
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

int pick_bn(const IgemmParams& p) {
  for (int bn : {128, 64, 32, 16})
    if (p.N % bn == 0 && (size_t)p.T * bn * p.C * 2 <= 100 * 1024) return bn;
  return 0;
}

}  // namespace

bool igemm_fprop_halo_supported(const IgemmParams& p) {
  if (p.w_layout != 1) return false;
  if (p.in_stride != 1 || p.H != p.OH || p.W != p.OW) return false;
  if (!((p.W == 32 || p.W == 16 || p.W == 8) && p.H >= 8 && p.H <= 32)) return false;
  if (!(p.C == 16 || p.C == 32 || p.C == 64 || p.C == 128) || p.N % 16 != 0) return false;
  if (p.out == nullptr || p.outf != nullptr) return false;
  for (int t = 0; t < p.T; ++t)
    if (p.dy[t] < -1 || p.dy[t] > 1 || p.dx[t] < -1 || p.dx[t] > 1) return false;
  if (p.stats != nullptr && p.NB / p.group_images > MAX_GROUPS) return false;
  if (p.bn_y != nullptr && (p.out_stride != 1 || p.OHf != p.OH || p.OWf != p.OW || (reinterpret_cast<uintptr_t>(p.bn_y) & 31))) return false;
  if (pick_bn(p) == 0) return false;
  if (!(p.T == 1 || p.T == 2 || p.T == 4 || p.T == 9)) return false;
  if ((reinterpret_cast<uintptr_t>(p.A) & 15) || (reinterpret_cast<uintptr_t>(p.Wt) & 15)) return false;
  if ((reinterpret_cast<uintptr_t>(p.out) & 31) || (reinterpret_cast<uintptr_t>(p.res) & 31)) return false;   // 32-byte epilogue accesses
  return true;
}

int igemm_fprop_halo(const IgemmParams& p, cudaStream_t st) {
  HaloParams q;
  memset(&q, 0, sizeof(q));
  q.out = p.out; q.res = p.res; q.bias = p.bias; q.stats = p.stats;
  q.NB = p.NB; q.H = p.H; q.W = p.W; q.C = p.C; q.N = p.N; q.T = p.T;
  q.P = p.W + 2;
  // rows needed: from the row above the tile's first slot to the row below its last slot
  int R = (3 * q.P + BM + 1 + q.P - 1) / q.P;
  if (R > p.H + 2) R = p.H + 2;
  q.R = R;
  q.tiles_per_img = ceil_div(p.H * q.P, BM);
  q.items = p.NB * q.tiles_per_img;
  q.BN = pick_bn(p);
  q.n_tiles = p.N / q.BN;
  q.OHf = p.OHf; q.OWf = p.OWf; q.out_stride = p.out_stride; q.out_off_y = p.out_off_y; q.out_off_x = p.out_off_x;
  q.group_images = p.group_images;
  q.groups = p.NB / p.group_images;
  q.a_plane_bytes = (uint32_t)R * q.P * 16;
  // the last tap of the last row may read up to (3P + 129) slots past the buffer start
  const uint32_t a_slots = (uint32_t)(3 * q.P + BM + 2 + PAD_SLOTS);
  const uint32_t a_need = (uint32_t)(p.C / 8 - 1) * q.a_plane_bytes + (a_slots > (uint32_t)(R * q.P + PAD_SLOTS) ? a_slots : (uint32_t)(R * q.P + PAD_SLOTS)) * 16;
  q.a_stage_bytes = (a_need + 1023) & ~1023u;
  q.b_plane_bytes = (uint32_t)q.BN * 16;
  q.b_tap_bytes = (uint32_t)(p.C / 8) * q.b_plane_bytes;
  q.b_bytes = (uint32_t)p.T * q.b_tap_bytes;
  const size_t b_alloc = (q.b_bytes + 1023) & ~(size_t)1023;
  int stages = (int)((196 * 1024 - b_alloc) / q.a_stage_bytes);
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages < 2) { sv_set_error("igemm_halo: tile does not fit"); return SV_ERR_UNSUPPORTED; }
  q.stages = stages;
  q.lag = stages - 1 < LAG ? stages - 1 : LAG;
  {
    static int nacc_env = -1, onec_env = -1;
    if (nacc_env < 0) { const char* e = getenv("SHOTVAE_HALO_NACC"); nacc_env = e ? atoi(e) : 0; }
    if (onec_env < 0) { const char* e = getenv("SHOTVAE_HALO_ONECOMMIT"); onec_env = e ? atoi(e) : 1; }
    int n_acc = 512 / q.BN;                   // whole TMEM: nothing else runs on the SM next to a 1-CTA/SM persistent kernel
    if (n_acc > MAX_ACC) n_acc = MAX_ACC;
    if (nacc_env > 0 && nacc_env < n_acc) n_acc = nacc_env;
    while (n_acc & (n_acc - 1)) --n_acc;      // power of two (tcgen05.alloc column counts)
    q.n_acc = n_acc;
    q.one_commit = onec_env;
  }
  int ctas = sm_count() / q.n_tiles;
  if (ctas < 1) ctas = 1;
  if (ctas > q.items) ctas = q.items;
  q.ctas_per_nt = ctas;
  memcpy(q.dy, p.dy, SV_MAX_TAPS);
  memcpy(q.dx, p.dx, SV_MAX_TAPS);

  q.A = p.A;
  q.Wp = p.Wt;
  q.bn_y = p.bn_y; q.bn_scale = p.bn_scale; q.bn_shift = p.bn_shift; q.bn_mean = p.bn_mean; q.bn_var = p.bn_var;
  q.bn_slope = p.bn_slope; q.bn_eps = p.bn_eps;
  q.planes = p.C / 8;
  q.planes_log2 = 0;
  while ((1 << q.planes_log2) < q.planes) ++q.planes_log2;
  q.chunks_per_row = p.W * q.planes;
  q.inv_P = (65536u + q.P - 1) / q.P;
  {
    static int trace = -1;
    if (trace < 0) { const char* e = getenv("SHOTVAE_HALO_TRACE"); trace = (e && e[0] == '1') ? 1 : 0; }
    q.trace = trace;
    static int ablate = -1;
    if (ablate < 0) { const char* e = getenv("SHOTVAE_HALO_ABLATE"); ablate = e ? atoi(e) : 0; }
    q.ablate = ablate;
  }
  const size_t smem = b_alloc + (size_t)stages * q.a_stage_bytes + 1024;
  const int grid = q.ctas_per_nt * q.n_tiles;
  const int kc = p.C / 16;
#define SV_HALO_LAUNCH(TT, KK, BB)                                                                                   \
  {                                                                                                                 \
    static bool configured = false;                                                                                 \
    if (!configured) {                                                                                              \
      cudaFuncSetAttribute(igemm_halo_kernel<TT, KK, BB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);  \
      configured = true;                                                                                            \
    }                                                                                                               \
    sv_launch_pdl(igemm_halo_kernel<TT, KK, BB>, dim3(grid), dim3(HL_THREADS), smem, st, q);                         \
    return sv_check_launch("igemm_halo");                                                                           \
  }
#define SV_HALO_CASE(TT, KK)                                                                                         \
  if (p.T == TT && kc == KK) {                                                                                       \
    if (p.bn_y != nullptr) SV_HALO_LAUNCH(TT, KK, true) else SV_HALO_LAUNCH(TT, KK, false)                            \
  }
#define SV_HALO_TAPS(TT) SV_HALO_CASE(TT, 1) SV_HALO_CASE(TT, 2) SV_HALO_CASE(TT, 4) SV_HALO_CASE(TT, 8)
  SV_HALO_TAPS(1) SV_HALO_TAPS(2) SV_HALO_TAPS(4) SV_HALO_TAPS(9)
#undef SV_HALO_TAPS
#undef SV_HALO_CASE
#undef SV_HALO_LAUNCH
  sv_set_error("igemm_halo: no instantiation for T=%d, C=%d", p.T, p.C);
  return SV_ERR_UNSUPPORTED;
}

extern "C" int sv_debug_halo_trace(long long* host_out /* [4][64][2] then [3][80] */) {
  if (cudaMemcpyFromSymbol(host_out, g_halo_trace, sizeof(long long) * 4 * 64 * 2) != cudaSuccess) return -2;
  if (cudaMemcpyFromSymbol(host_out + 4 * 64 * 2, g_halo_mma_trace, sizeof(long long) * 3 * 80) != cudaSuccess) return -2;
  return cudaMemcpyFromSymbol(host_out + 4 * 64 * 2 + 240, g_halo_marks, sizeof(long long) * 8) == cudaSuccess ? 0 : -2;
}
