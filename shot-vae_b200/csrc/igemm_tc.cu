// Implicit-GEMM convolution on the Blackwell tensor path: tcgen05.mma (bf16 x bf16 -> fp32 in TMEM),
// operands staged by TMA, persistent warp-specialised CTAs.
//
//   out[m, n] = sum_t sum_c A[pixel(m) + (dy_t, dx_t), c] * W[t][n][c]     (stride-1 gathers)
//
// * A tile  : 128 output pixels = a box (W_t x H_t x N_t images) of the NHWC activation tensor.  For
//             each filter tap the producer issues ONE 4-D TMA load of that box shifted by (dy, dx);
//             out-of-image elements are zero-filled by the TMA unit, which IS the convolution padding.
// * B tile  : [BN output channels][KB input channels] slice of the packed weights [T][N][C] (3-D TMA).
// * both tiles land in shared memory in the K-major 64B/128B-swizzled layout tcgen05.mma consumes;
//   one elected thread issues 128 x BN x 16 MMAs per 16 input channels; accumulators live in TMEM,
//   double buffered so the epilogue of tile i overlaps the main loop of tile i+1.
// * epilogue (4 warps): tcgen05.ld -> +bias, +residual -> bf16 NHWC store (any output stride: serves
//   transposed-conv / strided-dgrad phases) and per-channel sum / sum^2 for the next BatchNorm.
//
// warp roles: 1 = MMA issuer, 4..7 = epilogue, 0 / 2 / 3 / 8 / 9 / 10 = TMA producers (2 also allocates TMEM).
//
// MEASURED (tools/tma_probe.cu, profiles/r02_kernel_findings.md): tensor-map loads issued by ONE thread do not
// overlap -- each costs a full ~700-cycle round trip whatever the box size (4 KB or 32 KB) and however many
// stages are free, i.e. 23 B/cycle/SM for 16 KB boxes -- while loads issued from different warps run in parallel
// (2 issuers: 47, 4 issuers: 72 B/cycle/SM = the L2 fabric limit).  The A and B loads of consecutive k-blocks are
// therefore dealt round-robin to six producer warps; a stage's full barrier counts two arrivals (A and B).
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"
#include "igemm.h"

namespace {

constexpr int TC_THREADS = 384;
constexpr int N_PRODUCERS = 6;
constexpr int BM = 128;
constexpr int MAX_GROUPS = 4;
constexpr uint32_t SPIN_LIMIT = 1u << 28;

struct TcParams {
  bf16* out;
  const bf16* res;
  const float* bias;
  float* stats;
  int M, N, T, KC;           // rows, channels out, taps, k-blocks per tap
  int OH, OW, OHf, OWf, out_stride, out_off_y, out_off_x;
  int Wt, Ht, Nt;            // tile box
  int tiles_h;               // OH / Ht
  int m_tiles, n_tiles, BN, KB;
  int rows_per_group, groups;
  int stages;
  int pair_tiles;            // CL2: ceil(m_tiles / 2) * n_tiles tile pairs (adjacent m-tiles, same n-tile) shared by a CTA pair
  int in_stride;             // 1, or 2: the A boxes sample every second pixel (tensor-map element strides)
  int swizzle_bytes;         // 64 or 128
  int8_t dy[SV_MAX_TAPS];
  int8_t dx[SV_MAX_TAPS];
  // fused BatchNorm-backward statistics (input-gradient launches, see sv_igemm_args)
  const bf16* bn_y;
  const float* bn_scale;
  const float* bn_shift;
  const float* bn_mean;
  const float* bn_var;
  float bn_slope, bn_eps;
};

// ------------------------------------------------------------------------------------ PTX wrappers
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
// bounded spin: a protocol bug traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > SPIN_LIMIT) {
      printf("igemm_tc: mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// multicast variants (2-CTA cluster): the box lands at the same CTA-relative offset in every CTA of `mask`, and completes
// transaction bytes on the mbarrier at the same offset in each of them
__device__ __forceinline__ void tma_load_3d_mc(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tc_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask)
               : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
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

// K-major, swizzled operand tile: rows `row_bytes` apart, 8-row groups `8*row_bytes` apart
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t row_bytes) {
  const uint64_t layout = row_bytes == 128 ? 2ull : (row_bytes == 64 ? 4ull : 6ull);
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);            // start address
  d |= (uint64_t)1 << 16;                             // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)((8 * row_bytes) >> 4) << 32;        // stride byte offset: next 8-row group
  d |= (uint64_t)1 << 46;                             // descriptor version (sm_100)
  d |= layout << 61;
  return d;
}

// CL2: launched as clusters of two CTAs that work on adjacent m-tiles of the SAME channel tile in lockstep.  The B (weight) tile
// of a k-block is then identical for both: each CTA fetches half of its rows from L2 and multicasts them into both CTAs' shared
// memory, which halves the weight traffic out of L2 (the wide layers are bound by operand delivery, not by the tensor pipe).
// Protocol changes against the single-CTA kernel: a stage may be refilled only when BOTH CTAs' MMAs have read it (the MMA
// commit is multicast to both empty barriers, which count two arrivals); cluster barriers after the mbarrier initialisation
// and before exit (no CTA may leave while its peer can still write into it).  MMAs, TMEM and the epilogue stay per CTA.
template <bool CL2>
__device__ __forceinline__ void tc_kernel_body(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcParams& p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t full_bar[8], empty_bar[8], tmem_full[2], tmem_empty[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_stat[MAX_GROUPS][2][256];
  __shared__ __align__(16) float s_coef[MAX_GROUPS][4][256];     // {scale, shift, rstd, -mean*rstd} of the current channel tile

  pdl_trigger();
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // warp-uniform by construction
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t a_bytes = BM * p.KB * 2, b_bytes = p.BN * p.KB * 2;
  const uint32_t stage_bytes = (a_bytes + b_bytes + 1023) & ~1023u;
  // tile walk: single CTA: tile = mt * n_tiles + nt over the grid; CL2: the cluster walks tile PAIRS, CTA `rank` takes m-tile
  // 2 * pair_m + rank (a phantom tile past the end still runs -- its A boxes are out of bounds = zeros, its rows are not stored)
  const int rank = CL2 ? (int)cluster_ctarank() : 0;
  const int walk0 = CL2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int walk_step = CL2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int total_tiles = CL2 ? p.pair_tiles : p.m_tiles * p.n_tiles;
  const int KT = p.T * p.KC;
  const uint32_t tmem_cols = p.BN <= 32 ? 64 : (p.BN <= 64 ? 128 : (p.BN <= 128 ? 256 : 512));   // two accumulator stages

  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 2); mbar_init(&empty_bar[s], CL2 ? 2 : 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < MAX_GROUPS * 2 * 256; i += TC_THREADS) (&s_stat[0][0][0])[i] = 0.f;
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (CL2) cluster_sync_all();      // the peer's barriers are initialised before anything is multicast to them
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  pdl_wait();

  const int prod = warp == 0 ? 0 : (warp == 2 ? 1 : (warp == 3 ? 2 : (warp >= 8 && warp <= 10 ? warp - 5 : -1)));
  if (prod >= 0) {
    // ===================================== TMA producers =====================================
    // load number 2*g is the A tile of global k-block g, 2*g + 1 its B tile; producer `prod` issues the loads
    // whose number is congruent to it mod N_PRODUCERS
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
      int g = 0;
      for (int tile = walk0; tile < total_tiles; tile += walk_step) {
        const int mq = tile / p.n_tiles, nt = tile - mq * p.n_tiles;
        const int mt = CL2 ? 2 * mq + rank : mq;
        const int img0 = (mt / p.tiles_h) * p.Nt;
        const int h0 = (mt % p.tiles_h) * p.Ht;
        for (int kb = 0; kb < KT; ++kb, ++g) {
          const int la = (2 * g) % N_PRODUCERS;
          const bool do_a = la == prod, do_b = (la + 1) % N_PRODUCERS == prod;
          if (!do_a && !do_b) continue;
          const int t = kb / p.KC, cb = kb - t * p.KC;
          const int stage = g % p.stages;
          // A parity wait only tells the current phase from the previous one.  This producer last waited three k-blocks ago,
          // for k-block g - 3 - stages to be consumed, so the barrier it looks at now is at most one phase behind what it asks
          // for as long as stages >= 4 (host-side guarantee) -- and it can never be ahead, because this stage cannot be
          // consumed again without the load issued below.  (Waiting at k-blocks a producer does NOT load for is wrong the
          // other way round: the thread blocks ~700 cycles in each TMA issue, the barrier can flip twice meanwhile, and the
          // stale parity then reads as "not yet" -> deadlock; MEASURED.)
          mbar_wait(&empty_bar[stage], (uint32_t)(((g / p.stages) & 1) ^ 1));
          uint8_t* sa = smem + (size_t)stage * stage_bytes;
          if (do_a) {
            mbar_expect_tx(&full_bar[stage], a_bytes);
            tma_load_4d(sa, &tmA, &full_bar[stage], cb * p.KB, (int)p.dx[t], p.in_stride * h0 + (int)p.dy[t], img0);
          }
          if (do_b) {
            mbar_expect_tx(&full_bar[stage], b_bytes);      // the whole B tile lands here: one half from each CTA of the pair when CL2
            if (CL2) {
              const uint32_t half = b_bytes >> 1;
              tma_load_3d_mc(sa + a_bytes + (uint32_t)rank * half, &tmB, &full_bar[stage], cb * p.KB, nt * p.BN + rank * (p.BN >> 1), t, (uint16_t)3);
            } else {
              tma_load_3d(sa + a_bytes, &tmB, &full_bar[stage], cb * p.KB, nt * p.BN, t);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    // whole warp runs the loop (warp-uniform descriptor arithmetic on the uniform datapath); one lane issues
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    const uint32_t row_bytes = p.KB * 2;
    const int ksteps = p.KB / 16;
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = walk0; tile < total_tiles; tile += walk_step) {
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.BN);
      const bool issuer = elect_one();      // the same lane issues the tile's MMAs and commits
      for (int kb = 0; kb < KT; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem) + (uint32_t)stage * stage_bytes;
        const uint64_t adesc = make_desc(sa, row_bytes), bdesc = make_desc(sa + a_bytes, row_bytes);
        if (issuer) {
          if (ksteps == 4) {
#pragma unroll
            for (int k = 0; k < 4; ++k) tc_mma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
          } else {
#pragma unroll
            for (int k = 0; k < 2; ++k) tc_mma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
          }
          if (CL2) tc_commit_mc(&empty_bar[stage], (uint16_t)3);   // both CTAs' producers write this slot: tell both
          else tc_commit(&empty_bar[stage]);                       // smem slot is free once these MMAs have read it
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      if (issuer) tc_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
      __syncwarp();
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4 && warp < 8) {
    // ===================================== epilogue ==========================================
    const int q = warp & 3;                  // TMEM lane quarter this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    const int ohw = p.OH * p.OW;
    const bool bnb = p.bn_y != nullptr;
    int coef_nt = -1;
    for (int tile = walk0; tile < total_tiles; tile += walk_step) {
      const int mq = tile / p.n_tiles, nt = tile - mq * p.n_tiles;
      const int mt = CL2 ? 2 * mq + rank : mq;
      if (bnb && nt != coef_nt) {
        // coefficients of this channel tile for every pass group (a CTA's tiles mostly share nt: refilled on change only)
        asm volatile("bar.sync 2, 128;" ::: "memory");
        for (int i = tid - 128; i < p.groups * p.BN; i += 128) {
          const int g = i / p.BN, c = i - g * p.BN;
          const size_t k = (size_t)g * p.N + nt * p.BN + c;
          const float rs = rsqrtf(p.bn_var[k] + p.bn_eps);
          s_coef[g][0][c] = p.bn_scale[k]; s_coef[g][1][c] = p.bn_shift[k]; s_coef[g][2][c] = rs; s_coef[g][3][c] = -p.bn_mean[k] * rs;
        }
        asm volatile("bar.sync 2, 128;" ::: "memory");
        coef_nt = nt;
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int m = mt * BM + q * 32 + lane;
      const bool row_ok = m < p.M;
      const int mm = row_ok ? m : 0;
      const int nb = mm / ohw, r = mm - nb * ohw;
      const int oh = r / p.OW, ow = r - oh * p.OW;
      const size_t pix = ((size_t)nb * p.OHf + oh * p.out_stride + p.out_off_y) * p.OWf + ow * p.out_stride + p.out_off_x;
      const int g = min((mt * BM) / p.rows_per_group, p.groups - 1);      // (clamped: a CL2 phantom tile lies past the last group)
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.BN);
      for (int c0 = 0; c0 < p.BN; c0 += 16) {
        uint32_t raw[16];
        tc_ld16(taddr + c0, raw);
        tc_ld_wait();
        float v[16], sq[16];
        const int n0 = nt * p.BN + c0;
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(raw[j]);
        if (p.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] += p.bias[n0 + j];
        }
        if (row_ok) {
          if (p.res != nullptr) {
            float rr[16];
            bf16x8 r0, r1;
            ld_global_32B(p.res + pix * p.N + n0, r0, r1);
            unpack8(r0, rr);
            unpack8(r1, rr + 8);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += rr[j];
          }
          const bf16x8 o0 = pack8(v), o1 = pack8(v + 8);
          st_global_32B(p.out + pix * p.N + n0, o0, o1);
          unpack8(o0, v);
          unpack8(o1, v + 8);
          if (bnb) {
            float yv[16];
            bf16x8 y0, y1;
            ld_global_32B(p.bn_y + pix * p.N + n0, y0, y1);
            unpack8(y0, yv);
            unpack8(y1, yv + 8);
            const float4* cf = reinterpret_cast<const float4*>(&s_coef[g][0][c0]);
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4)
              bn_bwd_terms4(yv + 4 * j4, cf[j4], cf[64 + j4], cf[128 + j4], cf[192 + j4], p.bn_slope, v + 4 * j4, sq + 4 * j4);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) { v[j] = 0.f; sq[j] = 0.f; }
        }
        if (p.stats != nullptr) {
          // 16 columns x 32 rows -> per-column sums by a butterfly transpose-reduce (20 shuffles instead of 160)
          if (!bnb) {
#pragma unroll
            for (int j = 0; j < 16; ++j) sq[j] = v[j] * v[j];
          }
          const float s1 = colsum16(v, lane), s2 = colsum16(sq, lane);
          if (lane < 16) {
            atomicAdd(&s_stat[g][0][c0 + lane], s1);
            atomicAdd(&s_stat[g][1][c0 + lane], s2);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      if (p.stats != nullptr && p.n_tiles > 1) {
        // several n-tiles share the smem accumulators by column index within the tile: flush per tile
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int et = tid - 128;
        for (int i = et; i < 2 * p.BN; i += 128) {
          const int s = i / p.BN, c = i - s * p.BN;
          const float val = s_stat[g][s][c];
          const size_t slot = bnb ? (size_t)(s * p.groups + g) : (size_t)(g * 2 + s);     // BatchNorm-backward pair: [2][G][N]
          if (val != 0.f) atomicAdd(&p.stats[slot * p.N + nt * p.BN + c], val);
          s_stat[g][s][c] = 0.f;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
    }
    if (p.stats != nullptr && p.n_tiles == 1) {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int et = tid - 128;
      for (int i = et; i < p.groups * 2 * p.BN; i += 128) {
        const int g = i / (2 * p.BN), rem = i - g * 2 * p.BN;
        const int s = rem / p.BN, c = rem - s * p.BN;
        const float val = s_stat[g][s][c];
        const size_t slot = bnb ? (size_t)(s * p.groups + g) : (size_t)(g * 2 + s);
        if (val != 0.f) atomicAdd(&p.stats[slot * p.N + c], val);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CL2) cluster_sync_all();      // the peer may still multicast into this CTA / arrive on its barriers until it is done too
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// ------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

struct TileGeom {
  int Wt, Ht, Nt;
};

bool tile_geom(int OH, int OW, TileGeom* g) {
  if (OW <= 0 || OW > BM || (BM % OW) != 0) return false;
  const int rows = BM / OW;
  if (OH >= rows) {
    if (OH % rows) return false;
    *g = TileGeom{OW, rows, 1};
    return true;
  }
  if (rows % OH) return false;
  *g = TileGeom{OW, OH, rows / OH};
  return true;
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


__global__ void __launch_bounds__(TC_THREADS, 1)
igemm_fprop_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ TcParams p) {
  tc_kernel_body<false>(tmA, tmB, p);
}

// launched with cluster dimension 2 (launch attribute)
__global__ void __launch_bounds__(TC_THREADS, 1)
igemm_fprop_tc_cl2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ TcParams p) {
  tc_kernel_body<true>(tmA, tmB, p);
}

// Up to four independent problems of identical tile shape in ONE grid (blockIdx.y selects the problem): the four
// output-parity phases of a transposed convolution at 1x1 ... 4x4 resolution are 16-64 CTA GEMMs each, latency
// bound on their K loop -- side by side they fill the machine instead of running one after the other.
struct TcBatch {
  CUtensorMap tmA[4];
  CUtensorMap tmB[4];
  TcParams p[4];
};

__global__ void __launch_bounds__(TC_THREADS, 1) igemm_fprop_tc_batched_kernel(const __grid_constant__ TcBatch b) {
  const int ph = blockIdx.y;
  tc_kernel_body<false>(b.tmA[ph], b.tmB[ph], b.p[ph]);
}

}  // namespace

// channel tile.  MEASURED (profiles/r02_kernel_findings.md section 6): the wide layers are bound by operand delivery into shared
// memory, not by the tensor pipe, and the operand bytes per MAC fall with the tile width (A tile 16 KB + B tile BN/8 KB per
// 128 x BN x 64 MACs): the widest tile (a multiple of 16 that divides N, <= 256 = one MMA, >= 4 shared-memory stages of 64-channel
// k-blocks) wins as long as the persistent grid stays filled -- C4 25.6 -> 24.9 ms/step against the 128-column cap
// (SHOTVAE_TC_BN=128 restores it).  Tiles wider than one MMA (320 columns as two N = 160 MMAs over one A tile, which needs
// 32-channel k-blocks and a single accumulator stage) were MEASURED twice as slow: the kernel is bound by the NUMBER of TMA loads.
static int next_bn_below(int N, int bn) {
  for (int b = bn - 16; b >= 16; b -= 16)
    if (N % b == 0) return b;
  return 0;
}

static int pick_bn(int N) {
  static int cap = -1;
  if (cap < 0) { const char* e = getenv("SHOTVAE_TC_BN"); cap = e ? atoi(e) : 256; if (cap < 64 || cap > 256) cap = 256; }
  if (N <= 128) return N;
  return next_bn_below(N, (N < cap ? N : cap) + 16);
}

bool igemm_fprop_tc_supported(const IgemmParams& p) {
  TileGeom g;
  if (p.w_layout != 0) return false;
  // stride 2 (the first conv / projection shortcut of a resolution block): same kernel, the tensor map of A traverses H and
  // W with element stride 2, so a box of (2 Wt) x (2 Ht) input pixels lands in shared memory as the Wt x Ht pixels the tile needs
  if (!(p.in_stride == 1 || p.in_stride == 2) || p.H != p.OH * p.in_stride || p.W != p.OW * p.in_stride) return false;
  if (!tile_geom(p.OH, p.OW, &g)) return false;
  if (g.Wt * p.in_stride > 256 || g.Ht * p.in_stride > 256) return false;
  if (p.C % 32 != 0 || p.N % 16 != 0) return false;
  if (pick_bn(p.N) == 0) return false;
  if (p.out == nullptr || p.outf != nullptr) return false;
  if (p.stats != nullptr && (p.rows_per_group % BM != 0 || p.NB / p.group_images > MAX_GROUPS)) return false;
  if (p.bn_y != nullptr && (p.stats == nullptr || (reinterpret_cast<uintptr_t>(p.bn_y) & 31))) return false;
  if (p.NB < g.Nt) return false;
  if ((reinterpret_cast<uintptr_t>(p.A) & 15) || (reinterpret_cast<uintptr_t>(p.Wt) & 15)) return false;
  if ((reinterpret_cast<uintptr_t>(p.out) & 31) || (reinterpret_cast<uintptr_t>(p.res) & 31)) return false;   // 32-byte epilogue accesses
  return get_encode() != nullptr;
}

static int max_active_clusters(size_t smem);

// cl2 (in/out): in = a 2-CTA cluster launch is possible for the caller; out = this problem should use it
static int tc_setup(const IgemmParams& p, TcParams& q, CUtensorMap& tmA, CUtensorMap& tmB, size_t& smem, int& grid, bool* cl2 = nullptr) {
  EncodeTiledFn encode = get_encode();
  if (!encode) { sv_set_error("cuTensorMapEncodeTiled unavailable"); return SV_ERR_UNSUPPORTED; }
  TileGeom g;
  if (!tile_geom(p.OH, p.OW, &g)) { sv_set_error("igemm_fprop_tc: unsupported tile geometry"); return SV_ERR_UNSUPPORTED; }
  memset(&q, 0, sizeof(q));
  q.out = p.out; q.res = p.res; q.bias = p.bias; q.stats = p.stats;
  q.M = p.M; q.N = p.N; q.T = p.T;
  // 64-channel k-blocks (128-byte swizzled rows); a ragged last block (C = 160: 64 + 64 + 32) reads out of
  // bounds along the channel dimension of BOTH tensor maps, which TMA fills with zeros
  q.KB = p.C >= 64 ? 64 : 32;
  q.KC = ceil_div(p.C, q.KB);
  q.OH = p.OH; q.OW = p.OW; q.OHf = p.OHf; q.OWf = p.OWf;
  q.out_stride = p.out_stride; q.out_off_y = p.out_off_y; q.out_off_x = p.out_off_x;
  q.Wt = g.Wt; q.Ht = g.Ht; q.Nt = g.Nt;
  q.tiles_h = p.OH / g.Ht;
  q.m_tiles = ceil_div(p.M, BM);
  // channel tile: whole N when small; otherwise split so that the persistent grid is filled
  int bn = pick_bn(p.N);
  while (bn > 32 && q.m_tiles * (p.N / bn) < sm_count()) {
    const int nb = next_bn_below(p.N, bn);
    if (nb < 32) break;
    bn = nb;
  }
  q.BN = bn;
  q.n_tiles = p.N / bn;
  q.rows_per_group = p.rows_per_group;
  q.groups = p.NB / p.group_images;
  q.bn_y = p.bn_y; q.bn_scale = p.bn_scale; q.bn_shift = p.bn_shift; q.bn_mean = p.bn_mean; q.bn_var = p.bn_var;
  q.bn_slope = p.bn_slope; q.bn_eps = p.bn_eps;
  q.swizzle_bytes = q.KB * 2;
  const size_t stage_bytes = ((size_t)BM * q.KB * 2 + (size_t)q.BN * q.KB * 2 + 1023) & ~(size_t)1023;
  int stages = (int)((192 * 1024) / stage_bytes);
  if (stages > 8) stages = 8;
  if (stages < 4) { sv_set_error("igemm_fprop_tc: tile too large (the producers' parity protocol needs >= 4 stages)"); return SV_ERR_UNSUPPORTED; }
  q.stages = stages;
  q.in_stride = p.in_stride;
  memcpy(q.dy, p.dy, SV_MAX_TAPS);
  memcpy(q.dx, p.dx, SV_MAX_TAPS);

  smem = (size_t)stages * stage_bytes + 1024;
  // CTA pairs with multicast weight tiles: wide channel tiles (that is where the weight traffic is), enough tile pairs to keep
  // every resident cluster busy for at least two rounds; SHOTVAE_TC_CLUSTER=0 turns it off (A/B switch)
  bool use_cl2 = false;
  if (cl2 != nullptr && *cl2) {
    static int on = -1;
    if (on < 0) { const char* e = getenv("SHOTVAE_TC_CLUSTER"); on = (e && e[0] == '0') ? 0 : 1; }
    q.pair_tiles = ceil_div(q.m_tiles, 2) * q.n_tiles;
    const int clusters = on ? max_active_clusters(smem) : 0;
    // MEASURED on C4 (profiles/r02_kernel_findings.md section 6): +10 % on the layers whose channel tile is the whole N (160-channel
    // block: 843 -> 929 TFLOP/s), -4 % where N is split into several tiles (320 / 640 channels) -> only the former take it
    use_cl2 = clusters >= 32 && q.BN >= 128 && q.n_tiles == 1 && q.pair_tiles >= 2 * clusters;
    if (use_cl2) grid = 2 * clusters;
    *cl2 = use_cl2;
  }
  const CUtensorMapSwizzle sw = q.KB == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  {
    cuuint64_t dims[4] = {(cuuint64_t)p.C, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.NB};
    cuuint64_t strides[3] = {(cuuint64_t)p.C * 2, (cuuint64_t)p.W * p.C * 2, (cuuint64_t)p.H * p.W * p.C * 2};
    const cuuint32_t is = (cuuint32_t)p.in_stride;    // box extents are in tensor elements: ceil(box / stride) pixels are loaded
    cuuint32_t box[4] = {(cuuint32_t)q.KB, (cuuint32_t)g.Wt * is, (cuuint32_t)g.Ht * is, (cuuint32_t)g.Nt};
    cuuint32_t es[4] = {1, is, is, 1};
    CUresult r = encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(p.A), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { sv_set_error("cuTensorMapEncodeTiled(A) failed: %d", (int)r); return SV_ERR_CUDA; }
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)p.C, (cuuint64_t)p.N, (cuuint64_t)p.T};
    cuuint64_t strides[2] = {(cuuint64_t)p.C * 2, (cuuint64_t)p.N * p.C * 2};
    cuuint32_t box[3] = {(cuuint32_t)q.KB, (cuuint32_t)(use_cl2 ? q.BN / 2 : q.BN), 1};     // CL2: each CTA of the pair fetches half the rows
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<bf16*>(p.Wt), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { sv_set_error("cuTensorMapEncodeTiled(B) failed: %d", (int)r); return SV_ERR_CUDA; }
  }
  if (!use_cl2) {
    const int total = q.m_tiles * q.n_tiles;
    grid = total < sm_count() ? total : sm_count();
  }
  return SV_OK;
}

static void tc_configure() {
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(igemm_fprop_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(igemm_fprop_tc_cl2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(igemm_fprop_tc_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    configured = true;
  }
}

// how many 2-CTA clusters of the CL2 kernel can be resident at once (1 CTA per SM, pairs inside one GPC): the persistent grid
// must not exceed it, or the surplus clusters would run as a second wave
static int max_active_clusters(size_t smem) {
  static int cached = -1;
  static size_t cached_smem = 0;
  if (cached >= 0 && cached_smem == smem) return cached;
  tc_configure();
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(sm_count() & ~1);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, igemm_fprop_tc_cl2_kernel, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
  if (n > sm_count() / 2) n = sm_count() / 2;
  cached = n; cached_smem = smem;
  return n;
}

int igemm_fprop_tc(const IgemmParams& p, cudaStream_t st) {
  TcParams q;
  CUtensorMap tmA, tmB;
  size_t smem;
  int grid;
  bool cl2 = true;
  const int rc = tc_setup(p, q, tmA, tmB, smem, grid, &cl2);
  if (rc != SV_OK) return rc;
  tc_configure();
  if (cl2) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = sv_pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    cudaLaunchKernelEx(&cfg, igemm_fprop_tc_cl2_kernel, tmA, tmB, q);
    return sv_check_launch("igemm_fprop_tc_cl2");
  }
  sv_launch_pdl(igemm_fprop_tc_kernel, dim3(grid), dim3(TC_THREADS), smem, st, tmA, tmB, q);
  return sv_check_launch("igemm_fprop_tc");
}

int igemm_fprop_tc_batch(const IgemmParams* ps, int n, cudaStream_t st) {
  if (n < 1 || n > 4) { sv_set_error("igemm_fprop_tc_batch: 1..4 problems"); return SV_ERR_ARG; }
  TcBatch b;
  memset(&b, 0, sizeof(b));
  size_t smem = 0;
  int grid = 0;
  for (int i = 0; i < n; ++i) {
    size_t s;
    int g;
    const int rc = tc_setup(ps[i], b.p[i], b.tmA[i], b.tmB[i], s, g);
    if (rc != SV_OK) return rc;
    if (i > 0 && (s != smem || b.p[i].stages != b.p[0].stages)) { sv_set_error("igemm_fprop_tc_batch: problems differ in tile shape"); return SV_ERR_ARG; }
    smem = s;
    if (g > grid) grid = g;
  }
  tc_configure();
  sv_launch_pdl(igemm_fprop_tc_batched_kernel, dim3(grid, n), dim3(TC_THREADS), smem, st, b);
  return sv_check_launch("igemm_fprop_tc_batch");
}
