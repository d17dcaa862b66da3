// Weight-gradient implicit GEMM for wide layers (WRN-28-10: 160 / 320 / 640 channels, PreActResNet18: 256 / 512) on
// tcgen05 with TMA-fed operands:
//
//   dW[t][n][c] = sum over output pixels s of  G[s][n] * A[s + (dy_t, dx_t)][c]          (stride-1 convolutions)
//
// Both operands are tiles of NHWC tensors -- 128 pixels x 64 channels, 128-byte rows, loaded by 4-D tensor-map boxes
// with the 128-byte swizzle (the activation box shifted by the tap offset, out-of-image pixels zero-filled by TMA).
// Such a tile IS an MN-major UMMA operand whose K dimension runs over its pixel rows:
//   A operand (M = 128 output channels n): two 64-channel boxes of the output gradient G, 16 KB apart (LBO);
//   B operand (N = cb <= 256 input channels c): ceil(cb / 64) boxes of the shifted activation;
//   8-row groups are 1 KB apart (SBO); one MMA consumes K = 16 pixels, a 128-pixel block 8 MMAs.
// A CTA owns a "unit" = one 128-channel n-tile x up to floor(512 / cb) (tap, c-block) pairs, whose FP32 accumulators
// (pairs x cb columns) stay in TMEM for the whole launch, and a slice of the pixel blocks (split-K over CTAs); the
// partial sums are written once at the end and summed by sv_wgrad_reduce.
//
// warp roles: 1 = MMA issuer, 4..7 = final epilogue, 0 / 2 / 3 / 8..12 = eight TMA producers (2 allocates TMEM).  Tensor-map
// loads issued by one thread do not overlap (tools/tma_probe.cu: ~700 cycles each whatever the box size), so every box of a
// stage has its own producer: producer 4*stage + b loads box b of EVERY use of that stage (G box b and A box b).  Owning
// a fixed slot matters for correctness, not only for speed: an mbarrier parity wait can only tell the current phase from
// the previous one, so a thread must neither skip uses of a barrier (it could run two phases ahead -- found with
// compute-sanitizer as an over-arrival) nor wait on uses it does not load for (it blocks in its own TMA issue meanwhile and
// can fall two phases behind -- found as a deadlock in igemm_tc.cu).
#include <cuda.h>
#include "common.cuh"
#include "igemm.h"

namespace {

constexpr int WT_THREADS = 544;          // 17 warps: MMA, 4 epilogue, up to 12 TMA producers
constexpr int MAX_A_STAGES = 3;
constexpr int BLK = 128;                 // pixels per block = channels per n-tile
constexpr uint32_t BOX_BYTES = 64 * BLK * 2;
constexpr uint32_t SPIN_LIMIT = 1u << 28;

struct WtcParams {
  float* partial;
  int N, C, T;
  int PB;                    // pixel blocks
  int Nt, Ht, tiles_h;       // box geometry (images, rows per block; blocks per image column)
  int cb, c_blocks, boxes;   // channels per MMA (N of the instruction), C / cb, ceil(cb / 64)
  int pairs, ppu, units_per_nt, splits;
  int a_stages;              // activation stages in shared memory: 3 while they fit (<= 3 boxes per stage), else 2
  int in_stride;             // 1, or 2: the activation boxes sample every second pixel (tensor-map element strides)
  int8_t dy[SV_MAX_TAPS];
  int8_t dx[SV_MAX_TAPS];
};

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
      printf("wgrad_tc: mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
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

// MN-major, 128-byte swizzled operand: 64-element (128 B) rows along M/N, one row per K index; LBO = distance between
// 64-element column blocks, SBO = distance between groups of 8 K rows
__device__ __forceinline__ uint64_t make_desc_mn128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;       // SWIZZLE_128B
  return d;
}

__global__ void __launch_bounds__(WT_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmA, const __grid_constant__ WtcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t g_full[2], g_empty[2], a_full[MAX_A_STAGES], a_empty[MAX_A_STAGES], done_bar;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t g_stage = 2 * BOX_BYTES, a_stage = (uint32_t)p.boxes * BOX_BYTES;
  const uint32_t g_base = smem_u32(smem), a_base = g_base + 2 * g_stage;

  const int unit = blockIdx.x / p.splits, slice = blockIdx.x - unit * p.splits;
  const int nt = unit / p.units_per_nt, u = unit - nt * p.units_per_nt;
  const int pair0 = u * p.ppu, npair = min(p.ppu, p.pairs - pair0);
  const int pb0 = (int)((long long)slice * p.PB / p.splits), pb1 = (int)((long long)(slice + 1) * p.PB / p.splits);

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) { mbar_init(&g_full[s], 2); mbar_init(&g_empty[s], 1); }
    for (int s = 0; s < p.a_stages; ++s) { mbar_init(&a_full[s], (uint32_t)p.boxes); mbar_init(&a_empty[s], 1); }
    mbar_init(&done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  // producer 4 * stage + b owns box slot b of activation stage `stage` (and, for stage < 2 and b < 2, of the G stage)
  const int prod = warp == 0 ? 0 : (warp == 2 ? 1 : (warp == 3 ? 2 : (warp >= 8 && warp <= 16 ? warp - 5 : -1)));
  if (prod >= 0) {
    // ===================================== TMA producers =====================================
    const int ps = prod >> 2, pbx = prod & 3;       // the stage and the box slot this producer owns
    if (lane == 0 && ps < p.a_stages && (pbx < p.boxes || (pbx < 2 && ps < 2))) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmG) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
      int gi = 0, ai = 0;       // G stages / A stages filled so far
      for (int pb = pb0; pb < pb1; ++pb, ++gi) {
        const int img0 = (pb / p.tiles_h) * p.Nt, h0 = (pb % p.tiles_h) * p.Ht;
        if (ps < 2 && (gi & 1) == ps && pbx < 2) {
          mbar_wait(&g_empty[ps], (uint32_t)(((gi >> 1) & 1) ^ 1));
          mbar_expect_tx(&g_full[ps], BOX_BYTES);
          tma_load_4d(g_base + ps * g_stage + pbx * BOX_BYTES, &tmG, &g_full[ps], nt * BLK + 64 * pbx, 0, h0, img0);
        }
        for (int pr = 0; pr < npair; ++pr, ++ai) {
          if (ai % p.a_stages != ps || pbx >= p.boxes) continue;
          const int pair = pair0 + pr, t = pair / p.c_blocks, cblk = pair - t * p.c_blocks;
          mbar_wait(&a_empty[ps], (uint32_t)(((ai / p.a_stages) & 1) ^ 1));
          mbar_expect_tx(&a_full[ps], BOX_BYTES);
          tma_load_4d(a_base + ps * a_stage + pbx * BOX_BYTES, &tmA, &a_full[ps], cblk * p.cb + 64 * pbx, (int)p.dx[t], p.in_stride * h0 + (int)p.dy[t], img0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(p.cb >> 3) << 17) | ((uint32_t)(BLK >> 4) << 24);
    const bool issuer = elect_one();
    int gi = 0, ai = 0;
    for (int pb = pb0; pb < pb1; ++pb, ++gi) {
      const int gs = gi & 1;
      mbar_wait(&g_full[gs], (uint32_t)((gi >> 1) & 1));
      tc_fence_after();
      const uint64_t gdesc = make_desc_mn128(g_base + gs * g_stage, BOX_BYTES, 1024);
      for (int pr = 0; pr < npair; ++pr, ++ai) {
        const int as = ai % p.a_stages;
        mbar_wait(&a_full[as], (uint32_t)((ai / p.a_stages) & 1));
        tc_fence_after();
        const uint64_t adesc = make_desc_mn128(a_base + as * a_stage, BOX_BYTES, 1024);
        const uint32_t d_tmem = tmem_base + (uint32_t)(pr * p.cb);
        if (issuer) {
#pragma unroll
          for (int ks = 0; ks < BLK / 16; ++ks)       // 16 pixel rows = 2 KB further into both tiles
            tc_mma_bf16(d_tmem, gdesc + (uint64_t)(ks * 128), adesc + (uint64_t)(ks * 128), idesc, (pb != pb0 || ks != 0) ? 1u : 0u);
          tc_commit(&a_empty[as]);
          if (pr == npair - 1) tc_commit(&g_empty[gs]);
        }
        __syncwarp();
      }
    }
    if (issuer) tc_commit(&done_bar);
    __syncwarp();
  } else if (warp >= 4 && warp < 8) {
    // ===================================== final epilogue ====================================
    const int q = warp & 3;
    mbar_wait(&done_bar, 0);
    tc_fence_after();
    const int n = nt * BLK + q * 32 + lane;
    const size_t row = ((size_t)slice * p.N + (size_t)min(n, p.N - 1)) * (size_t)(p.T * p.C);
    for (int pr = 0; pr < npair; ++pr) {
      const int pair = pair0 + pr, t = pair / p.c_blocks, cblk = pair - t * p.c_blocks;
      float* dst = p.partial + row + (size_t)t * p.C + (size_t)cblk * p.cb;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(pr * p.cb);
      for (int c0 = 0; c0 < p.cb; c0 += 16) {
        uint32_t raw[16];
        tc_ld16(taddr + c0, raw);
        tc_ld_wait();
        if (n < p.N) {
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4*>(dst + c0 + j) = make_float4(__uint_as_float(raw[j]), __uint_as_float(raw[j + 1]), __uint_as_float(raw[j + 2]),
                                                                  __uint_as_float(raw[j + 3]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

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
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
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

// channels per MMA: the largest multiple of 16 that divides C and is <= 256
int pick_cb(int C) {
  for (int cb = 256; cb >= 16; cb -= 16)
    if (C % cb == 0) return cb;
  return 0;
}

bool geometry(const WgradParams& p, WtcParams& q) {
  // pixel blocks tile the OUTPUT grid (the rows of G); the activation box of a block covers in_stride x as many input pixels
  if (p.OW <= 0 || p.OW > BLK || BLK % p.OW) return false;
  const int rows = BLK / p.OW;
  if (p.OH >= rows) {
    if (p.OH % rows) return false;
    q.Ht = rows; q.Nt = 1;
  } else {
    if (rows % p.OH) return false;
    q.Ht = p.OH; q.Nt = rows / p.OH;
    if (p.NB % q.Nt) return false;
  }
  if (p.OW * p.in_stride > 256 || q.Ht * p.in_stride > 256) return false;
  q.in_stride = p.in_stride;
  q.tiles_h = p.OH / q.Ht;
  q.PB = p.M / BLK;
  q.N = p.N; q.C = p.C; q.T = p.T;
  q.cb = pick_cb(p.C);
  if (q.cb == 0) return false;
  q.c_blocks = p.C / q.cb;
  q.boxes = ceil_div(q.cb, 64);
  q.a_stages = q.boxes <= 3 ? MAX_A_STAGES : 2;      // 2 G stages (64 KB) + 3 x 48 KB or 2 x 64 KB of activation stages
  q.pairs = p.T * q.c_blocks;
  q.ppu = 512 / q.cb;
  q.units_per_nt = ceil_div(q.pairs, q.ppu);
  const int units = ceil_div(p.N, BLK) * q.units_per_nt;
  int splits = sm_count() / units;
  if (splits > q.PB) splits = q.PB;
  if (splits < 1) splits = 1;
  q.splits = splits;
  memcpy(q.dy, p.dy, SV_MAX_TAPS);
  memcpy(q.dx, p.dx, SV_MAX_TAPS);
  return true;
}

}  // namespace

bool wgrad_tc_supported(const WgradParams& p) {
  if (!(p.in_stride == 1 || p.in_stride == 2) || p.H != p.OH * p.in_stride || p.W != p.OW * p.in_stride) return false;
  if (p.C < 64 || p.N < 64 || p.C % 16 || p.N % 16) return false;       // narrow layers: the halo-tile kernel
  if (p.M % BLK) return false;
  if ((reinterpret_cast<uintptr_t>(p.A) & 15) || (reinterpret_cast<uintptr_t>(p.Gr) & 15) || (reinterpret_cast<uintptr_t>(p.partial) & 15)) return false;
  WtcParams q;
  if (!geometry(p, q)) return false;
  return get_encode() != nullptr;
}

int wgrad_tc_splits(const WgradParams& p) {
  WtcParams q;
  return geometry(p, q) ? q.splits : 0;
}

int wgrad_tc(const WgradParams& p, cudaStream_t st) {
  WtcParams q;
  memset(&q, 0, sizeof(q));
  if (!geometry(p, q)) { sv_set_error("wgrad_tc: unsupported geometry"); return SV_ERR_UNSUPPORTED; }
  if (p.splits != q.splits) { sv_set_error("wgrad_tc: caller must use %d splits (sv_igemm_wgrad_splits), got %d", q.splits, p.splits); return SV_ERR_ARG; }
  q.partial = p.partial;
  EncodeTiledFn encode = get_encode();
  if (!encode) { sv_set_error("cuTensorMapEncodeTiled unavailable"); return SV_ERR_UNSUPPORTED; }
  CUtensorMap tmG, tmA;
  cuuint32_t box[4] = {64, (cuuint32_t)p.OW, (cuuint32_t)q.Ht, (cuuint32_t)q.Nt};
  cuuint32_t es[4] = {1, 1, 1, 1};
  {
    cuuint64_t dims[4] = {(cuuint64_t)p.N, (cuuint64_t)p.OW, (cuuint64_t)p.OH, (cuuint64_t)p.NB};
    cuuint64_t strides[3] = {(cuuint64_t)p.N * 2, (cuuint64_t)p.OW * p.N * 2, (cuuint64_t)p.OH * p.OW * p.N * 2};
    CUresult r = encode(&tmG, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(p.Gr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { sv_set_error("cuTensorMapEncodeTiled(G) failed: %d", (int)r); return SV_ERR_CUDA; }
  }
  {
    cuuint64_t dims[4] = {(cuuint64_t)p.C, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.NB};
    cuuint64_t strides[3] = {(cuuint64_t)p.C * 2, (cuuint64_t)p.W * p.C * 2, (cuuint64_t)p.H * p.W * p.C * 2};
    const cuuint32_t is = (cuuint32_t)p.in_stride;    // box extents are in tensor elements: ceil(box / stride) pixels are loaded
    cuuint32_t abox[4] = {64, (cuuint32_t)p.OW * is, (cuuint32_t)q.Ht * is, (cuuint32_t)q.Nt};
    cuuint32_t aes[4] = {1, is, is, 1};
    CUresult r = encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(p.A), dims, strides, abox, aes, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { sv_set_error("cuTensorMapEncodeTiled(A) failed: %d", (int)r); return SV_ERR_CUDA; }
  }
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024);
    configured = true;
  }
  const size_t smem = (size_t)(2 * 2 + q.a_stages * q.boxes) * BOX_BYTES + 1024;
  const int grid = ceil_div(p.N, BLK) * q.units_per_nt * q.splits;
  wgrad_tc_kernel<<<grid, WT_THREADS, smem, st>>>(tmG, tmA, q);
  return sv_check_launch("wgrad_tc");
}
