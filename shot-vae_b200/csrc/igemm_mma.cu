// Implicit-GEMM convolution kernels on the legacy tensor path (mma.sync m16n8k16, bf16 -> fp32).
//
// These cover every shape (any C, N multiple of 16, any tap table, strided gather / strided store)
// and are the correctness anchor for the tcgen05/TMA kernels in igemm_tc.cu, which take over the
// shapes they support.  Operands are gathered with zero-filling cp.async into padded shared-memory
// tiles, three pipeline stages deep.
#include "common.cuh"
#include "igemm.h"

namespace {

__device__ __forceinline__ void ldmatrix_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma_bf16(float* c, const uint32_t* a, const uint32_t* b) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// ------------------------------------------------------------------------------------------------
// fprop-like gather GEMM
// ------------------------------------------------------------------------------------------------
template <int BN, int BK>
__device__ __forceinline__ void fprop_mma_body(const IgemmParams& p) {
  constexpr int BM = 128, STAGES = 3, LDS = BK + 8, CPR = BK / 8;
  constexpr int WARPS_N = (BN >= 64) ? 2 : 1;
  constexpr int WARPS_M = 8 / WARPS_N;
  constexpr int WM = BM / WARPS_M, WN = BN / WARPS_N;
  constexpr int MI = WM / 16, NJ = WN / 8;
  constexpr int A_ITERS = (BM * CPR) / 256;
  constexpr int B_ITERS = (BN * CPR + 255) / 256;
  constexpr int LDC = BN + 8;
  static_assert(A_ITERS >= 1 && NJ % 2 == 0, "tile config");

  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ float s_stat[2][BN];
  bf16* sA = reinterpret_cast<bf16*>(smem_raw);
  bf16* sB = sA + STAGES * BM * LDS;
  float* sC = reinterpret_cast<float*>(smem_raw);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int warp_m = warp % WARPS_M, warp_n = warp / WARPS_M;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int ohw = p.OH * p.OW;
  const bf16* __restrict__ Ag = p.A;
  const bf16* __restrict__ Wg = p.Wt;

  int a_pix[A_ITERS], a_ih[A_ITERS], a_iw[A_ITERS];
  bool a_ok[A_ITERS];
#pragma unroll
  for (int i = 0; i < A_ITERS; ++i) {
    const int row = (tid + i * 256) / CPR;
    const int m = m0 + row;
    a_ok[i] = m < p.M;
    const int mm = a_ok[i] ? m : 0;
    const int nb = mm / ohw, r = mm - nb * ohw;
    const int oh = r / p.OW, ow = r - oh * p.OW;
    a_pix[i] = nb * p.H * p.W;
    a_ih[i] = oh * p.in_stride;
    a_iw[i] = ow * p.in_stride;
  }
  const int KC = p.C / BK;
  const int KT = p.T * KC;

  auto load_stage = [&](int stage, int kiter) {
    const int t = kiter / KC, cb = kiter - t * KC;
    const int dy = p.dy[t], dx = p.dx[t];
#pragma unroll
    for (int i = 0; i < A_ITERS; ++i) {
      const int q = tid + i * 256, row = q / CPR, ch = q % CPR;
      const int ih = a_ih[i] + dy, iw = a_iw[i] + dx;
      const bool ok = a_ok[i] && ih >= 0 && ih < p.H && iw >= 0 && iw < p.W;
      const bf16* src = ok ? Ag + ((size_t)(a_pix[i] + ih * p.W + iw) * p.C + cb * BK + ch * 8) : Ag;
      cp_async16(sA + ((size_t)stage * BM + row) * LDS + ch * 8, src, ok);
    }
#pragma unroll
    for (int i = 0; i < B_ITERS; ++i) {
      const int q = tid + i * 256;
      if (q < BN * CPR) {
        const int row = q / CPR, ch = q % CPR;
        const int n = n0 + row;
        const bool ok = n < p.N;
        const bf16* src = ok ? Wg + ((size_t)(t * p.N + n) * p.C + cb * BK + ch * 8) : Wg;
        cp_async16(sB + ((size_t)stage * BN + row) * LDS + ch * 8, src, ok);
      }
    }
  };

  float acc[MI][NJ][4];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NJ; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < KT) load_stage(s, s);
    cp_async_commit();
  }
  for (int k = 0; k < KT; ++k) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    if (k + STAGES - 1 < KT) load_stage((k + STAGES - 1) % STAGES, k + STAGES - 1);
    cp_async_commit();
    const int stage = k % STAGES;
    const bf16* tA = sA + (size_t)stage * BM * LDS;
    const bf16* tB = sB + (size_t)stage * BN * LDS;
#pragma unroll
    for (int ks = 0; ks < BK / 16; ++ks) {
      uint32_t af[MI][4], bfr[NJ][2];
#pragma unroll
      for (int mi = 0; mi < MI; ++mi) {
        const int row = warp_m * WM + mi * 16 + (lane & 15);
        const int col = ks * 16 + (lane >> 4) * 8;
        ldmatrix_x4(af[mi][0], af[mi][1], af[mi][2], af[mi][3], tA + row * LDS + col);
      }
#pragma unroll
      for (int nj = 0; nj < NJ / 2; ++nj) {
        const int row = warp_n * WN + nj * 16 + (lane & 7) + (lane >> 4) * 8;
        const int col = ks * 16 + ((lane >> 3) & 1) * 8;
        ldmatrix_x4(bfr[2 * nj][0], bfr[2 * nj][1], bfr[2 * nj + 1][0], bfr[2 * nj + 1][1], tB + row * LDS + col);
      }
#pragma unroll
      for (int mi = 0; mi < MI; ++mi)
#pragma unroll
        for (int nj = 0; nj < NJ; ++nj) mma_bf16(acc[mi][nj], af[mi], bfr[nj]);
    }
  }
  cp_async_wait<0>();
  __syncthreads();

  // ---- epilogue: accumulators -> smem (fp32) -> coalesced 16-byte rows ----
#pragma unroll
  for (int mi = 0; mi < MI; ++mi)
#pragma unroll
    for (int nj = 0; nj < NJ; ++nj) {
      const int row = warp_m * WM + mi * 16 + (lane >> 2);
      const int col = warp_n * WN + nj * 8 + (lane & 3) * 2;
      *reinterpret_cast<float2*>(&sC[row * LDC + col]) = make_float2(acc[mi][nj][0], acc[mi][nj][1]);
      *reinterpret_cast<float2*>(&sC[(row + 8) * LDC + col]) = make_float2(acc[mi][nj][2], acc[mi][nj][3]);
    }
  if (tid < BN) { s_stat[0][tid] = 0.f; s_stat[1][tid] = 0.f; }
  __syncthreads();

  constexpr int CPO = BN / 8;  // 16-byte output chunks per row
  const int cc = tid % CPO;
  const int nbase = n0 + cc * 8;
  const bool n_ok = nbase < p.N;
  const int m_last = min(m0 + BM, p.M) - 1;
  const bool one_group = (p.stats != nullptr) && (m0 / p.rows_per_group == m_last / p.rows_per_group);
  float s1[8], s2[8], bias8[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    s1[j] = 0.f; s2[j] = 0.f;
    bias8[j] = (p.bias != nullptr && n_ok) ? p.bias[nbase + j] : 0.f;
  }
  for (int q = tid; q < BM * CPO; q += 256) {
    const int row = q / CPO;
    const int m = m0 + row;
    if (m >= p.M || !n_ok) continue;
    const int nb = m / ohw, r = m - nb * ohw;
    const int oh = r / p.OW, ow = r - oh * p.OW;
    const size_t pix = ((size_t)nb * p.OHf + oh * p.out_stride + p.out_off_y) * p.OWf + ow * p.out_stride + p.out_off_x;
    float v[8];
    const float4 c0 = *reinterpret_cast<const float4*>(&sC[row * LDC + cc * 8]);
    const float4 c1 = *reinterpret_cast<const float4*>(&sC[row * LDC + cc * 8 + 4]);
    v[0] = c0.x; v[1] = c0.y; v[2] = c0.z; v[3] = c0.w; v[4] = c1.x; v[5] = c1.y; v[6] = c1.z; v[7] = c1.w;
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] += bias8[j];
    if (p.res != nullptr) {
      float rr[8];
      unpack8(*reinterpret_cast<const bf16x8*>(p.res + pix * p.N + nbase), rr);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += rr[j];
    }
    if (p.outf != nullptr) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (nbase + j < p.n_valid) p.outf[pix * p.n_valid + nbase + j] = v[j];
    }
    if (p.out != nullptr) {
      const bf16x8 o = pack8(v);
      *reinterpret_cast<bf16x8*>(p.out + pix * p.N + nbase) = o;
      unpack8(o, v);  // statistics of the values actually stored
    }
    if (p.stats != nullptr) {
      if (one_group) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { s1[j] += v[j]; s2[j] += v[j] * v[j]; }
      } else {
        const int g = nb / p.group_images;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          atomicAdd(&p.stats[(size_t)(g * 2 + 0) * p.N + nbase + j], v[j]);
          atomicAdd(&p.stats[(size_t)(g * 2 + 1) * p.N + nbase + j], v[j] * v[j]);
        }
      }
    }
  }
  if (one_group) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(&s_stat[0][cc * 8 + j], s1[j]);
      atomicAdd(&s_stat[1][cc * 8 + j], s2[j]);
    }
    __syncthreads();
    if (tid < BN && n0 + tid < p.N) {
      const int g = m0 / p.rows_per_group;
      atomicAdd(&p.stats[(size_t)(g * 2 + 0) * p.N + n0 + tid], s_stat[0][tid]);
      atomicAdd(&p.stats[(size_t)(g * 2 + 1) * p.N + n0 + tid], s_stat[1][tid]);
    }
  }
}

template <int BN, int BK>
__global__ void __launch_bounds__(256) igemm_fprop_mma_kernel(const __grid_constant__ IgemmParams p) {
  pdl_trigger();
  pdl_wait();
  fprop_mma_body<BN, BK>(p);
}

// up to four problems of identical geometry in one grid (blockIdx.z = problem): the output-parity phases of a transposed
// convolution whose shape the tcgen05 kernels do not cover (the last decoder layer, 64 -> 3 channels)
struct MmaBatch {
  IgemmParams p[4];
};
template <int BN, int BK>
__global__ void __launch_bounds__(256) igemm_fprop_mma_batched_kernel(const __grid_constant__ MmaBatch b) {
  pdl_trigger();
  pdl_wait();
  fprop_mma_body<BN, BK>(b.p[blockIdx.z]);
}

template <int BN, int BK>
int launch_fprop_batch(const IgemmParams* ps, int n, cudaStream_t st) {
  constexpr int BM = 128, STAGES = 3, LDS = BK + 8;
  constexpr size_t pipe = (size_t)STAGES * (BM + BN) * LDS * sizeof(bf16);
  constexpr size_t epi = (size_t)BM * (BN + 8) * sizeof(float);
  constexpr size_t smem = pipe > epi ? pipe : epi;
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(igemm_fprop_mma_batched_kernel<BN, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured = true;
  }
  MmaBatch b;
  for (int i = 0; i < n; ++i) b.p[i] = ps[i];
  dim3 grid(ceil_div(ps[0].M, BM), ceil_div(ps[0].N, BN), n);
  sv_launch_pdl(igemm_fprop_mma_batched_kernel<BN, BK>, dim3(grid), dim3(256), smem, st, b);
  return sv_check_launch("igemm_fprop_mma_batched");
}

template <int BN, int BK>
int launch_fprop(const IgemmParams& p, cudaStream_t st) {
  constexpr int BM = 128, STAGES = 3, LDS = BK + 8;
  constexpr size_t pipe = (size_t)STAGES * (BM + BN) * LDS * sizeof(bf16);
  constexpr size_t epi = (size_t)BM * (BN + 8) * sizeof(float);
  constexpr size_t smem = pipe > epi ? pipe : epi;
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(igemm_fprop_mma_kernel<BN, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured = true;
  }
  dim3 grid(ceil_div(p.M, BM), ceil_div(p.N, BN));
  sv_launch_pdl(igemm_fprop_mma_kernel<BN, BK>, dim3(grid), dim3(256), smem, st, p);
  return sv_check_launch("igemm_fprop_mma");
}

// ------------------------------------------------------------------------------------------------
// wgrad-like GEMM: part[s][n][v] = sum_m Gr[m][n] * A[gather(m, tap(v)), chan(v)]
// ------------------------------------------------------------------------------------------------
template <int BMN>
__global__ void __launch_bounds__(256) igemm_wgrad_mma_kernel(const WgradParams p) {
  pdl_trigger();
  pdl_wait();
  constexpr int BV = 128, BKP = 32, STAGES = 3;
  constexpr int LDG = BMN + 8, LDA = BV + 8;
  constexpr int WARPS_M = (BMN >= 32) ? 2 : 1;
  constexpr int WARPS_N = 8 / WARPS_M;
  constexpr int WM = BMN / WARPS_M, WN = BV / WARPS_N;
  constexpr int MI = WM / 16, NJ = WN / 8;
  constexpr int GCPR = BMN / 8;                        // 16B chunks per G row
  constexpr int G_ITERS = (BKP * GCPR + 255) / 256;
  static_assert(NJ % 2 == 0 && MI >= 1, "tile config");

  extern __shared__ __align__(16) unsigned char smem_raw[];
  bf16* sG = reinterpret_cast<bf16*>(smem_raw);        // [STAGES][BKP][LDG]
  bf16* sA = sG + STAGES * BKP * LDG;                  // [STAGES][BKP][LDA]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int warp_m = warp % WARPS_M, warp_n = warp / WARPS_M;
  const int split = blockIdx.x;
  const int v0 = blockIdx.y * BV, n0 = blockIdx.z * BMN;
  const int TC = p.T * p.C;
  const int ohw = p.OH * p.OW;
  const int m_begin = split * p.rows_per_split;
  const int m_end = min(m_begin + p.rows_per_split, p.M);
  const bf16* __restrict__ Ag = p.A;
  const bf16* __restrict__ Gg = p.Gr;

  // this thread's fixed virtual-column chunk of the A tile
  const int aj = tid & 15;
  const int av = v0 + aj * 8;
  const bool av_ok = av < TC;
  const int at = av_ok ? av / p.C : 0;
  const int ac = av_ok ? av - at * p.C : 0;
  const int ady = p.dy[at], adx = p.dx[at];

  auto load_stage = [&](int stage, int k0) {
#pragma unroll
    for (int i = 0; i < G_ITERS; ++i) {
      const int q = tid + i * 256;
      if (q < BKP * GCPR) {
        const int row = q / GCPR, ch = q % GCPR;
        const int m = k0 + row, n = n0 + ch * 8;
        const bool ok = m < m_end && n < p.N;
        const bf16* src = ok ? Gg + (size_t)m * p.N + n : Gg;
        cp_async16(sG + ((size_t)stage * BKP + row) * LDG + ch * 8, src, ok);
      }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int row = (tid >> 4) + i * 16;
      const int m = k0 + row;
      bool ok = av_ok && m < m_end;
      const bf16* src = Ag;
      if (ok) {
        const int nb = m / ohw, r = m - nb * ohw;
        const int oh = r / p.OW, ow = r - oh * p.OW;
        const int ih = oh * p.in_stride + ady, iw = ow * p.in_stride + adx;
        ok = ih >= 0 && ih < p.H && iw >= 0 && iw < p.W;
        if (ok) src = Ag + ((size_t)(nb * p.H + ih) * p.W + iw) * p.C + ac;
      }
      cp_async16(sA + ((size_t)stage * BKP + row) * LDA + aj * 8, src, ok);
    }
  };

  float acc[MI][NJ][4];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NJ; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;

  const int KT = (m_end > m_begin) ? (m_end - m_begin + BKP - 1) / BKP : 0;
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < KT) load_stage(s, m_begin + s * BKP);
    cp_async_commit();
  }
  for (int k = 0; k < KT; ++k) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    if (k + STAGES - 1 < KT) load_stage((k + STAGES - 1) % STAGES, m_begin + (k + STAGES - 1) * BKP);
    cp_async_commit();
    const int stage = k % STAGES;
    const bf16* tG = sG + (size_t)stage * BKP * LDG;
    const bf16* tA = sA + (size_t)stage * BKP * LDA;
#pragma unroll
    for (int ks = 0; ks < BKP / 16; ++ks) {
      uint32_t af[MI][4], bfr[NJ][2];
#pragma unroll
      for (int mi = 0; mi < MI; ++mi) {
        const int kk = ks * 16 + (lane & 7) + (lane >> 4) * 8;
        const int nn = warp_m * WM + mi * 16 + ((lane >> 3) & 1) * 8;
        ldmatrix_x4_trans(af[mi][0], af[mi][1], af[mi][2], af[mi][3], tG + kk * LDG + nn);
      }
#pragma unroll
      for (int nj = 0; nj < NJ / 2; ++nj) {
        const int kk = ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int vv = warp_n * WN + nj * 16 + (lane >> 4) * 8;
        ldmatrix_x4_trans(bfr[2 * nj][0], bfr[2 * nj][1], bfr[2 * nj + 1][0], bfr[2 * nj + 1][1], tA + kk * LDA + vv);
      }
#pragma unroll
      for (int mi = 0; mi < MI; ++mi)
#pragma unroll
        for (int nj = 0; nj < NJ; ++nj) mma_bf16(acc[mi][nj], af[mi], bfr[nj]);
    }
  }
  cp_async_wait<0>();

  float* out = p.partial + (size_t)split * p.N * TC;
#pragma unroll
  for (int mi = 0; mi < MI; ++mi)
#pragma unroll
    for (int nj = 0; nj < NJ; ++nj) {
      const int n = n0 + warp_m * WM + mi * 16 + (lane >> 2);
      const int v = v0 + warp_n * WN + nj * 8 + (lane & 3) * 2;
      if (v < TC) {
        if (n < p.N) *reinterpret_cast<float2*>(&out[(size_t)n * TC + v]) = make_float2(acc[mi][nj][0], acc[mi][nj][1]);
        if (n + 8 < p.N)
          *reinterpret_cast<float2*>(&out[(size_t)(n + 8) * TC + v]) = make_float2(acc[mi][nj][2], acc[mi][nj][3]);
      }
    }
}

template <int BMN>
int launch_wgrad(const WgradParams& p, cudaStream_t st) {
  constexpr int BV = 128, BKP = 32, STAGES = 3;
  constexpr size_t smem = (size_t)STAGES * BKP * ((BMN + 8) + (BV + 8)) * sizeof(bf16);
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(igemm_wgrad_mma_kernel<BMN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured = true;
  }
  dim3 grid(p.splits, ceil_div(p.T * p.C, BV), ceil_div(p.N, BMN));
  sv_launch_pdl(igemm_wgrad_mma_kernel<BMN>, dim3(grid), dim3(256), smem, st, p);
  return sv_check_launch("igemm_wgrad_mma");
}

struct TapIdx {
  int8_t v[SV_MAX_TAPS];
};

// grid.y slices the splits: every thread sums up to 16 partial slices (independent loads in flight) and
// adds the result to the FP32 gradient with one atomic (which also gives the += accumulate semantics)
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ grad, int splits, int N, int C,
                                    int T, int n_real, int c_real, long long sn, long long sc, long long st, TapIdx ti) {
  pdl_trigger();
  pdl_wait();
  const long long total = (long long)n_real * T * c_real;
  const long long TC = (long long)T * C;
  const int k0 = blockIdx.y * 16, k1 = min(k0 + 16, splits);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c_real);
    const long long r = i / c_real;
    const int t = (int)(r % T);
    const int n = (int)(r / T);
    const float* src = partial + (long long)n * TC + (long long)t * C + c;
    float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (int k = k0; k < k1; ++k) s[k & 3] += src[(long long)k * N * TC];
    atomicAdd(&grad[n * sn + c * sc + ti.v[t] * st], (s[0] + s[1]) + (s[2] + s[3]));
  }
}

// Several weight tensors in ONE launch (sv_wgrad_reduce_batched): the descriptors travel as kernel parameters, block ranges
// [first[i], first[i+1]) belong to descriptor i; inside a range the blocks are (x = element slice, y = slice of 16 partial sums)
// exactly like the grid of wgrad_reduce_kernel.
struct ReduceDesc {          // == sv_wgrad_reduce_desc (include/shotvae.h)
  const float* partial;
  float* grad;
  long long sn, sc, st;
  int splits, N, C, T, n_real, c_real;
  int8_t tap[SV_MAX_TAPS];
};
static_assert(sizeof(ReduceDesc) == sizeof(sv_wgrad_reduce_desc), "sv_wgrad_reduce_desc layout");
constexpr int RB_MAX = 40;   // 40 x 80 B + block table = 3.4 KB of the 4 KB parameter space
struct ReduceBatch {
  ReduceDesc d[RB_MAX];
  int first[RB_MAX + 1];
  int n;
};

__global__ void __launch_bounds__(256) wgrad_reduce_batched_kernel(const __grid_constant__ ReduceBatch b) {
  int i = 0;
  while (i + 1 < b.n && (int)blockIdx.x >= b.first[i + 1]) ++i;
  const ReduceDesc& d = b.d[i];
  const int lb = blockIdx.x - b.first[i], nblk = b.first[i + 1] - b.first[i];
  const int ysl = (d.splits + 15) / 16, bxn = nblk / ysl;
  const int bx = lb % bxn, by = lb / bxn;
  const long long total = (long long)d.n_real * d.T * d.c_real;
  const long long TC = (long long)d.T * d.C;
  const int k0 = by * 16, k1 = min(k0 + 16, d.splits);
  for (long long e = bx * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)bxn * blockDim.x) {
    const int c = (int)(e % d.c_real);
    const long long r = e / d.c_real;
    const int t = (int)(r % d.T);
    const int n = (int)(r / d.T);
    const float* src = d.partial + (long long)n * TC + (long long)t * d.C + c;
    float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (int k = k0; k < k1; ++k) s[k & 3] += src[(long long)k * d.N * TC];
    atomicAdd(&d.grad[n * d.sn + c * d.sc + d.tap[t] * d.st], (s[0] + s[1]) + (s[2] + s[3]));
  }
}

__global__ void pack_weight_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int N, int C, int T, int n_real,
                                   int c_real, long long sn, long long sc, long long st, TapIdx ti, int layout) {
  const long long total = (long long)T * N * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long r = i / C;
    const int n = (int)(r % N);
    const int t = (int)(r / N);
    float v = 0.f;
    if (n < n_real && c < c_real) v = src[n * sn + c * sc + ti.v[t] * st];
    if (layout == 2) {                 // fp32 [T][N][C]: operands of the parity-grade FP32 kernels (igemm_f32.cu)
      reinterpret_cast<float*>(dst)[i] = v;
      continue;
    }
    const long long o = layout == 0 ? i : ((((long long)t * (C >> 3) + (c >> 3)) * N + n) << 3) + (c & 7);
    dst[o] = __float2bfloat16(v);
  }
}

// one launch for every weight tensor of the network: blockIdx.y selects the pack
struct PackDesc {
  const float* src;
  bf16* dst;
  long long sn, sc, st;
  int N, C, T, n_real, c_real, layout;
  int8_t tap[SV_MAX_TAPS];
};

__global__ void pack_weights_batched_kernel(const PackDesc* __restrict__ table) {
  const PackDesc d = table[blockIdx.y];
  // One thread per (output channel n, input channel c) PAIR, all taps in an inner loop: the taps of a pair are
  // adjacent in the source (OIHW / IOHW weights), so the thread's T reads stay inside one or two 32-byte sectors
  // (L1 hits after the first) instead of T different threads re-fetching the sector from L2.  The pair index is
  // decoded so that consecutive threads write consecutive destination elements:
  //   layout 0  [T][N][C]        : c fastest
  //   layout 1  [T][C/8][N][8]   : (c % 8) fastest, then n, then c / 8
  const long long pairs = (long long)d.N * d.C;
  const long long want = (pairs + 256 * 4 - 1) / (256 * 4);
  const int nblk = (int)(want < (long long)gridDim.x ? want : (long long)gridDim.x);
  if ((int)blockIdx.x >= nblk) return;
  const long long NC = (long long)d.N * d.C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < pairs; i += (long long)nblk * blockDim.x) {
    int n, c;
    long long o;           // destination offset of tap 0; taps are NC elements apart in both layouts
    if (d.layout == 0 || d.layout == 2) {
      c = (int)(i % d.C);
      n = (int)(i / d.C);
      o = i;
    } else {
      const int c_lo = (int)(i & 7);
      const long long r = i >> 3;
      n = (int)(r % d.N);
      const int c_hi = (int)(r / d.N);
      c = c_hi * 8 + c_lo;
      o = i;               // ((c_hi * N + n) * 8 + c_lo) == i by construction
    }
    const bool live = n < d.n_real && c < d.c_real;
    const float* src = d.src + n * d.sn + c * d.sc;
    for (int t = 0; t < d.T; ++t) {
      const float v = live ? src[d.tap[t] * d.st] : 0.f;
      if (d.layout == 2) reinterpret_cast<float*>(d.dst)[(long long)t * NC + o] = v;     // fp32 [T][N][C] (FP32 mode)
      else d.dst[(long long)t * NC + o] = __float2bfloat16(v);
    }
  }
}

}  // namespace

extern "C" int sv_sizeof_pack_desc(void) { return (int)sizeof(PackDesc); }

extern "C" int sv_pack_weights_batched(const void* table_dev, int32_t n_packs, int32_t blocks_per_pack, void* stream) {
  SV_REQUIRE(table_dev && n_packs > 0 && blocks_per_pack > 0, "sv_pack_weights_batched: bad arguments");
  pack_weights_batched_kernel<<<dim3(blocks_per_pack, n_packs), 256, 0, (cudaStream_t)stream>>>((const PackDesc*)table_dev);
  return sv_check_launch("pack_weights_batched");
}

int igemm_fprop_mma_any(const IgemmParams& p, const IgemmParams* batch, int nb, cudaStream_t st);

int igemm_fprop_mma(const IgemmParams& p, cudaStream_t st) { return igemm_fprop_mma_any(p, nullptr, 0, st); }

// batch[0..nb): problems of the same geometry (NB, H, W, C, N, OH, OW, T) -> one grid
int igemm_fprop_mma_batch(const IgemmParams* batch, int nb, cudaStream_t st) { return igemm_fprop_mma_any(batch[0], batch, nb, st); }

int igemm_fprop_mma_any(const IgemmParams& p, const IgemmParams* batch, int nb, cudaStream_t st) {
  const bool k32 = (p.C % 32) == 0;
  const int N = p.N;
  // long reductions (the decoder's input-gradient GEMMs: K = taps x C up to 4096) are latency bound on the number of
  // pipeline steps: 64-channel k-blocks halve them
  const bool k64 = (p.C % 64) == 0 && p.T * p.C >= 1024;
#define SV_DISPATCH(BNV)                                                     \
  if (batch != nullptr) {                                                    \
    if (k64) return launch_fprop_batch<BNV, 64>(batch, nb, st);              \
    return k32 ? launch_fprop_batch<BNV, 32>(batch, nb, st) : launch_fprop_batch<BNV, 16>(batch, nb, st); \
  }                                                                          \
  if (k64) return launch_fprop<BNV, 64>(p, st);                              \
  return k32 ? launch_fprop<BNV, 32>(p, st) : launch_fprop<BNV, 16>(p, st);
  // widest tile that still gives about one CTA per SM: the decoder's input-gradient GEMMs have M = 256 ... 4096
  // rows and K up to 4096, and ran 50-100 us on 16-64 CTAs (latency bound) with the widest tile
  const int m_tiles = ceil_div(p.M, 128);
  auto enough = [&](int bn) { return m_tiles * (N / bn) >= 120; };
  if (N % 128 == 0 && enough(128)) { SV_DISPATCH(128) }
  if (N % 64 == 0 && enough(64)) { SV_DISPATCH(64) }
  if (N % 32 == 0 && enough(32)) { SV_DISPATCH(32) }
  if (N % 16 == 0 && (enough(16) || N % 32 != 0)) { SV_DISPATCH(16) }
  if (N % 32 == 0) { SV_DISPATCH(32) }
  SV_DISPATCH(16)
#undef SV_DISPATCH
}

int igemm_wgrad_mma(const WgradParams& p, cudaStream_t st) {
  if (p.N % 128 == 0) return launch_wgrad<128>(p, st);
  if (p.N % 64 == 0) return launch_wgrad<64>(p, st);
  if (p.N % 32 == 0) return launch_wgrad<32>(p, st);
  return launch_wgrad<16>(p, st);
}

extern "C" int sv_wgrad_reduce(const float* partial, float* grad, int32_t splits, int32_t N, int32_t C, int32_t T,
                               int32_t n_real, int32_t c_real, int64_t sn, int64_t sc, int64_t st,
                               const int8_t* tap_index, void* stream) {
  SV_REQUIRE(T >= 1 && T <= SV_MAX_TAPS && splits >= 1, "sv_wgrad_reduce: bad T/splits");
  TapIdx ti;
  memcpy(ti.v, tap_index, T);
  const long long total = (long long)n_real * T * c_real;
  const int blocks = (int)((total + 255) / 256 > 148 * 8 ? 148 * 8 : (total + 255) / 256);
  sv_launch_pdl(wgrad_reduce_kernel, dim3(blocks, (splits + 15) / 16), dim3(256), 0, (cudaStream_t)stream, partial, grad, splits, N, C, T, n_real, c_real,
                                                                                          sn, sc, st, ti);
  return sv_check_launch("wgrad_reduce");
}

extern "C" int sv_sizeof_wgrad_reduce_desc() { return (int)sizeof(sv_wgrad_reduce_desc); }

extern "C" int sv_wgrad_reduce_batched(const sv_wgrad_reduce_desc* descs, int32_t n, void* stream) {
  SV_REQUIRE(descs != nullptr && n >= 0, "sv_wgrad_reduce_batched: bad arguments");
  for (int i0 = 0; i0 < n; i0 += RB_MAX) {
    ReduceBatch b;
    b.n = n - i0 < RB_MAX ? n - i0 : RB_MAX;
    int blocks = 0;
    for (int i = 0; i < b.n; ++i) {
      memcpy(&b.d[i], &descs[i0 + i], sizeof(ReduceDesc));
      const ReduceDesc& d = b.d[i];
      SV_REQUIRE(d.partial && d.grad && d.T >= 1 && d.T <= SV_MAX_TAPS && d.splits >= 1, "sv_wgrad_reduce_batched: bad descriptor");
      const long long total = (long long)d.n_real * d.T * d.c_real;
      const int bx = (int)((total + 255) / 256 > 148 * 4 ? 148 * 4 : (total + 255) / 256);
      b.first[i] = blocks;
      blocks += (bx > 0 ? bx : 1) * ((d.splits + 15) / 16);
    }
    b.first[b.n] = blocks;
    wgrad_reduce_batched_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(b);
    const int rc = sv_check_launch("wgrad_reduce_batched");
    if (rc != 0) return rc;
  }
  return 0;
}

extern "C" int sv_pack_weight(const float* src, void* dst, int32_t N, int32_t C, int32_t T, int32_t n_real,
                              int32_t c_real, int64_t sn, int64_t sc, int64_t st, const int8_t* tap_index, int32_t layout,
                              void* stream) {
  SV_REQUIRE(T >= 1 && T <= SV_MAX_TAPS, "sv_pack_weight: bad T");
  SV_REQUIRE(layout == 0 || layout == 2 || (layout == 1 && C % 8 == 0), "sv_pack_weight: bad layout");
  TapIdx ti;
  memcpy(ti.v, tap_index, T);
  const long long total = (long long)T * N * C;
  const int blocks = (int)((total + 255) / 256 > 148 * 8 ? 148 * 8 : (total + 255) / 256);
  pack_weight_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, (bf16*)dst, N, C, T, n_real, c_real, sn, sc, st, ti, layout);
  return sv_check_launch("pack_weight");
}
