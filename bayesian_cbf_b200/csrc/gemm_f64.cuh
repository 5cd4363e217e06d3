// FP64 tensor-core (DMMA.8x8x4) tiled GEMM used by the blocked Cholesky, the triangular inverse and the
// Linv products.  Row-major operands, 128x128x16 CTA tiles, 3-stage cp.async pipeline, 8 warps (2x4),
// 64x32 warp tiles -> 32 DMMA per k4-step per warp against 12 LDS.64 fragment loads; a 32x128x16 variant (1x8 warps)
// for problems of a few tiles, where the machine is otherwise idle and the time is one CTA's latency.
//
//   C[M,N] = alpha * opA(A) * opB(B) + beta * C        (batched over blockIdx.z with element strides)
//
//   A_KC = true : A stored (M,K) row-major (k contiguous)        false: A stored (K,M) (i.e. A^T is given)
//   B_KC = true : B stored (N,K) row-major ("NT", C = A B^T)      false: B stored (K,N) row-major ("NN")
//
// Triangular awareness is expressed as per-tile k-ranges and tile masks (all in units of 128-blocks):
//   kTriNone            every tile, k in [0,K)
//   kTriLowerOut        only tiles with tj <= ti are computed (SYRK trailing update of the Cholesky)
//   kTriALower          A is lower triangular (M == K):    k in [0, (ti+1)*128)
//   kTriAUpper          opA(A) is upper triangular:        k in [ti*128, K)
//   kTriBLower          B (K,N) is lower triangular:       k in [tj*128, K)
// Shared-memory tiles are padded so that the DMMA fragment loads (8 rows x 4 k per instruction) are
// bank-conflict free: row stride = 20 doubles (k-contiguous) / 132 doubles (mn-contiguous), both = 4 mod 16.
#pragma once
#include "common.cuh"

namespace bcbf {

enum GemmTri { kTriNone = 0, kTriLowerOut = 1, kTriALower = 2, kTriAUpper = 3, kTriBLower = 4 };

struct GemmArgs {
  const double* A;
  const double* B;
  double* C;
  int lda, ldb, ldc;
  int M, N, K;
  long long sA, sB, sC;  // batch strides in elements
  double alpha, beta;
  int tri;
  int mtiles, ntiles;
};

constexpr int kGemmBM = 128, kGemmBN = 128, kGemmBK = 16, kGemmStages = 3, kGemmThreads = 256;
constexpr int kGemmBMSmall = 32;         // row tile of the small-problem variant (see launch_gemm)
constexpr int kStrideKC = kGemmBK + 4;   // 20
constexpr int kStrideMN = kGemmBM + 4;   // 132
constexpr int kTileElems = 128 * kStrideKC;  // 2560 doubles >= 16*132 = 2112
__host__ __device__ constexpr int gemm_tile_elems(int rows) {
  return rows * kStrideKC > kGemmBK * (rows + 4) ? rows * kStrideKC : kGemmBK * (rows + 4);
}
__host__ __device__ constexpr int gemm_smem_bytes(int bm) {
  return kGemmStages * (gemm_tile_elems(bm) + gemm_tile_elems(kGemmBN)) * (int)sizeof(double);
}
constexpr int kGemmSmemBytes = gemm_smem_bytes(kGemmBM);  // 122880

template <bool KC, int R>
__device__ __forceinline__ void gemm_load_tile(double* s, const double* g, int ld, int r0, int k0, int rmax,
                                               int kmax, int tid) {
  // KC: tile rows r0..r0+R-1 (limit rmax) x k0..k0+15 (limit kmax), k contiguous in memory.
  // !KC: memory is (K, R): rows k0..k0+15 of length R, r contiguous.
#pragma unroll
  for (int i = 0; i < R * 8 / kGemmThreads; ++i) {
    int c = tid + i * kGemmThreads;
    if (KC) {
      int row = c >> 3, kc = (c & 7) * 2;
      bool ok = (r0 + row < rmax) && (k0 + kc < kmax);
      const double* src = g + (long long)(r0 + row) * ld + (k0 + kc);
      cp_async16(s + row * kStrideKC + kc, ok ? src : g, ok);
    } else {
      int k = c / (R / 2), rc = (c % (R / 2)) * 2;
      bool ok = (r0 + rc < rmax) && (k0 + k < kmax);
      const double* src = g + (long long)(k0 + k) * ld + (r0 + rc);
      cp_async16(s + k * (R + 4) + rc, ok ? src : g, ok);
    }
  }
}

// BM = 128: 2 x 4 warps, 64 x 32 warp tiles.  BM = 32 (small problems: a handful of 128 x 128 tiles would leave most of
// the 148 SMs idle while each CTA spends ~17 us per 128 k): 1 x 8 warps, 32 x 16 warp tiles, four times the CTAs, two
// CTAs per SM.  BN stays 128, so a CTA still owns complete rows of a 128-wide panel: the in-place panel solve of the
// Cholesky (C == A, N == K == 128) remains race-free.  Triangular k-ranges are in units of 128-blocks in both variants.
template <bool A_KC, bool B_KC, int BM>
__global__ void __launch_bounds__(kGemmThreads, BM == kGemmBM ? 1 : 2) gemm_f64_kernel(GemmArgs p) {
  constexpr int WARPS_M = BM == kGemmBM ? 2 : 1, WARPS_N = 8 / WARPS_M;
  constexpr int WTM = BM / WARPS_M, WTN = kGemmBN / WARPS_N, MI = WTM / 8, NJ = WTN / 8;
  constexpr int TA = gemm_tile_elems(BM), TB = gemm_tile_elems(kGemmBN);
  constexpr int SA_MN = BM + 4;            // row stride of an (k, m)-ordered A tile: = 4 mod 16 for BM = 32, 128
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp / WARPS_N, wn = warp % WARPS_N;
  int ti, tj;
  if (p.tri == kTriLowerOut && BM == kGemmBM) {
    // linear index over the lower triangle, largest rows first is not needed (uniform K)
    int t = blockIdx.x;
    int r = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
    while ((long long)(r + 1) * (r + 2) / 2 <= t) ++r;
    while ((long long)r * (r + 1) / 2 > t) --r;
    ti = r;
    tj = t - r * (r + 1) / 2;
  } else {
    ti = blockIdx.x / p.ntiles;
    tj = blockIdx.x % p.ntiles;
  }
  const int m0 = ti * BM, n0 = tj * kGemmBN;
  const int bi = m0 / kGemmBM;             // 128-block row of this tile
  if (p.tri == kTriLowerOut && tj > bi) return;   // (small variant: rectangular grid, tiles above the diagonal idle)
  const double* A = p.A + (long long)blockIdx.z * p.sA;
  const double* B = p.B + (long long)blockIdx.z * p.sB;
  double* C = p.C + (long long)blockIdx.z * p.sC;
  int kbeg = 0, kend = p.K;
  if (p.tri == kTriALower) kend = min(p.K, (bi + 1) * kGemmBM);
  if (p.tri == kTriAUpper) kbeg = bi * kGemmBM;
  if (p.tri == kTriBLower) kbeg = tj * kGemmBN;
  const int nk = (kend - kbeg + kGemmBK - 1) / kGemmBK;

  double acc[MI][NJ][2];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  auto stageA = [&](int s) { return smem + s * (TA + TB); };
  auto stageB = [&](int s) { return smem + s * (TA + TB) + TA; };
  auto load = [&](int kt, int s) {
    int k0 = kbeg + kt * kGemmBK;
    gemm_load_tile<A_KC, BM>(stageA(s), A, p.lda, m0, k0, p.M, kend, tid);
    gemm_load_tile<B_KC, kGemmBN>(stageB(s), B, p.ldb, n0, k0, p.N, kend, tid);
  };

#pragma unroll
  for (int s = 0; s < kGemmStages - 1; ++s) {
    if (s < nk) load(s, s);
    cp_async_commit();
  }
  const int lr = lane >> 2, lk = lane & 3;
  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<kGemmStages - 2>();
    __syncthreads();
    {
      int nxt = kt + kGemmStages - 1;
      if (nxt < nk) load(nxt, nxt % kGemmStages);
      cp_async_commit();
    }
    const double* As = stageA(kt % kGemmStages);
    const double* Bs = stageB(kt % kGemmStages);
#pragma unroll
    for (int k4 = 0; k4 < kGemmBK / 4; ++k4) {
      double a[MI], b[NJ];
      const int kk = k4 * 4 + lk;
#pragma unroll
      for (int i = 0; i < MI; ++i) {
        int row = wm * WTM + i * 8 + lr;
        a[i] = A_KC ? As[row * kStrideKC + kk] : As[kk * SA_MN + row];
      }
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        int col = wn * WTN + j * 8 + lr;
        b[j] = B_KC ? Bs[col * kStrideKC + kk] : Bs[kk * kStrideMN + col];
      }
#pragma unroll
      for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
  }
  cp_async_wait<0>();
  __syncthreads();  // every operand load of this CTA has landed: in-place C (== A or B block) is now safe

#pragma unroll
  for (int i = 0; i < MI; ++i) {
    int row = m0 + wm * WTM + i * 8 + lr;
    if (row >= p.M) continue;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      int col = n0 + wn * WTN + j * 8 + lk * 2;
      if (col >= p.N) continue;
      double* dst = C + (long long)row * p.ldc + col;
      double2 v;
      v.x = p.alpha * acc[i][j][0];
      v.y = p.alpha * acc[i][j][1];
      if (p.beta != 0.0) {
        double2 old = *reinterpret_cast<const double2*>(dst);
        v.x += p.beta * old.x;
        v.y += p.beta * old.y;
      }
      *reinterpret_cast<double2*>(dst) = v;
    }
  }
}

// 0: choose by problem size (default), 1: always 128-row tiles, 2: always the small variant (tests, A/B timing)
inline int& gemm_tile_policy() {
  static int policy = 0;
  return policy;
}

// Host launcher.  Returns a cudaError_t-like int through BCBF conventions in the callers.
template <bool A_KC, bool B_KC>
inline cudaError_t launch_gemm(GemmArgs a, int batch, cudaStream_t stream) {
  a.mtiles = (a.M + kGemmBM - 1) / kGemmBM;
  a.ntiles = (a.N + kGemmBN - 1) / kGemmBN;
  if (a.mtiles <= 0 || a.ntiles <= 0 || batch <= 0) return cudaSuccess;
  long long tiles = (a.tri == kTriLowerOut) ? (long long)a.mtiles * (a.mtiles + 1) / 2
                                             : (long long)a.mtiles * a.ntiles;
  // under half a wave of 128 x 128 tiles: quarter the row tile (4x the CTAs, 2 per SM)
  const int policy = gemm_tile_policy();
  const bool small = policy == 2 || (policy == 0 && tiles * batch <= 74);
  if (small) {
    constexpr int smem = gemm_smem_bytes(kGemmBMSmall);
    cudaError_t e = cudaFuncSetAttribute(gemm_f64_kernel<A_KC, B_KC, kGemmBMSmall>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    a.mtiles = (a.M + kGemmBMSmall - 1) / kGemmBMSmall;
    dim3 grid((unsigned)((long long)a.mtiles * a.ntiles), 1, (unsigned)batch);
    gemm_f64_kernel<A_KC, B_KC, kGemmBMSmall><<<grid, kGemmThreads, smem, stream>>>(a);
    ++g_launch_count;
    return cudaGetLastError();
  }
  cudaError_t e = cudaFuncSetAttribute(gemm_f64_kernel<A_KC, B_KC, kGemmBM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       kGemmSmemBytes);
  if (e != cudaSuccess) return e;
  dim3 grid((unsigned)tiles, 1, (unsigned)batch);
  gemm_f64_kernel<A_KC, B_KC, kGemmBM><<<grid, kGemmThreads, kGemmSmemBytes, stream>>>(a);
  ++g_launch_count;
  return cudaGetLastError();
}

}  // namespace bcbf
