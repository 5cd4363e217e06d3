// Fused construction of the control-affine Gram matrices.
//   gram_train : Kb[i,j] = s exp(-1/2 |(x_i-x_j)/l|^2) * (uh_i^T B uh_j)           (reference :370-372)
//   cross_gram : Kstar[i,q] = s exp(-1/2 |(x_i-xq_q)/l|^2)                          (reference :536, :1051)
//   rbf_blocks : k, dk/dx1, d2k/dx1 dx2 in closed form                              (reference :465-477)
// One CTA produces a 64x64 tile (256 threads, 4x4 outputs per thread); the scaled inputs x/l and the
// G = UH B rows of the tile are staged in shared memory; every output row segment is written with 16-byte
// stores, 64 consecutive doubles (512 B) per tile row.  HBM-write bound; exp() is the co-limiter.
#include "../../include/bcbf.h"
#include "common.cuh"

namespace bcbf {

constexpr int kGT = 64;       // tile edge
constexpr int kMaxN = BCBF_MAX_N_DIM;
constexpr int kMaxP = BCBF_MAX_P_DIM;

struct GramParams {
  const double* X1;  // rows (a, n)
  const double* X2;  // cols (c, n)
  const double* UH;   // rows: (a, p) or null
  const double* UH2;  // cols: (c, p) or null (train mode: == UH)
  double inv_ls[kMaxN];       // 1 / lengthscale, filled on the host when the caller's lengthscale lives in host memory
  double Bm[kMaxP * kMaxP];   // B, likewise
  const double* ls_dev;       // non-null: lengthscale (n) in DEVICE memory, read by the kernel itself (no host fetch,
  const double* B_dev;        // no stream synchronisation: the torch-tensor call path); same for B (p, p)
  double scale;
  int a, c, n, p;
  double* out;
  int ld;
  int rows_out, cols_out;  // padded extents to fill (>= a, >= c)
  int pad_identity;        // train mode: identity on the pad diagonal
  int vec_ok;              // 16-byte stores allowed (even ld, aligned base)
};

// Every kernel starts by staging 1 / lengthscale and B in shared memory, from the parameter block or from device memory.
__device__ __forceinline__ void load_hyper(const GramParams& P, double* inv_ls, double* Bm) {
  const int t = threadIdx.x;
  if (t < kMaxN) inv_ls[t] = t < P.n ? (P.ls_dev ? 1.0 / P.ls_dev[t] : P.inv_ls[t]) : 0.0;
  if (t >= 32 && t < 32 + kMaxP * kMaxP) {
    const int e = t - 32;
    Bm[e] = e < P.p * P.p ? (P.B_dev ? P.B_dev[e] : P.Bm[e]) : 0.0;
  }
  __syncthreads();
}
#define BCBF_LOAD_HYPER(P)                                   \
  __shared__ double h_inv_ls[kMaxN], h_Bm[kMaxP * kMaxP];    \
  load_hyper(P, h_inv_ls, h_Bm)

// ---- one Gram entry; the ONE definition every kernel that needs Kb[i,j] uses, so that the train Gram, the compensated
// residual (gram_resid_kernel) and the ensembles see bit-identical values.  Every operation is spelled out (no
// compiler-chosen contraction): d2 = fma chain over the state dimensions of (x_i/l - x_j/l)^2, k = s * exp(-d2/2),
// S = fma chain over q of g_i[q] * uh_j[q], entry = k * S.
template <int NN>
__device__ __forceinline__ double rbf_entry(const double* __restrict__ xr, const double* __restrict__ xc, int n,
                                            double scale) {
  const int nn = NN > 0 ? NN : n;
  double d2 = 0.0;
#pragma unroll
  for (int d = 0; d < kMaxN; ++d)
    if (d < nn) {
      const double df = __dsub_rn(xr[d], xc[d]);
      d2 = __fma_rn(df, df, d2);
    }
  return __dmul_rn(scale, exp(__dmul_rn(-0.5, d2)));
}
template <int PP>
__device__ __forceinline__ double ub_entry(const double* __restrict__ g, const double* __restrict__ u, int p) {
  const int pp = PP > 0 ? PP : p;
  double ub = 0.0;
#pragma unroll
  for (int q = 0; q < kMaxP; ++q)
    if (q < pp) ub = __fma_rn(g[q], u[q], ub);
  return ub;
}

template <bool TRAIN>
__global__ void __launch_bounds__(256) gram_kernel(GramParams P) {
  BCBF_LOAD_HYPER(P);
  __shared__ double xr[kGT][kMaxN + 1], xc[kGT][kMaxN + 1];
  __shared__ double gr[kGT][kMaxP], uc[kGT][kMaxP];
  const int tid = threadIdx.x;
  const int r0 = blockIdx.y * kGT, c0 = blockIdx.x * kGT;
  const int n = P.n, p = P.p;
  for (int idx = tid; idx < kGT * n; idx += 256) {
    int r = idx / n, d = idx % n;
    xr[r][d] = (r0 + r < P.a) ? __dmul_rn(P.X1[(long long)(r0 + r) * n + d], h_inv_ls[d]) : 0.0;
    xc[r][d] = (c0 + r < P.c) ? __dmul_rn(P.X2[(long long)(c0 + r) * n + d], h_inv_ls[d]) : 0.0;
  }
  if (TRAIN) {
    for (int idx = tid; idx < kGT * p; idx += 256) {
      int r = idx / p, q = idx % p;
      gr[r][q] = (r0 + r < P.a) ? g_entry(P.UH + (long long)(r0 + r) * p, h_Bm, p, q) : 0.0;
      uc[r][q] = (c0 + r < P.c) ? P.UH2[(long long)(c0 + r) * p + q] : 0.0;
    }
  }
  __syncthreads();
  const int ty = tid >> 4, tx = tid & 15;  // 16 x 16 threads; thread owns rows ty*4..+3, cols tx*4..+3
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int rl = ty * 4 + i, row = r0 + rl;
    if (row >= P.rows_out) continue;
    double v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int cl = tx * 4 + j, col = c0 + cl;
      double out = 0.0;
      if (row < P.a && col < P.c) {
        out = rbf_entry<0>(xr[rl], xc[cl], n, P.scale);
        if (TRAIN) out = __dmul_rn(out, ub_entry<0>(gr[rl], uc[cl], p));
      } else if (TRAIN && P.pad_identity && row == col) {
        out = 1.0;
      }
      v[j] = out;
    }
    const int col = c0 + tx * 4;
    double* dst = P.out + (long long)row * P.ld + col;
    if (P.vec_ok && col + 3 < P.cols_out) {
      *reinterpret_cast<double2*>(dst) = make_double2(v[0], v[1]);
      *reinterpret_cast<double2*>(dst + 2) = make_double2(v[2], v[3]);
    } else {
      for (int j = 0; j < 4; ++j)
        if (col + j < P.cols_out) dst[j] = v[j];
    }
  }
}

// lower tile index t -> (bi, bj), bj <= bi, row-major over the lower triangle
__device__ __forceinline__ void lower_tile(int t, int& bi, int& bj) {
  bi = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
  while ((long long)(bi + 1) * (bi + 2) / 2 <= t) ++bi;
  while ((long long)bi * (bi + 1) / 2 > t) --bi;
  bj = t - bi * (bi + 1) / 2;
}

// ---- the train Gram of the fit path: compile-time state / control dimensions, operands of the thread's 4x4 outputs in
// registers (the generic kernel above re-reads shared memory inside its run-time loops: 2.8 ms at N = 16384), and with
// LOWER only the 64x64 tiles on or below the diagonal are produced (half the exps and half the bytes: the factorisation
// reads nothing else; diagonal tiles are written whole).  Entry (i, j), j <= i, is rbf(i,j) * (g_i . uh_j); the upper
// part of a full matrix mirrors it (entry(j, i) evaluated as rbf(j,i) * (g_j . uh_i), the same bits as its mirror image).
template <int NN, int PP, bool LOWER>
__global__ void __launch_bounds__(256) gram_train_kernel(GramParams P) {
  BCBF_LOAD_HYPER(P);
  __shared__ double xr[kGT][kMaxN + 1], xc[kGT][kMaxN + 1];
  __shared__ double gr[kGT][kMaxP + 1], ur[kGT][kMaxP + 1], gc[kGT][kMaxP + 1], uc[kGT][kMaxP + 1];
  const int tid = threadIdx.x;
  int bi, bj;
  if (LOWER) lower_tile(blockIdx.x, bi, bj);
  else { bi = blockIdx.y; bj = blockIdx.x; }
  const int r0 = bi * kGT, c0 = bj * kGT;
  const int n = NN > 0 ? NN : P.n, p = PP > 0 ? PP : P.p;
  for (int idx = tid; idx < kGT * n; idx += 256) {
    int r = idx / n, d = idx % n;
    xr[r][d] = (r0 + r < P.a) ? __dmul_rn(P.X1[(long long)(r0 + r) * n + d], h_inv_ls[d]) : 0.0;
    xc[r][d] = (c0 + r < P.a) ? __dmul_rn(P.X1[(long long)(c0 + r) * n + d], h_inv_ls[d]) : 0.0;
  }
  for (int idx = tid; idx < kGT * p; idx += 256) {
    int r = idx / p, q = idx % p;
    const bool vr = r0 + r < P.a, vc = c0 + r < P.a;
    gr[r][q] = vr ? g_entry(P.UH + (long long)(r0 + r) * p, h_Bm, p, q) : 0.0;
    ur[r][q] = vr ? P.UH[(long long)(r0 + r) * p + q] : 0.0;
    gc[r][q] = vc ? g_entry(P.UH + (long long)(c0 + r) * p, h_Bm, p, q) : 0.0;
    uc[r][q] = vc ? P.UH[(long long)(c0 + r) * p + q] : 0.0;
  }
  __syncthreads();
  const int ty = tid >> 4, tx = tid & 15;
  // strictly-lower tile: rows give g, columns give uh.  strictly-upper tile (full mode only): the mirror image, rows
  // give uh and columns give g (same products in the same order).  diagonal tile: chosen per entry.
  const bool upper = bi < bj, diag = bi == bj;
  double xa[4][kMaxN], xb[4][kMaxN], ga[4][kMaxP], ub[4][kMaxP];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int d = 0; d < kMaxN; ++d)
      if (d < n) { xa[i][d] = xr[ty * 4 + i][d]; xb[i][d] = xc[tx * 4 + i][d]; }
#pragma unroll
    for (int q = 0; q < kMaxP; ++q)
      if (q < p) {
        ga[i][q] = upper ? ur[ty * 4 + i][q] : gr[ty * 4 + i][q];
        ub[i][q] = upper ? gc[tx * 4 + i][q] : uc[tx * 4 + i][q];
      }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int rl = ty * 4 + i, row = r0 + rl;
    double v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int cl = tx * 4 + j, col = c0 + cl;
      double out = 0.0;
      if (row < P.a && col < P.a) {
        out = rbf_entry<NN>(xa[i], xb[j], n, P.scale);
        double s = ub_entry<PP>(ga[i], ub[j], p);
        if (diag && col > row) s = ub_entry<PP>(ur[rl], gc[cl], p);
        out = __dmul_rn(out, s);
      } else if (row == col) {
        out = 1.0;   // identity on the pad diagonal
      }
      v[j] = out;
    }
    if (row < P.rows_out) {
      double* dst = P.out + (long long)row * P.ld + c0 + tx * 4;
      *reinterpret_cast<double2*>(dst) = make_double2(v[0], v[1]);
      *reinterpret_cast<double2*>(dst + 2) = make_double2(v[2], v[3]);
    }
  }
}

// ---- compensated residual  R = Y - (Kb + jscale diag(jitter)) alpha  with Kb re-evaluated on the fly --------------------
// The iterative refinement of alpha = Kb^-1 Y (reference: cholesky_solve, control_affine_model.py:545) needs the residual
// against the matrix that was factorised, in more than working precision: Kb's condition number is 1e10..1e13 at the
// bench shapes and the posterior mean is a sum with 1e6-fold cancellation.  Kb is not kept (the factor overwrites it);
// its entries are recomputed here with the same device functions as gram_train_kernel — bit-identical — and every
// product Kb[i,k] alpha[k,c] is accumulated exactly-rounded twice: TwoProduct (one FMA) + TwoSum into a (hi, lo) pair per
// row and column (Ogita-Rump-Oishi Dot2: the result is as if computed in ~106-bit arithmetic and rounded once).
// One CTA = 64 rows x one column range; thread = 4 x 4 entries per 64 x 64 tile; the 16 threads of a row are merged by
// shuffles, the column ranges by a fixed-order (hi, lo) sum in gram_resid_finalize_kernel: deterministic.
__device__ __forceinline__ void dd_add_prod(double& hi, double& lo, double a, double b) {
  const double pr = __dmul_rn(a, b);
  const double pe = __fma_rn(a, b, -pr);            // a*b = pr + pe exactly
  const double s = __dadd_rn(hi, pr);
  const double bb = __dsub_rn(s, hi);
  const double se = __dadd_rn(__dsub_rn(hi, __dsub_rn(s, bb)), __dsub_rn(pr, bb));   // hi + pr = s + se exactly
  hi = s;
  lo = __dadd_rn(lo, __dadd_rn(pe, se));
}
__device__ __forceinline__ void dd_add(double& hi, double& lo, double h2, double l2) {
  const double s = __dadd_rn(hi, h2);
  const double bb = __dsub_rn(s, hi);
  const double se = __dadd_rn(__dsub_rn(hi, __dsub_rn(s, bb)), __dsub_rn(h2, bb));
  hi = s;
  lo = __dadd_rn(lo, __dadd_rn(l2, se));
}

constexpr int kResMaxC = 4;   // columns of alpha handled per pass (n <= 4 in one pass; larger n loops)

// STORED: Kb is read from P.out (lower triangle valid, leading dimension P.ld: what bcbf_gram_train_lower wrote, so the
// same bits as the on-the-fly evaluation) instead of being re-evaluated — the residual is then bound by the 2 N^2 x 8 B of
// the two triangle passes and the Dot2 arithmetic, not by N^2 exponentials.  Same accumulation order: same result bits.
template <int NN, int PP, bool STORED = false>
__global__ void __launch_bounds__(256) gram_resid_kernel(GramParams P, const double* __restrict__ jitter, double jscale,
                                                         const double* __restrict__ alpha, int lda, int cfirst, int nc,
                                                         int tiles_per_split, double* __restrict__ partial) {
  BCBF_LOAD_HYPER(P);
  __shared__ double xr[kGT][kMaxN + 1], xc[kGT][kMaxN + 1];
  __shared__ double gr[kGT][kMaxP + 1], ur[kGT][kMaxP + 1], gc[kGT][kMaxP + 1], uc[kGT][kMaxP + 1];
  __shared__ double al[kGT][kResMaxC];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int bi = blockIdx.x, r0 = bi * kGT;
  const int n = NN > 0 ? NN : P.n, p = PP > 0 ? PP : P.p;
  const int ntiles = (P.a + kGT - 1) / kGT;
  const int t0 = blockIdx.y * tiles_per_split, t1 = min(ntiles, t0 + tiles_per_split);
  if (!STORED) {
    for (int idx = tid; idx < kGT * n; idx += 256) {
      int r = idx / n, d = idx % n;
      xr[r][d] = (r0 + r < P.a) ? __dmul_rn(P.X1[(long long)(r0 + r) * n + d], h_inv_ls[d]) : 0.0;
    }
    for (int idx = tid; idx < kGT * p; idx += 256) {
      int r = idx / p, q = idx % p;
      const bool vr = r0 + r < P.a;
      gr[r][q] = vr ? g_entry(P.UH + (long long)(r0 + r) * p, h_Bm, p, q) : 0.0;
      ur[r][q] = vr ? P.UH[(long long)(r0 + r) * p + q] : 0.0;
    }
  }
  double hi[4][kResMaxC], lo[4][kResMaxC];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int c = 0; c < kResMaxC; ++c) hi[i][c] = lo[i][c] = 0.0;
  for (int bj = t0; bj < t1; ++bj) {
    const int c0 = bj * kGT;
    __syncthreads();
    if (!STORED) {
      for (int idx = tid; idx < kGT * n; idx += 256) {
        int r = idx / n, d = idx % n;
        xc[r][d] = (c0 + r < P.a) ? __dmul_rn(P.X1[(long long)(c0 + r) * n + d], h_inv_ls[d]) : 0.0;
      }
      for (int idx = tid; idx < kGT * p; idx += 256) {
        int r = idx / p, q = idx % p;
        const bool vc = c0 + r < P.a;
        gc[r][q] = vc ? g_entry(P.UH + (long long)(c0 + r) * p, h_Bm, p, q) : 0.0;
        uc[r][q] = vc ? P.UH[(long long)(c0 + r) * p + q] : 0.0;
      }
    }
    for (int idx = tid; idx < kGT * kResMaxC; idx += 256) {
      int r = idx / kResMaxC, c = idx % kResMaxC;
      al[r][c] = (c0 + r < P.a && c < nc) ? alpha[(long long)(c0 + r) * lda + cfirst + c] : 0.0;
    }
    __syncthreads();
    const bool upper = bi < bj, diag = bi == bj;
    if (STORED) {
      // entry (row, col) of the symmetric matrix from its stored lower triangle; the 4 x 4 block of a thread is fetched
      // before the arithmetic so that the 16 loads are in flight together
      double kv[4][4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int row = r0 + ty * 4 + i, col = c0 + tx * 4 + j;
          double k = 0.0;
          if (row < P.a && col < P.a)
            k = (col <= row) ? __ldg(P.out + (long long)row * P.ld + col) : __ldg(P.out + (long long)col * P.ld + row);
          kv[i][j] = k;
        }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int cl = tx * 4 + j, col = c0 + cl;
        double av[kResMaxC];
#pragma unroll
        for (int c = 0; c < kResMaxC; ++c) av[c] = al[cl][c];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int row = r0 + ty * 4 + i;
          if (row < P.a && col < P.a) {
            double k = kv[i][j];
            if (diag && col == row && jitter != nullptr) k = __fma_rn(jscale, jitter[row], k);
#pragma unroll
            for (int c = 0; c < kResMaxC; ++c)
              if (c < nc) dd_add_prod(hi[i][c], lo[i][c], k, av[c]);
          }
        }
      }
      continue;
    }
    double xa[4][kMaxN], xb[4][kMaxN], ga[4][kMaxP], ub[4][kMaxP];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
      for (int d = 0; d < kMaxN; ++d)
        if (d < n) { xa[i][d] = xr[ty * 4 + i][d]; xb[i][d] = xc[tx * 4 + i][d]; }
#pragma unroll
      for (int q = 0; q < kMaxP; ++q)
        if (q < p) {
          ga[i][q] = upper ? ur[ty * 4 + i][q] : gr[ty * 4 + i][q];
          ub[i][q] = upper ? gc[tx * 4 + i][q] : uc[tx * 4 + i][q];
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int cl = tx * 4 + j, col = c0 + cl;
      double av[kResMaxC];
#pragma unroll
      for (int c = 0; c < kResMaxC; ++c) av[c] = al[cl][c];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int rl = ty * 4 + i, row = r0 + rl;
        if (row < P.a && col < P.a) {
          double k = rbf_entry<NN>(xa[i], xb[j], n, P.scale);
          double s = ub_entry<PP>(ga[i], ub[j], p);
          if (diag && col > row) s = ub_entry<PP>(ur[rl], gc[cl], p);
          k = __dmul_rn(k, s);
          if (diag && col == row && jitter != nullptr) k = __fma_rn(jscale, jitter[row], k);   // as potf2_inv_kernel adds it
#pragma unroll
          for (int c = 0; c < kResMaxC; ++c)
            if (c < nc) dd_add_prod(hi[i][c], lo[i][c], k, av[c]);
        }
      }
    }
  }
  // merge the 16 threads (tx) that share a row: xor-shuffles inside the half-warp, fixed pattern
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int c = 0; c < kResMaxC; ++c) {
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        const double h2 = __shfl_xor_sync(0xffffffffu, hi[i][c], o);
        const double l2 = __shfl_xor_sync(0xffffffffu, lo[i][c], o);
        dd_add(hi[i][c], lo[i][c], h2, l2);
      }
      if (tx == 0 && c < nc) {
        double* dst = partial + (((long long)blockIdx.y * gridDim.x * kGT + r0 + ty * 4 + i) * kResMaxC + c) * 2;
        dst[0] = hi[i][c];
        dst[1] = lo[i][c];
      }
    }
}

__global__ void gram_resid_finalize_kernel(const double* __restrict__ partial, int nsplit, int rows_padded, int N,
                                           const double* __restrict__ Y, int ldy, int cfirst, int nc,
                                           double* __restrict__ R, int ldr) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows_padded * kResMaxC) return;
  const int row = idx / kResMaxC, c = idx % kResMaxC;
  if (c >= nc) return;
  double hi = 0.0, lo = 0.0;
  if (row < N) {
    for (int s = 0; s < nsplit; ++s) {
      const double* src = partial + (((long long)s * rows_padded + row) * kResMaxC + c) * 2;
      dd_add(hi, lo, src[0], src[1]);
    }
    // y - (hi + lo): y and hi agree to many digits once alpha is close, so y - hi is exact (Sterbenz) or nearly so
    R[(long long)row * ldr + cfirst + c] = __dsub_rn(__dsub_rn(Y[(long long)row * ldy + cfirst + c], hi), lo);
  }
}

// out[i,j] = K[i,j] * (uh1_i^T B uh2_j): the control-affine weighting of a data-kernel matrix that some OTHER module
// evaluated (the plug-in contract of HetergeneousMatrixVariateKernel: any data_covar_module).  One thread per entry.
__global__ void ca_weight_kernel(const double* __restrict__ K, int ldk, const double* __restrict__ UH1,
                                 const double* __restrict__ UH2, GramParams P, double* __restrict__ out, int ldo) {
  BCBF_LOAD_HYPER(P);
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)P.a * P.c) return;
  const int i = (int)(idx / P.c), j = (int)(idx % P.c), p = P.p;
  double s = 0.0;
  for (int q = 0; q < p; ++q) s = __fma_rn(g_entry(UH1 + (long long)i * p, h_Bm, p, q), UH2[(long long)j * p + q], s);
  out[(long long)i * ldo + j] = __dmul_rn(K[(long long)i * ldk + j], s);
}

// k, dk/dx1 (a,c,n), d2k/dx1dx2 (a,c,n,n); one thread per (i,j) pair — small-b API path only.
__global__ void rbf_blocks_kernel(GramParams P, double* K, double* dK, double* d2K) {
  BCBF_LOAD_HYPER(P);
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)P.a * P.c) return;
  const int i = (int)(idx / P.c), j = (int)(idx % P.c), n = P.n;
  double w[kMaxN];  // (x1 - x2) / l^2
  double d2 = 0.0;
  for (int d = 0; d < n; ++d) {
    double df = (P.X1[(long long)i * n + d] - P.X2[(long long)j * n + d]) * h_inv_ls[d];
    d2 = fma(df, df, d2);
    w[d] = df * h_inv_ls[d];
  }
  const double k = P.scale * exp(-0.5 * d2);
  if (K) K[idx] = k;
  if (dK)
    for (int d = 0; d < n; ++d) dK[idx * n + d] = -w[d] * k;
  if (d2K)
    for (int d = 0; d < n; ++d)
      for (int e = 0; e < n; ++e)
        d2K[(idx * n + d) * n + e] = ((d == e ? h_inv_ls[d] * h_inv_ls[d] : 0.0) - w[d] * w[e]) * k;
}


// ---- hyper-parameter gradient of the control-affine Gram matrix (the backward of gram_train, used by fit) ---------
// Given the adjoint  Gbar_ij = 1/2 [ (alpha A^-1)_i . alpha_j - n P_ij ],  P = Kb^-1  (never materialised: formed on the
// fly from P, alphaAi, alpha), accumulate  sum_ij Gbar_ij dKb_ij/dtheta  for theta = outputscale, lengthscale_d, B_ab:
//     dKb/ds = e S,   dKb/dl_d = Kb (dx_d / l_d)^2 / l_d,   dKb/dB_ab = s e uh_ia uh_jb,
// with e = exp(-1/2 |dx/l|^2), S = uh_i^T B uh_j.  One CTA per 64x64 tile writes a partial vector; a second kernel
// sums the partials in a fixed order (deterministic).  Replaces autograd through gpytorch's lazy kernel in
// ExactMarginalLogLikelihood (control_affine_model.py:309-325).
constexpr int kGradMaxOut = 1 + kMaxN + kMaxP * kMaxP;

__global__ void __launch_bounds__(256) gram_backward_kernel(GramParams P, const double* __restrict__ Pinv, int ldp,
                                                            const double* __restrict__ alphaAi,
                                                            const double* __restrict__ alpha, int lda, int nout_dim,
                                                            double* __restrict__ partial) {
  BCBF_LOAD_HYPER(P);
  __shared__ double xr[kGT][kMaxN + 1], xc[kGT][kMaxN + 1];
  __shared__ double ur[kGT][kMaxP], uc[kGT][kMaxP], gr[kGT][kMaxP];
  __shared__ double ar[kGT][kMaxN], ac[kGT][kMaxN];
  __shared__ double red[8][kGradMaxOut];
  const int tid = threadIdx.x;
  const int r0 = blockIdx.y * kGT, c0 = blockIdx.x * kGT;
  const int n = P.n, p = P.p, nd = nout_dim;
  for (int idx = tid; idx < kGT * n; idx += 256) {
    int r = idx / n, d = idx % n;
    xr[r][d] = (r0 + r < P.a) ? P.X1[(long long)(r0 + r) * n + d] * h_inv_ls[d] : 0.0;
    xc[r][d] = (c0 + r < P.c) ? P.X1[(long long)(c0 + r) * n + d] * h_inv_ls[d] : 0.0;
  }
  for (int idx = tid; idx < kGT * nd; idx += 256) {
    int r = idx / nd, d = idx % nd;
    ar[r][d] = (r0 + r < P.a) ? alphaAi[(long long)(r0 + r) * lda + d] : 0.0;
    ac[r][d] = (c0 + r < P.c) ? alpha[(long long)(c0 + r) * lda + d] : 0.0;
  }
  for (int idx = tid; idx < kGT * p; idx += 256) {
    int r = idx / p, q = idx % p;
    double g = 0.0, u = 0.0;
    if (r0 + r < P.a) {
      u = P.UH[(long long)(r0 + r) * p + q];
      for (int t = 0; t < p; ++t) g += P.UH[(long long)(r0 + r) * p + t] * h_Bm[t * p + q];
    }
    ur[r][q] = u;
    gr[r][q] = g;
    uc[r][q] = (c0 + r < P.c) ? P.UH[(long long)(c0 + r) * p + q] : 0.0;
  }
  __syncthreads();
  double acc[kGradMaxOut];
#pragma unroll
  for (int t = 0; t < kGradMaxOut; ++t) acc[t] = 0.0;
  const int ty = tid >> 4, tx = tid & 15;
  for (int i = 0; i < 4; ++i) {
    const int rl = ty * 4 + i, row = r0 + rl;
    if (row >= P.a) continue;
    for (int j = 0; j < 4; ++j) {
      const int cl = tx * 4 + j, col = c0 + cl;
      if (col >= P.c) continue;
      double d2 = 0.0, dd[kMaxN];
      for (int d = 0; d < n; ++d) {
        double df = xr[rl][d] - xc[cl][d];
        dd[d] = df * df;
        d2 += dd[d];
      }
      const double e = exp(-0.5 * d2);
      double S = 0.0;
      for (int q = 0; q < p; ++q) S = fma(gr[rl][q], uc[cl][q], S);
      double aa = 0.0;
      for (int d = 0; d < nd; ++d) aa = fma(ar[rl][d], ac[cl][d], aa);
      const double gbar = 0.5 * (aa - (double)nd * Pinv[(long long)row * ldp + col]);
      const double ge = gbar * e;
      acc[0] += ge * S;                                       // d/d outputscale
      const double gk = ge * S * P.scale;                     // Gbar * Kb
      for (int d = 0; d < n; ++d) acc[1 + d] += gk * dd[d] * h_inv_ls[d];   // (dx/l)^2 / l
      const double gs = ge * P.scale;
      for (int a = 0; a < p; ++a)
        for (int b = 0; b < p; ++b) acc[1 + kMaxN + a * kMaxP + b] += gs * ur[rl][a] * uc[cl][b];
    }
  }
  // CTA reduction: warp shuffles then 8 warp rows in shared memory, fixed order
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int t = 0; t < kGradMaxOut; ++t) {
    double v = warp_sum(acc[t]);
    if (lane == 0) red[warp][t] = v;
  }
  __syncthreads();
  if (tid < kGradMaxOut) {
    double v = 0.0;
    for (int w = 0; w < 8; ++w) v += red[w][tid];
    partial[((long long)blockIdx.y * gridDim.x + blockIdx.x) * kGradMaxOut + tid] = v;
  }
}

__global__ void gram_backward_finalize_kernel(const double* __restrict__ partial, int nblocks, double* __restrict__ out) {
  // one warp per output entry, strided fixed-order accumulation + shuffle tree: deterministic
  const int t = blockIdx.x, lane = threadIdx.x;
  double v = 0.0;
  for (int b = lane; b < nblocks; b += 32) v += partial[(long long)b * kGradMaxOut + t];
  v = warp_sum(v);
  if (lane == 0) out[t] = v;
}

// Small parameter blocks (lengthscale, B) may live on the host or on the device.  Host pointers are copied into the
// kernel's parameter block; device pointers are handed to the kernel, which reads them itself — no copy back, no stream
// synchronisation on the torch-tensor call path (round 1 fetched them with a D2H copy + sync per call: the dominant cost of
// the small-N regimes).
static int resolve_small(double* dst, const double** dev_out, const double* src, int count) {
  cudaPointerAttributes attr{};
  cudaError_t e = cudaPointerGetAttributes(&attr, src);
  if (e != cudaSuccess) { cudaGetLastError(); attr.type = cudaMemoryTypeUnregistered; }
  if (attr.type == cudaMemoryTypeUnregistered || attr.type == cudaMemoryTypeHost) {
    for (int i = 0; i < count; ++i) dst[i] = src[i];
    *dev_out = nullptr;
  } else {
    *dev_out = src;
  }
  return BCBF_OK;
}

static int fill_common(GramParams& P, const double* lengthscale, double outputscale, int n, cudaStream_t stream) {
  (void)stream;
  double ls[kMaxN] = {1, 1, 1, 1, 1, 1, 1, 1};
  int rc = resolve_small(ls, &P.ls_dev, lengthscale, n);
  if (rc != BCBF_OK) return rc;
  for (int d = 0; d < n; ++d) P.inv_ls[d] = 1.0 / ls[d];
  P.scale = outputscale;
  P.n = n;
  return BCBF_OK;
}

}  // namespace bcbf

using namespace bcbf;

// dispatch on the compile-time (n, p) pairs of the reference's systems (unicycle 3/3, pendulum 2/2); anything else runs
// the same kernel with run-time extents
template <bool LOWER>
static void launch_gram_train(const GramParams& P, dim3 grid, cudaStream_t stream) {
  if (P.n == 3 && P.p == 3) gram_train_kernel<3, 3, LOWER><<<grid, 256, 0, stream>>>(P);
  else if (P.n == 2 && P.p == 2) gram_train_kernel<2, 2, LOWER><<<grid, 256, 0, stream>>>(P);
  else gram_train_kernel<0, 0, LOWER><<<grid, 256, 0, stream>>>(P);
}

static int gram_train_impl(const double* X, const double* UH, const double* Bmat, const double* lengthscale,
                           double outputscale, int N, int n, int p, double* Kb, int ld, int Npad, bool lower,
                           cudaStream_t stream) {
  BCBF_REQUIRE(X && UH && Bmat && lengthscale && Kb, "bcbf_gram_train: null pointer");
  BCBF_REQUIRE(n >= 1 && n <= kMaxN && p >= 1 && p <= kMaxP, "bcbf_gram_train: n=%d (<=%d) p=%d (<=%d)", n, kMaxN,
               p, kMaxP);
  BCBF_REQUIRE(N >= 1 && Npad >= N && Npad % kGT == 0 && ld >= Npad && ld % 2 == 0 &&
                   (reinterpret_cast<uintptr_t>(Kb) & 15) == 0,
               "bcbf_gram_train: N=%d Npad=%d (multiple of %d) ld=%d (even), Kb 16-byte aligned", N, Npad, kGT, ld);
  GramParams P{};
  int rc = fill_common(P, lengthscale, outputscale, n, stream);
  if (rc != BCBF_OK) return rc;
  if ((rc = resolve_small(P.Bm, &P.B_dev, Bmat, p * p)) != BCBF_OK) return rc;
  P.X1 = X; P.X2 = X; P.UH = UH; P.UH2 = UH; P.a = N; P.c = N; P.p = p;
  P.out = Kb; P.ld = ld; P.rows_out = Npad; P.cols_out = Npad; P.pad_identity = 1; P.vec_ok = 1;
  const int nt = Npad / kGT;
  if (lower) launch_gram_train<true>(P, dim3(nt * (nt + 1) / 2), stream);
  else launch_gram_train<false>(P, dim3(nt, nt), stream);
  BCBF_LAUNCH_CHECK();
  return BCBF_OK;
}

extern "C" int bcbf_gram_train(const double* X, const double* UH, const double* Bmat, const double* lengthscale,
                               double outputscale, int N, int n, int p, double* Kb, int ld, int Npad,
                               void* stream_) {
  return gram_train_impl(X, UH, Bmat, lengthscale, outputscale, N, n, p, Kb, ld, Npad, false,
                         static_cast<cudaStream_t>(stream_));
}

extern "C" int bcbf_gram_train_lower(const double* X, const double* UH, const double* Bmat, const double* lengthscale,
                                     double outputscale, int N, int n, int p, double* Kb, int ld, int Npad,
                                     void* stream_) {
  return gram_train_impl(X, UH, Bmat, lengthscale, outputscale, N, n, p, Kb, ld, Npad, true,
                         static_cast<cudaStream_t>(stream_));
}

// column splits of the residual: enough CTAs for ~6 waves of the machine, at most one split per column tile
static void resid_splits(int N, int* nsplit, int* tiles_per_split) {
  const int T = (N + kGT - 1) / kGT;
  int S = (888 + T - 1) / T;
  if (S > T) S = T;
  if (S < 1) S = 1;
  const int tps = (T + S - 1) / S;
  *tiles_per_split = tps;
  *nsplit = (T + tps - 1) / tps;
}

extern "C" long long bcbf_gram_resid_scratch_elems(int N) {
  int S, tps;
  resid_splits(N, &S, &tps);
  const long long T = (N + kGT - 1) / kGT;
  return (long long)S * T * kGT * kResMaxC * 2;
}

extern "C" int bcbf_gram_resid(const double* X, const double* UH, const double* Bmat, const double* lengthscale,
                               double outputscale, int N, int n, int p, const double* jitter, double jitter_scale,
                               const double* alpha, int lda, const double* Y, int ldy, int nc, double* R, int ldr,
                               double* scratch, long long scratch_elems, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BCBF_REQUIRE(X && UH && Bmat && lengthscale && alpha && Y && R && scratch, "bcbf_gram_resid: null pointer");
  BCBF_REQUIRE(n >= 1 && n <= kMaxN && p >= 1 && p <= kMaxP && N >= 1 && nc >= 1 && nc <= kMaxN && lda >= nc &&
                   ldy >= nc && ldr >= nc,
               "bcbf_gram_resid: N=%d n=%d p=%d nc=%d lda=%d ldy=%d ldr=%d", N, n, p, nc, lda, ldy, ldr);
  BCBF_REQUIRE(scratch_elems >= bcbf_gram_resid_scratch_elems(N), "bcbf_gram_resid: scratch too small (%lld < %lld)",
               scratch_elems, bcbf_gram_resid_scratch_elems(N));
  GramParams P{};
  int rc = fill_common(P, lengthscale, outputscale, n, stream);
  if (rc != BCBF_OK) return rc;
  if ((rc = resolve_small(P.Bm, &P.B_dev, Bmat, p * p)) != BCBF_OK) return rc;
  P.X1 = X; P.X2 = X; P.UH = UH; P.UH2 = UH; P.a = N; P.c = N; P.p = p;
  int S, tps;
  resid_splits(N, &S, &tps);
  const int T = (N + kGT - 1) / kGT;
  for (int cfirst = 0; cfirst < nc; cfirst += kResMaxC) {
    const int ncp = (nc - cfirst) < kResMaxC ? (nc - cfirst) : kResMaxC;
    dim3 grid(T, S);
    if (n == 3 && p == 3) gram_resid_kernel<3, 3><<<grid, 256, 0, stream>>>(P, jitter, jitter_scale, alpha, lda, cfirst, ncp, tps, scratch);
    else if (n == 2 && p == 2) gram_resid_kernel<2, 2><<<grid, 256, 0, stream>>>(P, jitter, jitter_scale, alpha, lda, cfirst, ncp, tps, scratch);
    else gram_resid_kernel<0, 0><<<grid, 256, 0, stream>>>(P, jitter, jitter_scale, alpha, lda, cfirst, ncp, tps, scratch);
    BCBF_LAUNCH_CHECK();
    gram_resid_finalize_kernel<<<ceil_div((long long)T * kGT * kResMaxC, 256), 256, 0, stream>>>(
        scratch, S, T * kGT, N, Y, ldy, cfirst, ncp, R, ldr);
    BCBF_LAUNCH_CHECK();
  }
  return BCBF_OK;
}

extern "C" int bcbf_gram_resid_stored(const double* Kb, int ldk, int N, const double* jitter, double jitter_scale,
                                      const double* alpha, int lda, const double* Y, int ldy, int nc, double* R, int ldr,
                                      double* scratch, long long scratch_elems, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BCBF_REQUIRE(Kb && alpha && Y && R && scratch, "bcbf_gram_resid_stored: null pointer");
  BCBF_REQUIRE(N >= 1 && ldk >= N && nc >= 1 && nc <= kMaxN && lda >= nc && ldy >= nc && ldr >= nc,
               "bcbf_gram_resid_stored: N=%d ldk=%d nc=%d lda=%d ldy=%d ldr=%d", N, ldk, nc, lda, ldy, ldr);
  BCBF_REQUIRE(scratch_elems >= bcbf_gram_resid_scratch_elems(N), "bcbf_gram_resid_stored: scratch too small (%lld < %lld)",
               scratch_elems, bcbf_gram_resid_scratch_elems(N));
  GramParams P{};
  P.a = N; P.c = N; P.n = 1; P.p = 1;
  P.out = const_cast<double*>(Kb);   // read only in the STORED instantiation
  P.ld = ldk;
  int S, tps;
  resid_splits(N, &S, &tps);
  const int T = (N + kGT - 1) / kGT;
  for (int cfirst = 0; cfirst < nc; cfirst += kResMaxC) {
    const int ncp = (nc - cfirst) < kResMaxC ? (nc - cfirst) : kResMaxC;
    dim3 grid(T, S);
    gram_resid_kernel<1, 1, true><<<grid, 256, 0, stream>>>(P, jitter, jitter_scale, alpha, lda, cfirst, ncp, tps, scratch);
    BCBF_LAUNCH_CHECK();
    gram_resid_finalize_kernel<<<ceil_div((long long)T * kGT * kResMaxC, 256), 256, 0, stream>>>(
        scratch, S, T * kGT, N, Y, ldy, cfirst, ncp, R, ldr);
    BCBF_LAUNCH_CHECK();
  }
  return BCBF_OK;
}

extern "C" int bcbf_cross_gram(const double* X, const double* Xq, const double* lengthscale, double outputscale,
                               int N, int Q, int n, double* Kstar, int ldks, int Npad, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BCBF_REQUIRE(X && Xq && lengthscale && Kstar, "bcbf_cross_gram: null pointer");
  BCBF_REQUIRE(n >= 1 && n <= kMaxN, "bcbf_cross_gram: n=%d", n);
  BCBF_REQUIRE(N >= 1 && Q >= 1 && Npad >= N && ldks >= Q && ldks % 2 == 0, "bcbf_cross_gram: N=%d Q=%d ldks=%d",
               N, Q, ldks);
  GramParams P{};
  int rc = fill_common(P, lengthscale, outputscale, n, stream);
  if (rc != BCBF_OK) return rc;
  P.X1 = X; P.X2 = Xq; P.UH = nullptr; P.a = N; P.c = Q; P.p = 0;
  P.out = Kstar; P.ld = ldks; P.rows_out = Npad; P.cols_out = ldks; P.pad_identity = 0; P.vec_ok = 1;
  dim3 grid(ceil_div(ldks, kGT), ceil_div(Npad, kGT));
  gram_kernel<false><<<grid, 256, 0, stream>>>(P);
  BCBF_LAUNCH_CHECK();
  return BCBF_OK;
}

extern "C" int bcbf_gram_ca(const double* X1, const double* UH1, int a, const double* X2, const double* UH2, int c,
                            const double* Bmat, const double* lengthscale, double outputscale, int n, int p,
                            double* out, int ld, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BCBF_REQUIRE(X1 && X2 && lengthscale && out, "bcbf_gram_ca: null pointer");
  BCBF_REQUIRE((UH1 == nullptr) == (UH2 == nullptr), "bcbf_gram_ca: UH1 and UH2 must both be given or both be NULL");
  BCBF_REQUIRE(n >= 1 && n <= kMaxN && a >= 1 && c >= 1 && ld >= c, "bcbf_gram_ca: a=%d c=%d n=%d ld=%d", a, c, n, ld);
  BCBF_REQUIRE(UH1 == nullptr || (Bmat && p >= 1 && p <= kMaxP), "bcbf_gram_ca: p=%d / Bmat", p);
  GramParams P{};
  int rc = fill_common(P, lengthscale, outputscale, n, stream);
  if (rc != BCBF_OK) return rc;
  if (UH1 && (rc = resolve_small(P.Bm, &P.B_dev, Bmat, p * p)) != BCBF_OK) return rc;
  P.X1 = X1; P.X2 = X2; P.UH = UH1; P.UH2 = UH2; P.a = a; P.c = c; P.p = UH1 ? p : 0;
  P.out = out; P.ld = ld; P.rows_out = a; P.cols_out = c; P.pad_identity = 0;
  dim3 grid(ceil_div(c, kGT), ceil_div(a, kGT));
  // 16-byte stores need an even leading dimension and an aligned base; otherwise the scalar tail path is taken
  P.vec_ok = (ld % 2 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  if (UH1) gram_kernel<true><<<grid, 256, 0, stream>>>(P);
  else gram_kernel<false><<<grid, 256, 0, stream>>>(P);
  BCBF_LAUNCH_CHECK();
  return BCBF_OK;
}

extern "C" int bcbf_ca_weight(const double* K, int ldk, const double* UH1, int a, const double* UH2, int c,
                              const double* Bmat, int p, double* out, int ldo, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BCBF_REQUIRE(K && UH1 && UH2 && Bmat && out, "bcbf_ca_weight: null pointer");
  BCBF_REQUIRE(a >= 1 && c >= 1 && p >= 1 && p <= kMaxP && ldk >= c && ldo >= c, "bcbf_ca_weight: a=%d c=%d p=%d ldk=%d ldo=%d",
               a, c, p, ldk, ldo);
  GramParams P{};
  int rc = resolve_small(P.Bm, &P.B_dev, Bmat, p * p);
  if (rc != BCBF_OK) return rc;
  P.a = a; P.c = c; P.p = p;
  ca_weight_kernel<<<ceil_div((long long)a * c, 256), 256, 0, stream>>>(K, ldk, UH1, UH2, P, out, ldo);
  BCBF_LAUNCH_CHECK();
  return BCBF_OK;
}

extern "C" int bcbf_rbf_blocks(const double* X1, const double* X2, const double* lengthscale, double outputscale,
                               int a, int c, int n, double* K, double* dK, double* d2K, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BCBF_REQUIRE(X1 && X2 && lengthscale, "bcbf_rbf_blocks: null pointer");
  BCBF_REQUIRE(n >= 1 && n <= kMaxN && a >= 1 && c >= 1, "bcbf_rbf_blocks: a=%d c=%d n=%d", a, c, n);
  GramParams P{};
  int rc = fill_common(P, lengthscale, outputscale, n, stream);
  if (rc != BCBF_OK) return rc;
  P.X1 = X1; P.X2 = X2; P.a = a; P.c = c;
  long long total = (long long)a * c;
  rbf_blocks_kernel<<<ceil_div(total, 256), 256, 0, stream>>>(P, K, dK, d2K);
  BCBF_LAUNCH_CHECK();
  return BCBF_OK;
}

extern "C" int bcbf_gram_train_backward(const double* X, const double* UH, const double* Bmat, const double* lengthscale,
                                        double outputscale, int N, int n, int p, const double* Pinv, int ldp,
                                        const double* alphaAi, const double* alpha, int lda, int nout, double* partial,
                                        long long partial_elems, double* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BCBF_REQUIRE(X && UH && Bmat && lengthscale && Pinv && alphaAi && alpha && partial && out,
               "bcbf_gram_train_backward: null pointer");
  BCBF_REQUIRE(n >= 1 && n <= kMaxN && p >= 1 && p <= kMaxP && nout >= 1 && nout <= kMaxN && N >= 1 && ldp >= N &&
                   lda >= nout,
               "bcbf_gram_train_backward: N=%d n=%d p=%d nout=%d ldp=%d lda=%d", N, n, p, nout, ldp, lda);
  GramParams P{};
  int rc = fill_common(P, lengthscale, outputscale, n, stream);
  if (rc != BCBF_OK) return rc;
  if ((rc = resolve_small(P.Bm, &P.B_dev, Bmat, p * p)) != BCBF_OK) return rc;
  P.X1 = X; P.X2 = X; P.UH = UH; P.UH2 = UH; P.a = N; P.c = N; P.p = p;
  dim3 grid(ceil_div(N, kGT), ceil_div(N, kGT));
  const long long nblocks = (long long)grid.x * grid.y;
  BCBF_REQUIRE(partial_elems >= nblocks * kGradMaxOut, "bcbf_gram_train_backward: partial buffer too small (%lld < %lld)",
               partial_elems, nblocks * kGradMaxOut);
  gram_backward_kernel<<<grid, 256, 0, stream>>>(P, Pinv, ldp, alphaAi, alpha, lda, nout, partial);
  BCBF_LAUNCH_CHECK();
  gram_backward_finalize_kernel<<<kGradMaxOut, 32, 0, stream>>>(partial, (int)nblocks, out);
  BCBF_LAUNCH_CHECK();
  return BCBF_OK;
}

extern "C" int bcbf_gram_backward_layout(int* out_elems, int* max_n, int* max_p) {
  if (out_elems) *out_elems = kGradMaxOut;
  if (max_n) *max_n = kMaxN;
  if (max_p) *max_p = kMaxP;
  return BCBF_OK;
}
