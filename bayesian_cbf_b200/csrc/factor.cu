// Blocked FP64 Cholesky (right-looking, 128-wide panels) and triangular inverse.
//   potf2_inv_kernel : one CTA factorises a 128x128 diagonal block in shared memory and inverts it
//   panel / trailing : DMMA GEMMs from gemm_f64.cuh  (A_ik <- A_ik Dinv_k^T ;  A_ij -= L_ik L_jk^T)
//   trtri            : divide-and-conquer  inv([[L11,0],[L21,L22]]) = [[X11,0],[-X22 L21 X11, X22]],
//                      every level is two batched triangular-aware DMMA GEMMs.
// Replaces torch.linalg.cholesky / cholesky_solve / linalg.solve on the factor
// (reference control_affine_model.py:907-911, 545, 565, 1053).
#include "../../include/bcbf.h"
#include "gemm_f64.cuh"

namespace bcbf {

constexpr int kPad = kBlk + 1;  // 129: conflict-free row and column walks of the 128x128 smem block

constexpr int kSB = 16;  // sub-panel width of the in-CTA factorisation (16: the unrolled register code stays ~2k SASS
                         // instructions and lives in the instruction cache; at 32 it was 30k and thrashed it)
constexpr int kSBT = kSB / 4;             // 4x4 register tiles per sub-block edge
constexpr int kTElems = (kBlk / kSB - 1) * kSB * (kSB + 1);  // scratch: up to 112 x 17 doubles

// X(r,c) of the inverse under construction: strictly-lower entries live transposed in the upper triangle of `a`,
// the diagonal in xd.
__device__ __forceinline__ double inv_get(const double* a, const double* xd, int r, int c) {
  return r == c ? xd[r] : (r > c ? a[c * kPad + r] : 0.0);
}

// One CTA factorises a 128x128 diagonal block in shared memory and inverts it.  Blocked by 32 columns:
//   A1  the 32x32 diagonal sub-block is factorised by ONE WARP in registers (lane = row), pivots and multipliers
//       exchanged with warp shuffles — no block barrier inside the 32-step chain;
//   A2  the rows below are solved against it, one thread per row (forward substitution in registers);
//   A3  the trailing sub-matrix gets the rank-32 update with 4x4 register tiles on all 256 threads;
//   B   inverse: the four 32x32 diagonal blocks by four warps (lane = column), then the off-diagonal blocks level by
//       level, X_ij = -X_ii (sum_k L_ik X_kj), as small register-tiled products.
// 12 + 8 block barriers instead of 128 * 5.  Batched over blockIdx.x (ensemble of small factors).
__global__ void __launch_bounds__(256, 1)
potf2_inv_kernel(double* __restrict__ A_, int ld, int k0, const double* __restrict__ jitter_, int N, double jscale,
                 double* __restrict__ dinv_, int* __restrict__ info_, long long sA, long long sD, long long sJ) {
  extern __shared__ __align__(16) double sm[];
  double* a = sm;                 // [128][129]
  double* xd = sm + kBlk * kPad;  // [128] diagonal of the inverse
  double* T = xd + kBlk;          // [kTElems] scratch: panel solve / off-diagonal inverse blocks, row stride kSB+1
  __shared__ int failed;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* A = A_ + (long long)blockIdx.x * sA;
  double* dinv = dinv_ + (long long)blockIdx.x * sD;
  const double* jitter = jitter_ ? jitter_ + (long long)blockIdx.x * sJ : nullptr;
  int* info = info_ + blockIdx.x;
  if (tid == 0) failed = 0;
  // load the block: branch-free 16-byte loads, 8 in flight per thread; the upper triangle is masked to zero
#pragma unroll 8
  for (int it = 0; it < kBlk * kBlk / 2 / 256; ++it) {
    const int idx2 = it * 256 + tid, r = idx2 >> 6, c = (idx2 & 63) * 2;
    const double2 v = *reinterpret_cast<const double2*>(A + (long long)(k0 + r) * ld + (k0 + c));
    a[r * kPad + c] = (c <= r) ? v.x : 0.0;
    a[r * kPad + c + 1] = (c + 1 <= r) ? v.y : 0.0;
  }
  __syncthreads();
  if (tid < kBlk && jitter != nullptr && (k0 + tid) < N)   // one rounding, as gram_resid_kernel re-creates the diagonal
    a[tid * kPad + tid] = __fma_rn(jscale, jitter[k0 + tid], a[tid * kPad + tid]);
  __syncthreads();

  // ======================= Phase A: L L^T = block =======================================================
  for (int pnl = 0; pnl < kBlk / kSB; ++pnl) {
    const int c0 = pnl * kSB;
    if (warp == 0) {
      // ---- A1: factor the 16x16 diagonal sub-block in REGISTERS (lane = row, lanes >= 16 idle): pivots and multipliers
      //      travel by warp shuffle, no shared-memory round trip and no barrier on the dependent chain; one rsqrt per
      //      pivot, no divisions.  Fully unrolled so that every register index is a compile-time constant.
      double r[kSB];
      const int rl = lane & (kSB - 1);
#pragma unroll
      for (int k = 0; k < kSB; ++k) r[k] = (k <= rl) ? a[(c0 + rl) * kPad + c0 + k] : 0.0;
      int bad = 0;
      double myinv = 0.0;  // 1 / L[lane][lane]
#pragma unroll
      for (int j = 0; j < kSB; ++j) {
        double djj = __shfl_sync(0xffffffffu, r[j], j);
        if (!(djj > 0.0)) {  // also catches NaN; uniform across the warp
          if (bad == 0) bad = c0 + j + 1;
          djj = __longlong_as_double(0x7ff8000000000000LL);
        }
        const double inv = rsqrt(djj);
        if (rl == j) { r[j] = djj * inv; myinv = inv; }   // sqrt(djj) to within an ulp
        else if (rl > j) r[j] = r[j] * inv;
#pragma unroll
        for (int k = 0; k < kSB; ++k) {
          if (k > j) {  // compile-time after unrolling
            const double lkj = __shfl_sync(0xffffffffu, r[j], k);
            if (rl >= k) r[k] = fma(-r[j], lkj, r[k]);
          }
        }
      }
      if (lane < kSB) {
#pragma unroll
        for (int k = 0; k < kSB; ++k)
          if (k <= lane) a[(c0 + lane) * kPad + c0 + k] = r[k];
      }
      if (lane == 0 && bad != 0) {
        atomicCAS(info, 0, k0 + bad);
        failed = 1;
      }
      // ---- A1b: invert it right away (lane = column j of X = L_D^-1), still in registers.  X is needed by the panel
      //      solve below AND is the diagonal block of the final inverse.
      double x[kSB];
#pragma unroll
      for (int i = 0; i < kSB; ++i) {
        double sacc = 0.0;
#pragma unroll
        for (int k = 0; k < kSB; ++k) {
          if (k < i) {
            const double lik = __shfl_sync(0xffffffffu, r[k], i);   // L[i][k]
            if (k >= rl) sacc = fma(lik, x[k], sacc);
          }
        }
        const double dii = __shfl_sync(0xffffffffu, myinv, i);
        x[i] = (i == rl) ? dii : (i > rl ? -sacc * dii : 0.0);
      }
      if (lane < kSB) {
#pragma unroll
        for (int i = 0; i < kSB; ++i) {
          if (i == lane) xd[c0 + lane] = x[i];
          else if (i > lane) a[(c0 + lane) * kPad + c0 + i] = x[i];   // X[i][j] (i > j) kept at a[j][i]
        }
      }
    }
    __syncthreads();
    if (failed) break;
    const int r0 = c0 + kSB, Tn = kBlk - r0;  // trailing extent
    {  // A2: panel <- panel * X^T  (P[i][j] = sum_{k <= j} A[i][c0+k] X[j][k]), 4 rows x 4 columns per thread
      const int ntask = (Tn / 4) * kSBT;
      for (int t = tid; t < ntask; t += 256) {
        const int ti = t / kSBT, tj = t % kSBT;
        double acc[4][4];
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
          for (int f = 0; f < 4; ++f) acc[e][f] = 0.0;
        const int jmax = 4 * tj + 3;
        for (int k = 0; k <= jmax; ++k) {
          double va[4], vx[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            va[e] = a[(r0 + 4 * ti + e) * kPad + c0 + k];
            vx[e] = inv_get(a, xd, c0 + 4 * tj + e, c0 + k);   // X[j][k], zero for k > j
          }
#pragma unroll
          for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int f = 0; f < 4; ++f) acc[e][f] = fma(va[e], vx[f], acc[e][f]);
        }
        // every task reads A[i][c0 .. c0+jmax] of its own 4 rows only, but other tasks (other tj) read the same rows:
        // stage the results and write after the barrier
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
          for (int f = 0; f < 4; ++f) T[(4 * ti + e) * (kSB + 1) + 4 * tj + f] = acc[e][f];
      }
      __syncthreads();
      for (int t = tid; t < Tn * kSB; t += 256) {
        const int i = t / kSB, j = t % kSB;
        a[(r0 + i) * kPad + c0 + j] = T[i * (kSB + 1) + j];
      }
    }
    __syncthreads();
    {  // A3: trailing (lower) -= P P^T with 4x4 register tiles
      const int nt = Tn / 4, ntiles = nt * (nt + 1) / 2;
      for (int t = tid; t < ntiles; t += 256) {
        int ti = (int)((sqrtf(8.0f * t + 1.0f) - 1.0f) * 0.5f);
        while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
        while (ti * (ti + 1) / 2 > t) --ti;
        const int tj = t - ti * (ti + 1) / 2;
        const double* pa = a + (r0 + 4 * ti) * kPad + c0;
        const double* pb = a + (r0 + 4 * tj) * kPad + c0;
        double acc[4][4];
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
          for (int f = 0; f < 4; ++f) acc[e][f] = 0.0;
#pragma unroll 8
        for (int k = 0; k < kSB; ++k) {
          double va[4], vb[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            va[e] = pa[e * kPad + k];
            vb[e] = pb[e * kPad + k];
          }
#pragma unroll
          for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int f = 0; f < 4; ++f) acc[e][f] = fma(va[e], vb[f], acc[e][f]);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
          for (int f = 0; f < 4; ++f) {
            const int i = r0 + 4 * ti + e, j = r0 + 4 * tj + f;
            if (j <= i) a[i * kPad + j] -= acc[e][f];
          }
      }
    }
    __syncthreads();
  }
  __syncthreads();
  if (failed) {
    // leave NaNs so that nothing downstream looks plausible
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    for (int idx = tid; idx < kBlk * kBlk; idx += 256) {
      int r = idx >> 7, c = idx & 127;
      A[(long long)(k0 + r) * ld + (k0 + c)] = nan;
      dinv[idx] = nan;
    }
    return;
  }

  // ======================= Phase B: X = L^{-1}: the diagonal 32x32 blocks are already there (A1b) ===========
  for (int d = 1; d < kBlk / kSB; ++d) {  // B2: block (bi, bj) with bi - bj = d
    const int nblk = kBlk / kSB - d;
    // T_b = sum_{kb = bj}^{bi-1} L[bi, kb] X[kb, bj]      (32 x 32 each), 4x4 register tiles
    for (int t = tid; t < nblk * kSBT * kSBT; t += 256) {
      const int b = t / (kSBT * kSBT), tt = t % (kSBT * kSBT), ti = tt / kSBT, tj = tt % kSBT;
      const int bj = b, bi = b + d;
      double acc[4][4];
#pragma unroll
      for (int e = 0; e < 4; ++e)
#pragma unroll
        for (int f = 0; f < 4; ++f) acc[e][f] = 0.0;
      for (int k = bj * kSB; k < bi * kSB; ++k) {
        double va[4], vb[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          va[e] = a[(bi * kSB + 4 * ti + e) * kPad + k];
          vb[e] = inv_get(a, xd, k, bj * kSB + 4 * tj + e);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
          for (int f = 0; f < 4; ++f) acc[e][f] = fma(va[e], vb[f], acc[e][f]);
      }
#pragma unroll
      for (int e = 0; e < 4; ++e)
#pragma unroll
        for (int f = 0; f < 4; ++f) T[(b * kSB + 4 * ti + e) * (kSB + 1) + 4 * tj + f] = acc[e][f];
    }
    __syncthreads();
    // X[bi, bj] = -X[bi, bi] T_b
    for (int t = tid; t < nblk * kSBT * kSBT; t += 256) {
      const int b = t / (kSBT * kSBT), tt = t % (kSBT * kSBT), ti = tt / kSBT, tj = tt % kSBT;
      const int bj = b, bi = b + d;
      double acc[4][4];
#pragma unroll
      for (int e = 0; e < 4; ++e)
#pragma unroll
        for (int f = 0; f < 4; ++f) acc[e][f] = 0.0;
      for (int k = 0; k < kSB; ++k) {
        double va[4], vb[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          va[e] = inv_get(a, xd, bi * kSB + 4 * ti + e, bi * kSB + k);
          vb[e] = T[(b * kSB + k) * (kSB + 1) + 4 * tj + e];
        }
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
          for (int f = 0; f < 4; ++f) acc[e][f] = fma(va[e], vb[f], acc[e][f]);
      }
#pragma unroll
      for (int e = 0; e < 4; ++e)
#pragma unroll
        for (int f = 0; f < 4; ++f) {
          const int r = bi * kSB + 4 * ti + e, c = bj * kSB + 4 * tj + f;  // r > c always (d >= 1)
          a[c * kPad + r] = -acc[e][f];
        }
    }
    __syncthreads();
  }
#pragma unroll 4
  for (int idx2 = tid; idx2 < kBlk * kBlk / 2; idx2 += 256) {      // two columns per thread: 16-byte stores
    const int r = idx2 >> 6, c = (idx2 & 63) * 2;
    double2 l, x;
    l.x = (c <= r) ? a[r * kPad + c] : 0.0;
    l.y = (c + 1 <= r) ? a[r * kPad + c + 1] : 0.0;
    x.x = (c < r) ? a[c * kPad + r] : (c == r ? xd[r] : 0.0);
    x.y = (c + 1 < r) ? a[(c + 1) * kPad + r] : (c + 1 == r ? xd[r] : 0.0);
    *reinterpret_cast<double2*>(A + (long long)(k0 + r) * ld + (k0 + c)) = l;
    *reinterpret_cast<double2*>(dinv + r * kBlk + c) = x;
  }
}

// ======================================================================================================================
// potf2_inv2_kernel (round 2): the same 128x128 factor + inverse, re-organised around what the ncu source view of the kernel
// above showed (profiles/r02f_potf2_ncu_summary.json): 46 % of its 172 us went to a column-oriented inverse whose k-loops
// grow to 112 steps while 16..112 of 256 threads have work, 18 % to 7 warps waiting behind the single-warp 16x16 factor +
// inverse of every panel.
//   A  factor, 16-wide panels.  Warp 0 owns the diagonal blocks: it applies the pending rank-16 update to the NEXT
//      diagonal block itself and factorises it (registers, shuffles) WHILE warps 1..7 run the trailing update of the
//      current panel — the 16-pivot chain is off the other warps' path.  The panel solve is a forward substitution per
//      row (no inverse of the diagonal block needed), so the 16x16 inverses leave the serial chain:
//   A' all eight 16x16 diagonal inverses at once, one warp each.
//   B  inverse by recursive doubling, X21 = -X22 (L21 X11) for blocks of 16, 32, 64: three levels of two products on
//      8 x 8 DMMA tiles (inv_level).
// Same storage convention (L in the lower triangle of `a`, strictly-lower X transposed into the upper triangle, diagonal of
// X in xd), same outputs, same failure protocol.
constexpr int kTS = 65;                      // row stride of the 64 x 64 scratch of phase B
constexpr int kT2Elems = 64 * kTS;

// One level of the recursive-doubling inverse, X21 = -X22 (L21 X11) for every pair of half-size H, on the FP64 tensor
// pipe: 8 x 8 output tiles, one warp per tile at a time, DMMA.8x8x4 with the operands gathered straight from the packed
// storage (L below the diagonal, X transposed above it, diag(X) in xd).  Against the 4 x 4 register-tile FMA version this
// issues a quarter of the instructions per multiply-add, which is what a 2-warps-per-scheduler kernel is short of
// (round-2 ncu: phase B was 34 % of the kernel at 12x its FMA floor).  Two accumulator pairs (even / odd k-steps) halve
// the dependent DMMA chain; the triangular operands shorten the k range (k >= column tile for X11, k <= row tile for X22).
template <int H>
__device__ __forceinline__ void inv_level(double* __restrict__ a, const double* __restrict__ xd, double* __restrict__ T,
                                          int tid) {
  constexpr int NPR = kBlk / (2 * H), TT = H / 8, NT = NPR * TT * TT;
  const int warp = tid >> 5, lane = tid & 31, lr = lane >> 2, lk = lane & 3;
  for (int t = warp; t < NT; t += 8) {       // T_p = L21_p X11_p
    const int pr = t / (TT * TT), tt = t % (TT * TT), ti = tt / TT, tj = tt % TT;
    const int base = pr * 2 * H;
    const int row = base + H + 8 * ti + lr;  // A fragment: L21[row][k]
    const int col = base + 8 * tj + lr;      // B fragment: X11[k][col]
    double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;
#pragma unroll 4
    for (int k4 = 2 * tj; k4 < H / 4; k4 += 2) {
      const int ka = base + 4 * k4 + lk, kb = ka + 4;
      const double a0 = a[row * kPad + ka], a1 = a[row * kPad + kb];
      const double b0 = ka > col ? a[col * kPad + ka] : (ka == col ? xd[ka] : 0.0);
      const double b1 = kb > col ? a[col * kPad + kb] : (kb == col ? xd[kb] : 0.0);
      dmma884(c0, c1, a0, b0);
      dmma884(d0, d1, a1, b1);
    }
    double* dst = T + (pr * H + 8 * ti + lr) * kTS + 8 * tj + 2 * lk;
    dst[0] = c0 + d0;
    dst[1] = c1 + d1;
  }
  __syncthreads();
  for (int t = warp; t < NT; t += 8) {       // X21_p = -X22_p T_p
    const int pr = t / (TT * TT), tt = t % (TT * TT), ti = tt / TT, tj = tt % TT;
    const int base = pr * 2 * H;
    const int row = base + H + 8 * ti + lr;  // A fragment: X22[row][k], stored at a[k][row] for row > k
    double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;
#pragma unroll 4
    for (int k4 = 0; k4 < 2 * (ti + 1); k4 += 2) {
      const int kl = 4 * k4 + lk;            // k inside the pair's second half
      const int ka = base + H + kl, kb = ka + 4;
      const double a0 = row > ka ? a[ka * kPad + row] : (row == ka ? xd[row] : 0.0);
      const double a1 = row > kb ? a[kb * kPad + row] : (row == kb ? xd[row] : 0.0);
      const double b0 = T[(pr * H + kl) * kTS + 8 * tj + lr];
      const double b1 = T[(pr * H + kl + 4) * kTS + 8 * tj + lr];
      dmma884(c0, c1, a0, b0);
      dmma884(d0, d1, a1, b1);
    }
    // X21[r][c] (r in the second half, c in the first) lives at a[c][r]
    const int cc = base + 8 * tj + 2 * lk;
    a[cc * kPad + row] = -(c0 + d0);
    a[(cc + 1) * kPad + row] = -(c1 + d1);
  }
  __syncthreads();
}

// warp 0: (optionally) subtract the rank-16 update of the previous panel from the 16x16 diagonal block at c0, factorise it
// in registers (lane = row; pivots and multipliers by shuffle; one rsqrt per pivot, no division) and leave L_D in the
// lower triangle of the block and 1 / diag in invd[c0 ..].  Returns 0 or the 1-based index of a non-positive pivot.
__device__ __forceinline__ int factor_diag16(double* __restrict__ a, double* __restrict__ invd, int c0, int lane,
                                             bool apply_update) {
  // Both half-warps hold row rl = lane & 15 of the block (the upper half is a working copy: it halves the update below and
  // is otherwise redundant).  Entries above the diagonal (k > rl) are carried as don't-care values — never selected out —
  // so that the 16-pivot chain is straight-line code: they are read by no lane (pivot j comes from lane j's r[j],
  // multiplier L[k][j] from lane k's r[j], k > j) and are not stored.
  const int rl = lane & (kSB - 1), half = lane >> 4;
  double r[kSB];
#pragma unroll
  for (int k = 0; k < kSB; ++k) r[k] = a[(c0 + rl) * kPad + c0 + k];
  if (apply_update) {      // a[c0+rl][c0+k] -= sum_t P[c0+rl][t] P[c0+k][t],  P = panel columns c0-16 .. c0-1
    constexpr int HT = kSB / 2;          // each half-warp sums 8 of the 16 terms; the halves are added in a fixed order
    double prow[HT];
#pragma unroll
    for (int t = 0; t < HT; ++t) prow[t] = a[(c0 + rl) * kPad + c0 - kSB + half * HT + t];
#pragma unroll
    for (int k = 0; k < kSB; ++k) {
      double s = 0.0;
#pragma unroll
      for (int t = 0; t < HT; ++t) s = fma(prow[t], a[(c0 + k) * kPad + c0 - kSB + half * HT + t], s);
      const double o = __shfl_xor_sync(0xffffffffu, s, 16);
      r[k] -= (half == 0) ? (s + o) : (o + s);      // low half + high half in both copies: identical bits
    }
  }
  int bad = 0;
#pragma unroll
  for (int j = 0; j < kSB; ++j) {
    double djj = __shfl_sync(0xffffffffu, r[j], j);
    if (!(djj > 0.0)) {
      if (bad == 0) bad = c0 + j + 1;
      djj = __longlong_as_double(0x7ff8000000000000LL);
    }
    const double inv = rsqrt(djj);
    if (lane == j) invd[c0 + j] = inv;
    r[j] *= inv;                         // lane j: djj / sqrt(djj) = L[j][j]; lanes > j: L[rl][j]
#pragma unroll
    for (int k = 0; k < kSB; ++k) {
      if (k > j) {
        const double lkj = __shfl_sync(0xffffffffu, r[j], k);
        r[k] = fma(-r[j], lkj, r[k]);
      }
    }
  }
  if (lane < kSB) {
#pragma unroll
    for (int k = 0; k < kSB; ++k)
      if (k <= lane) a[(c0 + lane) * kPad + c0 + k] = r[k];
  }
  return bad;
}

__global__ void __launch_bounds__(256, 1)
potf2_inv2_kernel(double* __restrict__ A_, int ld, int k0, const double* __restrict__ jitter_, int N, double jscale,
                  double* __restrict__ dinv_, int* __restrict__ info_, long long sA, long long sD, long long sJ) {
  extern __shared__ __align__(16) double sm[];
  double* a = sm;                    // [128][129]
  double* xd = sm + kBlk * kPad;     // [128] diagonal of the inverse
  double* invd = xd + kBlk;          // [128] 1 / diag(L)
  double* T = invd + kBlk;           // [64][65] scratch of phase B
  __shared__ int failed;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* A = A_ + (long long)blockIdx.x * sA;
  double* dinv = dinv_ + (long long)blockIdx.x * sD;
  const double* jitter = jitter_ ? jitter_ + (long long)blockIdx.x * sJ : nullptr;
  int* info = info_ + blockIdx.x;
  if (tid == 0) failed = 0;
#pragma unroll 8
  for (int it = 0; it < kBlk * kBlk / 2 / 256; ++it) {
    const int idx2 = it * 256 + tid, r = idx2 >> 6, c = (idx2 & 63) * 2;
    const double2 v = *reinterpret_cast<const double2*>(A + (long long)(k0 + r) * ld + (k0 + c));
    a[r * kPad + c] = (c <= r) ? v.x : 0.0;
    a[r * kPad + c + 1] = (c + 1 <= r) ? v.y : 0.0;
  }
  __syncthreads();
  if (tid < kBlk && jitter != nullptr && (k0 + tid) < N)
    a[tid * kPad + tid] = __fma_rn(jscale, jitter[k0 + tid], a[tid * kPad + tid]);
  __syncthreads();

  // ======================= Phase A ===============================================================================
  if (warp == 0) {
    const int bad = factor_diag16(a, invd, 0, lane, false);
    if (lane == 0 && bad != 0) { atomicCAS(info, 0, k0 + bad); failed = 1; }
  }
  __syncthreads();
  for (int pnl = 0; pnl < kBlk / kSB; ++pnl) {
    if (failed) break;
    const int c0 = pnl * kSB, r0 = c0 + kSB, Tn = kBlk - r0;
    if (Tn == 0) break;
    // A2: forward substitution of every row below against L_D (thread = row; L_D and 1 / diag are broadcast reads)
    if (tid < Tn) {
      double* row = a + (r0 + tid) * kPad + c0;
      double v[kSB];
#pragma unroll
      for (int j = 0; j < kSB; ++j) v[j] = row[j];
#pragma unroll
      for (int j = 0; j < kSB; ++j) {
        const double pj = v[j] * invd[c0 + j];
        v[j] = pj;
#pragma unroll
        for (int jj = j + 1; jj < kSB; ++jj) v[jj] = fma(-pj, a[(c0 + jj) * kPad + c0 + j], v[jj]);
      }
#pragma unroll
      for (int j = 0; j < kSB; ++j) row[j] = v[j];
    }
    __syncthreads();
    if (warp == 0) {
      // next diagonal block: its share of the trailing update, then its factorisation — concurrently with A3 below
      const int bad = factor_diag16(a, invd, r0, lane, true);
      if (lane == 0 && bad != 0) { atomicCAS(info, 0, k0 + bad); failed = 1; }
    } else {
      // A3: trailing (lower) -= P P^T with 4x4 register tiles, minus the 16x16 diagonal block warp 0 owns (tiles t < 10)
      const int nt = Tn / 4, ntiles = nt * (nt + 1) / 2;
      for (int t = 10 + (tid - 32); t < ntiles; t += 224) {
        int ti = (int)((sqrtf(8.0f * t + 1.0f) - 1.0f) * 0.5f);
        while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
        while (ti * (ti + 1) / 2 > t) --ti;
        const int tj = t - ti * (ti + 1) / 2;
        const double* pa = a + (r0 + 4 * ti) * kPad + c0;
        const double* pb = a + (r0 + 4 * tj) * kPad + c0;
        double acc[4][4];
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
          for (int f = 0; f < 4; ++f) acc[e][f] = 0.0;
#pragma unroll 8
        for (int k = 0; k < kSB; ++k) {
          double va[4], vb[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            va[e] = pa[e * kPad + k];
            vb[e] = pb[e * kPad + k];
          }
#pragma unroll
          for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int f = 0; f < 4; ++f) acc[e][f] = fma(va[e], vb[f], acc[e][f]);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
          for (int f = 0; f < 4; ++f) {
            const int i = r0 + 4 * ti + e, j = r0 + 4 * tj + f;
            if (j <= i) a[i * kPad + j] -= acc[e][f];
          }
      }
    }
    __syncthreads();
  }
  __syncthreads();
  if (failed) {
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    for (int idx = tid; idx < kBlk * kBlk; idx += 256) {
      int r = idx >> 7, c = idx & 127;
      A[(long long)(k0 + r) * ld + (k0 + c)] = nan;
      dinv[idx] = nan;
    }
    return;
  }

  // ======================= Phase A': the eight 16x16 diagonal inverses, one warp each (lane = column of X) =============
  {
    const int c0 = warp * kSB, rl = lane & (kSB - 1);
    double r[kSB], x[kSB];
#pragma unroll
    for (int k = 0; k < kSB; ++k) r[k] = (k <= rl) ? a[(c0 + rl) * kPad + c0 + k] : 0.0;
    const double myinv = invd[c0 + rl];
#pragma unroll
    for (int i = 0; i < kSB; ++i) {
      double sacc = 0.0;
#pragma unroll
      for (int k = 0; k < kSB; ++k) {
        if (k < i) {
          const double lik = __shfl_sync(0xffffffffu, r[k], i);   // L[i][k]
          if (k >= rl) sacc = fma(lik, x[k], sacc);
        }
      }
      const double dii = __shfl_sync(0xffffffffu, myinv, i);
      x[i] = (i == rl) ? dii : (i > rl ? -sacc * dii : 0.0);
    }
    __syncwarp();
    if (lane < kSB) {
#pragma unroll
      for (int i = 0; i < kSB; ++i) {
        if (i == lane) xd[c0 + lane] = x[i];
        else if (i > lane) a[(c0 + lane) * kPad + c0 + i] = x[i];   // X[i][j] (i > j) kept at a[j][i]
      }
    }
  }
  __syncthreads();

  // ======================= Phase B: recursive doubling ============================================================
  inv_level<16>(a, xd, T, tid);
  inv_level<32>(a, xd, T, tid);
  inv_level<64>(a, xd, T, tid);

#pragma unroll 4
  for (int idx2 = tid; idx2 < kBlk * kBlk / 2; idx2 += 256) {      // two columns per thread: 16-byte stores
    const int r = idx2 >> 6, c = (idx2 & 63) * 2;
    double2 l, x;
    l.x = (c <= r) ? a[r * kPad + c] : 0.0;
    l.y = (c + 1 <= r) ? a[r * kPad + c + 1] : 0.0;
    x.x = (c < r) ? a[c * kPad + r] : (c == r ? xd[r] : 0.0);
    x.y = (c + 1 < r) ? a[(c + 1) * kPad + r] : (c + 1 == r ? xd[r] : 0.0);
    *reinterpret_cast<double2*>(A + (long long)(k0 + r) * ld + (k0 + c)) = l;
    *reinterpret_cast<double2*>(dinv + r * kBlk + c) = x;
  }
}

__global__ void zero_upper_blocks_kernel(double* __restrict__ A, int ld, int nb, long long sA) {
  // one CTA per strictly-upper 128x128 block (bi < bj); blockIdx.y = batch
  A += (long long)blockIdx.y * sA;
  int t = blockIdx.x;
  int bj = (int)((sqrt(8.0 * t + 1.0) + 1.0) * 0.5);
  while ((long long)bj * (bj - 1) / 2 > t) --bj;
  while ((long long)(bj + 1) * bj / 2 <= t) ++bj;
  int bi = t - bj * (bj - 1) / 2;
  double2 z = make_double2(0.0, 0.0);
  for (int idx = threadIdx.x; idx < kBlk * kBlk / 2; idx += blockDim.x) {
    int r = idx >> 6, c = (idx & 63) * 2;
    *reinterpret_cast<double2*>(A + (long long)(bi * kBlk + r) * ld + bj * kBlk + c) = z;
  }
}

__global__ void scatter_diag_blocks_kernel(const double* __restrict__ dinv, double* __restrict__ Linv, int ld,
                                           long long sD, long long sL) {
  const double* src = dinv + (long long)blockIdx.y * sD + (long long)blockIdx.x * kBlk * kBlk;
  double* dst = Linv + (long long)blockIdx.y * sL + (long long)blockIdx.x * kBlk * (ld + 1);
  for (int idx = threadIdx.x; idx < kBlk * kBlk / 2; idx += blockDim.x) {
    int r = idx >> 6, c = (idx & 63) * 2;
    *reinterpret_cast<double2*>(dst + (long long)r * ld + c) = *reinterpret_cast<const double2*>(src + r * kBlk + c);
  }
}

}  // namespace bcbf

using namespace bcbf;

// 0: by problem size (default), 1: always 128 x 128 tiles, 2: always 32 x 128 tiles (tests and A/B timing)
extern "C" int bcbf_set_gemm_tile_policy(int policy) {
  BCBF_REQUIRE(policy >= 0 && policy <= 2, "bcbf_set_gemm_tile_policy: %d not in 0..2", policy);
  gemm_tile_policy() = policy;
  return BCBF_OK;
}

extern "C" long long bcbf_dinv_elems(int Npad) { return (long long)(Npad / kBlk) * kBlk * kBlk; }

// Look-ahead of the single-matrix factorisation: the serial sweep (single-CTA diagonal factorisations, panel solves) and
// the update of the next block column run on an internal HIGH-priority stream; the bulk trailing update stays on the
// caller's stream underneath it.  Stream priorities matter: both kernels want a whole SM's shared memory, and only a
// higher-priority stream gets the SM that a retiring GEMM CTA frees.  One set per device, created on demand.
struct LookAhead {
  cudaStream_t side = nullptr;   // high priority: critical path
  cudaStream_t side2 = nullptr;  // high priority: panel solves / in-block updates below the next diagonal block
  cudaEvent_t panels_done = nullptr, t2_done = nullptr, start = nullptr, mini_done = nullptr, rest_done = nullptr;
};
static LookAhead g_look[64];

static LookAhead* get_lookahead() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  LookAhead& l = g_look[dev & 63];
  if (l.side == nullptr) {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);   // hi = numerically least = greatest priority
    if (cudaStreamCreateWithPriority(&l.side, cudaStreamNonBlocking, hi) != cudaSuccess) { l.side = nullptr; return nullptr; }
    if (cudaStreamCreateWithPriority(&l.side2, cudaStreamNonBlocking, hi) != cudaSuccess ||
        cudaEventCreateWithFlags(&l.mini_done, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&l.rest_done, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&l.panels_done, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&l.start, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&l.t2_done, cudaEventDisableTiming) != cudaSuccess) {
      cudaStreamDestroy(l.side);
      l.side = nullptr;
      return nullptr;
    }
  }
  return &l;
}

// Trailing updates of bcbf_potrf with at least kPotrfI8MinRows rows run on the int8 tensor cores (csrc/ozaki.cu);
// bcbf_set_potrf_i8(0) keeps everything on the FP64 pipe.
static int g_potf2_variant = 1;   // 1: potf2_inv2_kernel (round 2), 0: potf2_inv_kernel (round 1); bcbf_set_potf2_variant
static int g_potrf_i8 = 1;
constexpr int kPotrfI8MinRows = 1024;
extern "C" int bcbf_set_potf2_variant(int v) {
  g_potf2_variant = v ? 1 : 0;
  return BCBF_OK;
}
extern "C" int bcbf_set_potrf_i8(int on) {
  g_potrf_i8 = on ? 1 : 0;
  return BCBF_OK;
}

// R independent factorisations of equal size (R = 1: the single-matrix entry point).  Strides in elements:
// sA between matrices, sD between dinv blocks sets, sJ between jitter vectors; info is int[R].
static int potrf_impl(double* A, int ld, int Npad, int N, const double* jitter, double jitter_scale, double* dinv,
                      int* info, int R, long long sA, long long sD, long long sJ, cudaStream_t stream) {
  BCBF_REQUIRE(A && dinv && info, "bcbf_potrf: null pointer");
  BCBF_REQUIRE(Npad > 0 && Npad % kBlk == 0 && ld >= Npad && ld % 2 == 0 && N <= Npad && N >= 0,
               "bcbf_potrf: Npad=%d must be a positive multiple of %d, ld=%d >= Npad and even, N=%d <= Npad", Npad,
               kBlk, ld, N);
  const int nb = Npad / kBlk;
  const int smem = (kBlk * kPad + kBlk + kTElems) * (int)sizeof(double);
  const int smem2 = (kBlk * kPad + 2 * kBlk + kT2Elems) * (int)sizeof(double);
  BCBF_CUDA(cudaFuncSetAttribute(potf2_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  BCBF_CUDA(cudaFuncSetAttribute(potf2_inv2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
  BCBF_CUDA(cudaMemsetAsync(info, 0, sizeof(int) * (size_t)R, stream));  // before the look-ahead fork below
  // Two-level blocking: outer block columns of kOuter = 512; inside one, a right-looking sweep over 128-wide panels
  // whose updates stay within the block column (K = 128, small); the bulk of the N^3/3 flops runs in ONE trailing
  // SYRK per outer block with K = 512, where the DMMA GEMM's pipeline fill is amortised over 32 k-steps.
  constexpr int kOuter = 4;  // in units of 128-blocks
  // look-ahead only pays (and is only wired) for one large matrix; ensembles are already parallel over R
  LookAhead* look = (R == 1 && nb > 2 * kOuter) ? get_lookahead() : nullptr;
  cudaStream_t cs = stream;   // critical-path stream: sweep + next-block-column update
  if (look) {
    // order the critical stream after everything already queued on the caller's stream, and reset t2_done
    BCBF_CUDA(cudaEventRecord(look->start, stream));
    BCBF_CUDA(cudaStreamWaitEvent(look->side, look->start, 0));
    BCBF_CUDA(cudaEventRecord(look->t2_done, stream));
    cs = look->side;
  }
  for (int J = 0; J < nb; J += kOuter) {
    const int jend = (J + kOuter < nb) ? J + kOuter : nb;  // exclusive, in blocks
    bool rest_pending = false;  // a "rest" step of this outer block is in flight on the second critical stream
    for (int k = J; k < jend; ++k) {
      const int k0 = k * kBlk;
      double* dk = dinv + (long long)k * kBlk * kBlk;
      if (g_potf2_variant == 1)
        potf2_inv2_kernel<<<R, 256, smem2, cs>>>(A, ld, k0, jitter, N, jitter_scale, dk, info, sA, sD, sJ);
      else
        potf2_inv_kernel<<<R, 256, smem, cs>>>(A, ld, k0, jitter, N, jitter_scale, dk, info, sA, sD, sJ);
      BCBF_LAUNCH_CHECK();
      const int rows = Npad - (k0 + kBlk);
      if (rows <= 0) break;
      double* panel = A + (long long)(k0 + kBlk) * ld + k0;
      const int w = (jend - (k + 1)) * kBlk;
      if (look && w > 0) {
        // Inner look-ahead.  The next diagonal factorisation (one CTA, ~170 us, the serial chain of the whole
        // algorithm) needs only row block k+1: its panel solve and the update of its diagonal block ("mini", here);
        // the solve and update of all rows below ("rest") run on a second high-priority stream next to it.
        if (rest_pending) BCBF_CUDA(cudaStreamWaitEvent(cs, look->rest_done, 0));   // rest(k-1) wrote row block k+1
        GemmArgs g{};
        g.A = panel; g.lda = ld; g.B = dk; g.ldb = kBlk; g.C = panel; g.ldc = ld;
        g.M = kBlk; g.N = kBlk; g.K = kBlk; g.alpha = 1.0; g.beta = 0.0; g.tri = kTriNone;
        BCBF_CUDA((launch_gemm<true, true>(g, 1, cs)));
        GemmArgs u{};
        u.A = panel; u.lda = ld; u.B = panel; u.ldb = ld;
        u.C = A + (long long)(k0 + kBlk) * (ld + 1); u.ldc = ld;
        u.M = kBlk; u.N = kBlk; u.K = kBlk; u.alpha = -1.0; u.beta = 1.0; u.tri = kTriNone;
        BCBF_CUDA((launch_gemm<true, true>(u, 1, cs)));
        BCBF_CUDA(cudaEventRecord(look->mini_done, cs));
        if (rows > kBlk) {
          BCBF_CUDA(cudaStreamWaitEvent(look->side2, look->mini_done, 0));
          double* below = panel + (long long)kBlk * ld;
          GemmArgs g2 = g;
          g2.A = below; g2.C = below; g2.M = rows - kBlk;
          BCBF_CUDA((launch_gemm<true, true>(g2, 1, look->side2)));
          GemmArgs u2 = u;
          u2.A = below; u2.B = panel;
          u2.C = A + (long long)(k0 + 2 * kBlk) * ld + (k0 + kBlk);
          u2.M = rows - kBlk; u2.N = w;
          BCBF_CUDA((launch_gemm<true, true>(u2, 1, look->side2)));
          BCBF_CUDA(cudaEventRecord(look->rest_done, look->side2));
          rest_pending = true;
        }
        continue;
      }
      if (rest_pending) {
        BCBF_CUDA(cudaStreamWaitEvent(cs, look->rest_done, 0));
        rest_pending = false;
      }
      GemmArgs g{};
      // panel <- panel * Dinv_k^T       (L_ik = A_ik L_kk^{-T})
      g.A = panel; g.lda = ld; g.B = dk; g.ldb = kBlk; g.C = panel; g.ldc = ld;
      g.M = rows; g.N = kBlk; g.K = kBlk; g.alpha = 1.0; g.beta = 0.0; g.tri = kTriNone;
      g.sA = sA; g.sB = sD; g.sC = sA;
      BCBF_CUDA((launch_gemm<true, true>(g, R, cs)));
      // inside the outer block column: columns (k+1)*128 .. jend*128, all rows below  -= panel panel^T
      if (w > 0) {
        GemmArgs u{};
        u.A = panel; u.lda = ld; u.B = panel; u.ldb = ld;
        u.C = A + (long long)(k0 + kBlk) * (ld + 1); u.ldc = ld;
        u.M = rows; u.N = w; u.K = kBlk; u.alpha = -1.0; u.beta = 1.0; u.tri = kTriNone;
        u.sA = u.sB = u.sC = sA;
        BCBF_CUDA((launch_gemm<true, true>(u, R, cs)));
      }
    }
    if (rest_pending) BCBF_CUDA(cudaStreamWaitEvent(cs, look->rest_done, 0));
    const int c1 = jend * kBlk, rows = Npad - c1;
    if (rows > 0) {
      // trailing (lower tiles) -= P P^T,  P = A[c1:, J*128 : c1]   (K up to 512)
      const double* P = A + (long long)c1 * ld + (long long)J * kBlk;
      const int Kp = c1 - J * kBlk;
      const int strip = rows < kOuter * kBlk ? rows : kOuter * kBlk;   // columns the NEXT outer block will factor
      if (look && rows > strip) {
        // look-ahead: (t1) update only the next block column on the caller's stream, so that its serial sweep
        // (single-CTA diagonal factorisations) can start at once, and (t2) update the rest on a low-priority side
        // stream underneath it.  t1 of the next outer block waits for t2 (same region).
        BCBF_CUDA(cudaStreamWaitEvent(cs, look->t2_done, 0));   // previous t2 (no-op the first time)
        GemmArgs t1{};
        t1.A = P; t1.lda = ld; t1.B = P; t1.ldb = ld;
        t1.C = A + (long long)c1 * (ld + 1); t1.ldc = ld;
        t1.M = rows; t1.N = strip; t1.K = Kp; t1.alpha = -1.0; t1.beta = 1.0; t1.tri = kTriNone;
        BCBF_CUDA((launch_gemm<true, true>(t1, 1, cs)));
        BCBF_CUDA(cudaEventRecord(look->panels_done, cs));
        BCBF_CUDA(cudaStreamWaitEvent(stream, look->panels_done, 0));
        const double* P2 = P + (long long)strip * ld;
        if (g_potrf_i8 && rows - strip >= kPotrfI8MinRows && Kp % 32 == 0) {
          // the bulk of the N^3/3 flops: lower tiles -= P2 P2^T on the int8 tensor cores (bcbf_oz_update, FP64-accurate,
          // one short-lived CTA per tile so that the high-priority sweep keeps finding free SMs)
          int rc = bcbf_oz_update(rows - strip, rows - strip, Kp, -1.0, P2, ld, P2, ld,
                                  A + (long long)(c1 + strip) * (ld + 1), ld, 1, stream);
          if (rc) return rc;
        } else {
          GemmArgs t2{};
          t2.A = P2; t2.lda = ld; t2.B = P2; t2.ldb = ld;
          t2.C = A + (long long)(c1 + strip) * (ld + 1); t2.ldc = ld;
          t2.M = rows - strip; t2.N = rows - strip; t2.K = Kp; t2.alpha = -1.0; t2.beta = 1.0; t2.tri = kTriLowerOut;
          BCBF_CUDA((launch_gemm<true, true>(t2, 1, stream)));
        }
        BCBF_CUDA(cudaEventRecord(look->t2_done, stream));
      } else {
        if (look) BCBF_CUDA(cudaStreamWaitEvent(cs, look->t2_done, 0));
        GemmArgs t{};
        t.A = P; t.lda = ld; t.B = P; t.ldb = ld;
        t.C = A + (long long)c1 * (ld + 1); t.ldc = ld;
        t.M = rows; t.N = rows; t.K = Kp; t.alpha = -1.0; t.beta = 1.0; t.tri = kTriLowerOut;
        t.sA = t.sB = t.sC = sA;
        BCBF_CUDA((launch_gemm<true, true>(t, R, cs)));
      }
    }
  }
  if (look) {  // rejoin: everything the critical stream did becomes visible to the caller's stream
    BCBF_CUDA(cudaEventRecord(look->panels_done, cs));
    BCBF_CUDA(cudaStreamWaitEvent(stream, look->panels_done, 0));
  }
  if (nb > 1) {
    zero_upper_blocks_kernel<<<dim3(nb * (nb - 1) / 2, R), 256, 0, stream>>>(A, ld, nb, sA);
    BCBF_LAUNCH_CHECK();
  }
  return BCBF_OK;
}

extern "C" int bcbf_potrf(double* A, int ld, int Npad, int N, const double* jitter, double jitter_scale,
                          double* dinv, int* info, void* stream_) {
  ::bcbf::ScratchScope scratch_scope(static_cast<cudaStream_t>(stream_));
  return potrf_impl(A, ld, Npad, N, jitter, jitter_scale, dinv, info, 1, 0, 0, 0, static_cast<cudaStream_t>(stream_));
}

extern "C" int bcbf_potrf_batched(double* A, int ld, int Npad, int N, const double* jitter, double jitter_scale,
                                  double* dinv, int* info, int R, void* stream_) {
  ::bcbf::ScratchScope scratch_scope(static_cast<cudaStream_t>(stream_));
  BCBF_REQUIRE(R >= 1, "bcbf_potrf_batched: R=%d", R);
  return potrf_impl(A, ld, Npad, N, jitter, jitter_scale, dinv, info, R, (long long)ld * Npad,
                    bcbf_dinv_elems(Npad), N, static_cast<cudaStream_t>(stream_));
}

extern "C" int bcbf_check_info(const int* info, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int h = 0;
  BCBF_CUDA(cudaMemcpyAsync(&h, info, sizeof(int), cudaMemcpyDeviceToHost, stream));
  BCBF_CUDA(cudaStreamSynchronize(stream));
  if (h != 0) {
    set_last_error("cholesky: the leading minor of order %d is not positive-definite", h);
    return BCBF_ERR_NOT_PD;
  }
  return BCBF_OK;
}

// Levels of the divide-and-conquer triangular inverse with half-size h >= kTrtriI8MinH run their two products per pair on
// the int8 tensor cores (csrc/ozaki.cu); smaller levels and ensembles stay on the DMMA GEMM.  bcbf_set_trtri_i8(0) turns
// it off (all levels on the FP64 pipe).
static int g_trtri_i8 = 1;
constexpr int kTrtriI8MinH = 2048;
extern "C" int bcbf_set_trtri_i8(int on) {
  g_trtri_i8 = on ? 1 : 0;
  return BCBF_OK;
}

// R independent triangular inverses (stride sL between factor-sized matrices, sD between dinv block sets).
static int trtri_impl(const double* L, const double* dinv, double* Linv, double* scratch, int ld, int Npad, int R,
                      long long sL, long long sD, cudaStream_t stream) {
  BCBF_REQUIRE(L && dinv && Linv && scratch, "bcbf_trtri: null pointer");
  BCBF_REQUIRE(Npad > 0 && Npad % kBlk == 0 && ld >= Npad && ld % 2 == 0, "bcbf_trtri: bad Npad=%d / ld=%d", Npad, ld);
  const int nb = Npad / kBlk;
  if (R == 1) BCBF_CUDA(cudaMemsetAsync(Linv, 0, sizeof(double) * (size_t)ld * Npad, stream));
  else BCBF_CUDA(cudaMemsetAsync(Linv, 0, sizeof(double) * (size_t)sL * R, stream));
  scatter_diag_blocks_kernel<<<dim3(nb, R), 256, 0, stream>>>(dinv, Linv, ld, sD, sL);
  BCBF_LAUNCH_CHECK();
  for (int hb = 1; hb < nb; hb *= 2) {
    const int h = hb * kBlk;
    // pairs g: rows [r0, r0+h) (X11) and [r0+h, r0+2h) clipped to Npad (X22), r0 = g*2h
    const int npairs_full = Npad / (2 * h);                          // both halves complete
    const int rem = Npad - npairs_full * 2 * h;                      // leftover rows after the full pairs
    const int partial_rows = rem > h ? rem - h : 0;                  // clipped second half of the last pair
    // a single matrix batches over the pairs of a level; an ensemble batches over the matrices and loops the pairs
    const int npass = (R == 1) ? 2 : npairs_full + (partial_rows > 0 ? 1 : 0);
    for (int pass = 0; pass < npass; ++pass) {
      int batch, M2;
      long long r0, stride;
      if (R == 1) {
        batch = pass == 0 ? npairs_full : (partial_rows > 0 ? 1 : 0);
        r0 = pass == 0 ? 0 : (long long)npairs_full * 2 * h;
        M2 = pass == 0 ? h : partial_rows;
        stride = (long long)2 * h * (ld + 1);
      } else {
        batch = R;
        r0 = (long long)pass * 2 * h;
        M2 = pass < npairs_full ? h : partial_rows;
        stride = sL;
      }
      if (batch == 0) continue;
      if (R == 1 && g_trtri_i8 && h >= kTrtriI8MinH && h <= bcbf_oz_max_npad() && M2 % kBlk == 0) {
        // large levels (94 % of the flops): the two products on the int8 tensor cores (bcbf_oz_gemm, FP64-accurate)
        for (int b = 0; b < batch; ++b) {
          const long long o = r0 + (long long)b * 2 * h;
          int rc = bcbf_oz_gemm(M2, h, h, 1.0, L + (o + h) * ld + o, ld, Linv + o * (ld + 1), ld,
                                scratch + (o + h) * ld + o, ld, /*B lower triangular*/ 2, stream);
          if (rc) return rc;
          rc = bcbf_oz_gemm(M2, h, M2, -1.0, Linv + (o + h) * (ld + 1), ld, scratch + (o + h) * ld + o, ld,
                            Linv + (o + h) * ld + o, ld, /*A lower triangular*/ 1, stream);
          if (rc) return rc;
        }
        continue;
      }
      GemmArgs t{};  // T = L21 * X11
      t.A = L + (r0 + h) * ld + r0; t.lda = ld;
      t.B = Linv + r0 * (ld + 1); t.ldb = ld;
      t.C = scratch + (r0 + h) * ld + r0; t.ldc = ld;
      t.M = M2; t.N = h; t.K = h; t.alpha = 1.0; t.beta = 0.0; t.tri = kTriBLower;
      t.sA = t.sB = t.sC = stride;
      BCBF_CUDA((launch_gemm<true, false>(t, batch, stream)));
      GemmArgs x{};  // X21 = -X22 * T
      x.A = Linv + (r0 + h) * (ld + 1); x.lda = ld;
      x.B = scratch + (r0 + h) * ld + r0; x.ldb = ld;
      x.C = Linv + (r0 + h) * ld + r0; x.ldc = ld;
      x.M = M2; x.N = h; x.K = M2; x.alpha = -1.0; x.beta = 0.0; x.tri = kTriALower;
      x.sA = x.sB = x.sC = stride;
      BCBF_CUDA((launch_gemm<true, false>(x, batch, stream)));
    }
  }
  return BCBF_OK;
}

extern "C" int bcbf_trtri(const double* L, const double* dinv, double* Linv, double* scratch, int ld, int Npad,
                          void* stream_) {
  ::bcbf::ScratchScope scratch_scope(static_cast<cudaStream_t>(stream_));
  return trtri_impl(L, dinv, Linv, scratch, ld, Npad, 1, 0, 0, static_cast<cudaStream_t>(stream_));
}

extern "C" int bcbf_trtri_batched(const double* L, const double* dinv, double* Linv, double* scratch, int ld, int Npad,
                                  int R, void* stream_) {
  ::bcbf::ScratchScope scratch_scope(static_cast<cudaStream_t>(stream_));
  BCBF_REQUIRE(R >= 1, "bcbf_trtri_batched: R=%d", R);
  return trtri_impl(L, dinv, Linv, scratch, ld, Npad, R, (long long)ld * Npad, bcbf_dinv_elems(Npad),
                    static_cast<cudaStream_t>(stream_));
}

static int trmm_impl(const double* A, int lda, int Npad, int trans, const double* B, int ldb, int ncols, double alpha,
                     double beta, double* C, int ldc, int R, long long sA, long long sB, long long sC,
                     cudaStream_t stream);

extern "C" int bcbf_trmm_lower(const double* A, int lda, int Npad, int trans, const double* B, int ldb, int ncols,
                               double alpha, double beta, double* C, int ldc, void* stream_) {
  return trmm_impl(A, lda, Npad, trans, B, ldb, ncols, alpha, beta, C, ldc, 1, 0, 0, 0,
                   static_cast<cudaStream_t>(stream_));
}

extern "C" int bcbf_trmm_lower_batched(const double* A, int lda, int Npad, int trans, const double* B, int ldb,
                                       int ncols, double alpha, double beta, double* C, int ldc, int R,
                                       void* stream_) {
  BCBF_REQUIRE(R >= 1, "bcbf_trmm_lower_batched: R=%d", R);
  return trmm_impl(A, lda, Npad, trans, B, ldb, ncols, alpha, beta, C, ldc, R, (long long)lda * Npad,
                   (long long)ldb * Npad, (long long)ldc * Npad, static_cast<cudaStream_t>(stream_));
}

static int trmm_impl(const double* A, int lda, int Npad, int trans, const double* B, int ldb, int ncols, double alpha,
                     double beta, double* C, int ldc, int R, long long sA, long long sB, long long sC,
                     cudaStream_t stream) {
  BCBF_REQUIRE(A && B && C, "bcbf_trmm_lower: null pointer");
  BCBF_REQUIRE(Npad > 0 && Npad % kBlk == 0 && lda >= Npad && lda % 2 == 0 && ldb % 2 == 0 && ldc % 2 == 0 &&
                   ncols > 0 && ncols % 2 == 0 && ldb >= ncols && ldc >= ncols,
               "bcbf_trmm_lower: Npad=%d lda=%d ldb=%d ldc=%d ncols=%d (leading dims and ncols must be even)", Npad,
               lda, ldb, ldc, ncols);
  GemmArgs g{};
  g.A = A; g.lda = lda; g.B = B; g.ldb = ldb; g.C = C; g.ldc = ldc;
  g.M = Npad; g.N = ncols; g.K = Npad; g.alpha = alpha; g.beta = beta;
  g.sA = sA; g.sB = sB; g.sC = sC;
  if (!trans) {
    g.tri = kTriALower;
    BCBF_CUDA((launch_gemm<true, false>(g, R, stream)));
  } else {
    g.tri = kTriAUpper;
    BCBF_CUDA((launch_gemm<false, false>(g, R, stream)));
  }
  return BCBF_OK;
}

// General row-major C(M,N) = alpha * op(A) op(B) + beta * C on the DMMA GEMM (host glue for the small-batch API
// paths: v^T v', kb*^T alpha, Linv^T Linv).  op(A) is M x K: transa = 0 -> A stored (M,K); 1 -> stored (K,M).
// op(B) is K x N: transb = 0 -> B stored (K,N); 1 -> stored (N,K).
static int gemm_impl(int transa, int transb, int M, int N, int K, double alpha, const double* A, int lda, long long sA,
                     const double* B, int ldb, long long sB, double beta, double* C, int ldc, long long sC, int R,
                     cudaStream_t stream);

extern "C" int bcbf_gemm(int transa, int transb, int M, int N, int K, double alpha, const double* A, int lda,
                         const double* B, int ldb, double beta, double* C, int ldc, void* stream_) {
  return gemm_impl(transa, transb, M, N, K, alpha, A, lda, 0, B, ldb, 0, beta, C, ldc, 0, 1,
                   static_cast<cudaStream_t>(stream_));
}

// R independent products with element strides sA / sB / sC between consecutive operands (ensembles).
extern "C" int bcbf_gemm_batched(int transa, int transb, int M, int N, int K, double alpha, const double* A, int lda,
                                 long long sA, const double* B, int ldb, long long sB, double beta, double* C, int ldc,
                                 long long sC, int R, void* stream_) {
  BCBF_REQUIRE(R >= 1 && sA % 2 == 0 && sB % 2 == 0 && sC % 2 == 0, "bcbf_gemm_batched: R=%d / odd batch stride", R);
  return gemm_impl(transa, transb, M, N, K, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, R,
                   static_cast<cudaStream_t>(stream_));
}

static int gemm_impl(int transa, int transb, int M, int N, int K, double alpha, const double* A, int lda, long long sA,
                     const double* B, int ldb, long long sB, double beta, double* C, int ldc, long long sC, int R,
                     cudaStream_t stream) {
  BCBF_REQUIRE(A && B && C, "bcbf_gemm: null pointer");
  BCBF_REQUIRE(M > 0 && N > 0 && K > 0 && M % 2 == 0 && N % 2 == 0 && K % 2 == 0,
               "bcbf_gemm: M=%d N=%d K=%d must be positive and even (pad with zeros)", M, N, K);
  BCBF_REQUIRE(lda % 2 == 0 && ldb % 2 == 0 && ldc % 2 == 0, "bcbf_gemm: leading dimensions must be even");
  BCBF_REQUIRE(((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B) | reinterpret_cast<uintptr_t>(C)) & 15) == 0,
               "bcbf_gemm: operands must be 16-byte aligned");
  GemmArgs g{};
  g.A = A; g.lda = lda; g.B = B; g.ldb = ldb; g.C = C; g.ldc = ldc;
  g.M = M; g.N = N; g.K = K; g.alpha = alpha; g.beta = beta; g.tri = kTriNone;
  g.sA = sA; g.sB = sB; g.sC = sC;
  if (!transa && !transb) BCBF_CUDA((launch_gemm<true, false>(g, R, stream)));
  else if (!transa && transb) BCBF_CUDA((launch_gemm<true, true>(g, R, stream)));
  else if (transa && !transb) BCBF_CUDA((launch_gemm<false, false>(g, R, stream)));
  else BCBF_CUDA((launch_gemm<false, true>(g, R, stream)));
  return BCBF_OK;
}
