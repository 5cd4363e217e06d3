// Ensembles of small independent MVGPs (one per rollout; BASELINE configs[4]: 4096 rollouts, N <= 200 training points
// each, one query state per control step).  Every rollout has its own training set, hyper-parameters and factor.
//
//   ens_gram_kernel      Kb_r = k_r(X_r, X_r) o (UH_r B_r UH_r^T), padded to Npad with the identity    (batched Gram)
//   ens_prep_kernel      G_r = UH_r B_r,  Y_r = Xdot_r - UH_r C_r
//   ens_w_kernel         W_r = alpha_r (.) G_r
//   ens_posterior_kernel one CTA per rollout: k*(x_r), frakB = k* G, M_k = C^T + k*^T W,  V = L^-1 frakB streamed from HBM
//                        (transposed storage, lower triangle only, coalesced), B_k = s B - V^T V.
// The per-step posterior is HBM-bound: 4 N^2 bytes of L^-1 per rollout (its own factor, no reuse between rollouts)
// against N^2 p flops — 160 KB vs 0.12 MFLOP at N = 200.  Factorisation / inverse / alpha reuse the batched DMMA
// kernels of factor.cu.  Replaces, per rollout and control step, the b = 1 custom_predict calls of
// ControllerCLFBayesian.control (unicycle_move_to_pose.py:880-920 -> control_affine_model.py:931-961, 983-1096).
#include "../../include/bcbf.h"
#include "common.cuh"

namespace bcbf {

constexpr int kEN = BCBF_MAX_N_DIM, kEP = BCBF_MAX_P_DIM;

__global__ void __launch_bounds__(256)
ens_gram_kernel(const double* __restrict__ X, const double* __restrict__ UH, const double* __restrict__ ls,
                const double* __restrict__ scale, const double* __restrict__ Bm, int N, int Npad, int n, int p,
                double* __restrict__ Kb) {
  constexpr int T = 64;
  __shared__ double xr[T][kEN + 1], xc[T][kEN + 1], gr[T][kEP], uc[T][kEP], il[kEN], Bs[kEP * kEP];
  const int r = blockIdx.z, tid = threadIdx.x;
  const int r0 = blockIdx.y * T, c0 = blockIdx.x * T;
  X += (long long)r * N * n;
  UH += (long long)r * N * p;
  Kb += (long long)r * Npad * Npad;
  if (tid < n) il[tid] = 1.0 / ls[(long long)r * n + tid];
  if (tid < p * p) Bs[tid] = Bm[(long long)r * p * p + tid];
  __syncthreads();
  for (int idx = tid; idx < T * n; idx += 256) {
    int i = idx / n, d = idx % n;
    xr[i][d] = (r0 + i < N) ? X[(long long)(r0 + i) * n + d] * il[d] : 0.0;
    xc[i][d] = (c0 + i < N) ? X[(long long)(c0 + i) * n + d] * il[d] : 0.0;
  }
  for (int idx = tid; idx < T * p; idx += 256) {
    int i = idx / p, q = idx % p;
    double g = 0.0;
    if (r0 + i < N)
      for (int t = 0; t < p; ++t) g += UH[(long long)(r0 + i) * p + t] * Bs[t * p + q];
    gr[i][q] = g;
    uc[i][q] = (c0 + i < N) ? UH[(long long)(c0 + i) * p + q] : 0.0;
  }
  __syncthreads();
  const double s = scale[r];
  const int ty = tid >> 4, tx = tid & 15;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int rl = ty * 4 + i, row = r0 + rl;
    if (row >= Npad) continue;
    double v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int cl = tx * 4 + j, col = c0 + cl;
      double out = 0.0;
      if (row < N && col < N) {
        double d2 = 0.0;
        for (int d = 0; d < n; ++d) {
          double df = xr[rl][d] - xc[cl][d];
          d2 = fma(df, df, d2);
        }
        double ub = 0.0;
        for (int q = 0; q < p; ++q) ub = fma(gr[rl][q], uc[cl][q], ub);
        out = s * exp(-0.5 * d2) * ub;
      } else if (row == col) {
        out = 1.0;
      }
      v[j] = out;
    }
    const int col = c0 + tx * 4;
    if (col + 3 < Npad) {
      double* dst = Kb + (long long)row * Npad + col;
      *reinterpret_cast<double2*>(dst) = make_double2(v[0], v[1]);
      *reinterpret_cast<double2*>(dst + 2) = make_double2(v[2], v[3]);
    }
  }
}

__global__ void ens_prep_kernel(const double* __restrict__ UH, const double* __restrict__ Xdot,
                                const double* __restrict__ Bm, const double* __restrict__ C, int R, int N, int Npad,
                                int n, int p, int ldy, double* __restrict__ G, double* __restrict__ Y) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)R * Npad) return;
  const int r = (int)(idx / Npad), i = (int)(idx % Npad);
  double uh[kEP];
  for (int j = 0; j < p; ++j) uh[j] = (i < N) ? UH[((long long)r * N + i) * p + j] : 0.0;
  for (int j = 0; j < p; ++j) {
    double g = 0.0;
    for (int t = 0; t < p; ++t) g = fma(uh[t], Bm[((long long)r * p + t) * p + j], g);
    G[idx * p + j] = g;
  }
  for (int c = 0; c < ldy; ++c) {
    double y = 0.0;
    if (i < N && c < n) {
      y = Xdot[((long long)r * N + i) * n + c];
      for (int t = 0; t < p; ++t) y -= uh[t] * C[((long long)r * p + t) * n + c];
    }
    Y[idx * ldy + c] = y;
  }
}

__global__ void ens_w_kernel(const double* __restrict__ alpha, int ldy, const double* __restrict__ G, long long rows,
                             int n, int p, double* __restrict__ W) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  for (int c = 0; c < n; ++c)
    for (int j = 0; j < p; ++j) W[(i * n + c) * p + j] = alpha[i * ldy + c] * G[i * p + j];
}

// One CTA (128 threads) per rollout; each thread owns TWO adjacent training rows.  The factor inverse is stored
// TRANSPOSED (LinvT[k][i] = L^-1[i][k]) so that at step k the threads of a warp read one contiguous 512-byte segment of
// row k with 16-byte loads: every byte of the lower triangle is fetched once, coalesced, 8 independent loads in flight
// per thread, up to 6 CTAs per SM.  Each thread keeps its own running V_i (2 x p values) — no per-row reduction; only
// the final p(p+1)/2 + n p sums are reduced across the CTA (fixed order: deterministic).
constexpr int kEnsThreads = 128;
__global__ void __launch_bounds__(kEnsThreads, 6)
ens_posterior_kernel(const double* __restrict__ LinvT, const double* __restrict__ X, const double* __restrict__ G,
                     const double* __restrict__ W, const double* __restrict__ ls, const double* __restrict__ scale,
                     const double* __restrict__ Bm, const double* __restrict__ C, const double* __restrict__ xq, int N,
                     int Npad, int n, int p, double* __restrict__ Mk, double* __restrict__ Bk) {
  extern __shared__ __align__(16) double sm[];
  double* fb = sm;                 // [Npad][kEP] frakB rows (k* G), padded to 4 columns
  constexpr int NW = kEnsThreads / 32;
  __shared__ double red[NW][kEN * kEP + kEP * (kEP + 1) / 2];
  __shared__ double xs[kEN], il[kEN];
  const int r = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int np = n * p;
  LinvT += (long long)r * Npad * Npad;
  X += (long long)r * N * n;
  G += (long long)r * Npad * p;
  W += (long long)r * Npad * np;
  if (tid < n) {
    il[tid] = 1.0 / ls[(long long)r * n + tid];
    xs[tid] = xq[(long long)r * n + tid];
  }
  for (int i = tid; i < NW * (kEN * kEP + kEP * (kEP + 1) / 2); i += kEnsThreads) (&red[0][0])[i] = 0.0;
  __syncthreads();
  const double s = scale[r];
  // ---- k*, frakB and the mean partial sums (rows strided over the CTA) --------------------------------------------
  for (int i0 = 0; i0 < Npad; i0 += kEnsThreads) {
    const int i = i0 + tid;
    double kv = 0.0;
    if (i < N) {
      double d2 = 0.0;
      for (int d = 0; d < n; ++d) {
        double df = (X[(long long)i * n + d] - xs[d]) * il[d];
        d2 = fma(df, df, d2);
      }
      kv = s * exp(-0.5 * d2);
    }
    if (i < Npad) {
#pragma unroll
      for (int q = 0; q < kEP; ++q) fb[i * kEP + q] = (q < p) ? kv * G[(long long)i * p + q] : 0.0;
    }
    for (int c0 = 0; c0 < np; c0 += 4) {  // mean: 4 columns of W at a time, warp tree, one smem slot per warp
      double m4[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        double t = (i < N && c0 + e < np) ? kv * W[(long long)i * np + c0 + e] : 0.0;
        m4[e] = warp_sum(t);
      }
      if (lane == 0) {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (c0 + e < np) red[warp][c0 + e] += m4[e];
      }
    }
  }
  __syncthreads();
  // ---- V_i = sum_{k <= i} L^-1[i][k] frakB[k],  S += V_i V_i^T ---------------------------------------------------
  for (int i0 = 0; i0 < N; i0 += 2 * kEnsThreads) {
    const int i = i0 + 2 * tid;                                   // rows i, i+1
    const bool live = i < N;                                      // (Npad is even and row N.. of L^-1 is the identity
                                                                  //  pad against zero frakB rows: harmless)
    const int kmax = min(N - 1, i0 + 2 * ((warp + 1) * 32) - 1);  // last row of this warp: uniform bound per warp
    double v0[kEP], v1[kEP];
#pragma unroll
    for (int q = 0; q < kEP; ++q) v0[q] = v1[q] = 0.0;
    const double* col = LinvT + i;
    int k = 0;
    for (; k + 8 <= kmax + 1; k += 8) {
      double2 l8[8];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        l8[u] = (live && k + u <= i + 1) ? __ldg(reinterpret_cast<const double2*>(col + (long long)(k + u) * Npad))
                                         : make_double2(0.0, 0.0);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const double4 f = *reinterpret_cast<const double4*>(fb + (k + u) * kEP);
        v0[0] = fma(l8[u].x, f.x, v0[0]); v0[1] = fma(l8[u].x, f.y, v0[1]);
        v0[2] = fma(l8[u].x, f.z, v0[2]); v0[3] = fma(l8[u].x, f.w, v0[3]);
        v1[0] = fma(l8[u].y, f.x, v1[0]); v1[1] = fma(l8[u].y, f.y, v1[1]);
        v1[2] = fma(l8[u].y, f.z, v1[2]); v1[3] = fma(l8[u].y, f.w, v1[3]);
      }
    }
    for (; k <= kmax; ++k) {
      const double2 l = (live && k <= i + 1) ? __ldg(reinterpret_cast<const double2*>(col + (long long)k * Npad))
                                              : make_double2(0.0, 0.0);
      const double4 f = *reinterpret_cast<const double4*>(fb + k * kEP);
      v0[0] = fma(l.x, f.x, v0[0]); v0[1] = fma(l.x, f.y, v0[1]);
      v0[2] = fma(l.x, f.z, v0[2]); v0[3] = fma(l.x, f.w, v0[3]);
      v1[0] = fma(l.y, f.x, v1[0]); v1[1] = fma(l.y, f.y, v1[1]);
      v1[2] = fma(l.y, f.z, v1[2]); v1[3] = fma(l.y, f.w, v1[3]);
    }
    int e = 0;
#pragma unroll
    for (int q = 0; q < kEP; ++q)
#pragma unroll
      for (int t = q; t < kEP; ++t) {
        // columns >= p of frakB are zero; rows >= N give V = 0.  Reduced per 256-row chunk so that the pair sums are
        // not live (20 registers) across the streaming loop.
        const double pr = warp_sum(fma(v0[q], v0[t], v1[q] * v1[t]));
        if (lane == 0) red[warp][kEN * kEP + e] += pr;
        ++e;
      }
  }
  __syncthreads();
  if (tid < np) {
    double t = 0.0;
    for (int w = 0; w < NW; ++w) t += red[w][tid];
    const int c = tid / p, j = tid % p;  // Mk[c][j] = C[j][c] + ...
    Mk[(long long)r * np + tid] = C[((long long)r * p + j) * n + c] + t;
  }
  if (tid < p * p) {
    const int a = tid / p, b = tid % p, q = a < b ? a : b, t2 = a < b ? b : a;
    int e = 0;  // index of the (q, t2) pair in the kEP-wide upper-triangular enumeration used above
    for (int qq = 0; qq < q; ++qq) e += kEP - qq;
    e += t2 - q;
    double t = 0.0;
    for (int w = 0; w < NW; ++w) t += red[w][kEN * kEP + e];
    Bk[(long long)r * p * p + tid] = s * Bm[(long long)r * p * p + tid] - t;
  }
}


// Batched twin of gram_backward_kernel (gram.cu): per rollout r, sum_ij Gbar_ij dKb_ij/dtheta with
// Gbar = 1/2 (alphaAi alpha^T - nout Kb^-1); per-tile partials, then a fixed-order sum per rollout (deterministic).
// Output layout per rollout: [d/d outputscale, d/d lengthscale (kEN), d/dB (kEP x kEP)].
constexpr int kEGrad = 1 + kEN + kEP * kEP;

__global__ void __launch_bounds__(256)
ens_gram_backward_kernel(const double* __restrict__ X, const double* __restrict__ UH, const double* __restrict__ ls,
                         const double* __restrict__ scale, const double* __restrict__ Bm,
                         const double* __restrict__ Pinv, const double* __restrict__ alphaAi,
                         const double* __restrict__ alpha, int N, int Npad, int n, int p, int nd,
                         double* __restrict__ partial) {
  constexpr int T = 64;
  __shared__ double xr[T][kEN + 1], xc[T][kEN + 1], ur[T][kEP], uc[T][kEP], gr[T][kEP], ar[T][kEN], ac[T][kEN];
  __shared__ double il[kEN], Bs[kEP * kEP], red[8][kEGrad];
  const int r = blockIdx.z, tid = threadIdx.x;
  const int r0 = blockIdx.y * T, c0 = blockIdx.x * T;
  X += (long long)r * N * n;
  UH += (long long)r * N * p;
  Pinv += (long long)r * Npad * Npad;
  alphaAi += (long long)r * N * nd;
  alpha += (long long)r * N * nd;
  if (tid < n) il[tid] = 1.0 / ls[(long long)r * n + tid];
  if (tid < p * p) Bs[tid] = Bm[(long long)r * p * p + tid];
  __syncthreads();
  for (int idx = tid; idx < T * n; idx += 256) {
    int i = idx / n, d = idx % n;
    xr[i][d] = (r0 + i < N) ? X[(long long)(r0 + i) * n + d] * il[d] : 0.0;
    xc[i][d] = (c0 + i < N) ? X[(long long)(c0 + i) * n + d] * il[d] : 0.0;
  }
  for (int idx = tid; idx < T * nd; idx += 256) {
    int i = idx / nd, d = idx % nd;
    ar[i][d] = (r0 + i < N) ? alphaAi[(long long)(r0 + i) * nd + d] : 0.0;
    ac[i][d] = (c0 + i < N) ? alpha[(long long)(c0 + i) * nd + d] : 0.0;
  }
  for (int idx = tid; idx < T * p; idx += 256) {
    int i = idx / p, q = idx % p;
    double g = 0.0, u = 0.0;
    if (r0 + i < N) {
      u = UH[(long long)(r0 + i) * p + q];
      for (int t = 0; t < p; ++t) g += UH[(long long)(r0 + i) * p + t] * Bs[t * p + q];
    }
    ur[i][q] = u;
    gr[i][q] = g;
    uc[i][q] = (c0 + i < N) ? UH[(long long)(c0 + i) * p + q] : 0.0;
  }
  __syncthreads();
  const double s = scale[r];
  double acc[kEGrad];
#pragma unroll
  for (int t = 0; t < kEGrad; ++t) acc[t] = 0.0;
  const int ty = tid >> 4, tx = tid & 15;
  for (int i = 0; i < 4; ++i) {
    const int rl = ty * 4 + i, row = r0 + rl;
    if (row >= N) continue;
    for (int j = 0; j < 4; ++j) {
      const int cl = tx * 4 + j, col = c0 + cl;
      if (col >= N) continue;
      double d2 = 0.0, dd[kEN];
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
      const double gbar = 0.5 * (aa - (double)nd * Pinv[(long long)row * Npad + col]);
      const double ge = gbar * e;
      acc[0] += ge * S;
      const double gk = ge * S * s;
      for (int d = 0; d < n; ++d) acc[1 + d] += gk * dd[d] * il[d];
      const double gs = ge * s;
      for (int a = 0; a < p; ++a)
        for (int b = 0; b < p; ++b) acc[1 + kEN + a * kEP + b] += gs * ur[rl][a] * uc[cl][b];
    }
  }
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int t = 0; t < kEGrad; ++t) {
    double v = warp_sum(acc[t]);
    if (lane == 0) red[warp][t] = v;
  }
  __syncthreads();
  if (tid < kEGrad) {
    double v = 0.0;
    for (int w = 0; w < 8; ++w) v += red[w][tid];
    const long long blk = ((long long)r * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    partial[blk * kEGrad + tid] = v;
  }
}

__global__ void ens_gram_backward_finalize_kernel(const double* __restrict__ partial, int nblocks, double* __restrict__ out) {
  // grid (kEGrad, R), one warp each
  const int t = blockIdx.x, r = blockIdx.y, lane = threadIdx.x;
  double v = 0.0;
  for (int b = lane; b < nblocks; b += 32) v += partial[((long long)r * nblocks + b) * kEGrad + t];
  v = warp_sum(v);
  if (lane == 0) out[(long long)r * kEGrad + t] = v;
}

// LinvT[r][k][i] = Linv[r][i][k]  (32x32 smem tiles)
__global__ void ens_transpose_kernel(const double* __restrict__ A, double* __restrict__ At, int Npad) {
  __shared__ double t[32][33];
  A += (long long)blockIdx.z * Npad * Npad;
  At += (long long)blockIdx.z * Npad * Npad;
  const int x = blockIdx.x * 32 + threadIdx.x, y0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) t[j][threadIdx.x] = A[(long long)(y0 + j) * Npad + x];
  __syncthreads();
  const int xo = blockIdx.y * 32 + threadIdx.x, yo0 = blockIdx.x * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) At[(long long)(yo0 + j) * Npad + xo] = t[threadIdx.x][j];
}

}  // namespace bcbf

using namespace bcbf;

extern "C" int bcbf_ens_gram(const double* X, const double* UH, const double* lengthscale, const double* outputscale,
                             const double* Bmat, int R, int N, int n, int p, double* Kb, int Npad, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BCBF_REQUIRE(X && UH && lengthscale && outputscale && Bmat && Kb, "bcbf_ens_gram: null pointer");
  BCBF_REQUIRE(R >= 1 && N >= 1 && Npad >= N && Npad % 4 == 0 && n >= 1 && n <= kEN && p >= 1 && p <= kEP,
               "bcbf_ens_gram: R=%d N=%d Npad=%d n=%d p=%d", R, N, Npad, n, p);
  dim3 grid(ceil_div(Npad, 64), ceil_div(Npad, 64), R);
  ens_gram_kernel<<<grid, 256, 0, stream>>>(X, UH, lengthscale, outputscale, Bmat, N, Npad, n, p, Kb);
  BCBF_LAUNCH_CHECK();
  return BCBF_OK;
}

extern "C" int bcbf_ens_prep(const double* UH, const double* Xdot, const double* Bmat, const double* C, int R, int N,
                             int Npad, int n, int p, int ldy, double* G, double* Y, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BCBF_REQUIRE(UH && Xdot && Bmat && C && G && Y, "bcbf_ens_prep: null pointer");
  BCBF_REQUIRE(R >= 1 && N >= 1 && Npad >= N && ldy >= n && n <= kEN && p <= kEP, "bcbf_ens_prep: bad sizes");
  ens_prep_kernel<<<ceil_div((long long)R * Npad, 128), 128, 0, stream>>>(UH, Xdot, Bmat, C, R, N, Npad, n, p, ldy, G, Y);
  BCBF_LAUNCH_CHECK();
  return BCBF_OK;
}

extern "C" int bcbf_ens_w(const double* alpha, int ldy, const double* G, int R, int Npad, int n, int p, double* W,
                          void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BCBF_REQUIRE(alpha && G && W && R >= 1 && Npad >= 1, "bcbf_ens_w: bad arguments");
  ens_w_kernel<<<ceil_div((long long)R * Npad, 128), 128, 0, stream>>>(alpha, ldy, G, (long long)R * Npad, n, p, W);
  BCBF_LAUNCH_CHECK();
  return BCBF_OK;
}

extern "C" int bcbf_ens_transpose(const double* A, double* At, int Npad, int R, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BCBF_REQUIRE(A && At && A != At && Npad > 0 && Npad % 32 == 0 && R >= 1, "bcbf_ens_transpose: bad arguments");
  ens_transpose_kernel<<<dim3(Npad / 32, Npad / 32, R), dim3(32, 8), 0, stream>>>(A, At, Npad);
  BCBF_LAUNCH_CHECK();
  return BCBF_OK;
}

extern "C" int bcbf_ens_posterior(const double* Linv, const double* X, const double* G, const double* W,
                                  const double* lengthscale, const double* outputscale, const double* Bmat,
                                  const double* C, const double* xq, int R, int N, int Npad, int n, int p, double* Mk,
                                  double* Bk, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BCBF_REQUIRE(Linv && X && G && W && lengthscale && outputscale && Bmat && C && xq && Mk && Bk,
               "bcbf_ens_posterior: null pointer");
  BCBF_REQUIRE(R >= 1 && N >= 1 && Npad >= N && n >= 1 && n <= kEN && p >= 1 && p <= kEP,
               "bcbf_ens_posterior: R=%d N=%d Npad=%d n=%d p=%d", R, N, Npad, n, p);
  const int smem = (int)sizeof(double) * Npad * kEP;
  BCBF_REQUIRE(smem <= 200 * 1024, "bcbf_ens_posterior: Npad=%d too large for the per-rollout kernel", Npad);
  BCBF_CUDA(cudaFuncSetAttribute(ens_posterior_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  ens_posterior_kernel<<<R, kEnsThreads, smem, stream>>>(Linv, X, G, W, lengthscale, outputscale, Bmat, C, xq, N, Npad, n, p,
                                                 Mk, Bk);
  BCBF_LAUNCH_CHECK();
  return BCBF_OK;
}

extern "C" int bcbf_ens_gram_backward(const double* X, const double* UH, const double* lengthscale,
                                      const double* outputscale, const double* Bmat, const double* Pinv,
                                      const double* alphaAi, const double* alpha, int R, int N, int Npad, int n, int p,
                                      int nout, double* partial, long long partial_elems, double* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BCBF_REQUIRE(X && UH && lengthscale && outputscale && Bmat && Pinv && alphaAi && alpha && partial && out,
               "bcbf_ens_gram_backward: null pointer");
  BCBF_REQUIRE(R >= 1 && N >= 1 && Npad >= N && n >= 1 && n <= kEN && p >= 1 && p <= kEP && nout >= 1 && nout <= kEN,
               "bcbf_ens_gram_backward: R=%d N=%d Npad=%d n=%d p=%d nout=%d", R, N, Npad, n, p, nout);
  dim3 grid(ceil_div(N, 64), ceil_div(N, 64), R);
  const long long nblocks = (long long)grid.x * grid.y;
  BCBF_REQUIRE(partial_elems >= nblocks * R * kEGrad, "bcbf_ens_gram_backward: partial buffer too small");
  ens_gram_backward_kernel<<<grid, 256, 0, stream>>>(X, UH, lengthscale, outputscale, Bmat, Pinv, alphaAi, alpha, N, Npad,
                                                     n, p, nout, partial);
  BCBF_LAUNCH_CHECK();
  ens_gram_backward_finalize_kernel<<<dim3(kEGrad, R), 32, 0, stream>>>(partial, (int)nblocks, out);
  BCBF_LAUNCH_CHECK();
  return BCBF_OK;
}
