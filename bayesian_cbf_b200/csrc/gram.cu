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
  double inv_ls[kMaxN];
  double Bm[kMaxP * kMaxP];
  double scale;
  int a, c, n, p;
  double* out;
  int ld;
  int rows_out, cols_out;  // padded extents to fill (>= a, >= c)
  int pad_identity;        // train mode: identity on the pad diagonal
  int vec_ok;              // 16-byte stores allowed (even ld, aligned base)
};

template <bool TRAIN>
__global__ void __launch_bounds__(256) gram_kernel(GramParams P) {
  __shared__ double xr[kGT][kMaxN + 1], xc[kGT][kMaxN + 1];
  __shared__ double gr[kGT][kMaxP], uc[kGT][kMaxP];
  const int tid = threadIdx.x;
  const int r0 = blockIdx.y * kGT, c0 = blockIdx.x * kGT;
  const int n = P.n, p = P.p;
  for (int idx = tid; idx < kGT * n; idx += 256) {
    int r = idx / n, d = idx % n;
    xr[r][d] = (r0 + r < P.a) ? P.X1[(long long)(r0 + r) * n + d] * P.inv_ls[d] : 0.0;
    xc[r][d] = (c0 + r < P.c) ? P.X2[(long long)(c0 + r) * n + d] * P.inv_ls[d] : 0.0;
  }
  if (TRAIN) {
    for (int idx = tid; idx < kGT * p; idx += 256) {
      int r = idx / p, q = idx % p;
      double g = 0.0;
      if (r0 + r < P.a)
        for (int t = 0; t < p; ++t) g += P.UH[(long long)(r0 + r) * p + t] * P.Bm[t * p + q];
      gr[r][q] = g;
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
        double d2 = 0.0;
        for (int d = 0; d < n; ++d) {
          double df = xr[rl][d] - xc[cl][d];
          d2 = fma(df, df, d2);
        }
        out = P.scale * exp(-0.5 * d2);
        if (TRAIN) {
          double ub = 0.0;
          for (int q = 0; q < p; ++q) ub = fma(gr[rl][q], uc[cl][q], ub);
          out *= ub;
        }
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

// k, dk/dx1 (a,c,n), d2k/dx1dx2 (a,c,n,n); one thread per (i,j) pair — small-b API path only.
__global__ void rbf_blocks_kernel(GramParams P, double* K, double* dK, double* d2K) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)P.a * P.c) return;
  const int i = (int)(idx / P.c), j = (int)(idx % P.c), n = P.n;
  double w[kMaxN];  // (x1 - x2) / l^2
  double d2 = 0.0;
  for (int d = 0; d < n; ++d) {
    double df = (P.X1[(long long)i * n + d] - P.X2[(long long)j * n + d]) * P.inv_ls[d];
    d2 = fma(df, df, d2);
    w[d] = df * P.inv_ls[d];
  }
  const double k = P.scale * exp(-0.5 * d2);
  if (K) K[idx] = k;
  if (dK)
    for (int d = 0; d < n; ++d) dK[idx * n + d] = -w[d] * k;
  if (d2K)
    for (int d = 0; d < n; ++d)
      for (int e = 0; e < n; ++e)
        d2K[(idx * n + d) * n + e] = ((d == e ? P.inv_ls[d] * P.inv_ls[d] : 0.0) - w[d] * w[e]) * k;
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
  __shared__ double xr[kGT][kMaxN + 1], xc[kGT][kMaxN + 1];
  __shared__ double ur[kGT][kMaxP], uc[kGT][kMaxP], gr[kGT][kMaxP];
  __shared__ double ar[kGT][kMaxN], ac[kGT][kMaxN];
  __shared__ double red[8][kGradMaxOut];
  const int tid = threadIdx.x;
  const int r0 = blockIdx.y * kGT, c0 = blockIdx.x * kGT;
  const int n = P.n, p = P.p, nd = nout_dim;
  for (int idx = tid; idx < kGT * n; idx += 256) {
    int r = idx / n, d = idx % n;
    xr[r][d] = (r0 + r < P.a) ? P.X1[(long long)(r0 + r) * n + d] * P.inv_ls[d] : 0.0;
    xc[r][d] = (c0 + r < P.c) ? P.X1[(long long)(c0 + r) * n + d] * P.inv_ls[d] : 0.0;
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
      for (int t = 0; t < p; ++t) g += P.UH[(long long)(r0 + r) * p + t] * P.Bm[t * p + q];
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
      for (int d = 0; d < n; ++d) acc[1 + d] += gk * dd[d] * P.inv_ls[d];   // (dx/l)^2 / l
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

// Small parameter blocks (lengthscale, B) may live on the host or on the device.  Host pointers are read directly;
// device pointers are fetched on the stream (one short synchronisation — the torch-tensor call path).
static int fetch_small(double* dst, const double* src, int count, cudaStream_t stream) {
  cudaPointerAttributes attr{};
  cudaError_t e = cudaPointerGetAttributes(&attr, src);
  if (e != cudaSuccess) { cudaGetLastError(); attr.type = cudaMemoryTypeUnregistered; }
  if (attr.type == cudaMemoryTypeUnregistered || attr.type == cudaMemoryTypeHost) {
    for (int i = 0; i < count; ++i) dst[i] = src[i];
    return BCBF_OK;
  }
  e = cudaMemcpyAsync(dst, src, sizeof(double) * count, cudaMemcpyDeviceToHost, stream);
  if (e != cudaSuccess) return cuda_fail(e, "copy hyper-parameters", __FILE__, __LINE__);
  e = cudaStreamSynchronize(stream);
  if (e != cudaSuccess) return cuda_fail(e, "sync hyper-parameters", __FILE__, __LINE__);
  return BCBF_OK;
}

static int fill_common(GramParams& P, const double* lengthscale, double outputscale, int n, cudaStream_t stream) {
  double ls[kMaxN];
  int rc = fetch_small(ls, lengthscale, n, stream);
  if (rc != BCBF_OK) return rc;
  for (int d = 0; d < n; ++d) P.inv_ls[d] = 1.0 / ls[d];
  P.scale = outputscale;
  P.n = n;
  return BCBF_OK;
}

}  // namespace bcbf

using namespace bcbf;

extern "C" int bcbf_gram_train(const double* X, const double* UH, const double* Bmat, const double* lengthscale,
                               double outputscale, int N, int n, int p, double* Kb, int ld, int Npad,
                               void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BCBF_REQUIRE(X && UH && Bmat && lengthscale && Kb, "bcbf_gram_train: null pointer");
  BCBF_REQUIRE(n >= 1 && n <= kMaxN && p >= 1 && p <= kMaxP, "bcbf_gram_train: n=%d (<=%d) p=%d (<=%d)", n, kMaxN,
               p, kMaxP);
  BCBF_REQUIRE(N >= 1 && Npad >= N && Npad % 2 == 0 && ld >= Npad && ld % 2 == 0,
               "bcbf_gram_train: N=%d Npad=%d ld=%d", N, Npad, ld);
  GramParams P{};
  int rc = fill_common(P, lengthscale, outputscale, n, stream);
  if (rc != BCBF_OK) return rc;
  if ((rc = fetch_small(P.Bm, Bmat, p * p, stream)) != BCBF_OK) return rc;
  P.X1 = X; P.X2 = X; P.UH = UH; P.UH2 = UH; P.a = N; P.c = N; P.p = p;
  P.out = Kb; P.ld = ld; P.rows_out = Npad; P.cols_out = Npad; P.pad_identity = 1; P.vec_ok = 1;
  dim3 grid(ceil_div(Npad, kGT), ceil_div(Npad, kGT));
  gram_kernel<true><<<grid, 256, 0, stream>>>(P);
  BCBF_LAUNCH_CHECK();
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
  if (UH1 && (rc = fetch_small(P.Bm, Bmat, p * p, stream)) != BCBF_OK) return rc;
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
  if ((rc = fetch_small(P.Bm, Bmat, p * p, stream)) != BCBF_OK) return rc;
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
