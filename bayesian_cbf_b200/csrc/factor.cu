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

__global__ void __launch_bounds__(256, 1)
potf2_inv_kernel(double* __restrict__ A, int ld, int k0, const double* __restrict__ jitter, int N, double jscale,
                 double* __restrict__ dinv, int* __restrict__ info) {
  extern __shared__ __align__(16) double sm[];
  double* a = sm;                 // [128][129]
  double* xd = sm + kBlk * kPad;  // [128] reciprocal diagonal
  __shared__ int failed;
  const int tid = threadIdx.x;
  if (tid == 0) failed = 0;
  for (int idx = tid; idx < kBlk * kBlk; idx += 256) {
    int r = idx >> 7, c = idx & 127;
    double v = 0.0;
    if (c <= r) {
      v = A[(long long)(k0 + r) * ld + (k0 + c)];
      if (c == r && jitter != nullptr && (k0 + r) < N) v += jscale * jitter[k0 + r];
    }
    a[r * kPad + c] = v;
  }
  __syncthreads();
  for (int j = 0; j < kBlk; ++j) {
    const double ajj = a[j * kPad + j];
    if (!(ajj > 0.0)) {  // also catches NaN; uniform across the CTA
      if (tid == 0) {
        atomicCAS(info, 0, k0 + j + 1);
        failed = 1;
      }
      break;
    }
    const double d = sqrt(ajj);
    __syncthreads();  // everyone has read a[j][j]
    if (tid < kBlk) {
      if (tid > j) a[tid * kPad + j] /= d;
      else if (tid == j) a[j * kPad + j] = d;
    }
    __syncthreads();
    const int i = j + 1 + (tid >> 1);
    if (i < kBlk) {
      const double lij = a[i * kPad + j];
      for (int k = j + 1 + (tid & 1); k <= i; k += 2) a[i * kPad + k] -= lij * a[k * kPad + j];
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
  // ---- X = L^{-1}: X[i][j] (i > j) is kept at a[j][i] (strict upper part), diagonal in xd -----------
  if (tid < kBlk) xd[tid] = 1.0 / a[tid * kPad + tid];
  __syncthreads();
  {
    const int j = tid >> 1, h = tid & 1;
    for (int i = 1; i < kBlk; ++i) {
      double s = 0.0;
      if (j < i) {
        // sum_{k=j}^{i-1} L[i][k] X[k][j],  X[j][j] = xd[j]
        for (int k = j + h; k < i; k += 2) {
          double xkj = (k == j) ? xd[j] : a[j * kPad + k];
          s += a[i * kPad + k] * xkj;
        }
      }
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      __syncthreads();  // all reads of column i of the upper part (none yet) / row i done before the write
      if (j < i && h == 0) a[j * kPad + i] = -s * xd[i];
      __syncthreads();
    }
  }
  for (int idx = tid; idx < kBlk * kBlk; idx += 256) {
    int r = idx >> 7, c = idx & 127;
    double l = (c <= r) ? a[r * kPad + c] : 0.0;
    double x = (c < r) ? a[c * kPad + r] : (c == r ? xd[r] : 0.0);
    A[(long long)(k0 + r) * ld + (k0 + c)] = l;
    dinv[idx] = x;
  }
}

__global__ void zero_upper_blocks_kernel(double* __restrict__ A, int ld, int nb) {
  // one CTA per strictly-upper 128x128 block (bi < bj)
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

__global__ void scatter_diag_blocks_kernel(const double* __restrict__ dinv, double* __restrict__ Linv, int ld) {
  const double* src = dinv + (long long)blockIdx.x * kBlk * kBlk;
  double* dst = Linv + (long long)blockIdx.x * kBlk * (ld + 1);
  for (int idx = threadIdx.x; idx < kBlk * kBlk / 2; idx += blockDim.x) {
    int r = idx >> 6, c = (idx & 63) * 2;
    *reinterpret_cast<double2*>(dst + (long long)r * ld + c) = *reinterpret_cast<const double2*>(src + r * kBlk + c);
  }
}

}  // namespace bcbf

using namespace bcbf;

extern "C" long long bcbf_dinv_elems(int Npad) { return (long long)(Npad / kBlk) * kBlk * kBlk; }

extern "C" int bcbf_potrf(double* A, int ld, int Npad, int N, const double* jitter, double jitter_scale,
                          double* dinv, int* info, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BCBF_REQUIRE(A && dinv && info, "bcbf_potrf: null pointer");
  BCBF_REQUIRE(Npad > 0 && Npad % kBlk == 0 && ld >= Npad && ld % 2 == 0 && N <= Npad && N >= 0,
               "bcbf_potrf: Npad=%d must be a positive multiple of %d, ld=%d >= Npad and even, N=%d <= Npad", Npad,
               kBlk, ld, N);
  const int nb = Npad / kBlk;
  const int smem = (kBlk * kPad + kBlk) * (int)sizeof(double);
  BCBF_CUDA(cudaFuncSetAttribute(potf2_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  BCBF_CUDA(cudaMemsetAsync(info, 0, sizeof(int), stream));
  for (int k = 0; k < nb; ++k) {
    const int k0 = k * kBlk;
    double* dk = dinv + (long long)k * kBlk * kBlk;
    potf2_inv_kernel<<<1, 256, smem, stream>>>(A, ld, k0, jitter, N, jitter_scale, dk, info);
    BCBF_LAUNCH_CHECK();
    const int rows = Npad - (k0 + kBlk);
    if (rows <= 0) break;
    double* panel = A + (long long)(k0 + kBlk) * ld + k0;
    GemmArgs g{};
    // panel <- panel * Dinv_k^T       (L_ik = A_ik L_kk^{-T})
    g.A = panel; g.lda = ld; g.B = dk; g.ldb = kBlk; g.C = panel; g.ldc = ld;
    g.M = rows; g.N = kBlk; g.K = kBlk; g.alpha = 1.0; g.beta = 0.0; g.tri = kTriNone;
    BCBF_CUDA((launch_gemm<true, true>(g, 1, stream)));
    // trailing (lower tiles) -= panel panel^T
    GemmArgs s{};
    s.A = panel; s.lda = ld; s.B = panel; s.ldb = ld;
    s.C = A + (long long)(k0 + kBlk) * (ld + 1); s.ldc = ld;
    s.M = rows; s.N = rows; s.K = kBlk; s.alpha = -1.0; s.beta = 1.0; s.tri = kTriLowerOut;
    BCBF_CUDA((launch_gemm<true, true>(s, 1, stream)));
  }
  if (nb > 1) {
    zero_upper_blocks_kernel<<<nb * (nb - 1) / 2, 256, 0, stream>>>(A, ld, nb);
    BCBF_LAUNCH_CHECK();
  }
  return BCBF_OK;
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

extern "C" int bcbf_trtri(const double* L, const double* dinv, double* Linv, double* scratch, int ld, int Npad,
                          void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BCBF_REQUIRE(L && dinv && Linv && scratch, "bcbf_trtri: null pointer");
  BCBF_REQUIRE(Npad > 0 && Npad % kBlk == 0 && ld >= Npad && ld % 2 == 0, "bcbf_trtri: bad Npad=%d / ld=%d", Npad, ld);
  const int nb = Npad / kBlk;
  BCBF_CUDA(cudaMemsetAsync(Linv, 0, sizeof(double) * (size_t)ld * Npad, stream));
  scatter_diag_blocks_kernel<<<nb, 256, 0, stream>>>(dinv, Linv, ld);
  BCBF_LAUNCH_CHECK();
  for (int hb = 1; hb < nb; hb *= 2) {
    const int h = hb * kBlk;
    // pairs g: rows [r0, r0+h) (X11) and [r0+h, r0+2h) clipped to Npad (X22), r0 = g*2h
    const int npairs_full = Npad / (2 * h);                          // both halves complete
    const int rem = Npad - npairs_full * 2 * h;                      // leftover rows after the full pairs
    const int partial_rows = rem > h ? rem - h : 0;                  // clipped second half of the last pair
    for (int pass = 0; pass < 2; ++pass) {
      const int batch = pass == 0 ? npairs_full : (partial_rows > 0 ? 1 : 0);
      if (batch == 0) continue;
      const long long r0 = pass == 0 ? 0 : (long long)npairs_full * 2 * h;
      const int M2 = pass == 0 ? h : partial_rows;
      const long long stride = (long long)2 * h * (ld + 1);
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

extern "C" int bcbf_trmm_lower(const double* A, int lda, int Npad, int trans, const double* B, int ldb, int ncols,
                               double alpha, double beta, double* C, int ldc, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BCBF_REQUIRE(A && B && C, "bcbf_trmm_lower: null pointer");
  BCBF_REQUIRE(Npad > 0 && Npad % kBlk == 0 && lda >= Npad && lda % 2 == 0 && ldb % 2 == 0 && ldc % 2 == 0 &&
                   ncols > 0 && ncols % 2 == 0 && ldb >= ncols && ldc >= ncols,
               "bcbf_trmm_lower: Npad=%d lda=%d ldb=%d ldc=%d ncols=%d (leading dims and ncols must be even)", Npad,
               lda, ldb, ldc, ncols);
  GemmArgs g{};
  g.A = A; g.lda = lda; g.B = B; g.ldb = ldb; g.C = C; g.ldc = ldc;
  g.M = Npad; g.N = ncols; g.K = Npad; g.alpha = alpha; g.beta = beta;
  if (!trans) {
    g.tri = kTriALower;
    BCBF_CUDA((launch_gemm<true, false>(g, 1, stream)));
  } else {
    g.tri = kTriAUpper;
    BCBF_CUDA((launch_gemm<false, false>(g, 1, stream)));
  }
  return BCBF_OK;
}

// General row-major C(M,N) = alpha * op(A) op(B) + beta * C on the DMMA GEMM (host glue for the small-batch API
// paths: v^T v', kb*^T alpha, Linv^T Linv).  op(A) is M x K: transa = 0 -> A stored (M,K); 1 -> stored (K,M).
// op(B) is K x N: transb = 0 -> B stored (K,N); 1 -> stored (N,K).
extern "C" int bcbf_gemm(int transa, int transb, int M, int N, int K, double alpha, const double* A, int lda,
                         const double* B, int ldb, double beta, double* C, int ldc, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BCBF_REQUIRE(A && B && C, "bcbf_gemm: null pointer");
  BCBF_REQUIRE(M > 0 && N > 0 && K > 0 && M % 2 == 0 && N % 2 == 0 && K % 2 == 0,
               "bcbf_gemm: M=%d N=%d K=%d must be positive and even (pad with zeros)", M, N, K);
  BCBF_REQUIRE(lda % 2 == 0 && ldb % 2 == 0 && ldc % 2 == 0, "bcbf_gemm: leading dimensions must be even");
  BCBF_REQUIRE(((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B) | reinterpret_cast<uintptr_t>(C)) & 15) == 0,
               "bcbf_gemm: operands must be 16-byte aligned");
  GemmArgs g{};
  g.A = A; g.lda = lda; g.B = B; g.ldb = ldb; g.C = C; g.ldc = ldc;
  g.M = M; g.N = N; g.K = K; g.alpha = alpha; g.beta = beta; g.tri = kTriNone;
  if (!transa && !transb) BCBF_CUDA((launch_gemm<true, false>(g, 1, stream)));
  else if (!transa && transb) BCBF_CUDA((launch_gemm<true, true>(g, 1, stream)));
  else if (transa && !transb) BCBF_CUDA((launch_gemm<false, false>(g, 1, stream)));
  else BCBF_CUDA((launch_gemm<false, true>(g, 1, stream)));
  return BCBF_OK;
}
