// C-ABI glue: error reporting, CBC epilogue kernel and the host-pointer model handle (include/bcbf.h).
#include "../../include/bcbf.h"
#include "common.cuh"

#include <cstdarg>
#include <cstring>
#include <mutex>
#include <new>

namespace bcbf {

static thread_local char g_err[512] = "";
unsigned long long g_launch_count = 0;

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

namespace {
struct ScratchState {
  std::recursive_mutex mu;
  int depth = 0;
  cudaEvent_t ev = nullptr;
  bool recorded = false;
  cudaStream_t last = nullptr;
};
ScratchState g_scratch[64];
}  // namespace

// A stream that is being captured into a CUDA graph may neither wait for an event recorded outside the capture nor
// publish one: captured calls are ordered by whoever replays the graph.
static bool stream_is_capturing(cudaStream_t s) {
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(s, &cs) != cudaSuccess) { cudaGetLastError(); return false; }
  return cs != cudaStreamCaptureStatusNone;
}

ScratchScope::ScratchScope(cudaStream_t s) : stream(s), dev(0) {
  if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
  ScratchState& st = g_scratch[dev & 63];
  st.mu.lock();
  if (st.depth++ == 0 && !stream_is_capturing(s)) {
    if (st.ev == nullptr) cudaEventCreateWithFlags(&st.ev, cudaEventDisableTiming);
    // the previous user of the scratch ran on another stream: order this call behind it on the device
    if (st.recorded && st.last != s && st.ev != nullptr) cudaStreamWaitEvent(s, st.ev, 0);
  }
}

ScratchScope::~ScratchScope() {
  ScratchState& st = g_scratch[dev & 63];
  if (--st.depth == 0 && st.ev != nullptr && !stream_is_capturing(stream)) {
    if (cudaEventRecord(st.ev, stream) == cudaSuccess) {
      st.recorded = true;
      st.last = stream;
    }
  }
  st.mu.unlock();
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  set_last_error("CUDA error %s (%s) at %s:%d: %s", cudaGetErrorName(e), cudaGetErrorString(e), file, line, what);
  return BCBF_ERR_CUDA;
}

// ---- relative-degree-1 CBC terms, one thread per constraint --------------------------------------------
__global__ void cbc1_terms_kernel(const double* __restrict__ Mk, const double* __restrict__ Bk,
                                  const double* __restrict__ Amat, const double* __restrict__ grad_h,
                                  const double* __restrict__ h, const double* __restrict__ Fbar, double gamma, int n,
                                  int p, int Q, double* __restrict__ bfe, double* __restrict__ e,
                                  double* __restrict__ Asq, double* __restrict__ A_socp, double* __restrict__ bfb,
                                  int* __restrict__ status) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Q) return;
  const int m = p - 1;
  double gh[BCBF_MAX_N_DIM];
  for (int r = 0; r < n; ++r) gh[r] = grad_h[(long long)q * n + r];
  // affine (mean) terms
  for (int j = 0; j < p; ++j) {
    double s = 0.0;
    for (int r = 0; r < n; ++r) {
      double f = Mk[((long long)q * n + r) * p + j];
      if (Fbar) f += Fbar[((long long)q * n + r) * p + j];
      s = fma(gh[r], f, s);
    }
    if (j == 0) e[q] = s + gamma * h[q];
    else bfe[(long long)q * m + (j - 1)] = s;
  }
  // quadratic (variance) terms: Asq = (gh^T A gh) * B_k
  double sA = 0.0;
  for (int r = 0; r < n; ++r)
    for (int c = 0; c < n; ++c) sA = fma(gh[r] * Amat[r * n + c], gh[c], sA);
  double M[BCBF_MAX_P_DIM][BCBF_MAX_P_DIM], Ls[BCBF_MAX_P_DIM][BCBF_MAX_P_DIM];
  for (int i = 0; i < p; ++i)
    for (int j = 0; j < p; ++j) {
      M[i][j] = sA * Bk[((long long)q * p + i) * p + j];
      Ls[i][j] = 0.0;
      if (Asq) Asq[((long long)q * p + i) * p + j] = M[i][j];
    }
  int st = 0;
  for (int j = 0; j < p; ++j) {
    double d = M[j][j];
    for (int k = 0; k < j; ++k) d -= Ls[j][k] * Ls[j][k];
    if (!(d > 0.0)) {
      if (st == 0) st = j + 1;
      d = __longlong_as_double(0x7ff8000000000000LL);
    }
    d = sqrt(d);
    Ls[j][j] = d;
    for (int i = j + 1; i < p; ++i) {
      double s = M[i][j];
      for (int k = 0; k < j; ++k) s -= Ls[i][k] * Ls[j][k];
      Ls[i][j] = s / d;
    }
  }
  if (status) status[q] = st;
  // A_socp = Ls^T[:, 1:]  (p, m);  bfb = Ls^T[:, 0]  (p)
  for (int i = 0; i < p; ++i) {
    if (bfb) bfb[(long long)q * p + i] = Ls[0][i];
    if (A_socp)
      for (int c = 0; c < m; ++c) A_socp[((long long)q * p + i) * m + c] = Ls[c + 1][i];
  }
}

// ---- batched factorisation of the SOCP matrix Asq = [[v, bfv^T/2],[bfv/2, V]] = Ls Ls^T (p x p, p <= 4) -------------
__global__ void socp_factor_kernel(const double* __restrict__ Asq, int p, int Q, double reg, double* __restrict__ A_socp,
                                   double* __restrict__ bfb, int* __restrict__ status) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Q) return;
  const int m = p - 1;
  double M[BCBF_MAX_P_DIM][BCBF_MAX_P_DIM], Ls[BCBF_MAX_P_DIM][BCBF_MAX_P_DIM];
  int st = 0;
  for (int attempt = 0; attempt < 2; ++attempt) {
    for (int i = 0; i < p; ++i)
      for (int j = 0; j < p; ++j) {
        M[i][j] = Asq[((long long)q * p + i) * p + j] + ((attempt == 1 && i == j) ? reg : 0.0);
        Ls[i][j] = 0.0;
      }
    st = 0;
    for (int j = 0; j < p; ++j) {
      double d = M[j][j];
      for (int k = 0; k < j; ++k) d -= Ls[j][k] * Ls[j][k];
      if (!(d > 0.0)) {
        if (st == 0) st = j + 1;
        d = __longlong_as_double(0x7ff8000000000000LL);
      }
      d = sqrt(d);
      Ls[j][j] = d;
      for (int i = j + 1; i < p; ++i) {
        double s = M[i][j];
        for (int k = 0; k < j; ++k) s -= Ls[i][k] * Ls[j][k];
        Ls[i][j] = s / d;
      }
    }
    if (st == 0 || reg <= 0.0) break;   // second attempt only when the caller asked for the singular fallback
  }
  if (status) status[q] = st;
  for (int i = 0; i < p; ++i) {
    bfb[(long long)q * p + i] = Ls[0][i];
    for (int c = 0; c < m; ++c) A_socp[((long long)q * p + i) * m + c] = Ls[c + 1][i];
  }
}

// ---- model-handle helper kernels -----------------------------------------------------------------------
__global__ void prep_train_kernel(const double* __restrict__ U, const double* __restrict__ Xdot, int N, int Npad, int n,
                                  int p, const double* __restrict__ Bm, const double* __restrict__ C,
                                  double* __restrict__ UH, double* __restrict__ G, double* __restrict__ Y, int ldy) {
  // UH = [1 | U] (N,p); G = UH B (Npad,p; pad rows 0); Y = Xdot - UH C (Npad, ldy; pad 0)
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Npad) return;
  const int m = p - 1;
  double uh[BCBF_MAX_P_DIM];
  if (i < N) {
    uh[0] = 1.0;
    for (int j = 0; j < m; ++j) uh[j + 1] = U[(long long)i * m + j];
    for (int j = 0; j < p; ++j) UH[(long long)i * p + j] = uh[j];
  }
  for (int j = 0; j < p; ++j) G[(long long)i * p + j] = (i < N) ? g_entry(uh, Bm, p, j) : 0.0;
  for (int r = 0; r < ldy; ++r) {
    double y = 0.0;
    if (i < N && r < n) {
      y = Xdot[(long long)i * n + r];
      for (int t = 0; t < p; ++t) y -= uh[t] * C[t * n + r];
    }
    Y[(long long)i * ldy + r] = y;
  }
}

__global__ void build_w_kernel(const double* __restrict__ alpha, int lda, const double* __restrict__ G, int Npad,
                               int n, int p, double* __restrict__ W) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Npad) return;
  for (int r = 0; r < n; ++r)
    for (int j = 0; j < p; ++j) W[((long long)i * n + r) * p + j] = alpha[(long long)i * lda + r] * G[(long long)i * p + j];
}

// ---- skinny triangular products: y = beta * y0 + alpha * op(T) x,  T lower triangular (Npad,Npad; ld, strictly upper part
// stored as zero), x / y / y0 (Npad, nc <= 8 columns, leading dimension ldx).  HBM-bound: T is read once (4 Npad^2 bytes).
// Used for alpha = Kb^-1 Y and its refinement step (control_affine_model.py:545).
constexpr int kMvMaxC = 8;
__global__ void __launch_bounds__(256) tri_mv_n_kernel(const double* __restrict__ T, int ld, int Npad,
                                                       const double* __restrict__ x, int ldx, int nc, double alpha,
                                                       double beta, const double* __restrict__ y0,
                                                       double* __restrict__ y) {
  const int row = blockIdx.x * 8 + threadIdx.x / 32, lane = threadIdx.x % 32;
  if (row >= Npad) return;
  double acc[kMvMaxC];
#pragma unroll
  for (int c = 0; c < kMvMaxC; ++c) acc[c] = 0.0;
  const double* t = T + (long long)row * ld;
  int k = lane;
  for (; k + 96 <= row; k += 128) {  // four independent row segments in flight per lane
    double v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = t[k + 32 * u];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int c = 0; c < kMvMaxC; ++c)
        if (c < nc) acc[c] = fma(v[u], x[(long long)(k + 32 * u) * ldx + c], acc[c]);
  }
  for (; k <= row; k += 32) {
    const double v = t[k];
#pragma unroll
    for (int c = 0; c < kMvMaxC; ++c)
      if (c < nc) acc[c] = fma(v, x[(long long)k * ldx + c], acc[c]);
  }
#pragma unroll
  for (int c = 0; c < kMvMaxC; ++c)
    if (c < nc) {
      const double s = warp_sum(acc[c]);
      if (lane == 0) y[(long long)row * ldx + c] = (y0 ? beta * y0[(long long)row * ldx + c] : 0.0) + alpha * s;
    }
}

// op(T) = T^T: thread per column j of a 128-column strip, rows split over blockIdx.y (mv_rows(Npad) rows per split, a
// function of Npad alone so that the summation order — and the result bits — depend on nothing else); partial[split][j][c]
static int mv_rows(int Npad) {
  // enough (strip, split) CTAs to cover the machine at small Npad: 1024 rows per split from Npad = 16384 down to 64 at
  // Npad <= 1024 (Npad = 256: 2 x 4 CTAs of 64 rows instead of 2 CTAs walking 256 rows, 26 us -> ~8 us)
  long long r = (long long)Npad * Npad / (128LL * 148LL);
  int rows = 64;
  while (rows * 2 <= r && rows < 1024) rows *= 2;
  return rows;
}
__global__ void __launch_bounds__(128) tri_mv_t_kernel(const double* __restrict__ T, int ld, int Npad,
                                                       const double* __restrict__ x, int ldx, int nc,
                                                       double* __restrict__ partial, int kMvRows) {
  __shared__ double xs[64 * kMvMaxC];
  const int j = blockIdx.x * 128 + threadIdx.x;
  const int i0 = max(blockIdx.y * kMvRows, blockIdx.x * 128), i1 = min(Npad, (blockIdx.y + 1) * kMvRows);
  double acc[kMvMaxC];
#pragma unroll
  for (int c = 0; c < kMvMaxC; ++c) acc[c] = 0.0;
  for (int ib = i0; ib < i1; ib += 64) {   // i0 and i1 are multiples of 64
    __syncthreads();
    for (int e = threadIdx.x; e < 64 * nc; e += 128) xs[(e / nc) * kMvMaxC + e % nc] = x[(long long)(ib + e / nc) * ldx + e % nc];
    __syncthreads();
    for (int r = 0; r < 64; r += 8) {
      double v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = (ib + r + u >= j) ? T[(long long)(ib + r + u) * ld + j] : 0.0;
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int c = 0; c < kMvMaxC; ++c)
          if (c < nc) acc[c] = fma(v[u], xs[(r + u) * kMvMaxC + c], acc[c]);
    }
  }
  for (int c = 0; c < nc; ++c) partial[((long long)blockIdx.y * Npad + j) * nc + c] = acc[c];
}

__global__ void tri_mv_t_finalize_kernel(const double* __restrict__ partial, int Npad, int nc, int nsplit, int ldx,
                                         double alpha, double beta, const double* __restrict__ y0,
                                         double* __restrict__ y, int kMvRows) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)Npad * nc) return;
  const int j = static_cast<int>(idx / nc), c = static_cast<int>(idx % nc);
  double s = 0.0;
  for (int sp = j / kMvRows; sp < nsplit; ++sp) s += partial[((long long)sp * Npad + j) * nc + c];
  y[(long long)j * ldx + c] = (y0 ? beta * y0[(long long)j * ldx + c] : 0.0) + alpha * s;
}

// ---- packed lower triangle of a factor-sized matrix (multi-GPU broadcast of L^-1) -----------------------------------------
// packed layout: block row after block row, block row i = a (128, 128 (i+1)) row-major matrix.  One CTA per 128x128 block
// of the full matrix: lower blocks are copied, strictly-upper blocks are zeroed on unpack.
__global__ void __launch_bounds__(256) pack_lower_kernel(const double* __restrict__ M, int ld, double* __restrict__ buf,
                                                         int unpack, double* __restrict__ Mout) {
  const int bi = blockIdx.y, bj = blockIdx.x;
  const long long off = (long long)kBlk * kBlk * bi * (bi + 1) / 2;      // start of block row bi in the packed buffer
  const int w = (bi + 1) * kBlk;
  for (int idx = threadIdx.x; idx < kBlk * kBlk / 2; idx += blockDim.x) {
    const int r = idx >> 6, c = (idx & 63) * 2;
    const long long m = (long long)(bi * kBlk + r) * ld + bj * kBlk + c;
    const long long b = off + (long long)r * w + bj * kBlk + c;
    if (!unpack) {
      if (bj <= bi) *reinterpret_cast<double2*>(buf + b) = *reinterpret_cast<const double2*>(M + m);
    } else {
      *reinterpret_cast<double2*>(Mout + m) = bj <= bi ? *reinterpret_cast<const double2*>(buf + b) : make_double2(0.0, 0.0);
    }
  }
}

__global__ void build_uh_kernel(const double* __restrict__ U, int Q, int p, double* __restrict__ UH) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Q) return;
  UH[(long long)i * p] = 1.0;
  for (int j = 1; j < p; ++j) UH[(long long)i * p + j] = U ? U[(long long)i * (p - 1) + (j - 1)] : 0.0;
}

}  // namespace bcbf

using namespace bcbf;

extern "C" const char* bcbf_last_error(void) { return g_err; }
extern "C" int bcbf_version(void) { return 100; }
extern "C" unsigned long long bcbf_launch_count(void) { return g_launch_count; }
extern "C" int bcbf_padded(int N) { return ((N + kBlk - 1) / kBlk) * kBlk; }

extern "C" long long bcbf_packed_lower_elems(int Npad) {
  const long long nb = Npad / kBlk;
  return (long long)kBlk * kBlk * nb * (nb + 1) / 2;
}

extern "C" int bcbf_pack_lower(const double* M, int ld, int Npad, double* buf, void* stream_) {
  BCBF_REQUIRE(M && buf && Npad > 0 && Npad % kBlk == 0 && ld >= Npad && ld % 2 == 0, "bcbf_pack_lower: Npad=%d ld=%d", Npad, ld);
  const int nb = Npad / kBlk;
  pack_lower_kernel<<<dim3(nb, nb), 256, 0, static_cast<cudaStream_t>(stream_)>>>(M, ld, buf, 0, nullptr);
  BCBF_LAUNCH_CHECK();
  return BCBF_OK;
}

extern "C" int bcbf_unpack_lower(const double* buf, int Npad, double* M, int ld, void* stream_) {
  BCBF_REQUIRE(M && buf && Npad > 0 && Npad % kBlk == 0 && ld >= Npad && ld % 2 == 0, "bcbf_unpack_lower: Npad=%d ld=%d", Npad, ld);
  const int nb = Npad / kBlk;
  pack_lower_kernel<<<dim3(nb, nb), 256, 0, static_cast<cudaStream_t>(stream_)>>>(nullptr, ld, const_cast<double*>(buf), 1, M);
  BCBF_LAUNCH_CHECK();
  return BCBF_OK;
}

extern "C" int bcbf_cbc1_terms(const double* Mk, const double* Bk, const double* Amat, const double* grad_h,
                               const double* h, const double* Fbar, double gamma, int n, int p, int Q, double* bfe,
                               double* e, double* Asq, double* A_socp, double* bfb, int* status, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BCBF_REQUIRE(Mk && Bk && Amat && grad_h && h && bfe && e, "bcbf_cbc1_terms: null pointer");
  BCBF_REQUIRE(n >= 1 && n <= BCBF_MAX_N_DIM && p >= 2 && p <= BCBF_MAX_P_DIM && Q >= 1, "bcbf_cbc1_terms: n=%d p=%d Q=%d",
               n, p, Q);
  cbc1_terms_kernel<<<ceil_div(Q, 128), 128, 0, stream>>>(Mk, Bk, Amat, grad_h, h, Fbar, gamma, n, p, Q, bfe, e, Asq,
                                                          A_socp, bfb, status);
  BCBF_LAUNCH_CHECK();
  return BCBF_OK;
}

extern "C" int bcbf_socp_factor(const double* Asq, int p, int Q, double reg, double* A_socp, double* bfb, int* status,
                                void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BCBF_REQUIRE(Asq && A_socp && bfb, "bcbf_socp_factor: null pointer");
  BCBF_REQUIRE(p >= 2 && p <= BCBF_MAX_P_DIM && Q >= 1, "bcbf_socp_factor: p=%d Q=%d", p, Q);
  socp_factor_kernel<<<ceil_div(Q, 128), 128, 0, stream>>>(Asq, p, Q, reg, A_socp, bfb, status);
  BCBF_LAUNCH_CHECK();
  return BCBF_OK;
}

// ======================================================================================================
// Model handle
// ======================================================================================================
struct bcbf_model {
  int device = 0;
  cudaStream_t stream = nullptr;
  bcbf_hyper hyp{};
  int N = 0, Npad = 0;
  bool fitted = false;
  // fitted state (device)
  double *X = nullptr, *UH = nullptr, *L = nullptr, *Linv = nullptr, *dinv = nullptr, *alpha = nullptr, *G = nullptr,
         *W = nullptr, *Y = nullptr, *hyp_dev = nullptr;
  int* info = nullptr;
  size_t cap_N = 0;  // capacity (Npad) of the factor-sized buffers
  // query scratch (device), grown on demand
  double *Xq = nullptr, *Uq = nullptr, *UHq = nullptr, *Kstar = nullptr, *Mk = nullptr, *Bk = nullptr, *mean = nullptr,
         *svar = nullptr;
  size_t cap_Q = 0, cap_K = 0;
  double fit_ms[5] = {0, 0, 0, 0, 0};
  // int8 tensor-core covariance path (csrc/ozaki.cu): digits of L^-1, built by fit or lazily by the first query
  int var_path = 0;
  void* oz_digits = nullptr;
  double* oz_rowscale = nullptr;
  size_t cap_oz = 0;
  bool oz_ready = false;
  int oz_sd = 7;            // digits per operand (7 default, 6 opt-in)
  double oz_split_ms = 0.0;
};

namespace {
constexpr int kLdY = 4;  // alpha / Y leading dimension (>= n, even) when n <= 4; generalised below

int ld_y(int n) { return (n + 1) / 2 * 2; }

template <class T>
int dev_alloc(T** p, size_t count) {
  if (*p) { BCBF_CUDA(cudaFree(*p)); *p = nullptr; }
  if (count == 0) return BCBF_OK;
  BCBF_CUDA(cudaMalloc(reinterpret_cast<void**>(p), sizeof(T) * count));
  return BCBF_OK;
}

int query_batch(int p, int var_path) {
  // queries per device batch: 4 waves of one CTA per SM (148 SMs) with the posterior kernel's TQ; the int8 path
  // tiles 64 frakB columns (64 / p queries) per CTA and balances best with a multiple of 148 column tiles
  if (var_path == 1) return 148 * (64 / p) * 6;
  const int tq = (p == 1) ? 96 : (p == 2 ? 64 : 32);
  return 148 * tq * 4;
}
}  // namespace

// digits of L^-1 for the int8 path (allocates on first use; `timed` records the device time in oz_split_ms)
static int ensure_oz_digits(bcbf_model* m, cudaStream_t s, bool timed) {
  if (m->oz_ready) return BCBF_OK;
  BCBF_REQUIRE(m->Npad <= bcbf_oz_max_npad(), "int8 covariance path supports Npad <= %d (got %d): use var_path 0",
               bcbf_oz_max_npad(), m->Npad);
  const size_t bytes = (size_t)bcbf_oz_factor_bytes_d(m->Npad, m->oz_sd);
  if (bytes > m->cap_oz) {
    if (m->oz_digits) { BCBF_CUDA(cudaFree(m->oz_digits)); m->oz_digits = nullptr; }
    if (m->oz_rowscale) { BCBF_CUDA(cudaFree(m->oz_rowscale)); m->oz_rowscale = nullptr; }
    m->cap_oz = 0;
    BCBF_CUDA(cudaMalloc(&m->oz_digits, bytes));
    BCBF_CUDA(cudaMalloc(reinterpret_cast<void**>(&m->oz_rowscale), sizeof(double) * (size_t)m->Npad));
    m->cap_oz = bytes;
  }
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (timed) {
    BCBF_CUDA(cudaEventCreate(&e0));
    BCBF_CUDA(cudaEventCreate(&e1));
    BCBF_CUDA(cudaEventRecord(e0, s));
  }
  int rc = bcbf_oz_split_factor_d(m->Linv, m->Npad, m->Npad, m->oz_digits, m->oz_rowscale, m->oz_sd, s);
  if (rc) {
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    return rc;
  }
  if (timed) {
    BCBF_CUDA(cudaEventRecord(e1, s));
    BCBF_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    BCBF_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    m->oz_split_ms = ms;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
  }
  m->oz_ready = true;
  return BCBF_OK;
}

extern "C" int bcbf_model_set_var_path(bcbf_model* m, int path) {
  BCBF_REQUIRE(m && (path == 0 || path == 1), "bcbf_model_set_var_path: path must be 0 (DMMA) or 1 (int8)");
  m->var_path = path;
  return BCBF_OK;
}
extern "C" int bcbf_model_get_var_path(bcbf_model* m) { return m ? m->var_path : -1; }
extern "C" int bcbf_model_set_oz_digits(bcbf_model* m, int digits) {
  BCBF_REQUIRE(m && (digits == 6 || digits == 7), "bcbf_model_set_oz_digits: 6 or 7");
  if (digits != m->oz_sd) {
    m->oz_sd = digits;
    m->oz_ready = false;      // re-split L^-1 at the next query
  }
  return BCBF_OK;
}
extern "C" double bcbf_model_oz_split_ms(bcbf_model* m) { return m ? m->oz_split_ms : 0.0; }

extern "C" int bcbf_model_create(bcbf_model** out, int device) {
  BCBF_REQUIRE(out, "bcbf_model_create: null out");
  BCBF_CUDA(cudaSetDevice(device));
  bcbf_model* m = new (std::nothrow) bcbf_model();
  BCBF_REQUIRE(m, "bcbf_model_create: out of host memory");
  m->device = device;
  cudaError_t e = cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { delete m; return cuda_fail(e, "cudaStreamCreate", __FILE__, __LINE__); }
  *out = m;
  return BCBF_OK;
}

extern "C" void bcbf_model_destroy(bcbf_model* m) {
  if (!m) return;
  cudaSetDevice(m->device);
  double** bufs[] = {&m->X, &m->UH, &m->L, &m->Linv, &m->dinv, &m->alpha, &m->G, &m->W, &m->Y, &m->hyp_dev,
                     &m->Xq, &m->Uq, &m->UHq, &m->Kstar, &m->Mk, &m->Bk, &m->mean, &m->svar};
  for (double** b : bufs)
    if (*b) cudaFree(*b);
  if (m->info) cudaFree(m->info);
  if (m->oz_digits) cudaFree(m->oz_digits);
  if (m->oz_rowscale) cudaFree(m->oz_rowscale);
  if (m->stream) cudaStreamDestroy(m->stream);
  delete m;
}

extern "C" int bcbf_model_alloc_state(bcbf_model* m, const bcbf_hyper* hyp, int N) {
  BCBF_REQUIRE(m && hyp, "bcbf_model_alloc_state: null pointer");
  BCBF_REQUIRE(N >= 1 && hyp->n >= 1 && hyp->n <= BCBF_MAX_N_DIM && hyp->p >= 1 && hyp->p <= BCBF_MAX_P_DIM,
               "bcbf_model_alloc_state: N=%d n=%d p=%d", N, hyp->n, hyp->p);
  BCBF_CUDA(cudaSetDevice(m->device));
  m->hyp = *hyp;
  const int Npad = bcbf_padded(N), n = hyp->n, p = hyp->p;
  if ((size_t)Npad > m->cap_N) {
    int rc;
    if ((rc = dev_alloc(&m->L, (size_t)Npad * Npad))) return rc;
    if ((rc = dev_alloc(&m->Linv, (size_t)Npad * Npad))) return rc;
    if ((rc = dev_alloc(&m->dinv, (size_t)bcbf_dinv_elems(Npad)))) return rc;
    if ((rc = dev_alloc(&m->X, (size_t)Npad * BCBF_MAX_N_DIM))) return rc;
    if ((rc = dev_alloc(&m->UH, (size_t)Npad * BCBF_MAX_P_DIM))) return rc;
    if ((rc = dev_alloc(&m->G, (size_t)Npad * BCBF_MAX_P_DIM))) return rc;
    if ((rc = dev_alloc(&m->alpha, (size_t)Npad * BCBF_MAX_N_DIM))) return rc;
    // Y followed by the scratch of bcbf_alpha_refine (two work copies, split partials of the transposed skinny
    // products, (hi, lo) partials of the compensated residual)
    if ((rc = dev_alloc(&m->Y, (size_t)Npad * BCBF_MAX_N_DIM +
                                   (size_t)bcbf_alpha_refine_scratch_elems(Npad, Npad, BCBF_MAX_N_DIM))))
      return rc;
    if ((rc = dev_alloc(&m->W, (size_t)Npad * BCBF_MAX_N_DIM * BCBF_MAX_P_DIM))) return rc;
    if ((rc = dev_alloc(&m->hyp_dev, (size_t)256))) return rc;
    if (!m->info) BCBF_CUDA(cudaMalloc(&m->info, sizeof(int)));
    m->cap_N = Npad;
  }
  m->N = N;
  m->Npad = Npad;
  if (Npad >= 4096 && Npad <= bcbf_oz_max_npad()) {  // bcbf_trtri's int8 levels: workspaces sized outside the timed fit
    int hmax = 2048;
    while (2 * hmax < Npad) hmax *= 2;
    int rc = bcbf_oz_gemm_reserve(hmax, hmax, hmax);
    if (rc) return rc;
    if ((rc = bcbf_oz_update_reserve(Npad, Npad, 512))) return rc;  // bcbf_potrf's trailing updates
  }
  // hyper-parameter block on the device: [lengthscale(8) | B(16) | C(32) | Ct(32) | A(64)]
  double hbuf[256];
  memset(hbuf, 0, sizeof(hbuf));
  for (int d = 0; d < n; ++d) hbuf[d] = hyp->lengthscale[d];
  for (int i = 0; i < p * p; ++i) hbuf[8 + i] = hyp->B[i];
  for (int i = 0; i < p * n; ++i) hbuf[24 + i] = hyp->C[i];
  for (int r = 0; r < n; ++r)
    for (int j = 0; j < p; ++j) hbuf[56 + r * p + j] = hyp->C[j * n + r];
  for (int i = 0; i < n * n; ++i) hbuf[88 + i] = hyp->A[i];
  BCBF_CUDA(cudaMemcpyAsync(m->hyp_dev, hbuf, sizeof(hbuf), cudaMemcpyHostToDevice, m->stream));
  BCBF_CUDA(cudaStreamSynchronize(m->stream));
  return BCBF_OK;
}

// y = beta * y0 + alpha * op(T) x  (y0 may be null); `partial` holds ceil(Npad / mv_rows(Npad)) * Npad * nc doubles
static int tri_mv(const double* T, int ld, int Npad, int trans, const double* x, int ldx, int nc, double alpha, double beta,
                  const double* y0, double* y, double* partial, cudaStream_t s) {
  if (!trans) {
    tri_mv_n_kernel<<<ceil_div(Npad, 8), 256, 0, s>>>(T, ld, Npad, x, ldx, nc, alpha, beta, y0, y);
    BCBF_LAUNCH_CHECK();
    return BCBF_OK;
  }
  const int rows = mv_rows(Npad), nsplit = ceil_div(Npad, rows);
  tri_mv_t_kernel<<<dim3(Npad / 128, nsplit), 128, 0, s>>>(T, ld, Npad, x, ldx, nc, partial, rows);
  BCBF_LAUNCH_CHECK();
  tri_mv_t_finalize_kernel<<<ceil_div((long long)Npad * nc, 256), 256, 0, s>>>(partial, Npad, nc, nsplit, ldx, alpha, beta,
                                                                               y0, y, rows);
  BCBF_LAUNCH_CHECK();
  return BCBF_OK;
}

extern "C" long long bcbf_alpha_refine_scratch_elems(int N, int Npad, int ldy) {
  return (long long)Npad * ldy * (2 + ceil_div(Npad, mv_rows(Npad))) + bcbf_gram_resid_scratch_elems(N);
}

// alpha = Kb^-1 Y by iterative refinement (include/bcbf.h): start alpha0 = Linv^T (Linv Y), then `iters` times
//   r = Y - (Kb + jscale diag(jitter)) alpha   [bcbf_gram_resid: Kb re-evaluated bit-identically, Dot2 accumulation]
//   alpha += Linv^T (Linv r)
// The explicit inverse carries a forward error ~ eps cond(L) (1e-8 relative at the bench shapes), so every step gains
// digits in the posterior mean (the explicit inverse of a cond ~1e11 matrix is that good and no better): measured at
// N = 16384, mean error 7e-7 -> ~2e-8 -> 2.5e-10 -> below 1e-12, so THREE steps are run (kRefineIters); the third is what
// makes the result independent of last-bit differences in the factor.
extern "C" int bcbf_alpha_refine_ws(const double* X, const double* UH, const double* Bmat, const double* lengthscale,
                                    double outputscale, int N, int n, int p, const double* jitter, double jitter_scale,
                                    const double* Linv, int ld, int Npad, const double* Y, int ldy, int nc, int iters,
                                    double* alpha, double* scratch, long long scratch_elems, double* kb_ws, int ldk,
                                    void* stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  BCBF_REQUIRE(X && UH && Bmat && lengthscale && Linv && Y && alpha && scratch, "bcbf_alpha_refine: null pointer");
  BCBF_REQUIRE(Npad > 0 && Npad % kBlk == 0 && N >= 1 && N <= Npad && ld >= Npad && nc >= 1 && nc <= ldy &&
                   ldy <= kMvMaxC && iters >= 0,
               "bcbf_alpha_refine: N=%d Npad=%d ld=%d nc=%d ldy=%d (<= %d) iters=%d", N, Npad, ld, nc, ldy, kMvMaxC, iters);
  BCBF_REQUIRE(scratch_elems >= bcbf_alpha_refine_scratch_elems(N, Npad, ldy), "bcbf_alpha_refine: scratch too small");
  BCBF_REQUIRE(!kb_ws || ldk >= Npad, "bcbf_alpha_refine_ws: ldk=%d < Npad=%d", ldk, Npad);
  double* t = scratch;
  double* r = t + (size_t)Npad * ldy;
  double* part = r + (size_t)Npad * ldy;
  double* rs = part + (size_t)ceil_div(Npad, mv_rows(Npad)) * Npad * ldy;
  const long long rs_elems = bcbf_gram_resid_scratch_elems(N);
  int rc;
  if ((rc = tri_mv(Linv, ld, Npad, 0, Y, ldy, ldy, 1.0, 0.0, nullptr, t, part, s))) return rc;
  if ((rc = tri_mv(Linv, ld, Npad, 1, t, ldy, ldy, 1.0, 0.0, nullptr, alpha, part, s))) return rc;
  // With a workspace the factorised matrix is written out once more (the same kernel as for the factorisation: the same
  // bits) and every residual streams it from HBM; without one every residual re-evaluates its N^2 entries.
  if (kb_ws && iters > 0 &&
      (rc = bcbf_gram_train_lower(X, UH, Bmat, lengthscale, outputscale, N, n, p, kb_ws, ldk, Npad, s)))
    return rc;
  for (int it = 0; it < iters; ++it) {
    BCBF_CUDA(cudaMemsetAsync(r, 0, sizeof(double) * (size_t)Npad * ldy, s));
    if (kb_ws)
      rc = bcbf_gram_resid_stored(kb_ws, ldk, N, jitter, jitter_scale, alpha, ldy, Y, ldy, nc, r, ldy, rs, rs_elems, s);
    else
      rc = bcbf_gram_resid(X, UH, Bmat, lengthscale, outputscale, N, n, p, jitter, jitter_scale, alpha, ldy, Y, ldy, nc,
                           r, ldy, rs, rs_elems, s);
    if (rc) return rc;
    if ((rc = tri_mv(Linv, ld, Npad, 0, r, ldy, ldy, 1.0, 0.0, nullptr, t, part, s))) return rc;
    if ((rc = tri_mv(Linv, ld, Npad, 1, t, ldy, ldy, 1.0, 1.0, alpha, alpha, part, s))) return rc;
  }
  return BCBF_OK;
}

extern "C" int bcbf_alpha_refine(const double* X, const double* UH, const double* Bmat, const double* lengthscale,
                                 double outputscale, int N, int n, int p, const double* jitter, double jitter_scale,
                                 const double* Linv, int ld, int Npad, const double* Y, int ldy, int nc, int iters,
                                 double* alpha, double* scratch, long long scratch_elems, void* stream_) {
  return bcbf_alpha_refine_ws(X, UH, Bmat, lengthscale, outputscale, N, n, p, jitter, jitter_scale, Linv, ld, Npad, Y, ldy,
                              nc, iters, alpha, scratch, scratch_elems, nullptr, 0, stream_);
}

constexpr int kRefineIters = 3;

static int finish_fit_from_factor(bcbf_model* m, const double* djit, double jitter_scale) {
  // alpha = Kb^-1 Y refined against the factorised matrix itself in compensated arithmetic (bcbf_alpha_refine); the
  // reference's cholesky_solve (control_affine_model.py:545) is the backward-stable version of the same solve.
  // W = alpha (.) G.
  const int n = m->hyp.n, p = m->hyp.p, Npad = m->Npad, ldy = ld_y(n);
  cudaStream_t s = m->stream;
  // the query buffer K* (>= Npad^2 doubles after a fit, idle until the first query) holds the copy of Kb the residuals read
  double* kb_ws = (m->Kstar && m->cap_K >= (size_t)Npad * Npad && Npad >= 1024) ? m->Kstar : nullptr;
  int rc = bcbf_alpha_refine_ws(m->X, m->UH, m->hyp.B, m->hyp.lengthscale, m->hyp.outputscale, m->N, n, p, djit,
                                jitter_scale, m->Linv, Npad, Npad, m->Y, ldy, n, kRefineIters, m->alpha,
                                m->Y + (size_t)Npad * ldy, bcbf_alpha_refine_scratch_elems(m->N, Npad, ldy), kb_ws, Npad, s);
  if (rc) return rc;
  build_w_kernel<<<ceil_div(Npad, 128), 128, 0, s>>>(m->alpha, ldy, m->G, Npad, n, p, m->W);
  BCBF_LAUNCH_CHECK();
  return BCBF_OK;
}

namespace {
// cudaEvents of one fit: destroyed on every exit path (the NOT_PD retry path returns early)
struct FitEvents {
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int create() {
    for (auto& e : ev) BCBF_CUDA(cudaEventCreate(&e));
    return BCBF_OK;
  }
  ~FitEvents() {
    for (auto& e : ev)
      if (e) cudaEventDestroy(e);
  }
};
}  // namespace

extern "C" int bcbf_model_fit(bcbf_model* m, const bcbf_hyper* hyp, const double* X, const double* U, const double* Xdot,
                              int N, const double* jitter, double jitter_scale) {
  BCBF_REQUIRE(m && hyp && X && U && Xdot, "bcbf_model_fit: null pointer");
  m->fitted = false;
  m->oz_ready = false;
  int rc = bcbf_model_alloc_state(m, hyp, N);
  if (rc) return rc;
  const int n = hyp->n, p = hyp->p, mm = p - 1, Npad = m->Npad, ldy = ld_y(n);
  cudaStream_t s = m->stream;
  // scratch for trtri (borrowed from the query K* buffer): allocated before the timed stages
  if (m->cap_K < (size_t)Npad * Npad) {
    if ((rc = dev_alloc(&m->Kstar, (size_t)Npad * Npad))) return rc;
    m->cap_K = (size_t)Npad * Npad;
  }
  FitEvents fe;
  if ((rc = fe.create())) return rc;
  cudaEvent_t* ev = fe.ev;
  // stage inputs: U and Xdot go through the (not yet used) Linv buffer, jitter through dinv
  double* dU = m->Linv;
  double* dXdot = m->Linv + (size_t)N * BCBF_MAX_P_DIM;
  double* djit = nullptr;
  BCBF_CUDA(cudaMemcpyAsync(m->X, X, sizeof(double) * (size_t)N * n, cudaMemcpyHostToDevice, s));
  if (mm > 0) BCBF_CUDA(cudaMemcpyAsync(dU, U, sizeof(double) * (size_t)N * mm, cudaMemcpyHostToDevice, s));
  BCBF_CUDA(cudaMemcpyAsync(dXdot, Xdot, sizeof(double) * (size_t)N * n, cudaMemcpyHostToDevice, s));
  if (jitter) {
    djit = m->W;  // W is rebuilt at the end of fit; N doubles of it carry the jitter until the refinement is done
    BCBF_CUDA(cudaMemcpyAsync(djit, jitter, sizeof(double) * (size_t)N, cudaMemcpyHostToDevice, s));
  }
  BCBF_CUDA(cudaEventRecord(ev[0], s));
  prep_train_kernel<<<ceil_div(Npad, 128), 128, 0, s>>>(dU, dXdot, N, Npad, n, p, m->hyp_dev + 8, m->hyp_dev + 24,
                                                        m->UH, m->G, m->Y, ldy);
  BCBF_LAUNCH_CHECK();
  rc = bcbf_gram_train_lower(m->X, m->UH, hyp->B, hyp->lengthscale, hyp->outputscale, N, n, p, m->L, Npad, Npad, s);
  if (rc) return rc;
  BCBF_CUDA(cudaEventRecord(ev[1], s));
  rc = bcbf_potrf(m->L, Npad, Npad, N, djit, jitter_scale, m->dinv, m->info, s);
  if (rc) return rc;
  BCBF_CUDA(cudaEventRecord(ev[2], s));
  rc = bcbf_check_info(m->info, s);
  if (rc) return rc;
  rc = bcbf_trtri(m->L, m->dinv, m->Linv, m->Kstar, Npad, Npad, s);
  if (rc) return rc;
  BCBF_CUDA(cudaEventRecord(ev[3], s));
  rc = finish_fit_from_factor(m, djit, jitter_scale);
  if (rc) return rc;
  m->oz_split_ms = 0.0;
  if (m->var_path == 1 && Npad <= bcbf_oz_max_npad() && (rc = ensure_oz_digits(m, s, true))) return rc;
  BCBF_CUDA(cudaEventRecord(ev[4], s));
  BCBF_CUDA(cudaStreamSynchronize(s));
  float ms;
  for (int i = 0; i < 4; ++i) {
    BCBF_CUDA(cudaEventElapsedTime(&ms, ev[i], ev[i + 1]));
    m->fit_ms[i] = ms;
  }
  BCBF_CUDA(cudaEventElapsedTime(&ms, ev[0], ev[4]));
  m->fit_ms[4] = ms;
  m->fitted = true;
  return BCBF_OK;
}

extern "C" int bcbf_model_fit_timing(bcbf_model* m, double out_ms[5]) {
  BCBF_REQUIRE(m && out_ms, "bcbf_model_fit_timing: null pointer");
  for (int i = 0; i < 5; ++i) out_ms[i] = m->fit_ms[i];
  return BCBF_OK;
}

extern "C" int bcbf_model_state(bcbf_model* m, int* N, int* Npad, double** L, double** Linv, double** alpha,
                                double** G, double** W, double** Xtrain) {
  BCBF_REQUIRE(m, "bcbf_model_state: null model");
  if (N) *N = m->N;
  if (Npad) *Npad = m->Npad;
  if (L) *L = m->L;
  if (Linv) *Linv = m->Linv;
  if (alpha) *alpha = m->alpha;
  if (G) *G = m->G;
  if (W) *W = m->W;
  if (Xtrain) *Xtrain = m->X;
  return BCBF_OK;
}

extern "C" int bcbf_model_adopt(bcbf_model* m) {
  // a rank that filled the buffers of bcbf_model_state from a broadcast declares them valid; its digits of L^-1 are
  // rebuilt by the next query
  BCBF_REQUIRE(m, "bcbf_model_adopt: null model");
  BCBF_REQUIRE(m->Npad > 0 && m->Linv && m->alpha && m->G && m->W && m->X,
               "bcbf_model_adopt: no state buffers (call bcbf_model_alloc_state first)");
  m->fitted = true;
  m->oz_ready = false;
  return BCBF_OK;
}

static int ensure_query_capacity(bcbf_model* m, int Qb) {
  const int Npad = m->Npad;
  const int ldks = ((Qb + 191) / 192) * 192;
  int rc;
  if ((size_t)Qb > m->cap_Q) {
    if ((rc = dev_alloc(&m->Xq, (size_t)Qb * BCBF_MAX_N_DIM))) return rc;
    if ((rc = dev_alloc(&m->Uq, (size_t)Qb * BCBF_MAX_P_DIM))) return rc;
    if ((rc = dev_alloc(&m->UHq, (size_t)Qb * BCBF_MAX_P_DIM))) return rc;
    if ((rc = dev_alloc(&m->Mk, (size_t)Qb * BCBF_MAX_N_DIM * BCBF_MAX_P_DIM))) return rc;
    if ((rc = dev_alloc(&m->Bk, (size_t)Qb * BCBF_MAX_P_DIM * BCBF_MAX_P_DIM))) return rc;
    if ((rc = dev_alloc(&m->mean, (size_t)Qb * BCBF_MAX_N_DIM))) return rc;
    if ((rc = dev_alloc(&m->svar, (size_t)Qb))) return rc;
    m->cap_Q = Qb;
  }
  if ((size_t)Npad * ldks > m->cap_K) {
    if ((rc = dev_alloc(&m->Kstar, (size_t)Npad * ldks))) return rc;
    m->cap_K = (size_t)Npad * ldks;
  }
  return BCBF_OK;
}

// One device batch: Xq/Uq device pointers (Qb queries) -> device outputs (any may be null)
static int query_batch_device(bcbf_model* m, const double* dXq, const double* dUq, int Qb, double* dmean, double* dsvar,
                              double* dMk, double* dBk, cudaStream_t s) {
  const int n = m->hyp.n, p = m->hyp.p, Npad = m->Npad, N = m->N;
  const int ldks = ((Qb + 191) / 192) * 192;
  int rc = bcbf_cross_gram(m->X, dXq, m->hyp.lengthscale, m->hyp.outputscale, N, Qb, n, m->Kstar, ldks, Npad, s);
  if (rc) return rc;
  double* Mk = dMk ? dMk : m->Mk;
  double* Bk = dBk ? dBk : m->Bk;
  const bool need_var = dBk || dsvar;
  const bool need_mean = dMk || dmean;
  // factors beyond the exact-int32-accumulation limit of the int8 kernel stay on the FP64 pipe
  const bool i8 = m->var_path == 1 && Npad <= bcbf_oz_max_npad();
  if (i8) {
    if (need_var && (rc = ensure_oz_digits(m, s, false))) return rc;
    rc = need_var ? bcbf_posterior_blocks_i8_d(m->oz_digits, m->oz_rowscale, Npad, m->Kstar, ldks, m->G, m->W,
                                               m->hyp_dev + 8, m->hyp_dev + 56, m->hyp.outputscale, n, p, Qb,
                                               need_mean ? Mk : nullptr, Bk, m->oz_sd, s)
                  : bcbf_posterior_blocks(m->Linv, Npad, Npad, m->Kstar, ldks, m->G, m->W, m->hyp_dev + 8,
                                          m->hyp_dev + 56, m->hyp.outputscale, n, p, Qb, Mk, nullptr, s);
  } else {
    rc = bcbf_posterior_blocks(m->Linv, Npad, Npad, m->Kstar, ldks, m->G, m->W, m->hyp_dev + 8, m->hyp_dev + 56,
                               m->hyp.outputscale, n, p, Qb, need_mean ? Mk : nullptr, need_var ? Bk : nullptr, s);
  }
  if (rc) return rc;
  if (dmean || dsvar) {
    build_uh_kernel<<<ceil_div(Qb, 128), 128, 0, s>>>(dUq, Qb, p, m->UHq);
    BCBF_LAUNCH_CHECK();
    rc = bcbf_contract_u(Mk, Bk, m->UHq, n, p, Qb, dmean, dsvar, s);
    if (rc) return rc;
  }
  return BCBF_OK;
}

extern "C" int bcbf_model_query_device(bcbf_model* m, const double* Xq, const double* Uq, int Q, double* mean,
                                       double* svar, double* Mk, double* Bk, void* stream_) {
  BCBF_REQUIRE(m && Xq, "bcbf_model_query_device: null pointer");
  if (!m->fitted) { set_last_error("bcbf_model_query_device: model is not fitted"); return BCBF_ERR_NOT_FITTED; }
  BCBF_CUDA(cudaSetDevice(m->device));
  cudaStream_t s = static_cast<cudaStream_t>(stream_);  // exactly the caller's stream (NULL = the default stream)
  const int n = m->hyp.n, p = m->hyp.p, mm = p - 1;
  const int QB = query_batch(p, m->var_path);
  int rc = ensure_query_capacity(m, Q < QB ? Q : QB);
  if (rc) return rc;
  for (int q0 = 0; q0 < Q; q0 += QB) {
    const int Qb = (Q - q0) < QB ? (Q - q0) : QB;
    rc = query_batch_device(m, Xq + (size_t)q0 * n, Uq ? Uq + (size_t)q0 * mm : nullptr, Qb,
                            mean ? mean + (size_t)q0 * n : nullptr, svar ? svar + q0 : nullptr,
                            Mk ? Mk + (size_t)q0 * n * p : nullptr, Bk ? Bk + (size_t)q0 * p * p : nullptr, s);
    if (rc) return rc;
  }
  return BCBF_OK;
}

extern "C" int bcbf_model_query(bcbf_model* m, const double* Xq, const double* Uq, int Q, double* mean, double* svar,
                                double* Mk, double* Bk) {
  BCBF_REQUIRE(m && Xq && Q >= 1, "bcbf_model_query: null pointer / empty query");
  if (!m->fitted) { set_last_error("bcbf_model_query: model is not fitted"); return BCBF_ERR_NOT_FITTED; }
  BCBF_CUDA(cudaSetDevice(m->device));
  cudaStream_t s = m->stream;
  const int n = m->hyp.n, p = m->hyp.p, mm = p - 1;
  const int QB = query_batch(p, m->var_path);
  int rc = ensure_query_capacity(m, Q < QB ? Q : QB);
  if (rc) return rc;
  for (int q0 = 0; q0 < Q; q0 += QB) {
    const int Qb = (Q - q0) < QB ? (Q - q0) : QB;
    BCBF_CUDA(cudaMemcpyAsync(m->Xq, Xq + (size_t)q0 * n, sizeof(double) * (size_t)Qb * n, cudaMemcpyHostToDevice, s));
    if (Uq && mm > 0)
      BCBF_CUDA(cudaMemcpyAsync(m->Uq, Uq + (size_t)q0 * mm, sizeof(double) * (size_t)Qb * mm, cudaMemcpyHostToDevice, s));
    rc = query_batch_device(m, m->Xq, Uq ? m->Uq : nullptr, Qb, mean ? m->mean : nullptr, svar ? m->svar : nullptr,
                            Mk ? m->Mk : nullptr, Bk ? m->Bk : nullptr, s);
    if (rc) return rc;
    if (mean) BCBF_CUDA(cudaMemcpyAsync(mean + (size_t)q0 * n, m->mean, sizeof(double) * (size_t)Qb * n, cudaMemcpyDeviceToHost, s));
    if (svar) BCBF_CUDA(cudaMemcpyAsync(svar + q0, m->svar, sizeof(double) * (size_t)Qb, cudaMemcpyDeviceToHost, s));
    if (Mk) BCBF_CUDA(cudaMemcpyAsync(Mk + (size_t)q0 * n * p, m->Mk, sizeof(double) * (size_t)Qb * n * p, cudaMemcpyDeviceToHost, s));
    if (Bk) BCBF_CUDA(cudaMemcpyAsync(Bk + (size_t)q0 * p * p, m->Bk, sizeof(double) * (size_t)Qb * p * p, cudaMemcpyDeviceToHost, s));
  }
  BCBF_CUDA(cudaStreamSynchronize(s));
  return BCBF_OK;
}
