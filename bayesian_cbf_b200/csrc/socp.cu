// Batched tiny second-order-cone programs: the per-control-step safety program of the Bayesian CLF/CBF controller
// (reference ControllerCLFBayesian.control, unicycle_move_to_pose.py:926-964, solved there with cvxpy + GUROBI on the host;
// SURVEY 8f-1).  One thread per problem:
//
//     minimise   sum_i w_i (y_i - r_i)^2 + q^T y              y in R^nv           (nv <= 4: [relaxation, u]; w_i >= 0)
//     subject to c_k^T y + d_k >= rho || A_k y + b_k ||_2      k = 0 .. K-1        (K <= 4 cones of dimension pc <= 4)
//
// Method: log-barrier interior point (barrier -log(t^2 - |z|^2) per cone), damped Newton with backtracking, preceded by
// a phase-I problem in (y, s) that either finds a strictly feasible point or proves infeasibility — the "feasibility
// decision" that ends a rollout in the reference (`raise ValueError(problem.status)`, :954-964).  All arithmetic in
// float64, no data-dependent randomness: the CPU restatement in oracle/socp_oracle.py follows it step for step.
#include "../../include/bcbf.h"
#include "common.cuh"

namespace bcbf {

constexpr int kSV = 4, kSK = 4, kSP = 4;  // maxima: variables, cones, cone dimension

// Sizes are template parameters so that, for the instantiated shapes, every loop unrolls and the whole problem lives in
// registers (one thread per problem; with run-time sizes the arrays fall to local memory and a Newton step costs ~10x).
// NV / KC / PC = 0 selects the run-time-sized fallback (maxima kSV / kSK / kSP).
template <int NV, int KC, int PC>
struct SocpProblem {
  static constexpr int MV = NV ? NV : kSV, MK = KC ? KC : kSK, MP = PC ? PC : kSP, MN = MV + 1;
  int nv_, K_, pc_;
  double rho;
  double w[MV], r[MV], q[MV];
  double c[MK][MV], d[MK];
  double A[MK][MP][MV], b[MK][MP];
  __device__ __forceinline__ int nv() const { return NV ? NV : nv_; }
  __device__ __forceinline__ int K() const { return KC ? KC : K_; }
  __device__ __forceinline__ int pc() const { return PC ? PC : pc_; }
};

// cone values at y (+ slack s added to every t): returns false if some cone is not strictly inside
template <class PT>
__device__ __forceinline__ bool socp_cones(const PT& P, const double* y, double s, double* t, double z[][PT::MP],
                                           double* D) {
  bool ok = true;
  for (int k = 0; k < P.K(); ++k) {
    double tk = P.d[k] + s;
    for (int i = 0; i < P.nv(); ++i) tk = fma(P.c[k][i], y[i], tk);
    double zz = 0.0;
    for (int q = 0; q < P.pc(); ++q) {
      double v = P.b[k][q];
      for (int i = 0; i < P.nv(); ++i) v = fma(P.A[k][q][i], y[i], v);
      v *= P.rho;
      z[k][q] = v;
      zz = fma(v, v, zz);
    }
    t[k] = tk;
    D[k] = tk * tk - zz;
    ok = ok && (tk > 0.0) && (D[k] > 0.0);
  }
  return ok;
}

// in-place Cholesky solve of the n x n SPD system H x = g (n <= kSN); returns false on a non-positive pivot
template <int MN>
__device__ __forceinline__ bool socp_solve(double H[][MN], double* g, int n) {
  for (int j = 0; j < n; ++j) {
    double dj = H[j][j];
    for (int k = 0; k < j; ++k) dj -= H[j][k] * H[j][k];
    if (!(dj > 0.0)) return false;
    dj = sqrt(dj);
    H[j][j] = dj;
    for (int i = j + 1; i < n; ++i) {
      double v = H[i][j];
      for (int k = 0; k < j; ++k) v -= H[i][k] * H[j][k];
      H[i][j] = v / dj;
    }
  }
  for (int i = 0; i < n; ++i) {
    double v = g[i];
    for (int k = 0; k < i; ++k) v -= H[i][k] * g[k];
    g[i] = v / H[i][i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double v = g[i];
    for (int k = i + 1; k < n; ++k) v -= H[k][i] * g[k];
    g[i] = v / H[i][i];
  }
  return true;
}

// Barrier value, gradient and Hessian of  tau * f(x) - sum_k log D_k  in the variables x = (y [, s]).
//   phase 1: f = s + eps1 * (sum w (y-r)^2 + q^T y) ;  phase 2: f = sum w (y-r)^2 + q^T y
template <class PT>
__device__ __forceinline__ double socp_merit(const PT& P, const double* x, bool phase1, double tau, double eps1) {
  double t[PT::MK], z[PT::MK][PT::MP], D[PT::MK];
  if (!socp_cones(P, x, phase1 ? x[P.nv()] : 0.0, t, z, D)) return __longlong_as_double(0x7ff0000000000000LL);
  double f = 0.0;
  for (int i = 0; i < P.nv(); ++i) f = fma(P.w[i] * (x[i] - P.r[i]), (x[i] - P.r[i]), f);
  for (int i = 0; i < P.nv(); ++i) f = fma(P.q[i], x[i], f);
  double val = phase1 ? tau * (x[P.nv()] + eps1 * f) : tau * f;
  for (int k = 0; k < P.K(); ++k) val -= log(D[k]);
  return val;
}

template <class PT>
__device__ __forceinline__ void socp_grad_hess(const PT& P, const double* x, bool phase1, double tau, double eps1,
                                               double* g, double H[][PT::MN]) {
  const int nv = P.nv(), n = nv + (phase1 ? 1 : 0);
  double t[PT::MK], z[PT::MK][PT::MP], D[PT::MK];
  socp_cones(P, x, phase1 ? x[nv] : 0.0, t, z, D);
  for (int i = 0; i < n; ++i) {
    g[i] = 0.0;
    for (int j = 0; j < n; ++j) H[i][j] = 0.0;
  }
  const double fs = phase1 ? tau * eps1 : tau;
  for (int i = 0; i < nv; ++i) {
    g[i] = fs * (2.0 * P.w[i] * (x[i] - P.r[i]) + P.q[i]);
    H[i][i] = 2.0 * fs * P.w[i];
  }
  if (phase1) g[nv] = tau;
  for (int k = 0; k < P.K(); ++k) {
    // u = (t, z),  q = t * dt/dx - sum_q z_q dz_q/dx ;  grad -= 2 q / D ;  Hess += 4 q q^T / D^2 - 2 (dt dt^T - dz^T dz) / D
    double dt[PT::MN], q[PT::MN];
    for (int i = 0; i < nv; ++i) dt[i] = P.c[k][i];
    if (phase1) dt[nv] = 1.0;
    for (int i = 0; i < n; ++i) {
      double v = t[k] * dt[i];
      if (i < nv)
        for (int e = 0; e < P.pc(); ++e) v -= z[k][e] * P.rho * P.A[k][e][i];
      q[i] = v;
    }
    const double iD = 1.0 / D[k];
    for (int i = 0; i < n; ++i) {
      g[i] -= 2.0 * q[i] * iD;
      for (int j = 0; j <= i; ++j) {
        double zz = 0.0;
        if (i < nv && j < nv)
          for (int e = 0; e < P.pc(); ++e) zz = fma(P.A[k][e][i], P.A[k][e][j], zz);
        H[i][j] += 4.0 * q[i] * q[j] * iD * iD - 2.0 * (dt[i] * dt[j] - P.rho * P.rho * zz) * iD;
      }
    }
  }
}

// One centering problem: damped Newton.  Returns the number of Newton steps taken; stops early in phase 1 as soon as the
// slack is negative (a strictly feasible y has been found).
template <class PT>
__device__ __forceinline__ int socp_center(const PT& P, double* x, bool phase1, double tau, double eps1, int max_newton,
                                           double center_tol) {
  const int n = P.nv() + (phase1 ? 1 : 0);
  int it = 0;
  for (; it < max_newton; ++it) {
    if (phase1 && x[P.nv()] < 0.0) break;
    double g[PT::MN], H[PT::MN][PT::MN], dx[PT::MN];
    socp_grad_hess(P, x, phase1, tau, eps1, g, H);
    for (int i = 0; i < n; ++i) dx[i] = -g[i];
    // tiny ridge keeps the factorisation safe when a direction is (numerically) unconstrained
    for (int i = 0; i < n; ++i) H[i][i] += 1e-14 * (1.0 + fabs(H[i][i]));
    if (!socp_solve<PT::MN>(H, dx, n)) break;
    double dec = 0.0;  // Newton decrement squared = -g^T dx
    for (int i = 0; i < n; ++i) dec -= g[i] * dx[i];
    if (!(dec > 1e-22)) break;   // also stops on NaN
    const double f0 = socp_merit(P, x, phase1, tau, eps1);
    double step = 1.0;
    double xn[PT::MN];
    bool moved = false;
    for (int ls = 0; ls < 60; ++ls) {
      for (int i = 0; i < n; ++i) xn[i] = fma(step, dx[i], x[i]);
      const double f1 = socp_merit(P, xn, phase1, tau, eps1);
      if (f1 <= f0 - 0.25 * step * dec) { moved = true; break; }
      step *= 0.5;
    }
    if (!moved) break;
    for (int i = 0; i < n; ++i) x[i] = xn[i];
    if (dec * 0.5 < center_tol) { ++it; break; }
  }
  return it;
}

template <int NV, int KC, int PC>
__global__ void __launch_bounds__(32) socp_solve_kernel(int Q, int nv, int K, int pc, double rho,
                                                        const double* __restrict__ w, int w_stride,
                                                        const double* __restrict__ r, const double* __restrict__ qlin,
                                                        const double* __restrict__ c,
                                                        const double* __restrict__ d, const double* __restrict__ A,
                                                        const double* __restrict__ b, double tol,
                                                        double* __restrict__ y_out, int* __restrict__ status,
                                                        int* __restrict__ iters) {
  using PT = SocpProblem<NV, KC, PC>;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= Q) return;
  PT P;
  P.nv_ = nv; P.K_ = K; P.pc_ = pc; P.rho = rho;
  for (int i = 0; i < P.nv(); ++i) {
    P.w[i] = w[(long long)p * w_stride + i];
    P.r[i] = r ? r[(long long)p * nv + i] : 0.0;
    P.q[i] = qlin ? qlin[(long long)p * nv + i] : 0.0;
  }
  for (int k = 0; k < P.K(); ++k) {
    P.d[k] = d[(long long)p * K + k];
    for (int i = 0; i < P.nv(); ++i) P.c[k][i] = c[((long long)p * K + k) * nv + i];
    for (int e = 0; e < P.pc(); ++e) {
      P.b[k][e] = b[((long long)p * K + k) * pc + e];
      for (int i = 0; i < P.nv(); ++i) P.A[k][e][i] = A[(((long long)p * K + k) * pc + e) * nv + i];
    }
  }
  double x[PT::MN];
  for (int i = 0; i < P.nv(); ++i) x[i] = P.r[i];
  int total = 0, st = 0;
  constexpr double kMu = 10.0;          // barrier parameter growth per outer iteration
  constexpr double kCenterTol = 1e-5;   // Newton decrement^2 / 2 at which an intermediate centering stops
  // ---- phase I: is y = r strictly feasible?  otherwise minimise the slack ----------------------------------------
  {
    double t[PT::MK], z[PT::MK][PT::MP], D[PT::MK];
    if (!socp_cones(P, x, 0.0, t, z, D)) {
      double s0 = 0.0, scale = 1.0;
      for (int k = 0; k < P.K(); ++k) {
        double zz = 0.0;
        for (int e = 0; e < P.pc(); ++e) zz = fma(z[k][e], z[k][e], zz);
        const double need = sqrt(zz) - (t[k]);   // t_k + s > |z_k|
        s0 = fmax(s0, need);
        scale = fmax(scale, fmax(fabs(t[k]), sqrt(zz)));
      }
      x[P.nv()] = s0 + 0.1 * scale + 1e-3;
      const double eps1 = 1e-6;
      double tau = 1.0 / scale;
      bool found = false;
      for (int outer = 0; outer < 60; ++outer) {
        total += socp_center(P, x, true, tau, eps1, 40, kCenterTol);
        if (x[P.nv()] < 0.0) { found = true; break; }
        if (2.0 * P.K() / tau < tol * scale) break;   // gap closed with s >= 0: no strictly feasible point
        tau *= kMu;
      }
      if (!found) st = 1;
    }
  }
  // ---- phase II ------------------------------------------------------------------------------------------------------
  if (st == 0) {
    double fscale = 1.0;
    for (int i = 0; i < P.nv(); ++i) fscale = fmax(fscale, fmax(P.w[i], fabs(P.q[i])));
    double tau = 1.0 / fscale;
    for (int outer = 0; outer < 80; ++outer) {
      const bool last = 2.0 * P.K() / tau < tol;
      total += socp_center(P, x, false, tau, 0.0, 40, last ? 1e-12 : kCenterTol);
      if (last) break;
      tau *= kMu;
    }
  }
  for (int i = 0; i < P.nv(); ++i)
    y_out[(long long)p * nv + i] = (st == 0) ? x[i] : __longlong_as_double(0x7ff8000000000000LL);
  status[p] = st;
  if (iters) iters[p] = total;
}

}  // namespace bcbf

using namespace bcbf;

extern "C" int bcbf_socp_solve(int Q, int nv, int K, int pc, double rho, const double* w, int w_per_problem,
                               const double* r, const double* c, const double* d, const double* A, const double* b,
                               double tol, double* y, int* status, int* iters, void* stream_) {
  return bcbf_socp_solve_lin(Q, nv, K, pc, rho, w, w_per_problem, r, nullptr, c, d, A, b, tol, y, status, iters, stream_);
}

extern "C" int bcbf_socp_solve_lin(int Q, int nv, int K, int pc, double rho, const double* w, int w_per_problem,
                                   const double* r, const double* q, const double* c, const double* d, const double* A,
                                   const double* b, double tol, double* y, int* status, int* iters, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BCBF_REQUIRE(w && c && d && A && b && y && status, "bcbf_socp_solve: null pointer");
  BCBF_REQUIRE(Q >= 1 && nv >= 1 && nv <= kSV && K >= 1 && K <= kSK && pc >= 1 && pc <= kSP && rho >= 0.0 && tol > 0.0,
               "bcbf_socp_solve: Q=%d nv=%d (<=%d) K=%d (<=%d) pc=%d (<=%d)", Q, nv, kSV, K, kSK, pc, kSP);
  // one thread per problem, 32-thread CTAs so that a few hundred problems still spread over the SMs
  const int ws = w_per_problem ? nv : 0;
  const dim3 grid(ceil_div(Q, 32)), block(32);
#define BCBF_SOCP_LAUNCH(NV, KC, PC) \
  socp_solve_kernel<NV, KC, PC><<<grid, block, 0, stream>>>(Q, nv, K, pc, rho, w, ws, r, q, c, d, A, b, tol, y, status, iters)
  if (nv == 3 && K == 3 && pc == 3) BCBF_SOCP_LAUNCH(3, 3, 3);        // unicycle: [relax, v, omega], CLC + 2 CBCs
  else if (nv == 3 && K == 2 && pc == 3) BCBF_SOCP_LAUNCH(3, 2, 3);   // unicycle, one obstacle
  else if (nv == 3 && K == 1 && pc == 3) BCBF_SOCP_LAUNCH(3, 1, 3);   // unicycle, CLC only
  else if (nv == 2 && K == 2 && pc == 2) BCBF_SOCP_LAUNCH(2, 2, 2);   // pendulum: [relax, u], CLC + CBC
  else BCBF_SOCP_LAUNCH(0, 0, 0);                                     // any other shape: run-time sizes
#undef BCBF_SOCP_LAUNCH
  BCBF_LAUNCH_CHECK();
  return BCBF_OK;
}
