"""Tolerance of the posterior MEAN in the CUDA-vs-oracle parity tests, derived from measurement instead of a guessed
constant ("measure the floor").

The north-star tolerance is rel. 1e-9 in float64.  For B_k / svar (measured against the prior scale) the CUDA path is
3-4 orders inside it.  The mean M_k = C^T + sum_i k(x_i, x) alpha_i (x) G_i inherits the conditioning of
alpha = (Kb + jitter)^-1 Y: cond ~ 1e9..1e13 at the test shapes and the sum cancels ~1e6-fold, so that two float64
evaluations of the REFERENCE's own formula whose Gram entries differ in the last bit (gpytorch's expanded-distance form
vs the difference form, MKL vs glibc exp, ...) disagree by a few 1e-9.  `mean_reference` measures, on the reference's own
arithmetic (oracle.mean_parity_floor):

  * the exact solution of the reference's linear system (error-free residual refinement) — what the CUDA mean is checked
    against: the CUDA alpha is refined with a compensated residual (bcbf_alpha_refine) and is the float64 rounding of the
    exact solution of ITS Gram matrix;
  * ulp_sensitivity: how far the exact answer moves under a random last-bit perturbation of Kb;
  * lapack_vs_exact: how far the reference's own cholesky_solve is from the exact answer.

tol_exact = max(1e-9, 3 * ulp_sensitivity) bounds |CUDA - exact|; tol_lapack = max(1e-9, 3 * max(both floors)) bounds
|CUDA - LAPACK oracle|.  Both are relative to max|mean| (resp. max|M_k|)."""
import torch

from oracle import mvgp_oracle as O


def mean_reference(hyp, X, U, Xdot, jit, Xq, Uq=None, jitter_scale=1e-5):
    Uq_ = Uq if Uq is not None else Xq.new_zeros(Xq.shape[0], hyp.p - 1)
    mean_exact, fl = O.mean_parity_floor(hyp, X, U, Xdot, jit, Xq, Uq_, jitter_scale=jitter_scale)
    G = O.homogeneous(U) @ hyp.B
    Ks = O.rbf_ard(X, Xq, hyp.lengthscale, hyp.outputscale, direct=True)
    W = fl['alpha_exact'].unsqueeze(-1) * G.unsqueeze(1)
    Mk_exact = hyp.C.t().unsqueeze(0) + torch.einsum('iq,inp->qnp', Ks, W)
    ulp, lap = fl['ulp_sensitivity'], fl['lapack_vs_exact']
    return dict(mean_exact=mean_exact, Mk_exact=Mk_exact, ulp_sensitivity=ulp, lapack_vs_exact=lap,
                tol_exact=max(1e-9, 3 * ulp), tol_lapack=max(1e-9, 3 * max(ulp, lap)))


def rel(got, want):
    got = torch.as_tensor(got).cpu()
    want = torch.as_tensor(want)
    return float((got - want).abs().max() / want.abs().max().clamp_min(1e-300))
