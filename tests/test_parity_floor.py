"""CPU tests of the exact-residual machinery that defines the parity floor of the posterior mean (oracle/mvgp_oracle.py:
ExactResidual, solve_exact, mean_parity_floor) and of the host-side stand-in of bcbf_alpha_refine."""
from fractions import Fraction

import numpy as np
import torch

from oracle import mvgp_oracle as O


def _problem(N, seed=0, Q=64):
    g = torch.Generator().manual_seed(seed)
    f = dict(generator=g, dtype=torch.float64)
    X = 4 * torch.rand(N, 3, **f) - 2
    U = 2 * torch.rand(N, 2, **f) - 1
    Xdot = torch.sin(X) + 0.01 * torch.randn(N, 3, **f)
    Ra, Rb = torch.randn(3, 3, **f), torch.randn(3, 3, **f)
    hyp = O.Hyper(torch.tensor([0.7, 0.9, 1.1], dtype=torch.float64), torch.tensor(1.3, dtype=torch.float64),
                  Ra @ Ra.T + torch.eye(3, dtype=torch.float64), Rb @ Rb.T + torch.eye(3, dtype=torch.float64),
                  0.1 * torch.randn(3, 3, **f))
    jit = torch.rand(N, **f)
    Xq = 4 * torch.rand(Q, 3, **f) - 2
    Uq = 2 * torch.rand(Q, 2, **f) - 1
    return X, U, Xdot, hyp, jit, Xq, Uq


def test_residual_exact_against_rational_arithmetic():
    X, U, Xdot, hyp, jit, _, _ = _problem(24)
    UH = O.homogeneous(U)
    Kbp = O.gram_train(hyp, X, UH, direct=True) + 1e-5 * torch.diag(jit)
    Y = O.residual_targets(hyp, UH, Xdot)
    alpha = torch.cholesky_solve(Y, torch.linalg.cholesky(Kbp))
    r = O.residual_exact(Kbp, alpha, Y)
    K, a, y = Kbp.tolist(), alpha.tolist(), Y.tolist()
    for i in range(24):
        for c in range(3):
            exact = Fraction(y[i][c]) - sum(Fraction(K[i][k]) * Fraction(a[k][c]) for k in range(24))
            scale = sum(abs(K[i][k] * a[k][c]) for k in range(24))
            # 4 x 4 slices of 19 bits: the dropped tail is < 2^-76 of rowscale * columnscale per term
            assert abs(float(Fraction(r[i, c].item()) - exact)) <= 1e-19 * scale + 2e-16 * abs(float(exact))


def test_residual_exact_beats_float64_and_matches_longdouble():
    X, U, Xdot, hyp, jit, _, _ = _problem(600, seed=1)
    UH = O.homogeneous(U)
    Kbp = O.gram_train(hyp, X, UH, direct=True) + 1e-5 * torch.diag(jit)
    Y = O.residual_targets(hyp, UH, Xdot)
    alpha = torch.cholesky_solve(Y, torch.linalg.cholesky(Kbp))
    r = O.residual_exact(Kbp, alpha, Y).numpy()
    rl = (Y.numpy().astype(np.longdouble) - Kbp.numpy().astype(np.longdouble) @ alpha.numpy().astype(np.longdouble))
    scale = (Kbp.abs() @ alpha.abs()).numpy()
    assert (np.abs(r - rl.astype(np.float64)) / scale).max() < 2.0 ** -60       # long double carries 64 bits
    r64 = (Y - Kbp @ alpha).numpy()
    assert (np.abs(r64 - r) / scale).max() > 2.0 ** -58                          # plain float64 is visibly worse


def test_solve_exact_converges_and_inverse_refinement_reaches_it(monkeypatch):
    from tests import fake_ops
    X, U, Xdot, hyp, jit, Xq, Uq = _problem(500, seed=2)
    UH = O.homogeneous(U)
    Kbp = O.gram_train(hyp, X, UH, direct=True) + 1e-5 * torch.diag(jit)
    L = torch.linalg.cholesky(Kbp)
    Y = O.residual_targets(hyp, UH, Xdot)
    a_e, rel = O.solve_exact(Kbp, L, Y)
    assert rel < 1e-15
    # the residual of the converged solution is at the level of rounding alpha itself to float64
    assert O.residual_exact(Kbp, a_e, Y).abs().max() < 4e-16 * (Kbp.abs() @ a_e.abs()).max()
    # host-side stand-in of bcbf_alpha_refine (explicit inverse + 2 refinement steps, long double residual)
    Linv = torch.linalg.solve_triangular(L, torch.eye(500, dtype=torch.float64), upper=False)
    a_r = fake_ops.alpha_refine(X, UH, hyp.B, hyp.lengthscale, hyp.outputscale, Linv, Y, jit, 1e-5, iters=2)
    a_0 = fake_ops.alpha_refine(X, UH, hyp.B, hyp.lengthscale, hyp.outputscale, Linv, Y, jit, 1e-5, iters=0)
    Ks = O.rbf_ard(X, Xq, hyp.lengthscale, hyp.outputscale, direct=True) * ((UH @ hyp.B) @ O.homogeneous(Uq).t())
    m_e = Ks.t() @ a_e
    err = lambda a: float((Ks.t() @ a - m_e).abs().max() / m_e.abs().max())
    assert err(a_r) < 1e-11 and err(a_r) <= err(a_0)


def test_mean_parity_floor_reports_reference_rounding_and_input_sensitivity():
    X, U, Xdot, hyp, jit, Xq, Uq = _problem(700, seed=3)
    mean_exact, fl = O.mean_parity_floor(hyp, X, U, Xdot, jit, Xq, Uq)
    assert fl['refinement_last_step'] < 1e-15
    assert 0 < fl['ulp_sensitivity'] < 1e-6 and 0 < fl['lapack_vs_exact'] < 1e-6
    # the exact mean agrees with the reference-restated path (posterior_blocks: LAPACK) to the reported LAPACK distance
    L = O.perturbed_cholesky(hyp, X, O.homogeneous(U), [jit], direct=True)
    _, _, mean, _ = O.posterior_blocks(hyp, X, U, Xdot, L, Xq, Uq, direct=True)
    d = float((mean - mean_exact).abs().max() / mean_exact.abs().max())
    assert d < 2 * fl['lapack_vs_exact'] + 1e-12
