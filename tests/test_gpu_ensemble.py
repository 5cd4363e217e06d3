"""Rollout ensembles (BASELINE configs[4]): batched fit + one-state-per-rollout posterior against the oracle, rollout by
rollout, on the same seeded inputs and jitter draws."""
import pytest
import torch

from oracle import mvgp_oracle as O

pytestmark = pytest.mark.gpu


def _ensemble(seed, R, N, n, m):
    g = torch.Generator().manual_seed(seed)
    f = dict(generator=g, dtype=torch.float64)
    p = m + 1
    X = 3 * (2 * torch.rand(R, N, n, **f) - 1)
    U = 2 * torch.rand(R, N, m, **f) - 1
    Xdot = torch.sin(X) + 0.05 * torch.randn(R, N, n, **f)
    ls = 0.6 + 0.6 * torch.rand(R, n, **f)
    s = 0.8 + torch.rand(R, **f)
    Ra, Rb = torch.randn(R, n, n, **f), torch.randn(R, p, p, **f)
    A = Ra @ Ra.transpose(1, 2) + torch.eye(n, dtype=torch.float64)
    B = Rb @ Rb.transpose(1, 2) + torch.eye(p, dtype=torch.float64)
    C = 0.2 * torch.randn(R, p, n, **f)
    jit = torch.rand(R, N, **f)
    xq = 3 * (2 * torch.rand(R, n, **f) - 1)
    return X, U, Xdot, ls, s, A, B, C, jit, xq


@pytest.mark.parametrize('R,N,n,m', [(7, 150, 3, 2), (5, 100, 2, 1), (3, 200, 3, 2), (2, 300, 3, 2)])
def test_ensemble_fit_and_posterior(R, N, n, m):
    from bayesian_cbf_b200.ensemble import MVGPEnsemble
    X, U, Xdot, ls, s, A, B, C, jit, xq = _ensemble(17, R, N, n, m)
    ens = MVGPEnsemble(n, m).fit(X, U, Xdot, ls, s, A, B, C, jitter=lambda t: jit)
    assert int(ens.tries_used.max()) == 1
    Mk, Bk = ens.posterior(xq)
    for r in range(R):
        hyp = O.Hyper(ls[r], s[r], A[r], B[r], C[r])
        Lref = O.perturbed_cholesky(hyp, X[r], O.homogeneous(U[r]), [jit[r]], direct=True)
        Mk_o, Bk_o = O.posterior_blocks(hyp, X[r], U[r], Xdot[r], Lref, xq[r:r + 1], direct=True)
        prior = float(s[r] * torch.linalg.matrix_norm(B[r], 2))
        assert (Bk[r].cpu() - Bk_o[0]).abs().max() / prior < 1e-9          # covariance, norm-wise vs prior scale
        assert (Mk[r].cpu() - Mk_o[0]).abs().max() < 1e-7 * max(1.0, Mk_o.abs().max().item())
        Lgot = ens.L[r, :N, :N].cpu()
        assert ((Lgot @ Lgot.T) - (Lref @ Lref.T)).abs().max() < 1e-12 * (Lref @ Lref.T).abs().max()


def test_ensemble_cbc_terms_match_single_model_path():
    from bayesian_cbf_b200.ensemble import MVGPEnsemble
    R, N, n, m = 4, 90, 3, 2
    X, U, Xdot, ls, s, A, B, C, jit, xq = _ensemble(23, R, N, n, m)
    ens = MVGPEnsemble(n, m).fit(X, U, Xdot, ls, s, A, B, C, jitter=lambda t: jit)
    Mk, Bk = ens.posterior(xq)
    g = torch.Generator().manual_seed(1)
    gh = torch.randn(R, n, generator=g, dtype=torch.float64).cuda()
    h = torch.randn(R, generator=g, dtype=torch.float64).cuda()
    bfe, e, Asq, A_socp, bfb, status = ens.cbc_terms(Mk, Bk, gh, h, 0.7)
    assert (status.cpu() == 0).all()
    for r in range(R):
        bfe_o, e_o, V, bfv, v = O.cbc1_terms_closed_form(Mk[r].cpu(), Bk[r].cpu(), A[r], gh[r].cpu(), h[r].cpu(), 0.7)
        A_o, bfb_o, _, _ = O.convert_cbc_terms_to_socp_terms(bfe_o, e_o, V, bfv, v, 0)
        assert (bfe[r].cpu() - bfe_o).abs().max() < 1e-12
        assert abs(e[r].item() - e_o.item()) < 1e-12
        assert (A_socp[r].cpu() - A_o).abs().max() < 1e-10
        assert (bfb[r].cpu() - bfb_o).abs().max() < 1e-10


def test_ensemble_log_marginal_matches_single_model_path():
    """Batched value + gradients of R log marginals == R calls of the single-model fused op (mll.py), which is itself
    checked against torch autograd of the dense density (tests/test_fit_mll.py)."""
    from bayesian_cbf_b200.ensemble import ensemble_log_marginal
    from bayesian_cbf_b200.mll import mvgp_log_marginal
    R, N, n, m = 4, 70, 3, 2
    X, U, Xdot, ls, s, A, B, C, _, _ = _ensemble(31, R, N, n, m)
    ls = 0.5 * ls            # short lengthscales keep the jitter-free Gram matrices well conditioned
    UH = torch.cat([torch.ones(R, N, 1, dtype=torch.float64), U], dim=2)
    leaves = [t.clone().cuda().requires_grad_(True) for t in (ls, s, A, B, C)]
    val = ensemble_log_marginal(*leaves, X.cuda(), UH.cuda(), Xdot.cuda())
    w = torch.linspace(0.5, 1.5, R, dtype=torch.float64).cuda()
    grads = torch.autograd.grad((w * val).sum(), leaves)
    for r in range(R):
        lv = [t[r].clone().cuda().requires_grad_(True) for t in (ls, s, A, B, C)]
        v1 = mvgp_log_marginal(*lv, X[r].cuda(), UH[r].cuda(), Xdot[r].cuda())
        g1 = torch.autograd.grad(v1, lv)
        assert abs(val[r].item() - v1.item()) < 1e-10 * abs(v1.item())
        for gb, gs in zip(grads, g1):
            ref = w[r] * gs
            assert (gb[r] - ref).abs().max() < 1e-9 * max(1.0, ref.abs().max().item())


def test_ensemble_hyperparameter_fit_lowers_every_loss():
    from bayesian_cbf_b200.ensemble import EnsembleHyperParameters, fit_ensemble_hyperparameters
    R, N, n, m = 6, 60, 3, 2
    X, U, _, _, _, _, _, _, _, _ = _ensemble(37, R, N, n, m)
    UH = torch.cat([torch.ones(R, N, 1, dtype=torch.float64), U], dim=2)
    Ftrue = torch.zeros(R, N, 3, 3, dtype=torch.float64)
    Ftrue[..., 0, 1] = X[..., 2].cos()
    Ftrue[..., 1, 1] = X[..., 2].sin()
    Ftrue[..., 2, 2] = 0.5
    Xdot = torch.einsum('rinp,rip->rin', Ftrue, UH)
    hp = EnsembleHyperParameters(R, n, m + 1, rank=1, seed=1)
    g = torch.Generator().manual_seed(2)
    first = fit_ensemble_hyperparameters(hp, X.cuda(), U.cuda(), Xdot.cuda(), training_iter=1, lr=1e-9, generator=g)
    last = fit_ensemble_hyperparameters(hp, X.cuda(), U.cuda(), Xdot.cuda(), training_iter=40, lr=0.05, generator=g)
    assert (last < first - 0.05).all(), (first, last)
