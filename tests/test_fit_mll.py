"""fit(): the fused log marginal likelihood and its hyper-parameter gradients (bayesian_cbf_b200/mll.py) against torch
autograd of the dense (N n)-dimensional density restated in the oracle (oracle/mvgp_oracle.py:mll_dense), that density
against torch.distributions.MultivariateNormal, and the whole Adam trajectory against the reference's own
_fit_with_warnings run over the dense gpytorch stand-in (oracle/gen_golden_fit.py).  Real gpytorch (a fork, absent here)
evaluates the same density with its own numerics: that last step stays "parity unpinned" (SURVEY 8c)."""
import numpy as np
import pytest
import torch

from oracle import mvgp_oracle as O
from tests import fake_ops


@pytest.fixture(params=['cpu-fakeops', pytest.param('cuda', marks=pytest.mark.gpu)])
def dev(request, monkeypatch):
    if request.param == 'cuda':
        yield 'cuda'
    else:
        with fake_ops.installed(monkeypatch):
            yield 'cpu'


def _problem(seed, N, n, m, ls0=0.7):
    g = torch.Generator().manual_seed(seed)
    f = dict(generator=g, dtype=torch.float64)
    p = m + 1
    X = 3 * (2 * torch.rand(N, n, **f) - 1)
    U = 2 * torch.rand(N, m, **f) - 1
    Xdot = torch.sin(X @ torch.randn(n, n, **f)) + 0.05 * torch.randn(N, n, **f)
    Ra, Rb = torch.randn(n, n, **f), torch.randn(p, p, **f)
    hyp = O.Hyper(lengthscale=ls0 + 0.4 * ls0 * torch.rand(n, **f), outputscale=torch.tensor(1.1, dtype=torch.float64),
                  A=Ra @ Ra.T + torch.eye(n, dtype=torch.float64), B=Rb @ Rb.T + torch.eye(p, dtype=torch.float64),
                  C=0.2 * torch.randn(p, n, **f))
    return X, U, Xdot, hyp


@pytest.mark.parametrize('N,n,m,ls0', [(90, 3, 2, 0.7), (200, 2, 1, 0.2)])
def test_mll_value_and_gradients(N, n, m, ls0, dev):
    import bayesian_cbf_b200.mll as mll
    X, U, Xdot, hyp = _problem(3, N, n, m, ls0)   # short lengthscales keep Kb (no jitter, no noise) well conditioned
    leaves = [t.clone().requires_grad_(True) for t in (hyp.lengthscale, hyp.outputscale, hyp.A, hyp.B, hyp.C)]
    ref = O.mll_dense(O.Hyper(*leaves), X, U, Xdot)
    gref = torch.autograd.grad(ref, leaves)
    ours_leaves = [t.clone().to(dev).requires_grad_(True) for t in (hyp.lengthscale, hyp.outputscale, hyp.A, hyp.B, hyp.C)]
    UH = O.homogeneous(U)
    val = mll.mvgp_log_marginal(*ours_leaves, X.to(dev), UH.to(dev), Xdot.to(dev))
    gours = torch.autograd.grad(val, ours_leaves)
    assert abs(val.item() - ref.item()) < 1e-9 * abs(ref.item())
    for name, a, b in zip(('lengthscale', 'outputscale', 'A', 'B', 'C'), gours, gref):
        err = (a.cpu() - b).abs().max().item() / max(b.abs().max().item(), 1e-12)
        assert err < 1e-7, (name, err)       # the adjoint carries Kb^-1: conditioning-limited, measured ~1e-10


def test_mll_dense_is_the_Nn_dimensional_normal_density():
    """the Kronecker form the fit maximises equals log N(vec Xdot; vec(UH C), Kb (x) A) of
    torch.distributions.MultivariateNormal — what gpytorch's ExactMarginalLogLikelihood evaluates (before / num_data)."""
    X, U, Xdot, hyp = _problem(9, 30, 3, 2, 0.3)
    UH = O.homogeneous(U)
    Kb = O.gram_train(hyp, X, UH)
    cov = O.torch_kron(Kb, hyp.A, batch_dims=0)                        # index order (point, r), r fastest
    mvn = torch.distributions.MultivariateNormal((UH @ hyp.C).reshape(-1), covariance_matrix=cov)
    want = mvn.log_prob(Xdot.reshape(-1)).item()
    got = O.mll_dense(hyp, X, U, Xdot).item()
    assert abs(got - want) < 1e-10 * abs(want), (got, want)


@pytest.mark.parametrize('case', ['ref_fit_unicycle_f64', 'ref_fit_pendulum_rank1_prior_f64'])
def test_fit_trajectory_against_the_reference(case, dev):
    """The reference's own _fit_with_warnings (control_affine_model.py:274-335), run over the dense gpytorch stand-in by
    oracle/gen_golden_fit.py, against `fit` here from the same initial raw parameters and the same target-noise draws:
    the loss of EVERY iteration and the final parameters must agree (Adam, MultiStepLR milestones, noise, loss
    normalisation, the Gamma lengthscale prior and parameter names included)."""
    from functools import partial
    from bayesian_cbf_b200.control_affine_model import ControlAffineExactGP, ControlAffineRegressor
    from tests.golden_util import T, load
    d = load(case)
    n, m, iters, rank = int(d['n']), int(d['m']), int(d['iters']), int(d['rank'])
    prior = tuple(d['prior'].tolist()) or None
    mc = partial(ControlAffineExactGP, rank=(None if rank < 0 else rank), gamma_length_scale_prior=prior)
    reg = ControlAffineRegressor(n, m, device=dev, model_class=mc)
    reg.model.double()
    names = dict(reg.model.named_parameters())
    want_names = {k[len('init/'):] for k in d.files if k.startswith('init/')}
    assert set(names) == want_names                      # state_dicts line up with the reference's parameter names
    with torch.no_grad():
        for k, prm in names.items():
            prm.copy_(T(d['init/' + k]).reshape(prm.shape).to(dev))
    reg.set_fit_noise_source(list(d['noise']))
    # The golden run had float64 as the default dtype (as the reference's unicycle scripts do,
    # unicycle_move_to_pose.py:50).  It matters: the MultiStepLR milestones are `(torch.tensor([.3,.6,.8,.9]) *
    # training_iter).tolist()` (reference :303-305, same expression here) and in float32 0.3 * 50 is 15.00000095, a
    # milestone that never fires.
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        reg.fit(T(d['X']), T(d['U']), T(d['Xdot']), training_iter=iters, lr=float(d['lr']))
    finally:
        torch.set_default_dtype(old)
    losses = torch.stack(reg.fit_losses).cpu().numpy()
    assert losses.shape == d['loss'].shape
    # measured ~1e-10; the later iterations inherit the conditioning of the log marginal's gradient through Adam
    assert np.abs(losses - d['loss']).max() < 1e-7 * max(1.0, np.abs(d['loss']).max()), np.abs(losses - d['loss'])
    for k, prm in names.items():
        want = d['final/' + k].reshape(prm.shape)
        assert np.abs(prm.detach().cpu().numpy() - want).max() < 1e-6 * max(1.0, np.abs(want).max()), k
    ls, s, A, B, C = reg._hyper64()
    for got, key in ((ls, 'final_lengthscale'), (A, 'final_A'), (B, 'final_B'), (C, 'final_C')):
        assert np.abs(got.cpu().numpy() - d[key]).max() < 1e-6 * max(1.0, np.abs(d[key]).max()), key
    assert abs(s - float(d['final_outputscale'])) < 1e-6


def test_cogp_fit_trajectory_against_the_reference(dev):
    """ControlAffineRegressorVector.fit (the CoGP comparator the reference's speed test fits for 50 iterations before
    timing, pendulum.py:1366): dense (N n)-dimensional log marginal on the CUDA Cholesky / inverse, covariance from
    HetergeneousCoregionalizationKernel — against the reference's own fit of the same class over the dense stand-in."""
    from bayesian_cbf_b200.control_affine_model import ControlAffineRegressorVector
    from tests.golden_util import T, load
    d = load('ref_fit_pendulum_cogp_vector_f64')
    n, m, iters = int(d['n']), int(d['m']), int(d['iters'])
    reg = ControlAffineRegressorVector(n, m, device=dev)
    reg.model.double()
    names = dict(reg.model.named_parameters())
    assert set(names) == {k[len('init/'):] for k in d.files if k.startswith('init/')}
    with torch.no_grad():
        for k, prm in names.items():
            prm.copy_(T(d['init/' + k]).reshape(prm.shape).to(dev))
    reg.set_fit_noise_source(list(d['noise']))
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        reg.fit(T(d['X']), T(d['U']), T(d['Xdot']), training_iter=iters, lr=float(d['lr']))
    finally:
        torch.set_default_dtype(old)
    losses = torch.stack(reg.fit_losses).cpu().numpy()
    assert np.abs(losses - d['loss']).max() < 1e-7 * max(1.0, np.abs(d['loss']).max()), np.abs(losses - d['loss'])
    for k, prm in names.items():
        want = d['final/' + k].reshape(prm.shape)
        assert np.abs(prm.detach().cpu().numpy() - want).max() < 1e-6 * max(1.0, np.abs(want).max()), k
    assert np.abs(reg._sigma64().cpu().numpy() - d['final_Sigma']).max() < 1e-6 * np.abs(d['final_Sigma']).max()
    # the fitted comparator predicts through the same class API as before
    mean, var = reg.custom_predict(T(d['X'])[:3].to(dev), T(d['U'])[:3].to(dev))
    assert mean.shape == (3, n) and torch.isfinite(mean).all() and torch.isfinite(var).all()


def test_fit_lowers_the_loss_and_predicts(dev):
    from bayesian_cbf_b200.control_affine_model import ControlAffineRegressorExact
    torch.manual_seed(0)
    N, n, m = 60, 2, 1
    X, U, _, _ = _problem(5, N, n, m)
    Ftrue = lambda X: torch.stack([torch.stack([X[:, 1], 0 * X[:, 0]], -1), torch.stack([-torch.sin(X[:, 0]), 1 + 0 * X[:, 0]], -1)], 1)
    UH = O.homogeneous(U)
    Xdot = torch.einsum('inp,ip->in', Ftrue(X), UH)
    reg = ControlAffineRegressorExact(n, m, device=dev)
    reg.model.double()

    def loss_now():
        ls, s, A, B, C = reg._hyper64()
        h = O.Hyper(ls.cpu(), torch.tensor(s, dtype=torch.float64), A.cpu(), B.cpu(), C.cpu())
        return -O.mll_dense(h, X, U, Xdot).item() / (N * n)

    reg.fit(X, U, Xdot, training_iter=0)
    before = loss_now()
    reg.fit(X, U, Xdot, training_iter=40, lr=0.05)
    after = loss_now()
    assert after < before - 0.05, (before, after)
    mean = reg.fu_func_mean(U.to(dev), X.to(dev))
    assert (mean.cpu() - Xdot).abs().max() < 0.2 * Xdot.abs().max()      # interpolates its own training data
    # checkpoint round trip (control_affine_model.py:862-874)
    sd = reg.state_dict()
    reg2 = ControlAffineRegressorExact(n, m, device=dev)
    reg2.model.double()
    reg2.load_state_dict(sd)
    for a, b in zip(reg._hyper64(), reg2._hyper64()):
        assert np.allclose(torch.as_tensor(a).cpu().numpy(), torch.as_tensor(b).cpu().numpy())


@pytest.mark.gpu
def test_captured_iteration_equals_eager_iteration():
    """fit(cuda_graph=True) replays value + gradients of an iteration from a CUDA graph; same kernels, same order: the loss
    sequence and the final parameters are bit-identical to the eager loop."""
    from bayesian_cbf_b200.control_affine_model import ControlAffineRegressorExact
    X, U, Xdot, _ = _problem(7, 150, 3, 2)
    out = {}
    for mode in (False, True):
        torch.manual_seed(3)
        reg = ControlAffineRegressorExact(3, 2, device='cuda')
        reg.model.double()
        torch.manual_seed(4)                         # the target-noise draws
        reg.fit(X, U, Xdot, training_iter=12, lr=0.05, cuda_graph=mode)
        out[mode] = (torch.stack(reg.fit_losses).cpu(), [p.detach().cpu().clone() for p in reg.model.parameters()])
    assert torch.equal(out[False][0], out[True][0])
    for a, b in zip(out[False][1], out[True][1]):
        assert torch.equal(a, b)
