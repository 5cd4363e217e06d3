"""The reference's own known-answer tests for this path, restated against the drop-in classes (CUDA: `-m gpu`; host logic
over tests/fake_ops.py otherwise), plus the methods of the regressor that only the reference's tests call.

  * tests/test_control_affine_kernel.py:14-111 (reference): HetergeneousMatrixVariateKernel built from FAKE task / data
    kernels must equal closed np.kron formulas on train, test and mixed rows — the plug-in contract of the class
    (it evaluates whatever modules it is handed, matrix_variate_multitask_kernel.py:99-204);
  * tests/test_control_affine_regression.py:151,184-198 (reference): `_predict_flatten`, gradient of `fu_func_mean`;
  * the reference's consumers (gp_algebra / cbc1 / cbc2 / convert_cbc_terms_to_socp_terms, imported unmodified) over OUR
    regressor reproduce the goldens its own regressor produced (tests/ref_consumers_check.py, where /root/reference exists).
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch
from scipy.linalg import block_diag

from tests import fake_ops
from tests.golden_util import T, load

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(params=['cpu-fakeops', pytest.param('cuda', marks=pytest.mark.gpu)])
def dev(request, monkeypatch):
    if request.param == 'cuda':
        yield 'cuda'
    else:
        with fake_ops.installed(monkeypatch):
            yield 'cpu'


# ---- reference tests/test_control_affine_kernel.py, restated ------------------------------------------------------------
def _fake_modules():
    from bayesian_cbf_b200.gp_modules import Kernel

    class ConstantIndexKernel(Kernel):              # reference test :14-28
        def __init__(self, A):
            super().__init__()
            self.A = A

        @property
        def raw_var(self):
            return self.A

        @property
        def covar_matrix(self):
            return self.A

        def forward(self, i, j):
            return self.A[i, j]

    class DataKernel(Kernel):                       # reference test :31-34
        def forward(self, x1, x2, **kw):
            return torch.exp(-((x1[:, None, :] - x2[None, :, :]) ** 2).sum(-1))

    return ConstantIndexKernel, DataKernel


def _np_data_kernel(x1, x2):
    return np.exp(-((x1[:, None, :] - x2[None, :, :]) ** 2).sum(-1))


def kernel_train(H, A, B, X):                       # reference test :37-39
    return np.kron(H @ np.kron(_np_data_kernel(X, X), B) @ H.T, A)


def kernel_test(A, B, X):                           # :42-43
    return np.kron(np.kron(_np_data_kernel(X, X), B), A)


def kernel_train_test(H, A, B, Xtrain, Xtest):      # :46-50
    k12 = np.kron(H @ np.kron(_np_data_kernel(Xtrain, Xtest), B), A)
    return np.vstack((np.hstack((kernel_train(H, A, B, Xtrain), k12)), np.hstack((k12.T, kernel_test(A, B, Xtest)))))


def _rand_psd(rng, n):
    R = rng.random((n, n))
    return R.T @ R + np.diag(np.abs(rng.random(n)))


@pytest.mark.parametrize('D,n,m,ntest', [(5, 1, 2, 1), (7, 3, 2, 3), (4, 2, 1, 2)])
def test_dynamics_model_kernel_with_fake_modules(dev, D, n, m, ntest):
    from bayesian_cbf_b200.control_affine_model import CatEncoder
    from bayesian_cbf_b200.matrix_variate_multitask_kernel import HetergeneousMatrixVariateKernel, MatrixVariateIndexKernel
    ConstantIndexKernel, DataKernel = _fake_modules()
    rng = np.random.default_rng(100 * D + n)
    U, X, Xtest = rng.random((D, m)), rng.random((D, n)), rng.random((ntest, n))
    A, B = _rand_psd(rng, n), _rand_psd(rng, 1 + m)
    UH = np.concatenate((np.ones((D, 1)), U), axis=1)
    H = block_diag(*UH[:, None, :])
    enc = lambda Xs, Us, M: np.concatenate([M * np.ones((Xs.shape[0], 1)), Xs,
                                            np.concatenate([M * np.ones((Xs.shape[0], 1)), Us * M], axis=1)], axis=1)
    t = lambda a: torch.from_numpy(a).to(dev)
    ker = HetergeneousMatrixVariateKernel(
        task_covar_module=MatrixVariateIndexKernel(ConstantIndexKernel(t(A)), ConstantIndexKernel(t(B))),
        data_covar_module=DataKernel(), decoder=CatEncoder(1, n, 1 + m))
    MXU, MXUtest = t(enc(X, U, 1)), t(enc(Xtest, np.zeros((ntest, m)), 0))
    both = torch.cat((MXU, MXUtest), dim=0)
    got = lambda a, b: ker(a, b).evaluate().detach().cpu().numpy()
    tol = dict(rel=1e-12, abs=1e-14)                      # float64 (the reference runs this test in float32)
    assert got(MXU, MXU) == pytest.approx(kernel_train(H, A, B, X), **tol)
    assert got(MXUtest, MXUtest) == pytest.approx(kernel_test(A, B, Xtest), **tol)
    assert got(both, both) == pytest.approx(kernel_train_test(H, A, B, X, Xtest), **tol)
    assert ker(both, both, diag=True).evaluate().cpu().numpy() == pytest.approx(
        np.diag(kernel_train_test(H, A, B, X, Xtest)), **tol)
    assert ker.num_tasks == n * (1 + m)


def test_plugin_path_equals_the_fused_stock_path(dev):
    """The stock ScaleKernel(RBFKernel) takes the fused bcbf_gram_ca blocks; the same kernel hidden behind a subclass takes
    the plug-in path (dense K from the module, then bcbf_ca_weight).  Both must give the reference's matrix."""
    from bayesian_cbf_b200.control_affine_model import ControlAffineRegressor
    from bayesian_cbf_b200.gp_modules import ScaleKernel
    d = load('ref_kernel_f64')
    reg = ControlAffineRegressor(int(d['n']), int(d['m']), device=dev)
    reg.model.double()
    reg.set_hyperparameters(d['h_lengthscale'], d['h_outputscale'], d['h_A'], d['h_B'], d['h_C'])
    cm = reg.model.covar_module
    both = torch.cat([T(d['MXU']), T(d['MXUte'])], dim=0).to(dev)
    fused = cm(both, both).evaluate()
    assert cm._stock_rbf() is not None

    class Opaque(ScaleKernel):                      # same arithmetic, but not the stock type: plug-in path
        pass

    opaque = Opaque(cm.data_covar_module.base_kernel)
    opaque.raw_outputscale = cm.data_covar_module.raw_outputscale
    cm.data_covar_module = opaque
    assert cm._stock_rbf() is None
    plug = cm(both, both).evaluate()
    want = np.asarray(d['K_mixed'])
    sc = np.abs(want).max()
    assert np.abs(plug.detach().cpu().numpy() - want).max() / sc < 1e-12
    assert np.abs(fused.detach().cpu().numpy() - want).max() / sc < 1e-12


# ---- regressor methods the reference's tests call -----------------------------------------------------------------------
def _fitted(dev, case='ref_predict_unicycle_f64'):
    from bayesian_cbf_b200.control_affine_model import ControlAffineRegressor
    d = load(case)
    reg = ControlAffineRegressor(int(d['n']), int(d['m']), device=dev)
    reg.model.double()
    reg.set_hyperparameters(d['hb_lengthscale'], d['hb_outputscale'], d['hb_A'], d['hb_B'], d['hb_C'])
    reg.fit(T(d['X']), T(d['U']), T(d['Xdot']), training_iter=0)
    reg.set_jitter_source([T(d['base_first_rand%d' % i]) for i in range(int(d['base_first_nrand']))])
    return reg, d


@pytest.mark.parametrize('case', ['ref_predict_unicycle_f64', 'ref_predict_pendulum_f64'])
def test_grad_gp_is_the_gradient_of_the_posterior(dev, case):
    """custom_predict(grad_gp=True) / _grad_fu_func_mean (reference :447-477, :759-771) against autograd through
    fu_func_mean (what the reference's commented-out check compares with, tests/test_control_affine_regression.py:184-191)
    and the Hessian of the scalar variance by double autograd."""
    reg, d = _fitted(dev, case)
    n = int(d['n'])
    Xt, Ut, Xtp, Utp = [T(d[k]).to(dev) for k in ('Xt', 'Ut', 'Xtp', 'Utp')]
    b = Xt.shape[0]
    got = reg._grad_fu_func_mean(Xt, Ut)
    assert got.shape == (b, n * n)
    one = reg._grad_fu_func_mean(Xt[0], Ut[0])
    assert one.shape == (n * n,) and torch.allclose(one, got[0], rtol=1e-12, atol=1e-14)
    for t in range(b):
        x = Xt[t].clone().requires_grad_(True)
        mean = reg.fu_func_mean(Ut[t], x)
        J = torch.stack([torch.autograd.grad(mean[r], x, retain_graph=True)[0] for r in range(n)])     # (r, d)
        assert (got[t].reshape(n, n) - J).abs().max() < 1e-9 * max(1.0, J.abs().max().item())
    # covariance of the gradient process: rows (t, d), columns (t', d')
    _, H = reg.custom_predict(Xt, Ut, Xtestp_in=Xtp, Utestp_in=Utp, grad_gp=True, scalar_var_only=True)
    assert H.shape == (b * n, b * n)
    for t, tp in ((0, 0), (1, 2)):
        x = Xt[t:t + 1].clone().requires_grad_(True)
        xp = Xtp[tp:tp + 1].clone().requires_grad_(True)
        _, sv = reg.custom_predict(x, Ut[t:t + 1], Xtestp_in=xp, Utestp_in=Utp[tp:tp + 1], scalar_var_only=True)
        g = torch.autograd.grad(sv.reshape(()), x, create_graph=True)[0].reshape(-1)
        Hx = torch.stack([torch.autograd.grad(g[dd], xp, retain_graph=True)[0].reshape(-1) for dd in range(n)])
        blk = H.reshape(b, n, b, n)[t, :, tp, :]
        assert (blk - Hx).abs().max() < 1e-8 * max(1.0, Hx.abs().max().item())
    _, full = reg.custom_predict(Xt, Ut, grad_gp=True)
    A = reg.get_kernel_param('A').double()
    assert full.shape == (1, b * n * n, b * n * n)
    _, Hs = reg.custom_predict(Xt, Ut, grad_gp=True, scalar_var_only=True)
    assert torch.allclose(full[0], torch.kron(Hs, A), rtol=1e-12, atol=1e-14)
    # prior only
    from bayesian_cbf_b200.control_affine_model import ControlAffineRegressor
    prior = ControlAffineRegressor(n, int(d['m']), device=dev)
    prior.model.double()
    m0, c0 = prior.custom_predict(Xt, Ut, grad_gp=True, scalar_var_only=True)
    assert float(m0.abs().max()) == 0.0 and c0.shape == (b * n, b * n)


def test_predict_flatten_and_cbf_func(dev):
    reg, d = _fitted(dev)
    n, m = int(d['n']), int(d['m'])
    Xt, Ut = T(d['Xt']), T(d['Ut'])                    # CPU inputs: results come back on the input's device
    b = Xt.shape[0]
    mean, cov = reg._predict_flatten(Xt.numpy(), Ut.numpy())
    assert mean.shape == (b, n) and cov.shape == (b, n, n, b) and mean.device.type == 'cpu'
    assert np.abs(mean.numpy() - d['base_first_out0']).max() < 1e-9 * np.abs(d['base_first_out0']).max()
    want = np.asarray(d['base_first_out1']).reshape(b * n, b * n).reshape(b, n, n, b)
    assert np.abs(cov.numpy() - want).max() < 1e-9 * np.abs(want).max()
    grad_h = torch.randn(1 + m, dtype=torch.float64, generator=torch.Generator().manual_seed(0))
    mh, ch = reg._cbf_func(Xt[:1], grad_h)
    assert ch is None and mh.shape == (1, n)
    assert torch.allclose(mh, grad_h @ reg.predict(Xt[:1], return_cov=False), rtol=1e-12, atol=1e-14)
    gh = torch.randn((1 + m) * n, dtype=torch.float64, generator=torch.Generator().manual_seed(1))
    _, cov_F = reg.predict(Xt[:1], return_cov=True)
    _, ch = reg._cbf_func(Xt[:1], gh.reshape(-1, 1)[:1 + m].reshape(1 + m), return_cov=False)
    assert ch is None


# ---- the reference's consumers over our regressor -----------------------------------------------------------------------
def _run_consumers(device):
    res = subprocess.run([sys.executable, os.path.join(ROOT, 'tests', 'ref_consumers_check.py'), device],
                         capture_output=True, text=True, timeout=900)
    assert res.returncode == 0 and 'REF_CONSUMERS_OK' in res.stdout, res.stdout[-2000:] + res.stderr[-4000:]


@pytest.mark.skipif(not os.path.isdir('/root/reference/bayes_cbf'), reason="needs the reference checkout (build container)")
def test_reference_consumers_over_our_regressor_host_logic():
    _run_consumers('cpu')


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.isdir('/root/reference/bayes_cbf'), reason="needs the reference checkout next to a GPU")
def test_reference_consumers_over_our_regressor_cuda():
    _run_consumers('cuda')
