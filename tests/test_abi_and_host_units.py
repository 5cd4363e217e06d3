"""CPU-side checks: the C-ABI library loads and exports every symbol include/bcbf.h declares (no compute calls), and
small host utilities behave like the reference's (misc.py doctest, CatEncoder, gp_algebra on analytic GPs)."""
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from bayesian_cbf_b200 import _lib
    header = open(os.path.join(ROOT, 'include', 'bcbf.h')).read()
    header = re.sub(r'/\*.*?\*/', '', header, flags=re.S)
    declared = set(re.findall(r'\b(bcbf_[a-z0-9_]+)\s*\(', header))
    assert len(declared) >= 25
    lib = _lib.load()
    for name in sorted(declared):
        assert hasattr(lib, name), "libbcbf.so does not export %s" % name
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    assert lib.bcbf_version() >= 100
    assert lib.bcbf_padded(129) == 256


def test_compute_without_gpu_fails_loudly():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from bayesian_cbf_b200 import ops
    from bayesian_cbf_b200.control_affine_model import ControlAffineRegressor
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.gram_train(torch.zeros(4, 2, dtype=torch.float64), torch.zeros(4, 2, dtype=torch.float64),
                       torch.eye(2, dtype=torch.float64), torch.ones(2, dtype=torch.float64), 1.0)
    reg = ControlAffineRegressor(2, 1, device='cpu')
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        reg.custom_predict(torch.zeros(1, 2))


def test_torch_kron_doctest():
    """bayes_cbf/misc.py:82-95."""
    from bayesian_cbf_b200.misc import torch_kron
    B, A = torch.rand(5, 3, 3), torch.rand(5, 2, 2)
    AB = torch_kron(A, B)
    assert torch.allclose(AB[1, :3, :3], A[1, 0, 0] * B[1])
    BA = torch_kron(B, A)
    assert torch.allclose(BA[1, :2, :2], B[1, 0, 0] * A[1])
    C, D = torch.rand(2, 3), torch.rand(3, 2)
    assert torch.allclose(torch_kron(C, D, batch_dims=0), torch.kron(C, D))


def test_cat_encoder_and_encode_from_xu():
    from bayesian_cbf_b200.control_affine_model import CatEncoder, ControlAffineExactGP, IdentityLikelihood
    enc, X = CatEncoder.from_data(torch.ones(4, 1), torch.rand(4, 3), torch.rand(4, 2))
    assert enc.sizes == [1, 3, 2] and X.shape == (4, 6)
    M, Xs, U = enc.decode(X)
    assert M.shape == (4, 1) and Xs.shape == (4, 3) and U.shape == (4, 2)
    gp = ControlAffineExactGP(3, 2, IdentityLikelihood())
    _, mxu = gp.encode_from_XU(torch.rand(5, 3), torch.rand(5, 2), 1)
    assert mxu.shape == (5, 1 + 3 + 3) and (mxu[:, 0] == 1).all() and (mxu[:, 4] == 1).all()
    _, mxt = gp.encode_from_XU(torch.rand(5, 3))
    assert (mxt[:, 0] == 0).all() and (mxt[:, 4:] == 0).all()
    gp.set_train_data(torch.rand(5, 3), torch.rand(5, 2), torch.rand(5, 3))
    assert gp.train_targets.shape == (15,)


class SimpleGP:
    """Analytic RBF GP of tests/test_gp_algebra.py:91-127 (reference)."""

    def __init__(self, m_true, ls):
        self.m_true, self.ls = m_true, ls

    def mean(self, x):
        return self.m_true @ x

    def knl(self, x, xp):
        d = x - xp
        return torch.exp(-0.5 * (d * d / self.ls ** 2).sum())


def test_gradient_gp_against_analytic_rbf():
    """tests/test_gp_algebra.py:130-142: GradientGP mean = m, knl = (L^-1 - L^-1 d d^T L^-1) k."""
    from bayesian_cbf_b200.gp_algebra import GaussianProcess, GradientGP
    torch.manual_seed(1)
    n = 3
    m_true, ls = torch.rand(n), 0.5 + torch.rand(n)
    sg = SimpleGP(m_true, ls)
    gp = GaussianProcess(sg.mean, sg.knl, shape=(1,), name="simple")
    ggp = GradientGP(gp, x_shape=(n,))
    x, xp = torch.rand(n), torch.rand(n)
    assert torch.allclose(ggp.mean(x), m_true)
    d = x - xp
    il2 = 1 / ls ** 2
    want = (torch.diag(il2) - torch.outer(il2 * d, il2 * d)) * sg.knl(x, xp)
    assert torch.allclose(ggp.knl(x, xp), want, rtol=1e-5, atol=1e-6)
    assert torch.allclose(ggp.knl(x, x), torch.diag(il2), rtol=1e-5, atol=1e-6)


def test_gp_algebra_affine_and_matmul():
    from bayesian_cbf_b200.gp_algebra import DeterministicGP, GaussianProcess
    n = 2
    K = torch.tensor([[2.0, 0.3], [0.3, 1.0]])
    f = GaussianProcess(lambda x: 2 * x, lambda x, xp: K * torch.exp(-(x - xp).pow(2).sum()), shape=(n,), name="f")
    d = DeterministicGP(lambda x: torch.tensor([1.0, -1.0]) * x.sum(), shape=(n,), name="d")
    x, xp = torch.tensor([0.3, 0.4]), torch.tensor([0.1, 0.9])
    e = d.t() @ f + DeterministicGP(lambda x: x[:1], shape=(1,)) * 3.0
    dm = d.mean(x)
    assert torch.allclose(e.mean(x), dm @ (2 * x) + 3.0 * x[:1])
    assert torch.allclose(e.knl(x, xp), dm @ (K * torch.exp(-(x - xp).pow(2).sum())) @ d.mean(xp))
    s = (f + f)
    assert torch.allclose(s.knl(x, x), 4 * K)
    assert torch.allclose((f * 2.0).knl(x, x), 4 * K)
    q = f.t() @ f
    assert q.shape == (1,)
    assert torch.allclose(q.mean(x), (2 * x) @ (2 * x) + K.trace())


def test_affine_and_quadratic_term_extraction():
    from bayesian_cbf_b200.misc import get_affine_terms, get_quadratic_terms
    A, b = torch.tensor([1.0, -2.0, 0.5]), torch.tensor(0.7)
    lin, const = get_affine_terms(lambda u: A @ u + b, torch.rand(3))
    assert torch.allclose(lin, A) and torch.allclose(const, b)
    Q = torch.tensor([[2.0, 0.5], [0.5, 1.0]])
    pvec, r = torch.tensor([0.3, -0.2]), torch.tensor(0.1)
    Qg, pg, rg = get_quadratic_terms(lambda u: u @ Q @ u + pvec @ u + r, torch.rand(2))
    assert torch.allclose(Qg, Q) and torch.allclose(pg, pvec) and torch.allclose(rg, r, atol=1e-6)


def test_safety_factors():
    from bayesian_cbf_b200.cbc1 import cbc1_safety_factor
    from bayesian_cbf_b200.cbc2 import cbc2_safety_factor
    assert abs(cbc1_safety_factor(0.01) - 2.3263478740408408) < 1e-12     # sqrt(2) erfinv(0.98)
    assert abs(cbc2_safety_factor(0.01) - np.sqrt(99.0)) < 1e-12
