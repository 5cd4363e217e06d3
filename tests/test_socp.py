"""The per-step safety SOCP: the CUDA batched barrier solver against its CPU restatement on the same
problems (decisions identical, solutions to 1e-7), and the restatement against scipy's SLSQP (independent method).
cvxpy / GUROBI — what the reference calls (unicycle_move_to_pose.py:926-964) — are not installable here."""
import numpy as np
import pytest
import torch
from scipy.optimize import minimize

from oracle import socp_oracle as S

RHO = 2.3263478740408408   # sqrt(2) erfinv(1 - 2 * 0.01)


def _problems(seed, Q, nv=3, K=3, pc=3):
    rng = np.random.default_rng(seed)
    c = rng.normal(size=(Q, K, nv))
    d = rng.normal(size=(Q, K)) + 0.5
    A = 0.3 * rng.normal(size=(Q, K, pc, nv))
    b = 0.3 * rng.normal(size=(Q, K, pc))
    # controller structure: variable 0 is the relaxation, it enters only the first (CLC) cone, with coefficient 1
    A[:, :, :, 0] = 0
    c[:, 0, 0] = 1
    c[:, 1:, 0] = 0
    w = np.array([0.33, 0.33, 0.33])[:nv]
    return w, c, d, A, b


def test_oracle_against_slsqp():
    w, c, d, A, b = _problems(0, 40)
    r = np.zeros(3)
    n_inf = 0
    for p in range(40):
        y, st, _ = S.solve(w, r, c[p], d[p], A[p], b[p], RHO)
        cons = [{'type': 'ineq', 'fun': (lambda y, k=k: c[p, k] @ y + d[p, k] - RHO * np.linalg.norm(A[p, k] @ y + b[p, k]))}
                for k in range(3)]
        best = None
        rng = np.random.default_rng(p)
        for x0 in (np.zeros(3), rng.normal(size=3), 3 * rng.normal(size=3)):
            res = minimize(lambda y: (w * y ** 2).sum(), x0, constraints=cons, method='SLSQP',
                           options=dict(ftol=1e-14, maxiter=500))
            if res.success and all(cn['fun'](res.x) > -1e-8 for cn in cons) and (best is None or res.fun < best.fun):
                best = res
        if st == 1:
            n_inf += 1
            assert best is None, "restatement says infeasible, SLSQP found %r" % (best.x,)
        else:
            assert all(cn['fun'](y) > -1e-9 for cn in cons)
            if best is not None:
                assert abs((w * y ** 2).sum() - best.fun) < 1e-6 * max(1.0, abs(best.fun))
    assert 5 < n_inf < 35        # the sample exercises both decisions


def test_oracle_known_answers():
    # no active constraint: y = r
    w = np.array([1.0, 2.0]); r = np.array([0.3, -0.2])
    y, st, _ = S.solve(w, r, np.array([[0.0, 0.0]]), np.array([5.0]), np.zeros((1, 2, 2)), np.zeros((1, 2)), 1.0)
    assert st == 0 and np.allclose(y, r, atol=1e-7)
    # half-space y0 >= 1 (A = 0, b = 0): projection of the origin
    y, st, _ = S.solve(np.array([1.0, 1.0]), np.zeros(2), np.array([[1.0, 0.0]]), np.array([-1.0]), np.zeros((1, 2, 2)),
                       np.zeros((1, 2)), 1.0)
    assert st == 0 and np.allclose(y, [1.0, 0.0], atol=1e-6)
    # contradictory half-spaces y0 >= 1 and -y0 >= 1
    y, st, _ = S.solve(np.array([1.0, 1.0]), np.zeros(2), np.array([[1.0, 0.0], [-1.0, 0.0]]), np.array([-1.0, -1.0]),
                       np.zeros((2, 2, 2)), np.zeros((2, 2)), 1.0)
    assert st == 1 and np.isnan(y).all()
    # unit ball ||y - (2,0)|| <= 1 written as 1 >= ||I y - (2,0)||: closest point to the origin is (1, 0)
    y, st, _ = S.solve(np.array([1.0, 1.0]), np.zeros(2), np.zeros((1, 2)), np.array([1.0]), np.eye(2)[None],
                       np.array([[-2.0, 0.0]]), 1.0)
    assert st == 0 and np.allclose(y, [1.0, 0.0], atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize('nv,K,pc', [(3, 3, 3), (2, 2, 2), (3, 1, 3), (4, 4, 4)])
def test_cuda_solver_matches_restatement(nv, K, pc):
    from bayesian_cbf_b200 import ops
    Q = 300
    w, c, d, A, b = _problems(7, Q, nv, K, pc)
    if nv == 4:
        w = np.array([0.33, 0.33, 0.33, 0.5])
    r = 0.1 * np.random.default_rng(1).normal(size=(Q, nv))
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    y, status, iters = ops.socp_solve(T(w), T(c), T(d), T(A), T(b), RHO, r=T(r))
    y, status = y.cpu().numpy(), status.cpu().numpy()
    n_inf = 0
    for p in range(Q):
        yo, st, _ = S.solve(w, r[p], c[p], d[p], A[p], b[p], RHO)
        assert st == status[p], "feasibility decision differs at problem %d" % p
        if st == 0:
            assert np.abs(y[p] - yo).max() < 1e-7 * max(1.0, np.abs(yo).max())
        else:
            n_inf += 1
            assert np.isnan(y[p]).all()
    assert n_inf < Q and (n_inf > 0 or K == 1)     # a single relaxed cone is always feasible
    assert int(iters.max()) < 2000


# ---- the reference's own known answer for this step: tests/test_optimizers.py:6-119 (the cvxopt documentation SOCP) ------
REF_C = np.array([-2., 1., 5.])
REF_A = [np.array([[-13., 3., 5.], [-12., 12., -6.]]), np.array([[-3., 6., 2.], [1., 9., 2.], [-1., -19., 3.]])]
REF_B = [np.array([-3., -2.]), np.array([0., 3., -42.])]
REF_CC = [np.array([-12., -6., 5.]), np.array([-3., 6., -10.])]
REF_D = [np.array(-12.), np.array(27.)]
REF_UOPT = np.array([-5.02, -5.77, -8.52])                 # printed by cvxopt, 3 significant digits
REF_NAMED = list(zip(("1", "2"), zip(REF_A, REF_B, REF_CC, REF_D)))


def _kkt_certificate(y):
    """Optimality certificate independent of any solver: both cones active, and c = sum_k lambda_k grad g_k(y) with
    lambda >= 0 for g_k(y) = c_k^T y + d_k - |A_k y + b_k|.  Returns (max |g_k|, stationarity residual, lambda); the
    multipliers are the first components of cvxopt's printed dual variables zq (1.34 and 1.02, reference test :22-31)."""
    g, grads = [], []
    for k in range(2):
        z = REF_A[k] @ y + REF_B[k]
        g.append(REF_CC[k] @ y + float(REF_D[k]) - np.linalg.norm(z))
        grads.append(REF_CC[k] - REF_A[k].T @ z / np.linalg.norm(z))
    Gm = np.array(grads).T
    lam, *_ = np.linalg.lstsq(Gm, REF_C, rcond=None)
    return np.abs(g).max(), np.linalg.norm(Gm @ lam - REF_C) / np.linalg.norm(REF_C), lam


def test_convert_socp_to_cvxopt_format_matches_the_reference_expectation():
    from bayesian_cbf_b200.optimizers import convert_socp_to_cvxopt_format
    c, Gqs, hqs = convert_socp_to_cvxopt_format(REF_C, REF_NAMED)
    exp_G = [np.array([[12., 13., 12.], [6., -3., -12.], [-5., -5., 6.]]),
             np.array([[3., 3., -1., 1.], [-6., -6., -9., 19.], [10., -2., -2., -3.]])]
    exp_h = [np.array([-12., -3., -2.]), np.array([27., 0., 3., -42.])]
    for G, h, eG, eh in zip(Gqs, hqs, exp_G, exp_h):      # reference test :97-103
        assert G.T == pytest.approx(eG)
        assert h.flatten() == pytest.approx(eh)
    assert c is REF_C


def test_restatement_reproduces_the_cvxopt_documentation_answer():
    A = np.zeros((2, 3, 3)); b = np.zeros((2, 3))
    A[0, :2], A[1] = REF_A[0], REF_A[1]
    b[0, :2], b[1] = REF_B[0], REF_B[1]
    y, st, _ = S.solve(np.zeros(3), np.zeros(3), np.array(REF_CC), np.array([float(d) for d in REF_D]), A, b, 1.0, 1e-10,
                       q=REF_C)
    assert st == 0
    assert y == pytest.approx(REF_UOPT, abs=1e-3, rel=1e-2)                  # the reference's tolerance (:18-21)
    gmax, resid, lam = _kkt_certificate(y)
    assert gmax < 1e-8 and resid < 1e-7 and (lam > 0).all()
    assert lam == pytest.approx([1.34, 1.02], abs=1e-2)                       # cvxopt's zq[0][0], zq[1][0]


def test_optimizer_front_door_host_logic(monkeypatch):
    from tests import fake_ops
    with fake_ops.installed(monkeypatch):
        from bayesian_cbf_b200 import optimizers as opt
        u = opt.optimizer_socp_cvxopt(np.random.rand(3), REF_C, REF_NAMED, device='cpu')
        assert u == pytest.approx(REF_UOPT, rel=1e-2)                        # reference test :116-119
        u2 = opt.optimizer_socp_cvxpy(np.zeros(3), REF_C, REF_NAMED, device='cpu')
        assert np.allclose(u, u2, atol=1e-6)
        with pytest.raises(opt.InfeasibleProblemError):
            opt.optimizer_socp_cvxopt(np.zeros(2), np.array([1., 0.]),
                                      [("a", (np.zeros((1, 2)), np.zeros(1), np.array([1., 0.]), np.array(-1.))),
                                       ("b", (np.zeros((1, 2)), np.zeros(1), np.array([-1., 0.]), np.array(-1.)))],
                                      device='cpu')
        # QPController's program (controllers.py:640-655): min |A u + b|^2 s.t. 0 <= c^T u + d, diagonal A
        A = np.diag([np.sqrt(10.), 1., 1.])
        u = opt.optimizer_qp_cvxpy(np.zeros(3), (A, np.zeros(3)), [('Safety', (np.array([1., 2., 0.]), np.array(-1.)))],
                                   device='cpu')
        res = minimize(lambda y: ((A @ y) ** 2).sum(), np.ones(3), method='SLSQP', options=dict(ftol=1e-15),
                       constraints=[{'type': 'ineq', 'fun': lambda y: y[0] + 2 * y[1] - 1}])
        assert np.abs(u - res.x).max() < 1e-6
        with pytest.raises(ValueError):
            opt.optimizer_qp_cvxpy(np.zeros(2), (np.ones((2, 2)), np.zeros(2)), [])


@pytest.mark.gpu
def test_cuda_solver_reproduces_the_cvxopt_documentation_answer():
    """bcbf_socp_solve_lin through the reference's call shape (optimizers.py:44-102) on the GPU."""
    from bayesian_cbf_b200 import optimizers as opt
    u = opt.optimizer_socp_cvxopt(np.random.rand(3), REF_C, REF_NAMED)
    assert u == pytest.approx(REF_UOPT, abs=1e-3, rel=1e-2)
    gmax, resid, lam = _kkt_certificate(u)
    assert gmax < 1e-8 and resid < 1e-7 and lam == pytest.approx([1.34, 1.02], abs=1e-2)
    A = np.zeros((2, 3, 3)); b = np.zeros((2, 3))
    A[0, :2], A[1] = REF_A[0], REF_A[1]
    b[0, :2], b[1] = REF_B[0], REF_B[1]
    yo, st, _ = S.solve(np.zeros(3), np.zeros(3), np.array(REF_CC), np.array([float(d) for d in REF_D]), A, b, 1.0, 1e-9,
                        q=REF_C)
    assert np.abs(u - yo).max() < 1e-9                                       # CUDA == restatement


@pytest.mark.gpu
def test_cuda_solver_with_linear_terms_matches_restatement():
    from bayesian_cbf_b200 import ops
    Q = 200
    w, c, d, A, b = _problems(11, Q)
    rng = np.random.default_rng(3)
    r, q = 0.1 * rng.normal(size=(Q, 3)), 0.5 * rng.normal(size=(Q, 3))
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    y, status, _ = ops.socp_solve(T(w), T(c), T(d), T(A), T(b), RHO, r=T(r), q=T(q))
    y, status = y.cpu().numpy(), status.cpu().numpy()
    for p in range(Q):
        yo, st, _ = S.solve(w, r[p], c[p], d[p], A[p], b[p], RHO, q=q[p])
        assert st == status[p]
        if st == 0:
            assert np.abs(y[p] - yo).max() < 1e-7 * max(1.0, np.abs(yo).max())
