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
