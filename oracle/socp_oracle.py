"""TEST INFRASTRUCTURE ONLY — CPU float64 restatement (numpy, one problem at a time) of the batched barrier solver in
bayesian_cbf_b200/csrc/socp.cu, which replaces the cvxpy + GUROBI call of the reference's ControllerCLFBayesian.control
(bayes_cbf/unicycle_move_to_pose.py:926-964; neither package is installable here, so parity for this step is
"unpinned against GUROBI" and pinned instead against (i) this restatement, step for step, and (ii) scipy's SLSQP on the
same problems in tests/test_socp.py).

    minimise sum_i w_i (y_i - r_i)^2 + q^T y   s.t.   c_k^T y + d_k >= rho ||A_k y + b_k||,  k < K

With w = 0 this is the linear-objective SOCP of the reference's optimizers.py:6-116; the reference-held known answer for it
(tests/test_optimizers.py:6-119, the cvxopt documentation example) pins this restatement and, through it, the CUDA solver
(tests/test_socp.py).
"""
import math

import numpy as np


MU = 10.0            # barrier parameter growth per outer iteration (socp.cu kMu)
CENTER_TOL = 1e-5    # intermediate centering tolerance on decrement^2 / 2 (socp.cu kCenterTol)


def _cones(c, d, A, b, rho, y, s):
    t = c @ y + d + s                      # (K,)
    z = rho * (A @ y + b)                  # (K, pc)
    D = t * t - (z * z).sum(1)
    return t, z, D, bool(np.all(t > 0) and np.all(D > 0))


def _merit(w, r, c, d, A, b, rho, x, phase1, tau, eps1, q=None):
    nv = len(w)
    y = x[:nv]
    t, z, D, ok = _cones(c, d, A, b, rho, y, x[nv] if phase1 else 0.0)
    if not ok:
        return math.inf
    f = float((w * (y - r) ** 2).sum())
    if q is not None:
        f += float(q @ y)
    val = tau * (x[nv] + eps1 * f) if phase1 else tau * f
    return val - float(np.log(D).sum())


def _grad_hess(w, r, c, d, A, b, rho, x, phase1, tau, eps1, qlin=None):
    nv = len(w)
    n = nv + (1 if phase1 else 0)
    y = x[:nv]
    t, z, D, _ = _cones(c, d, A, b, rho, y, x[nv] if phase1 else 0.0)
    g = np.zeros(n)
    H = np.zeros((n, n))
    fs = tau * eps1 if phase1 else tau
    g[:nv] = fs * (2.0 * w * (y - r) + (0.0 if qlin is None else qlin))
    H[np.arange(nv), np.arange(nv)] = 2.0 * fs * w
    if phase1:
        g[nv] = tau
    for k in range(len(d)):
        dt = np.zeros(n)
        dt[:nv] = c[k]
        if phase1:
            dt[nv] = 1.0
        q = t[k] * dt
        q[:nv] -= rho * (A[k].T @ z[k])
        iD = 1.0 / D[k]
        g -= 2.0 * q * iD
        ZZ = np.zeros((n, n))
        ZZ[:nv, :nv] = A[k].T @ A[k]
        H += 4.0 * np.outer(q, q) * iD * iD - 2.0 * (np.outer(dt, dt) - rho * rho * ZZ) * iD
    return g, H


def _center(w, r, c, d, A, b, rho, x, phase1, tau, eps1, max_newton, center_tol, q=None):
    nv = len(w)
    n = nv + (1 if phase1 else 0)
    it = 0
    while it < max_newton:
        if phase1 and x[nv] < 0.0:
            break
        g, H = _grad_hess(w, r, c, d, A, b, rho, x, phase1, tau, eps1, q)
        H = H + np.diag(1e-14 * (1.0 + np.abs(np.diag(H))))
        try:
            L = np.linalg.cholesky(H)
        except np.linalg.LinAlgError:
            break
        dx = -np.linalg.solve(L.T, np.linalg.solve(L, g))
        dec = float(-(g @ dx))
        if not dec > 1e-22:
            break
        f0 = _merit(w, r, c, d, A, b, rho, x, phase1, tau, eps1, q)
        step, moved = 1.0, False
        for _ in range(60):
            xn = x + step * dx
            if _merit(w, r, c, d, A, b, rho, xn, phase1, tau, eps1, q) <= f0 - 0.25 * step * dec:
                moved = True
                break
            step *= 0.5
        if not moved:
            break
        x[:] = xn
        if dec * 0.5 < center_tol:
            it += 1
            break
        it += 1
    return it


def solve(w, r, c, d, A, b, rho, tol=1e-9, q=None):
    """One problem.  w, r, q (nv,), c (K,nv), d (K,), A (K,pc,nv), b (K,pc).  Returns (y, status, newton_steps);
    status 0 = optimal, 1 = infeasible (y = nan)."""
    w, r, c, d, A, b = (np.asarray(v, dtype=np.float64) for v in (w, r, c, d, A, b))
    q = None if q is None else np.asarray(q, dtype=np.float64)
    nv, K = len(w), len(d)
    x = np.zeros(nv + 1)
    x[:nv] = r
    total, st = 0, 0
    t, z, D, ok = _cones(c, d, A, b, rho, x[:nv], 0.0)
    if not ok:
        zn = np.sqrt((z * z).sum(1))
        s0 = max(0.0, float((zn - t).max()))
        scale = max(1.0, float(np.maximum(np.abs(t), zn).max()))
        x[nv] = s0 + 0.1 * scale + 1e-3
        tau, found = 1.0 / scale, False
        for _ in range(60):
            total += _center(w, r, c, d, A, b, rho, x, True, tau, 1e-6, 40, CENTER_TOL, q)
            if x[nv] < 0.0:
                found = True
                break
            if 2.0 * K / tau < tol * scale:
                break
            tau *= MU
        if not found:
            st = 1
    if st == 0:
        y = x[:nv].copy()
        tau = 1.0 / max(1.0, float(w.max()), 0.0 if q is None else float(np.abs(q).max()))
        for _ in range(80):
            last = 2.0 * K / tau < tol
            total += _center(w, r, c, d, A, b, rho, y, False, tau, 0.0, 40, 1e-12 if last else CENTER_TOL, q)
            if last:
                break
            tau *= MU
        return y, 0, total
    return np.full(nv, np.nan), 1, total
