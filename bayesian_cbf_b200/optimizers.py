"""The reference's optimiser front door (bayes_cbf/optimizers.py:1-132) — same function names and call shapes — over the
batched CUDA cone solver (bcbf_socp_solve_lin, csrc/socp.cu) instead of cvxopt / cvxpy + GUROBI:

    optimizer_socp_cvxopt(u0, linear_objective, socp_constraints)          min c^T u  s.t.  bfc_k^T u + d_k >= |A_k u + bfb_k|
    optimizer_socp_cvxpy(u0, linear_objective, socp_constraints, solver)   the same program
    optimizer_qp_cvxpy(u0, (A, bfb), linear_constraints, solver)           min |A u + bfb|^2  s.t.  0 <= bfc_k^T u + d_k
    convert_socp_to_cvxopt_format(c, socp_constraints)                     (c, [G_k], [h_k]) with G = [-bfc^T; -A], h = [d; bfb]

`socp_constraints` is the reference's list of (name, (A, bfb, bfc, d)).  One problem per call here; the rollout ensembles
batch thousands of them through `ops.socp_solve` directly.  No CPU path: the program is solved on `device` (default cuda).
Limits of the kernel: at most 4 variables, 4 cones, 4 rows per cone.
"""
import numpy as np
import torch

from . import ops


class InfeasibleProblemError(ValueError):
    pass


def convert_socp_to_cvxopt_format(c, socp_constraints):
    """cvxopt's conic form of  d_k + bfc_k^T u >= |A_k u + bfb_k|:  h_k - G_k u in the second-order cone, first component
    the bound (reference :6-42)."""
    m = np.asarray(c).shape[-1]
    Gqs, hqs = [], []
    for _, (A, bfb, bfc, d) in socp_constraints:
        A = np.asarray(A, dtype=np.float64).reshape(-1, m)
        Gqs.append(np.vstack([-np.asarray(bfc, dtype=np.float64).reshape(1, m), -A]))
        hqs.append(np.concatenate([np.asarray(d, dtype=np.float64).reshape(1),
                                   np.asarray(bfb, dtype=np.float64).reshape(-1)]).reshape(-1, 1))
    return c, Gqs, hqs


def _stack_cones(nv, socp_constraints):
    cons = [(np.asarray(A, dtype=np.float64).reshape(-1, nv), np.asarray(bfb, dtype=np.float64).reshape(-1),
             np.asarray(bfc, dtype=np.float64).reshape(nv), float(np.asarray(d).reshape(())))
            for _, (A, bfb, bfc, d) in socp_constraints]
    K = len(cons)
    pc = max(1, max(A.shape[0] for A, _, _, _ in cons))
    As, bs = np.zeros((1, K, pc, nv)), np.zeros((1, K, pc))
    cs, ds = np.zeros((1, K, nv)), np.zeros((1, K))
    for k, (A, bfb, bfc, d) in enumerate(cons):         # cones with fewer rows: zero rows add nothing to the norm
        As[0, k, :A.shape[0]], bs[0, k, :A.shape[0]], cs[0, k], ds[0, k] = A, bfb, bfc, d
    return As, bs, cs, ds


def _solve(u0, w, r, q, socp_constraints, device, tol):
    u0 = np.asarray(u0)
    nv = u0.shape[-1]
    As, bs, cs, ds = _stack_cones(nv, socp_constraints)
    t = lambda a: None if a is None else torch.as_tensor(np.asarray(a, dtype=np.float64)).to(device).contiguous()
    y, status, _ = ops.socp_solve(t(np.asarray(w, dtype=np.float64).reshape(nv)), t(cs), t(ds), t(As), t(bs), 1.0,
                                  r=t(None if r is None else np.asarray(r).reshape(1, nv)),
                                  q=t(None if q is None else np.asarray(q).reshape(1, nv)), tol=tol)
    if int(status[0]) != 0:
        raise InfeasibleProblemError("Infeasible problem: primal infeasible")
    return y[0].cpu().numpy().astype(u0.dtype if u0.dtype.kind == 'f' else np.float64).reshape(-1)


def optimizer_socp_cvxopt(u0, linear_objective, socp_constraints, device='cuda', tol=1e-9):
    """min c^T u s.t. the cones (reference :44-88); raises InfeasibleProblemError like the reference."""
    return _solve(u0, np.zeros(np.asarray(u0).shape[-1]), None, linear_objective, list(socp_constraints), device, tol)


def optimizer_socp_cvxpy(u0, linear_objective, socp_constraints, solver='CUDA', device='cuda', tol=1e-9):
    """Same program as optimizer_socp_cvxopt (reference :91-102; `solver` is accepted and ignored)."""
    return optimizer_socp_cvxopt(u0, linear_objective, socp_constraints, device=device, tol=tol)


def optimizer_qp_cvxpy(u0, quadratic_objective, linear_constraints, solver='CUDA', device='cuda', tol=1e-9):
    """min |A u + bfb|^2 s.t. 0 <= bfc^T u + d (reference :105-116).  A must be diagonal with non-zero entries where it
    acts (the reference's QPController builds it from square roots of weights, controllers.py:640-646): the kernel's
    objective is sum_i w_i (u_i - r_i)^2."""
    A, bfb = (np.asarray(v, dtype=np.float64) for v in quadratic_objective)
    nv = np.asarray(u0).shape[-1]
    if A.shape != (nv, nv) or np.abs(A - np.diag(np.diag(A))).max() > 0:
        raise ValueError("optimizer_qp_cvxpy: only diagonal quadratic objectives are supported by the CUDA solver")
    a = np.diag(A)
    w = a * a
    r = np.where(a != 0, -bfb / np.where(a != 0, a, 1.0), 0.0)
    cones = [(name, (np.zeros((1, nv)), np.zeros(1), bfc, d)) for name, (bfc, d) in linear_constraints]
    return _solve(u0, w, r, None, cones, device, tol)
