"""Batched unicycle / Ackermann rollouts under the Bayesian CLF-CBF controller — the callers of the hot path for
BASELINE configs[2] (`unicycle_bayes_cbf_safe_obstacle`, prior-only) and configs[4] (ensembles of
`learning_helps_avoid_getting_stuck` rollouts), restated for R rollouts at once on the GPU
(reference bayes_cbf/unicycle_move_to_pose.py: AckermannDrive :200-292, CLFCartesian :520-612, ObstacleCBF :618-696,
ControllerCLFBayesian :801-998, PiecewiseLinearPlanner planner.py:19-64, sample_generator_trajectory sampling.py:49-75).

Per control step and rollout the reference builds three GP expressions (CLC + one CBC per obstacle), differentiates
them twice with autograd to get the affine / quadratic terms in u, factorises Asq and calls cvxpy + GUROBI.  Here the
terms come from the closed form (`bcbf_cbc1_terms`), the posterior from `MVGPEnsemble` when learning is
on, and the program is solved by the batched barrier kernel (`bcbf_socp_solve`).  The per-rollout state algebra below
(CLF / CBF values and gradients, planner) is elementwise glue on (R,)-shaped CUDA tensors.
"""
import math

import torch

from . import ops


def normalize_radians(theta):
    return (theta + math.pi) % (2 * math.pi) - math.pi


def ackermann_F(X, L):
    """F(x) = [f(x) | g(x)] (R,3,3) of the Ackermann drive: f = 0, g = [[cos th, 0],[sin th, 0],[0, 1/L]]."""
    th = X[:, 2]
    F = X.new_zeros(X.shape[0], 3, 3)
    F[:, 0, 1] = th.cos()
    F[:, 1, 1] = th.sin()
    F[:, 2, 2] = 1.0 / L
    return F


def cartesian2polar(X, Xg):
    dx, dy = Xg[:, 0] - X[:, 0], Xg[:, 1] - X[:, 1]
    rho = torch.sqrt(dx * dx + dy * dy)
    phi = torch.atan2(dy, dx)
    return rho, normalize_radians(X[:, 2] - phi), normalize_radians(Xg[:, 2] - phi)


class CLFCartesian:
    def __init__(self, Kp=(0.9, 1.5, 4.0)):
        self.Kp = [float(k) for k in Kp]

    def clf(self, X, Xg):
        rho, alpha, beta = cartesian2polar(X, Xg)
        return 0.5 * self.Kp[0] * rho ** 2 + self.Kp[1] * (1 - alpha.cos()) + self.Kp[2] * (1 - beta.cos())

    def grad_clf(self, X, Xg):
        dx, dy = Xg[:, 0] - X[:, 0], Xg[:, 1] - X[:, 1]
        rho, alpha, beta = cartesian2polar(X, Xg)
        r2 = rho ** 2
        k0, k1, k2 = self.Kp
        gx = -k0 * dx - k1 * alpha.sin() * dy / r2 - k2 * beta.sin() * dy / r2
        gy = -k0 * dy + k1 * alpha.sin() * dx / r2 + k2 * beta.sin() * dx / r2
        gt = k1 * alpha.sin()
        return torch.stack([gx, gy, gt], dim=1)

    def grad_clf_wrt_goal(self, X, Xg):
        dx, dy = Xg[:, 0] - X[:, 0], Xg[:, 1] - X[:, 1]
        rho, alpha, beta = cartesian2polar(X, Xg)
        r2 = rho ** 2
        k0, k1, k2 = self.Kp
        gx = k0 * dx + k1 * alpha.sin() * dy / r2 + k2 * beta.sin() * dy / r2
        gy = k0 * dy - k1 * alpha.sin() * dx / r2 - k2 * beta.sin() * dx / r2
        gt = k2 * beta.sin()
        return torch.stack([gx, gy, gt], dim=1)


class ObstacleCBF:
    """h(x) = w0 (|xy - c|^2 - r^2) + w1 cos(theta - atan2(xy - c))  (reference :618-696)."""

    def __init__(self, center, radius, term_weights=(0.5, 0.5)):
        self.center = [float(center[0]), float(center[1])]
        self.radius = float(radius)
        self.w = [float(term_weights[0]), float(term_weights[1])]

    def cbf(self, X):
        gx, gy = X[:, 0] - self.center[0], X[:, 1] - self.center[1]
        nrm = torch.sqrt(gx * gx + gy * gy)
        radial = gx * gx + gy * gy - self.radius ** 2
        heading = X[:, 2].cos() * gx / nrm + X[:, 2].sin() * gy / nrm
        return self.w[0] * radial + self.w[1] * heading

    def grad_cbf(self, X):
        gx, gy = X[:, 0] - self.center[0], X[:, 1] - self.center[1]
        r2 = gx * gx + gy * gy
        a = torch.atan2(gy, gx)
        th = X[:, 2]
        s = (a - th).sin()
        g0 = self.w[0] * 2 * gx + self.w[1] * s * gy / r2
        g1 = self.w[0] * 2 * gy - self.w[1] * s * gx / r2
        g2 = -self.w[1] * (th - a).sin()
        return torch.stack([g0, g1, g2], dim=1)


def obstacles_at_mid_from_start_and_goal(x, xg, term_weights=(0.5, 0.5)):
    """Two obstacles either side of the straight line (reference :1562-1570); R90 = [[0,-1],[1,0]]."""
    mx, my = (x[0] + xg[0]) / 2, (x[1] + xg[1]) / 2
    dx, dy = x[0] - xg[0], x[1] - xg[1]
    rx, ry = -dy / 3, dx / 3
    rad = math.hypot(dx, dy) / 4
    return [ObstacleCBF((mx + rx, my + ry), rad, term_weights), ObstacleCBF((mx - rx, my - ry), rad, term_weights)]


class PiecewiseLinearPlanner:
    """planner.py:19-64 (host scalars: every rollout of an ensemble shares start and goal)."""

    def __init__(self, x0, x_goal, numSteps, dt, frac_time_to_reach_goal=0.7):
        self.x0 = [float(v) for v in x0]
        self.xg = [float(v) for v in x_goal]
        self.numSteps, self.dt = int(numSteps), float(dt)
        d = [self.xg[0] - self.x0[0], self.xg[1] - self.x0[1]]
        nrm = math.hypot(*d)
        t2 = min(int(numSteps * frac_time_to_reach_goal), numSteps - 1)
        self.checkpoints = [(t2, [self.xg[0], self.xg[1], d[0] / nrm, d[1] / nrm]),
                            (self.numSteps, [self.xg[0], self.xg[1], math.cos(self.xg[2]), math.sin(self.xg[2])])]

    def _interval(self, t):
        prev_t, prev_x = 0, [self.x0[0], self.x0[1], math.cos(self.x0[2]), math.sin(self.x0[2])]
        for ct, cx in self.checkpoints:
            if t <= ct:
                break
            prev_t, prev_x = ct, cx
        return (ct, cx), (prev_t, prev_x)

    def _target_step(self, t):
        return min(t + max(int(0.1 * self.numSteps), 1), self.numSteps)

    def plan(self, t):
        t = self._target_step(t)
        (ct, cx), (pt, px) = self._interval(t)
        xp = [(c - p) * (t - pt) / (ct - pt) + p for c, p in zip(cx, px)]
        return [xp[0], xp[1], math.atan2(xp[3], xp[2])]

    def dot_plan(self, t):
        t = self._target_step(t)
        (ct, cx), (pt, px) = self._interval(t)
        xd = [(c - p) / ((ct - pt) * self.dt) for c, p in zip(cx, px)]
        den = xd[2] ** 2 + xd[3] ** 2
        return [xd[0], xd[1], (xd[2] - xd[3]) / den if den != 0 else float('nan')]


class BayesCBFController:
    """ControllerCLFBayesian.control (reference :926-964) for R rollouts at once.  `posterior(X) -> (Mk (R,3,3) or None,
    Bk (R,3,3), A (3,3) or (R,3,3))` gives the learned part; None -> the analytic prior of AckermannDrive.fu_func_gp
    (:262-275: B_k = I, A = diag(kernel_diag_A))."""

    def __init__(self, planner, clf, cbfs, cbf_gammas, model_L=1.0, kernel_diag_A=(1.0, 1.0, 1.0), clf_gamma=10.0,
                 cost_weights=(0.33, 0.33, 0.33), max_risk=1e-2, posterior=None):
        from .cbc1 import cbc1_safety_factor
        self.planner, self.clf, self.cbfs, self.cbf_gammas = planner, clf, list(cbfs), [float(g) for g in cbf_gammas]
        self.model_L, self.kdA = float(model_L), [float(v) for v in kernel_diag_A]
        self.clf_gamma, self.cost_weights = float(clf_gamma), [float(w) for w in cost_weights]
        self.rho = cbc1_safety_factor(max_risk)
        self.posterior = posterior
        self._const = {}

    def _constants(self, dev):
        """Device-resident constants, created once per device (also keeps host->device copies out of CUDA graphs)."""
        key = str(dev)
        if key not in self._const:
            f64 = dict(dtype=torch.float64, device=dev)
            self._const[key] = dict(
                A_prior=torch.diag(torch.tensor(self.kdA, **f64)), eye=torch.eye(3, **f64),
                w=torch.tensor([self.cost_weights[2], self.cost_weights[0], self.cost_weights[1]], **f64))
        return self._const[key]

    def constraint_terms(self, X, t, goal=None, dplan=None):
        """Cone terms of the CLC (k = 0) and the CBCs (k >= 1): c (R,K,3), d (R,K), A (R,K,3,3), b (R,K,3) in the
        variables y = [relax, u1, u2].  goal / dplan: optional (3,) device tensors holding planner.plan(t) /
        planner.dot_plan(t) (the CUDA-graph path keeps them in static buffers); default: computed from t."""
        R, dev = X.shape[0], X.device
        f64 = dict(dtype=torch.float64, device=dev)
        goal = (torch.tensor(self.planner.plan(t), **f64) if goal is None else goal).expand(R, 3)
        dplan = (torch.tensor(self.planner.dot_plan(t), **f64) if dplan is None else dplan).expand(R, 3)
        Fbar = ackermann_F(X, self.model_L)
        if self.posterior is None:
            Mk = torch.zeros(R, 3, 3, **f64)
            Bk = self._constants(dev)['eye'].expand(R, 3, 3).contiguous()
            Amat = self._constants(dev)['A_prior']
        else:
            Mk, Bk, Amat = self.posterior(X)
        K = 1 + len(self.cbfs)
        c = torch.zeros(R, K, 3, **f64)
        d = torch.zeros(R, K, **f64)
        A = torch.zeros(R, K, 3, 3, **f64)
        b = torch.zeros(R, K, 3, **f64)
        notpd = torch.zeros(R, dtype=torch.int32, device=dev)
        # CLC, negated (reference :880-899):  -(grad V^T F [1;u] + grad_g V^T xdot_plan + gamma V)
        gV = self.clf.grad_clf(X, goal)
        hval = -((self.clf.grad_clf_wrt_goal(X, goal) * dplan).sum(1) + self.clf_gamma * self.clf.clf(X, goal))
        rows = [(-gV, hval, 1.0)] + [(cbf.grad_cbf(X), cbf.cbf(X), g) for cbf, g in zip(self.cbfs, self.cbf_gammas)]
        for k, (gh, h, gamma) in enumerate(rows):
            if Amat.ndim == 3:   # per-rollout A: fold grad_h^T A grad_h into B_k (the kernel takes one shared A)
                sA = torch.einsum('rn,rnm,rm->r', gh, Amat, gh) / (gh * gh).sum(1).clamp_min(1e-300)
                Bk_k, A_k = Bk * sA.reshape(-1, 1, 1), self._constants(dev)['eye']
            else:
                Bk_k, A_k = Bk, Amat
            bfe, e, _, A_socp, bfb, status = ops.cbc1_terms(Mk.contiguous(), Bk_k.contiguous(), A_k.contiguous(),
                                                            gh.contiguous(), h.contiguous(), gamma, Fbar.contiguous())
            c[:, k, 1:] = bfe
            d[:, k] = e
            A[:, k, :, 1:] = A_socp
            b[:, k] = bfb
            notpd = torch.maximum(notpd, (status != 0).to(torch.int32))
        c[:, 0, 0] = 1.0   # the relaxation enters the CLC only
        # a posterior covariance that is not positive definite (the reference's torch.linalg.cholesky raises there,
        # :861) is reported apart from a genuinely infeasible program: `control` returns status 2 for it
        self.last_notpd = notpd
        return c, d, A, b

    def control(self, X, t, goal=None, dplan=None):
        """u (R,2), relax (R,), status (R,) [0 optimal, 1 infeasible, 2 posterior covariance not positive definite]."""
        c, d, A, b = self.constraint_terms(X, t, goal, dplan)
        w = self._constants(X.device)['w']
        y, status, _ = ops.socp_solve(w, c.contiguous(), d.contiguous(), A.contiguous(), b.contiguous(), self.rho)
        status = torch.where(self.last_notpd != 0, torch.full_like(status, 2), status)
        return y[:, 1:], y[:, 0], status


def rollout(controller, X0, steps, dt, true_L=12.0, on_step=None):
    """sample_generator_trajectory (sampling.py:49-75) for R rollouts: u = controller(x, t); x += F_true(x)[1;u] dt.
    A rollout whose program is infeasible stops there (the reference raises ValueError(problem.status)) and keeps its
    last state.  Returns dict(X (steps+1,R,3), U (steps,R,2), feasible (steps,R) bool, alive (R,) bool)."""
    R = X0.shape[0]
    X = X0.clone()
    alive = torch.ones(R, dtype=torch.bool, device=X.device)
    Xs, Us, Fs = [X.clone()], [], []
    for t in range(steps):
        u, relax, status = controller.control(X, t)
        ok = (status == 0) & alive
        u = torch.where(ok.unsqueeze(1), u, torch.zeros_like(u))
        UH = torch.cat([torch.ones(R, 1, dtype=X.dtype, device=X.device), u], dim=1)
        xdot = torch.einsum('rnp,rp->rn', ackermann_F(X, true_L), UH)
        if on_step is not None:
            on_step(t, X, u, xdot, ok)
        X = torch.where(ok.unsqueeze(1), X + xdot * dt, X)
        alive = ok
        Xs.append(X.clone())
        Us.append(u)
        Fs.append(ok)
    return dict(X=torch.stack(Xs), U=torch.stack(Us), feasible=torch.stack(Fs), alive=alive)


class EnsembleLearner:
    """LearnedShiftInvariantDynamics (reference :295-428) for R rollouts: every rollout records its own (x, u) pairs,
    and every `train_every_n_steps` steps all R per-rollout MVGPs are re-fitted in one batch on the residual
    xdot - F_prior(x)[1;u] over shift-invariant states [0, 0, theta] (:326-330, :340-386), subsampled at random to at most
    `max_train` points (the reference shuffles with numpy's global generator, :374-384; here a seeded torch generator, or
    explicit index arrays via `set_subsample_source` for parity tests).  With `adam_iters > 0` every refit first runs that
    many Adam steps on each rollout's own log marginal likelihood (the reference: 100, :386), all rollouts in the same
    launches (ensemble.fit_ensemble_hyperparameters, rank-one task covariances as in ControlAffineRegressorExactRankOne
    :301); with 0 the given hyper-parameters are held fixed.

    `query_raw_state` (default True, the reference's behaviour): the learned GP is trained on shift-invariant inputs but
    QUERIED at the raw state — `fu_func_gp` (:388-397) hands x itself to `learned_dynamics.fu_func_gp`; only
    f_func / g_func / custom_predict_fullmat wrap their argument.  False queries at [0, 0, theta], which is what the
    training inputs look like (the choice the round-1 throughput runs made)."""

    def __init__(self, R, dt, model_L=12.0, max_train=200, train_every_n_steps=400, lengthscale=(1.0, 1.0, 1.0),
                 outputscale=1.0, A=None, B=None, seed=0, device='cuda', adam_iters=0, lr=0.1, query_raw_state=True):
        self.R, self.dt, self.model_L = R, float(dt), float(model_L)
        self.max_train, self.every = int(max_train), int(train_every_n_steps)
        self.device = torch.device(device)
        f64 = dict(dtype=torch.float64, device=self.device)
        self.ls = torch.tensor(lengthscale, **f64).expand(R, 3).contiguous()
        self.s = torch.full((R,), float(outputscale), **f64)
        self.A = (torch.eye(3, **f64) if A is None else torch.as_tensor(A, **f64)).expand(R, 3, 3).contiguous()
        self.B = (torch.eye(3, **f64) if B is None else torch.as_tensor(B, **f64)).expand(R, 3, 3).contiguous()
        self.C = torch.zeros(R, 3, 3, **f64)
        self.ens = None                     # MVGPEnsemble, created at the first fit (CUDA only)
        self.fitted = False
        self.Xs, self.Us = [], []
        self.gen = torch.Generator().manual_seed(seed)
        self.refits = 0
        self.adam_iters, self.lr = int(adam_iters), float(lr)
        self.hp = None
        self.query_raw_state = bool(query_raw_state)
        self._subsample_source = None
        self._jitter_source = None

    def set_subsample_source(self, index_arrays):
        """Explicit subsample index arrays (one per refit that needs subsampling, consumed in order; the first `max_train`
        entries of each are used, like the reference's shuffled_indices[:max_train])."""
        self._subsample_source = None if index_arrays is None else iter(index_arrays)

    def set_jitter_source(self, draws):
        """Explicit U(0,1) factor-jitter draws, one (R, N) array per Cholesky attempt, instead of the seeded generator."""
        self._jitter_source = None if draws is None else iter(draws)

    @staticmethod
    def shift_invariant(X):
        Z = torch.zeros_like(X)
        Z[..., 2] = X[..., 2]
        return Z

    def record(self, t, X, U, xdot, ok):
        """on_step hook of `rollout` = LearnedShiftInvariantDynamics.train (:340-354): every n recorded steps train on
        what has been recorded so far, then record (x, u)."""
        if len(self.Xs) > 0 and len(self.Xs) % self.every == 0:
            self.fit()
        self.Xs.append(X.clone())
        self.Us.append(U.clone())

    def training_set(self):
        """(Xtr (R,T,3) shift-invariant, Utr (R,T,2), err (R,T,3)) from everything recorded so far (:343-348, :367-386):
        finite-difference Xdot minus the prior model's F(x)[1;u], subsampled to max_train points."""
        Xall = torch.stack(self.Xs, dim=1)                    # (R, T+1, 3)
        Uall = torch.stack(self.Us, dim=1)
        Xdot = (Xall[:, 1:] - Xall[:, :-1]) / self.dt         # finite differences (:348)
        Xtr, Utr = self.shift_invariant(Xall[:, :-1]), Uall[:, :-1]
        T = Xtr.shape[1]
        UH = torch.cat([torch.ones(self.R, T, 1, dtype=torch.float64, device=Xtr.device), Utr], dim=2)
        Fp = ackermann_F(Xtr.reshape(-1, 3), self.model_L).reshape(self.R, T, 3, 3)
        err = Xdot - torch.einsum('rtnp,rtp->rtn', Fp, UH)
        if T > self.max_train:
            if self._subsample_source is not None:
                idx = torch.as_tensor(next(self._subsample_source)).long()[:self.max_train]
            else:
                idx = torch.randperm(T, generator=self.gen)[:self.max_train]
            idx = idx.to(Xtr.device)
            Xtr, Utr, err = Xtr[:, idx], Utr[:, idx], err[:, idx]
        return Xtr.contiguous(), Utr.contiguous(), err.contiguous()

    def fit(self):
        from .ensemble import MVGPEnsemble
        if len(self.Xs) < 2:          # nothing to difference yet ("Nothing to fit", reference :357-359)
            return
        if self.ens is None:
            self.ens = MVGPEnsemble(3, 2, device=self.device)
        Xtr, Utr, err = self.training_set()
        if self.adam_iters > 0:
            from .ensemble import EnsembleHyperParameters, fit_ensemble_hyperparameters
            if self.hp is None:      # parameters persist across refits, like the reference's learned_dynamics object
                self.hp = EnsembleHyperParameters(self.R, 3, 3, rank=1, device=self.device)
            fit_ensemble_hyperparameters(self.hp, Xtr, Utr, err, training_iter=self.adam_iters, lr=self.lr,
                                         generator=self.gen)
            with torch.no_grad():
                ls, s, A, B, C = [t.detach().contiguous() for t in self.hp.constrained()]
            self.ls, self.s, self.A, self.B, self.C = ls, s, A, B, C
        if self._jitter_source is not None:
            draw = lambda t: torch.as_tensor(next(self._jitter_source), dtype=torch.float64).reshape(self.R, Xtr.shape[1])
        else:
            draw = lambda t: torch.rand(self.R, Xtr.shape[1], dtype=torch.float64, generator=self.gen)
        self.ens.fit(Xtr, Utr, err, self.ls, self.s, self.A, self.B, self.C, jitter=draw)
        self.fitted = True
        self.refits += 1

    def posterior(self, X):
        if not self.fitted:
            f64 = dict(dtype=torch.float64, device=self.device)
            Bk = self.B * self.s.reshape(-1, 1, 1)            # prior: k(x,x) B
            return torch.zeros(self.R, 3, 3, **f64), Bk.contiguous(), self.A
        Xq = X if self.query_raw_state else self.shift_invariant(X)
        Mk, Bk = self.ens.posterior(Xq.contiguous())
        return Mk, Bk, self.A


class GraphedRollout:
    """The control step of `rollout` captured once in a CUDA graph and replayed every step: the per-step work is ~150
    tiny launches (CLF / CBF algebra, cone assembly) around three kernels of ours (posterior, CBC terms, SOCP), i.e.
    launch-bound — exactly what a graph removes.  The planner's goal / derivative are scalars that change with t: they
    live in static device buffers refreshed from pinned host memory before each replay.  Semantics identical to
    `rollout` (an infeasible rollout stops and keeps its state)."""

    def __init__(self, controller, X0, dt, true_L=12.0):
        self.ctrl, self.dt, self.true_L = controller, float(dt), float(true_L)
        dev = X0.device
        self.R = X0.shape[0]
        self.X = X0.clone()
        self.alive = torch.ones(self.R, dtype=torch.bool, device=dev)
        self.goal = torch.zeros(3, dtype=torch.float64, device=dev)
        self.dplan = torch.zeros(3, dtype=torch.float64, device=dev)
        self._plans = None      # (T, 2, 3) device table of planner.plan / dot_plan, filled by `run`
        self.graph = None
        self.u = self.xdot = self.ok = None

    def _plan_table(self, steps):
        """planner.plan(t), planner.dot_plan(t) for t < steps as one device table: the per-step refresh of the static
        goal buffers is then a stream-ordered device copy (a re-used pinned staging buffer would race with the host
        running ahead of the GPU)."""
        if self._plans is None or self._plans.shape[0] < steps:
            tab = torch.tensor([[self.ctrl.planner.plan(t), self.ctrl.planner.dot_plan(t)] for t in range(steps)],
                               dtype=torch.float64)
            self._plans = tab.to(self.X.device)

    def _set_plan(self, t):
        self._plan_table(t + 1)
        self.goal.copy_(self._plans[t, 0])
        self.dplan.copy_(self._plans[t, 1])

    def _step_body(self):
        u, relax, status = self.ctrl.control(self.X, 0, self.goal, self.dplan)
        ok = (status == 0) & self.alive
        u = torch.where(ok.unsqueeze(1), u, torch.zeros_like(u))
        UH = torch.cat([torch.ones(self.R, 1, dtype=torch.float64, device=self.X.device), u], dim=1)
        xdot = torch.einsum('rnp,rp->rn', ackermann_F(self.X, self.true_L), UH)
        return u, xdot, ok

    def capture(self):
        self._set_plan(0)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):          # warm-up outside the capture (lazy initialisations, allocator)
            for _ in range(2):
                self._step_body()
        torch.cuda.current_stream().wait_stream(s)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.u, self.xdot, self.ok = self._step_body()
        return self

    def step(self, t, on_step=None):
        self._set_plan(t)
        self.graph.replay()
        u, xdot, ok = self.u, self.xdot, self.ok      # this replay's outputs (on_step may re-capture the graph)
        if on_step is not None:
            on_step(t, self.X, u, xdot, ok)
        # state update outside the graph: the learner may refit between steps, the graph reads self.X in place
        self.X.copy_(torch.where(ok.unsqueeze(1), self.X + xdot * self.dt, self.X))
        self.alive.copy_(ok)
        return u, ok

    def run(self, steps, on_step=None, record=True):
        self._plan_table(steps)
        Xs, Us, Fs = [self.X.clone()], [], []
        for t in range(steps):
            u, ok = self.step(t, on_step)
            if record:
                Xs.append(self.X.clone())
                Us.append(u.clone())
                Fs.append(ok.clone())
        out = dict(alive=self.alive.clone())
        if record:
            out.update(X=torch.stack(Xs), U=torch.stack(Us), feasible=torch.stack(Fs))
        return out
