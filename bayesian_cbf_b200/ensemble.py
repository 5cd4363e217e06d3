"""Ensembles of independent small MVGPs on one GPU — the rollout-ensemble regime of BASELINE configs[4]
(4096 `learning_helps_avoid_getting_stuck` rollouts, N <= 200 training points each, one posterior query per rollout and
control step; reference unicycle_move_to_pose.py:1948-1969, 880-998).

`MVGPEnsemble.fit` factorises all R rollout models in the same launches (batched fused Gram, batched blocked Cholesky
with the per-rollout 10x jitter retry of make_psd, batched triangular inverse, alpha); `posterior` answers one state
per rollout in one HBM-bound launch; `cbc_terms` turns the result into the SOCP constraint terms of each rollout's
control-barrier condition.  Rollouts are independent: multi-GPU = partition the rollouts (`sharding.shard_bounds`), no
exchange.
"""
import torch

from . import _lib, ops
from ._lib import check


def _ptr(t):
    return None if t is None else t.data_ptr()


class MVGPEnsemble:
    def __init__(self, n, m, device='cuda'):
        self.n, self.m, self.p = n, m, m + 1
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise RuntimeError("MVGPEnsemble runs on a CUDA device only (no CPU fallback)")
        self.R = 0

    def _t(self, x, shape):
        t = torch.as_tensor(x, dtype=torch.float64, device=self.device).contiguous()
        assert tuple(t.shape) == tuple(shape), (tuple(t.shape), tuple(shape))
        return t

    def fit(self, X, U, Xdot, lengthscale, outputscale, A, B, C, jitter=None, tries=10, perturb_init=1e-5,
            perturb_scale=10.0):
        """X (R,N,n), U (R,N,m), Xdot (R,N,n); per-rollout hyper-parameters lengthscale (R,n), outputscale (R,),
        A (R,n,n), B (R,p,p), C (R,p,n).  jitter: callable(try_index) -> (R,N) U(0,1) draws (default: torch.rand on the
        CPU generator, the reference's make_psd draw per attempt).  Rollouts whose Cholesky fails are retried with 10x
        the perturbation (only those), up to `tries` times; still failing -> RuntimeError like the reference."""
        lib = _lib.load()
        R, N, n = X.shape
        p = self.p
        self.R, self.N = R, N
        self.Npad = Npad = ops.padded(N)
        dev = self.device
        f64 = dict(dtype=torch.float64, device=dev)
        self.X = self._t(X, (R, N, n))
        U = self._t(U, (R, N, self.m))
        self.UH = torch.cat([torch.ones(R, N, 1, **f64), U], dim=2).contiguous()
        Xdot = self._t(Xdot, (R, N, n))
        self.ls = self._t(lengthscale, (R, n))
        self.s = self._t(outputscale, (R,))
        self.A = self._t(A, (R, n, n))
        self.B = self._t(B, (R, p, p))
        self.C = self._t(C, (R, p, n))
        st = torch.cuda.current_stream().cuda_stream
        L = torch.empty(R, Npad, Npad, **f64)
        dinv = torch.empty(R, lib.bcbf_dinv_elems(Npad), **f64)
        info = torch.zeros(R, dtype=torch.int32, device=dev)
        draw = jitter if jitter is not None else (lambda t: torch.rand(R, N, dtype=torch.float64))
        scale = torch.full((R,), perturb_init, **f64)
        pending = torch.ones(R, dtype=torch.bool, device=dev)
        self.tries_used = torch.zeros(R, dtype=torch.int32, device=dev)
        for t in range(tries):
            idx = torch.nonzero(pending).reshape(-1)
            r = idx.numel()
            # gather the still-failing rollouts into a dense sub-batch
            # NB: every gathered operand is bound to a name until the launch has been issued — a temporary whose
            # data_ptr() is taken inline is freed at once and its block may be handed to the next temporary
            Xs, UHs = self.X[idx].contiguous(), self.UH[idx].contiguous()
            lss, ss, Bs = self.ls[idx].contiguous(), self.s[idx].contiguous(), self.B[idx].contiguous()
            Ls = torch.empty(r, Npad, Npad, **f64)
            check(lib.bcbf_ens_gram(_ptr(Xs), _ptr(UHs), _ptr(lss), _ptr(ss), _ptr(Bs), r, N, n, p, _ptr(Ls), Npad, st))
            eps = draw(t).to(dev)[idx] * (scale[idx] / perturb_init).unsqueeze(1)   # per-rollout factor folded in
            eps = eps.contiguous()
            dsub = torch.empty(r, dinv.shape[1], **f64)
            isub = torch.zeros(r, dtype=torch.int32, device=dev)
            check(lib.bcbf_potrf_batched(_ptr(Ls), Npad, Npad, N, _ptr(eps), perturb_init, _ptr(dsub), _ptr(isub), r, st))
            ok = isub == 0
            good = idx[ok]
            L[good] = Ls[ok]
            dinv[good] = dsub[ok]
            self.tries_used[good] = t + 1
            pending[good] = False
            if not bool(pending.any()):
                break
            scale[idx[~ok]] *= perturb_scale
        if bool(pending.any()):
            bad = torch.nonzero(pending).reshape(-1).tolist()
            raise RuntimeError("linalg.cholesky: rollouts %s are not positive-definite after %d tries" % (bad[:8], tries))
        self.L = L
        self.Linv = torch.empty_like(L)
        scratch = torch.empty_like(L)
        check(lib.bcbf_trtri_batched(_ptr(L), _ptr(dinv), _ptr(self.Linv), _ptr(scratch), Npad, Npad, R, st))
        del scratch
        ldy = (n + 1) // 2 * 2
        self.G = torch.empty(R, Npad, p, **f64)
        Y = torch.empty(R, Npad, ldy, **f64)
        check(lib.bcbf_ens_prep(_ptr(self.UH), _ptr(Xdot), _ptr(self.B), _ptr(self.C), R, N, Npad, n, p, ldy,
                                _ptr(self.G), _ptr(Y), st))
        z = torch.empty_like(Y)
        self.alpha = torch.empty_like(Y)
        check(lib.bcbf_trmm_lower_batched(_ptr(self.Linv), Npad, Npad, 0, _ptr(Y), ldy, ldy, 1.0, 0.0, _ptr(z), ldy, R, st))
        check(lib.bcbf_trmm_lower_batched(_ptr(self.Linv), Npad, Npad, 1, _ptr(z), ldy, ldy, 1.0, 0.0, _ptr(self.alpha),
                                          ldy, R, st))
        self.LinvT = torch.empty_like(self.Linv)      # row k = column k of L^-1: coalesced streaming in `posterior`
        check(lib.bcbf_ens_transpose(_ptr(self.Linv), _ptr(self.LinvT), Npad, R, st))
        self.W = torch.empty(R, Npad, n * p, **f64)
        check(lib.bcbf_ens_w(_ptr(self.alpha), ldy, _ptr(self.G), R, Npad, n, p, _ptr(self.W), st))
        return self

    def posterior(self, xq, out=None):
        """xq (R,n): M_k (R,n,p), B_k (R,p,p) of F(x_r) under rollout r's model (no output jitter)."""
        xq = self._t(xq, (self.R, self.n))
        f64 = dict(dtype=torch.float64, device=self.device)
        Mk, Bk = out if out is not None else (torch.empty(self.R, self.n, self.p, **f64),
                                              torch.empty(self.R, self.p, self.p, **f64))
        check(_lib.load().bcbf_ens_posterior(_ptr(self.LinvT), _ptr(self.X), _ptr(self.G), _ptr(self.W), _ptr(self.ls),
                                             _ptr(self.s), _ptr(self.B), _ptr(self.C), _ptr(xq), self.R, self.N, self.Npad,
                                             self.n, self.p, _ptr(Mk), _ptr(Bk), torch.cuda.current_stream().cuda_stream))
        return Mk, Bk

    def posterior_bytes(self):
        """Algorithmic HBM bytes of one `posterior` launch: each rollout's lower-triangular L^-1 plus X, G, W rows."""
        N, n, p = self.N, self.n, self.p
        return self.R * (4 * N * (N + 1) + 8 * N * (n + p + n * p) + 8 * (n + n * p + p * p))

    def cbc_terms(self, Mk, Bk, grad_h, h, gamma, Fbar=None, A_index=None):
        """Relative-degree-1 CBC terms per rollout (closed form, bcbf_cbc1_terms); uses rollout 0's A unless every
        rollout has its own (then the scaling grad_h^T A grad_h is folded into B_k first)."""
        sA = torch.einsum('rn,rnm,rm->r', grad_h, self.A, grad_h)
        eye = torch.eye(self.n, dtype=torch.float64, device=self.device)
        # fold the per-rollout scale into B_k and pass the identity-normalised quadratic form
        gn = (grad_h * grad_h).sum(1).clamp_min(1e-300)
        Bs = Bk * (sA / gn).reshape(-1, 1, 1)
        return ops.cbc1_terms(Mk.contiguous(), Bs.contiguous(), eye, grad_h.contiguous(), h.contiguous(), gamma,
                              None if Fbar is None else Fbar.contiguous())


# =====================================================================================================================
# Per-rollout hyper-parameter fits: R log marginal likelihoods + gradients in the same launches
# =====================================================================================================================
def _small_cholesky(A):
    """Lower Cholesky factors of R small (n x n, n <= 8) SPD matrices with elementwise torch ops (column by column)."""
    n = A.shape[-1]
    L = torch.zeros_like(A)
    for j in range(n):
        d = A[:, j, j] - (L[:, j, :j] ** 2).sum(-1)
        L[:, j, j] = torch.sqrt(d)        # a non-positive pivot gives NaN, which the caller's NaN check reports (no sync here)
        if j + 1 < n:
            L[:, j + 1:, j] = (A[:, j + 1:, j] - (L[:, j + 1:, :j] * L[:, j:j + 1, :j]).sum(-1)) / L[:, j:j + 1, j]
    return L


def _small_lower_inverse(L):
    """Inverses of R small lower-triangular matrices by forward substitution on the identity."""
    n = L.shape[-1]
    X = torch.zeros_like(L)
    for i in range(n):
        e = torch.zeros(L.shape[0], n, dtype=L.dtype, device=L.device)
        e[:, i] = 1.0
        X[:, i, :] = (e - (L[:, i, :i].unsqueeze(-1) * X[:, :i, :]).sum(1)) / L[:, i, i].unsqueeze(-1)
    return X


class _EnsembleLogMarginal(torch.autograd.Function):
    """log N(vec Xdot_r; vec(UH_r C_r), Kb_r (x) A_r) for r < R (values (R,)) with the closed-form adjoints of mll.py,
    batched: ens Gram -> batched Cholesky (psd-safe escalation per call) -> batched inverse -> batched Kb^-1 = L^-T L^-1
    -> batched fused adjoint reduction (bcbf_ens_gram_backward)."""
    last_jitter = 0.0

    @staticmethod
    def forward(ctx, ls, s, A, B, C, X, UH, Xdot):
        import math
        lib = _lib.load()
        R, N, n = X.shape
        p = UH.shape[2]
        dev = X.device
        f64 = dict(dtype=torch.float64, device=dev)
        Npad = ops.padded(N)
        st = torch.cuda.current_stream().cuda_stream
        ls_d, s_d, B_d = ls.detach().contiguous(), s.detach().contiguous(), B.detach().contiguous()
        Y = (Xdot - UH @ C.detach()).contiguous()                       # (R,N,n)
        ones = torch.ones(R, N, **f64)
        dinv = torch.empty(R, lib.bcbf_dinv_elems(Npad), **f64)
        info = torch.zeros(R, dtype=torch.int32, device=dev)
        # start from the jitter the previous call needed (gpytorch's psd_safe_cholesky escalation, 1e-8, 1e-7, ...):
        # inside one Adam run the matrices barely move, and a failed first attempt costs a Gram + factorisation + sync
        jitter = _EnsembleLogMarginal.last_jitter
        for attempt in range(7):
            L = torch.empty(R, Npad, Npad, **f64)
            check(lib.bcbf_ens_gram(_ptr(X), _ptr(UH), _ptr(ls_d), _ptr(s_d), _ptr(B_d), R, N, n, p, _ptr(L), Npad, st))
            check(lib.bcbf_potrf_batched(_ptr(L), Npad, Npad, N, _ptr(ones) if jitter > 0 else None, jitter, _ptr(dinv),
                                         _ptr(info), R, st))
            if int(info.abs().max()) == 0:
                _EnsembleLogMarginal.last_jitter = jitter
                break
            if attempt == 6:
                raise _lib.NotPositiveDefiniteError(-3, "linalg.cholesky: ensemble log marginal: a Gram matrix is not "
                                                    "positive-definite even with jitter %g" % jitter)
            jitter = 1e-8 if jitter == 0.0 else jitter * 10
        Linv = torch.empty_like(L)
        scratch = torch.empty_like(L)
        check(lib.bcbf_trtri_batched(_ptr(L), _ptr(dinv), _ptr(Linv), _ptr(scratch), Npad, Npad, R, st))
        ldy = (n + 1) // 2 * 2
        Ypad = torch.zeros(R, Npad, ldy, **f64)
        Ypad[:, :N, :n] = Y
        z = torch.empty_like(Ypad)
        al = torch.empty_like(Ypad)
        check(lib.bcbf_trmm_lower_batched(_ptr(Linv), Npad, Npad, 0, _ptr(Ypad), ldy, ldy, 1.0, 0.0, _ptr(z), ldy, R, st))
        check(lib.bcbf_trmm_lower_batched(_ptr(Linv), Npad, Npad, 1, _ptr(z), ldy, ldy, 1.0, 0.0, _ptr(al), ldy, R, st))
        alpha = al[:, :N, :n].contiguous()                               # Kb^-1 Y
        zz = z[:, :N, :n]
        YtA = zz.transpose(1, 2) @ zz                                    # Y^T Kb^-1 Y   (R,n,n)
        La = _small_cholesky(A.detach())                                 # n x n glue, R at once, no host round trip
        Lai = _small_lower_inverse(La)
        Ai = Lai.transpose(1, 2) @ Lai
        quad = torch.einsum('rij,rji->r', Ai, YtA)
        logdetK = 2.0 * torch.log(torch.diagonal(L, dim1=1, dim2=2)[:, :N]).sum(1)
        logdetA = 2.0 * torch.log(torch.diagonal(La, dim1=1, dim2=2)).sum(1)
        value = -0.5 * (quad + n * logdetK + N * logdetA + N * n * math.log(2 * math.pi))
        # Kb^-1 = L^-T L^-1 for every rollout, then the fused adjoint reduction
        Pinv = scratch                                                   # reuse
        sL = Npad * Npad
        check(lib.bcbf_gemm_batched(1, 0, Npad, Npad, Npad, 1.0, _ptr(Linv), Npad, sL, _ptr(Linv), Npad, sL, 0.0,
                                    _ptr(Pinv), Npad, sL, R, st))
        alphaAi = (alpha @ Ai).contiguous()
        nblk = ((N + 63) // 64) ** 2
        egrad = 1 + _lib.MAX_N_DIM + _lib.MAX_P_DIM ** 2
        partial = torch.empty(R * nblk * egrad, **f64)
        out = torch.empty(R, egrad, **f64)
        check(lib.bcbf_ens_gram_backward(_ptr(X), _ptr(UH), _ptr(ls_d), _ptr(s_d), _ptr(B_d), _ptr(Pinv), _ptr(alphaAi),
                                         _ptr(alpha), R, N, Npad, n, p, n, _ptr(partial), partial.numel(), _ptr(out), st))
        g_s = out[:, 0].clone()
        g_ls = out[:, 1:1 + n].clone()
        g_B = out[:, 1 + _lib.MAX_N_DIM:].reshape(R, _lib.MAX_P_DIM, _lib.MAX_P_DIM)[:, :p, :p].clone()
        g_A = 0.5 * (Ai @ YtA @ Ai - N * Ai)
        g_C = UH.transpose(1, 2) @ alphaAi
        ctx.save_for_backward(g_ls, g_s, g_A, g_B, g_C)
        return value

    @staticmethod
    def backward(ctx, g):
        g_ls, g_s, g_A, g_B, g_C = ctx.saved_tensors
        return (g.unsqueeze(1) * g_ls, g * g_s, g.reshape(-1, 1, 1) * g_A, g.reshape(-1, 1, 1) * g_B,
                g.reshape(-1, 1, 1) * g_C, None, None, None)


def ensemble_log_marginal(lengthscale, outputscale, A, B, C, X, UH, Xdot):
    """(R,) log marginal likelihoods of R independent MVGPs; differentiable w.r.t. the five hyper-parameter batches."""
    for t in (lengthscale, outputscale, A, B, C, X, UH, Xdot):
        if not t.is_cuda:
            raise RuntimeError("ensemble_log_marginal runs on a CUDA device only (no CPU fallback)")
    return _EnsembleLogMarginal.apply(lengthscale, outputscale, A, B, C, X.contiguous(), UH.contiguous(), Xdot.contiguous())


class EnsembleHyperParameters(torch.nn.Module):
    """The reference's parameterisation (gpytorch names / constraints), one set per rollout: raw_lengthscale (R,n),
    raw_outputscale (R,), U/V covar_factor (R,n,rank) / (R,p,rank) and raw_var, mean constants (R,p,n).
    rank=1 is `ControlAffineRegressorExactRankOne`, the learned-dynamics class of the rollout recipes
    (unicycle_move_to_pose.py:301)."""

    def __init__(self, R, n, p, rank=1, device='cuda', seed=0):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        rn = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64).to(device)
        P = torch.nn.Parameter
        self.raw_lengthscale = P(torch.zeros(R, n, dtype=torch.float64, device=device))
        self.raw_outputscale = P(torch.zeros(R, dtype=torch.float64, device=device))
        self.U_factor, self.U_raw_var = P(rn(R, n, rank)), P(rn(R, n))
        self.V_factor, self.V_raw_var = P(rn(R, p, rank)), P(rn(R, p))
        self.C = P(torch.zeros(R, p, n, dtype=torch.float64, device=device))

    def constrained(self):
        sp = torch.nn.functional.softplus
        A = self.U_factor @ self.U_factor.transpose(1, 2) + torch.diag_embed(sp(self.U_raw_var))
        B = self.V_factor @ self.V_factor.transpose(1, 2) + torch.diag_embed(sp(self.V_raw_var))
        return sp(self.raw_lengthscale), sp(self.raw_outputscale), A, B, self.C


def fit_ensemble_hyperparameters(hp, X, U, Xdot, training_iter=100, lr=0.1, generator=None):
    """ControlAffineRegressor._fit_with_warnings (control_affine_model.py:274-335) for R rollouts at once: Adam +
    MultiStepLR on -(log marginal)/(N n) summed over rollouts (the rollouts share no parameter, so the sum optimises
    each one exactly as a separate fit would), fresh 1e-6 multiplicative target noise every iteration.  Returns the
    (R,) final losses."""
    R, N, n = X.shape
    UH = torch.cat([torch.ones(R, N, 1, dtype=torch.float64, device=X.device), U], dim=2).contiguous()
    opt = torch.optim.Adam(hp.parameters(), lr=lr)
    sched = torch.optim.lr_scheduler.MultiStepLR(opt, milestones=(torch.tensor([0.3, 0.6, 0.8, 0.90]) * training_iter).tolist())
    loss_r = None
    # the 1e-6 target noise of every iteration comes from a device generator seeded from the caller's (CPU) generator:
    # reproducible from the same seed without a 2.4 MB host draw + copy per iteration
    # (generator=None draws the seed from the global CPU stream: a run under torch.manual_seed stays reproducible)
    seed = int(torch.randint(2 ** 62, (1,), generator=generator))
    dgen = torch.Generator(device=X.device).manual_seed(seed)
    _EnsembleLogMarginal.last_jitter = 0.0
    for _ in range(training_iter):
        opt.zero_grad()
        noise = torch.rand(Xdot.shape, dtype=torch.float64, device=X.device, generator=dgen)
        ls, s, A, B, C = hp.constrained()
        logp = ensemble_log_marginal(ls, s, A, B, C, X, UH, Xdot * (1 + 1e-6 * noise))
        loss_r = -logp / (N * n)
        assert not torch.isnan(loss_r).any()
        loss_r.sum().backward()
        opt.step()
        sched.step()
    return loss_r.detach()
