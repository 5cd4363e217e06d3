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
