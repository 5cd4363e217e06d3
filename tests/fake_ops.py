"""TEST INFRASTRUCTURE ONLY — torch-CPU stand-ins for `bayesian_cbf_b200.ops` so that the HOST LOGIC of the drop-in API
(shapes, index orders, caching, retry loops, autograd wiring) can be exercised by the `-m "not gpu"` suite against the
reference-generated goldens.  The product never imports this module; on a GPU box the same tests run against the real
CUDA ops (the `cuda` parametrisation of tests/test_host_api.py)."""
import contextlib

import torch

from bayesian_cbf_b200._lib import NotPositiveDefiniteError

BLOCK = 128


def padded(N):
    return (N + BLOCK - 1) // BLOCK * BLOCK


def query_pad(Q):
    return (Q + 191) // 192 * 192


def _k(X1, X2, ls, s):
    d = (X1[:, None, :] - X2[None, :, :]) / ls
    return float(s) * torch.exp(-0.5 * (d * d).sum(-1))


def gram_train(X, UH, B, lengthscale, outputscale, Npad=None):
    N = X.shape[0]
    Npad = padded(N) if Npad is None else Npad
    Kb = torch.eye(Npad, dtype=torch.float64)
    Kb[:N, :N] = _k(X, X, lengthscale, outputscale) * (UH @ B @ UH.T)
    return Kb


def gram_train_lower(X, UH, B, lengthscale, outputscale, Npad=None):
    Kb = gram_train(X, UH, B, lengthscale, outputscale, Npad)
    return torch.tril(Kb) + torch.triu(torch.full_like(Kb, float('nan')), 1)   # storage above the diagonal: "untouched"


def alpha_refine(X, UH, B, lengthscale, outputscale, Linv, Ypad, jitter=None, jitter_scale=0.0, iters=3, store_kb=None):
    import numpy as np
    N = X.shape[0]
    Kb = _k(X, X, lengthscale, outputscale) * (UH @ B @ UH.T)
    if jitter is not None:
        Kb = Kb + jitter_scale * torch.diag(jitter)
    alpha = Linv.T @ (Linv @ Ypad)
    Kl, Yl = Kb.numpy().astype(np.longdouble), Ypad[:N].numpy().astype(np.longdouble)
    for _ in range(iters):
        r = torch.zeros_like(Ypad)
        r[:N] = torch.from_numpy((Yl - Kl @ alpha[:N].numpy().astype(np.longdouble)).astype(np.float64))
        alpha = alpha + Linv.T @ (Linv @ r)
    return alpha


def cross_gram(X, Xq, lengthscale, outputscale, Npad=None, ldks=None):
    N, Q = X.shape[0], Xq.shape[0]
    Npad = padded(N) if Npad is None else Npad
    ldks = query_pad(Q) if ldks is None else ldks
    Ks = torch.zeros(Npad, ldks, dtype=torch.float64)
    Ks[:N, :Q] = _k(X, Xq, lengthscale, outputscale)
    return Ks


def gram_ca(X1, X2, lengthscale, outputscale, UH1=None, UH2=None, B=None, rows_pad=None):
    K = _k(X1, X2, lengthscale, outputscale)
    if UH1 is not None:
        K = K * (UH1 @ B @ UH2.T)
    if rows_pad is not None and rows_pad > K.shape[0]:
        K = torch.cat([K, K.new_zeros(rows_pad - K.shape[0], K.shape[1])])
    return K


def ca_weight(K, UH1, UH2, B):
    return K * (UH1 @ B @ UH2.T)


def rbf_blocks(X1, X2, lengthscale, outputscale, grad=False, hess=False):
    K = _k(X1, X2, lengthscale, outputscale)
    il2 = 1.0 / lengthscale ** 2
    w = (X1[:, None, :] - X2[None, :, :]) * il2
    dK = -w * K[..., None] if grad else None
    d2K = (torch.diag(il2)[None, None] - w[..., :, None] * w[..., None, :]) * K[..., None, None] if hess else None
    return K, dK, d2K


def potrf_(A, N, jitter=None, jitter_scale=1e-5, check_pd=True, info_out=None):
    M = torch.tril(A)           # like bcbf_potrf, only the lower triangle is read
    M = M + torch.tril(M, -1).T
    if jitter is not None:
        M[:N, :N] += jitter_scale * torch.diag(jitter)
    L, info = torch.linalg.cholesky_ex(M)
    if int(info) != 0:
        raise NotPositiveDefiniteError(-3, "linalg.cholesky: the leading minor of order %d is not positive-definite" % int(info))
    A.copy_(L)
    return A, L      # the "dinv" slot carries L for the fake trtri


def check_info(info):
    return None          # the fake potrf_ raises at once


def trtri(L, dinv):
    return torch.linalg.solve_triangular(L, torch.eye(L.shape[0], dtype=torch.float64), upper=False)


def trmm_lower(A, Bm, trans=False, alpha=1.0):
    return alpha * ((A.T if trans else A) @ Bm)


def gemm(A, B, transa=False, transb=False, alpha=1.0, beta=0.0, C=None):
    out = alpha * ((A.T if transa else A) @ (B.T if transb else B))
    return out + beta * C if (C is not None and beta != 0.0) else out


def posterior_blocks(Linv, Kstar, G, W, Bmat, Ct, kss, n, p, Q, want_mean=True, want_cov=True):
    Ks = Kstar[:, :Q]
    Mk = (Ct.reshape(1, n, p) + (Ks.T @ W).reshape(Q, n, p)) if want_mean else None
    Bk = None
    if want_cov:
        V = Linv @ (Ks[:, :, None] * G[:, None, :]).reshape(Ks.shape[0], Q * p)
        V = V.reshape(-1, Q, p)
        Bk = kss * Bmat[None] - torch.einsum('iqa,iqb->qab', V, V)
    return Mk, Bk


def contract_u(Mk, Bk, UHq):
    mean = torch.einsum('qnp,qp->qn', Mk, UHq) if Mk is not None else None
    svar = torch.einsum('qa,qab,qb->q', UHq, Bk, UHq) if Bk is not None else None
    return mean, svar


def socp_factor(Asq, reg=0.0):
    Q, p, _ = Asq.shape
    L, info = torch.linalg.cholesky_ex(Asq)
    if reg > 0:
        bad = info != 0
        if bad.any():
            L2, info2 = torch.linalg.cholesky_ex(Asq + reg * torch.eye(p, dtype=Asq.dtype))
            L = torch.where(bad[:, None, None], L2, L)
            info = torch.where(bad, info2, info)
    Lt = L.transpose(1, 2)
    return Lt[:, :, 1:].contiguous(), Lt[:, :, 0].contiguous(), info.to(torch.int32)


def cbc1_terms(Mk, Bk, A, grad_h, h, gamma, Fbar=None):
    F = Mk if Fbar is None else Mk + Fbar
    row = torch.einsum('qn,qnp->qp', grad_h, F)
    e = row[:, 0] + gamma * h
    bfe = row[:, 1:]
    sA = torch.einsum('qn,nm,qm->q', grad_h, A, grad_h)
    Asq = sA[:, None, None] * Bk
    A_socp, bfb, status = socp_factor(Asq)
    return bfe, e, Asq, A_socp, bfb, status


def gram_train_backward(X, UH, B, lengthscale, outputscale, Pinv, alphaAi, alpha):
    N = X.shape[0]
    with torch.enable_grad():
        ls = lengthscale.detach().clone().requires_grad_(True)
        s = torch.tensor(float(outputscale), dtype=torch.float64, requires_grad=True)
        Bq = B.detach().clone().requires_grad_(True)
        d = (X[:, None, :] - X[None, :, :]) / ls
        Kb = s * torch.exp(-0.5 * (d * d).sum(-1)) * (UH @ Bq @ UH.T)
        Gbar = 0.5 * (alphaAi @ alpha.T - alpha.shape[1] * Pinv[:N, :N])
        g = torch.autograd.grad((Gbar * Kb).sum(), [s, ls, Bq])
    return g[0], g[1], g[2]


@contextlib.contextmanager
def installed(monkeypatch):
    """Route the host API through the CPU stand-ins and lift the CUDA-only guards (host-logic tests only)."""
    import bayesian_cbf_b200.control_affine_model as cam
    import bayesian_cbf_b200.controllers as ctl
    import bayesian_cbf_b200.gp_modules as gm
    from bayesian_cbf_b200 import ops
    me = globals()
    for name in ('padded', 'query_pad', 'gram_train', 'gram_train_lower', 'alpha_refine', 'ca_weight', 'cross_gram', 'gram_ca', 'rbf_blocks', 'potrf_', 'check_info', 'trtri',
                 'trmm_lower', 'gemm', 'posterior_blocks', 'contract_u', 'socp_factor', 'cbc1_terms',
                 'gram_train_backward', 'socp_solve', 'oz_max_npad', 'oz_split_factor', 'posterior_blocks_i8',
                 'oz_gemm_tn'):
        monkeypatch.setattr(ops, name, me[name])
    import bayesian_cbf_b200.mll as mll
    monkeypatch.setattr(mll, '_need_cuda', lambda *t: None)
    monkeypatch.setattr(cam, '_need_cuda', lambda device: None)
    monkeypatch.setattr(gm, '_need_cuda', lambda t, what: None)
    monkeypatch.setattr(ctl, '_compute_device', lambda dev: dev)
    yield


def socp_solve(w, c, d, A, b, rho, r=None, tol=1e-9, q=None):
    from oracle import socp_oracle as S
    Q, K, nv = c.shape
    y = torch.empty(Q, nv, dtype=torch.float64)
    status = torch.empty(Q, dtype=torch.int32)
    iters = torch.empty(Q, dtype=torch.int32)
    for p in range(Q):
        wp = (w[p] if w.ndim == 2 else w).numpy()
        rp = r[p].numpy() if r is not None else [0.0] * nv
        yo, st, it = S.solve(wp, rp, c[p].numpy(), d[p].numpy(), A[p].numpy(), b[p].numpy(), float(rho), tol,
                             q=None if q is None else q[p].numpy())
        y[p] = torch.from_numpy(yo)
        status[p], iters[p] = st, it
    return y, status, iters


# ---- int8 tensor-core path (csrc/ozaki.cu): on the CPU stand-in the "digits" are the inverse factor itself
def oz_max_npad():
    return 18432


def oz_split_factor(Linv, ndigits=7):
    return Linv, torch.ones(Linv.shape[0], dtype=torch.float64)


def posterior_blocks_i8(digits, rowscale, Kstar, G, W, Bmat, Ct, kss, n, p, Q, want_mean=True, want_cov=True, ndigits=7):
    return posterior_blocks(digits, Kstar, G, W, Bmat, Ct, kss, n, p, Q, want_mean=want_mean, want_cov=want_cov)


def oz_gemm_tn(A, B, alpha=1.0, lower=False):
    return alpha * (A.T @ B)
