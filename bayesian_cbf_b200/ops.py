"""Tensor-level wrappers over the C ABI (include/bcbf.h).  torch supplies device memory and the stream;
every wrapper refuses non-CUDA / non-float64 tensors instead of falling back."""
import torch

from . import _lib
from ._lib import BLOCK, check


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _req(*tensors):
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("bayesian_cbf_b200 ops need CUDA tensors (no CPU fallback); got device %s" % t.device)
        if t.dtype is not torch.float64:
            raise RuntimeError("bayesian_cbf_b200 ops compute in float64; got %s" % t.dtype)
        if not t.is_contiguous():
            raise RuntimeError("bayesian_cbf_b200 ops need contiguous tensors")


def padded(N):
    return (N + BLOCK - 1) // BLOCK * BLOCK


def query_pad(Q):
    return (Q + 191) // 192 * 192


def gram_train(X, UH, B, lengthscale, outputscale, Npad=None):
    """Kb (Npad, Npad): k(X,X) * (UH B UH^T) on [0,N)^2, identity on the pad diagonal."""
    _req(X, UH, B, lengthscale)
    N, n = X.shape
    p = UH.shape[1]
    Npad = padded(N) if Npad is None else Npad
    Kb = torch.empty(Npad, Npad, dtype=torch.float64, device=X.device)
    check(_lib.load().bcbf_gram_train(_ptr(X), _ptr(UH), _ptr(B), _ptr(lengthscale), float(outputscale), N, n, p,
                                      _ptr(Kb), Npad, Npad, _stream()))
    return Kb


def cross_gram(X, Xq, lengthscale, outputscale, Npad=None, ldks=None):
    """Kstar (Npad, ldks): k(X_i, Xq_j); pad rows/cols zero."""
    _req(X, Xq, lengthscale)
    N, n = X.shape
    Q = Xq.shape[0]
    Npad = padded(N) if Npad is None else Npad
    ldks = query_pad(Q) if ldks is None else ldks
    Ks = torch.empty(Npad, ldks, dtype=torch.float64, device=X.device)
    check(_lib.load().bcbf_cross_gram(_ptr(X), _ptr(Xq), _ptr(lengthscale), float(outputscale), N, Q, n, _ptr(Ks),
                                      ldks, Npad, _stream()))
    return Ks


def rbf_blocks(X1, X2, lengthscale, outputscale, grad=False, hess=False):
    _req(X1, X2, lengthscale)
    a, n = X1.shape
    c = X2.shape[0]
    K = torch.empty(a, c, dtype=torch.float64, device=X1.device)
    dK = torch.empty(a, c, n, dtype=torch.float64, device=X1.device) if grad else None
    d2K = torch.empty(a, c, n, n, dtype=torch.float64, device=X1.device) if hess else None
    check(_lib.load().bcbf_rbf_blocks(_ptr(X1), _ptr(X2), _ptr(lengthscale), float(outputscale), a, c, n, _ptr(K),
                                      _ptr(dK), _ptr(d2K), _stream()))
    return K, dK, d2K


def potrf_(A, N, jitter=None, jitter_scale=1e-5, check_pd=True):
    """In-place lower Cholesky of the padded matrix A (Npad,Npad) + jitter_scale*diag(jitter).
    Returns (A, dinv).  Raises NotPositiveDefiniteError (a RuntimeError) when a pivot is not positive."""
    _req(A, jitter)
    Npad = A.shape[0]
    lib = _lib.load()
    dinv = torch.empty(lib.bcbf_dinv_elems(Npad), dtype=torch.float64, device=A.device)
    info = torch.zeros(1, dtype=torch.int32, device=A.device)
    check(lib.bcbf_potrf(_ptr(A), A.stride(0), Npad, N, _ptr(jitter), float(jitter_scale), _ptr(dinv), _ptr(info),
                         _stream()))
    if check_pd:
        check(lib.bcbf_check_info(_ptr(info), _stream()))
    return A, dinv


def trtri(L, dinv):
    _req(L, dinv)
    Npad = L.shape[0]
    Linv = torch.empty_like(L)
    scratch = torch.empty_like(L)
    check(_lib.load().bcbf_trtri(_ptr(L), _ptr(dinv), _ptr(Linv), _ptr(scratch), L.stride(0), Npad, _stream()))
    return Linv


def trmm_lower(A, Bm, trans=False, alpha=1.0):
    """C = alpha * op(A) @ Bm with A (Npad,Npad) lower triangular; Bm (Npad, ncols)."""
    _req(A, Bm)
    Npad = A.shape[0]
    ncols = Bm.shape[1]
    ld = (ncols + 1) // 2 * 2
    if ld != ncols or Bm.stride(0) != ncols:
        Bp = torch.zeros(Npad, ld, dtype=torch.float64, device=A.device)
        Bp[:, :ncols] = Bm
    else:
        Bp = Bm
    C = torch.empty(Npad, ld, dtype=torch.float64, device=A.device)
    check(_lib.load().bcbf_trmm_lower(_ptr(A), A.stride(0), Npad, int(bool(trans)), _ptr(Bp), ld, ld, float(alpha),
                                      0.0, _ptr(C), ld, _stream()))
    return C[:, :ncols]


def posterior_blocks(Linv, Kstar, G, W, Bmat, Ct, kss, n, p, Q, want_mean=True, want_cov=True):
    _req(Linv, Kstar, G, W, Bmat, Ct)
    Npad = Linv.shape[0]
    dev = Linv.device
    Mk = torch.empty(Q, n, p, dtype=torch.float64, device=dev) if want_mean else None
    Bk = torch.empty(Q, p, p, dtype=torch.float64, device=dev) if want_cov else None
    check(_lib.load().bcbf_posterior_blocks(_ptr(Linv), Linv.stride(0), Npad, _ptr(Kstar), Kstar.stride(0), _ptr(G),
                                            _ptr(W), _ptr(Bmat), _ptr(Ct), float(kss), n, p, Q, _ptr(Mk), _ptr(Bk),
                                            _stream()))
    return Mk, Bk


def posterior_fu_var(Linv, Kstar, G, Bmat, UHq, kss, n, p):
    _req(Linv, Kstar, G, Bmat, UHq)
    Npad = Linv.shape[0]
    Q = UHq.shape[0]
    svar = torch.empty(Q, dtype=torch.float64, device=Linv.device)
    check(_lib.load().bcbf_posterior_fu(_ptr(Linv), Linv.stride(0), Npad, _ptr(Kstar), Kstar.stride(0), _ptr(G), None,
                                        _ptr(Bmat), None, _ptr(UHq), float(kss), n, p, Q, None, _ptr(svar), _stream()))
    return svar


def contract_u(Mk, Bk, UHq):
    _req(Mk, Bk, UHq)
    Q, p = UHq.shape
    n = Mk.shape[1] if Mk is not None else 1
    mean = torch.empty(Q, n, dtype=torch.float64, device=UHq.device) if Mk is not None else None
    svar = torch.empty(Q, dtype=torch.float64, device=UHq.device) if Bk is not None else None
    check(_lib.load().bcbf_contract_u(_ptr(Mk), _ptr(Bk), _ptr(UHq), n, p, Q, _ptr(mean), _ptr(svar), _stream()))
    return mean, svar


def cbc1_terms(Mk, Bk, A, grad_h, h, gamma, Fbar=None):
    _req(Mk, Bk, A, grad_h, h, Fbar)
    Q, n, p = Mk.shape
    m = p - 1
    dev = Mk.device
    f64 = dict(dtype=torch.float64, device=dev)
    bfe = torch.empty(Q, m, **f64)
    e = torch.empty(Q, **f64)
    Asq = torch.empty(Q, p, p, **f64)
    A_socp = torch.empty(Q, p, m, **f64)
    bfb = torch.empty(Q, p, **f64)
    status = torch.empty(Q, dtype=torch.int32, device=dev)
    check(_lib.load().bcbf_cbc1_terms(_ptr(Mk), _ptr(Bk), _ptr(A), _ptr(grad_h), _ptr(h), _ptr(Fbar), float(gamma), n,
                                      p, Q, _ptr(bfe), _ptr(e), _ptr(Asq), _ptr(A_socp), _ptr(bfb), _ptr(status),
                                      _stream()))
    return bfe, e, Asq, A_socp, bfb, status
