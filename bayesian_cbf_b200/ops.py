"""Tensor-level wrappers over the C ABI (include/bcbf.h).  torch supplies device memory and the stream;
every wrapper refuses non-CUDA / non-float64 tensors instead of falling back."""
import torch

from . import _lib
from ._lib import BLOCK, check


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    # the raw handle of torch's current stream; torch.cuda.current_stream() builds a Stream object per call (~10 us), which
    # at ~12 ops per small-N predict was a tenth of the whole call
    return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device())


def _req(*tensors, contiguous=True):
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("bayesian_cbf_b200 ops need CUDA tensors (no CPU fallback); got device %s" % t.device)
        if t.dtype is not torch.float64:
            raise RuntimeError("bayesian_cbf_b200 ops compute in float64; got %s" % t.dtype)
        if contiguous and not t.is_contiguous():
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


def gram_train_lower(X, UH, B, lengthscale, outputscale, Npad=None):
    """Kb for the factorisation: only the 64x64 tiles on/below the diagonal (+ identity on the pad diagonal) are written;
    the rest of the returned buffer is uninitialised (bcbf_potrf never reads it)."""
    _req(X, UH, B, lengthscale)
    N, n = X.shape
    p = UH.shape[1]
    Npad = padded(N) if Npad is None else Npad
    Kb = torch.empty(Npad, Npad, dtype=torch.float64, device=X.device)
    check(_lib.load().bcbf_gram_train_lower(_ptr(X), _ptr(UH), _ptr(B), _ptr(lengthscale), float(outputscale), N, n, p,
                                            _ptr(Kb), Npad, Npad, _stream()))
    return Kb


def gram_resid(X, UH, B, lengthscale, outputscale, alpha, Y, jitter=None, jitter_scale=0.0):
    """R = Y - (Kb + jitter_scale diag(jitter)) alpha with Kb re-evaluated on the fly and compensated (Dot2) accumulation
    (bcbf_gram_resid).  alpha, Y (N, nc) -> R (N, nc)."""
    _req(X, UH, B, lengthscale, alpha, Y, jitter)
    N, n = X.shape
    p = UH.shape[1]
    nc = alpha.shape[1]
    lib = _lib.load()
    R = torch.zeros(N, nc, dtype=torch.float64, device=X.device)
    ne = lib.bcbf_gram_resid_scratch_elems(N)
    scratch = torch.empty(ne, dtype=torch.float64, device=X.device)
    check(lib.bcbf_gram_resid(_ptr(X), _ptr(UH), _ptr(B), _ptr(lengthscale), float(outputscale), N, n, p, _ptr(jitter),
                              float(jitter_scale), _ptr(alpha), alpha.stride(0), _ptr(Y), Y.stride(0), nc, _ptr(R),
                              R.stride(0), _ptr(scratch), ne, _stream()))
    return R


def gram_resid_stored(Kb, alpha, Y, jitter=None, jitter_scale=0.0):
    """The same residual with Kb (>= N x N, lower triangle as gram_train_lower wrote it) read from memory
    (bcbf_gram_resid_stored): the same result bits as gram_resid."""
    _req(Kb, alpha, Y, jitter)
    N, nc = alpha.shape
    lib = _lib.load()
    R = torch.zeros(N, nc, dtype=torch.float64, device=Kb.device)
    ne = lib.bcbf_gram_resid_scratch_elems(N)
    scratch = torch.empty(ne, dtype=torch.float64, device=Kb.device)
    check(lib.bcbf_gram_resid_stored(_ptr(Kb), Kb.stride(0), N, _ptr(jitter), float(jitter_scale), _ptr(alpha),
                                     alpha.stride(0), _ptr(Y), Y.stride(0), nc, _ptr(R), R.stride(0), _ptr(scratch), ne,
                                     _stream()))
    return R


def alpha_refine(X, UH, B, lengthscale, outputscale, Linv, Ypad, jitter=None, jitter_scale=0.0, iters=3, store_kb=None):
    """alpha (Npad, nc) = (Kb + jitter)^-1 Y: explicit-inverse product + `iters` compensated refinement steps
    (bcbf_alpha_refine[_ws]; reference cholesky_solve, control_affine_model.py:545).  Ypad (Npad, nc), pad rows zero.
    store_kb: keep a copy of Kb for the residuals (default: when Npad >= 2048) — same result bits either way."""
    _req(X, UH, B, lengthscale, Linv, jitter)
    _req(Ypad, contiguous=False)
    N, n = X.shape
    p = UH.shape[1]
    Npad = Linv.shape[0]
    nc = Ypad.shape[1]
    ldy = (nc + 1) // 2 * 2
    Yp = torch.zeros(Npad, ldy, dtype=torch.float64, device=X.device)
    Yp[:, :nc] = Ypad
    alpha = torch.empty(Npad, ldy, dtype=torch.float64, device=X.device)
    lib = _lib.load()
    ne = lib.bcbf_alpha_refine_scratch_elems(N, Npad, ldy)
    scratch = torch.empty(ne, dtype=torch.float64, device=X.device)
    if store_kb is None:
        store_kb = Npad >= 2048
    kb = torch.empty(Npad, Npad, dtype=torch.float64, device=X.device) if (store_kb and iters > 0) else None
    check(lib.bcbf_alpha_refine_ws(_ptr(X), _ptr(UH), _ptr(B), _ptr(lengthscale), float(outputscale), N, n, p, _ptr(jitter),
                                   float(jitter_scale), _ptr(Linv), Linv.stride(0), Npad, _ptr(Yp), ldy, nc, int(iters),
                                   _ptr(alpha), _ptr(scratch), ne, _ptr(kb), Npad, _stream()))
    return alpha[:, :nc]


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


def gram_ca(X1, X2, lengthscale, outputscale, UH1=None, UH2=None, B=None, rows_pad=None):
    """out[i, j] = k(x1_i, x2_j) * (uh1_i^T B uh2_j)  (a, c) — kb*, kb** and, with one-hot uh2, frakB.
    UH1 = UH2 = None gives the plain data kernel.  rows_pad > a appends zero rows (factor-sized operands)."""
    _req(X1, X2, lengthscale, UH1, UH2, B)
    a, n = X1.shape
    c = X2.shape[0]
    p = 0 if UH1 is None else UH1.shape[1]
    ld = (c + 1) // 2 * 2
    if rows_pad is None or rows_pad == a:
        buf = torch.empty(a, ld, dtype=torch.float64, device=X1.device)
    else:
        buf = torch.zeros(rows_pad, ld, dtype=torch.float64, device=X1.device)
    check(_lib.load().bcbf_gram_ca(_ptr(X1), _ptr(UH1), a, _ptr(X2), _ptr(UH2), c, _ptr(B), _ptr(lengthscale),
                                   float(outputscale), n, p, _ptr(buf), ld, _stream()))
    return buf[:, :c]


def ca_weight(K, UH1, UH2, B):
    """K[i,j] * (uh1_i^T B uh2_j) for a data-kernel matrix K (a,c) evaluated by some other module (bcbf_ca_weight)."""
    _req(K, UH1, UH2, B)
    a, c = K.shape
    out = torch.empty(a, c, dtype=torch.float64, device=K.device)
    check(_lib.load().bcbf_ca_weight(_ptr(K), K.stride(0), _ptr(UH1), a, _ptr(UH2), c, _ptr(B), UH1.shape[1], _ptr(out),
                                     c, _stream()))
    return out


def _even_pad(t, rows, cols):
    """Zero-padded contiguous copy of a 2-D tensor with the requested (even) extents; no copy when it fits."""
    if t.shape == (rows, cols) and t.is_contiguous() and t.data_ptr() % 16 == 0:
        return t
    out = torch.zeros(rows, cols, dtype=torch.float64, device=t.device)
    out[:t.shape[0], :t.shape[1]] = t
    return out


def gemm(A, B, transa=False, transb=False, alpha=1.0, beta=0.0, C=None):
    """alpha * op(A) @ op(B) + beta * C on the FP64 tensor-core GEMM (bcbf_gemm).  2-D float64 CUDA tensors of any
    shape: operands are zero-padded to even extents as the kernel requires."""
    _req(A, B, C, contiguous=False)
    M, K = (A.shape[1], A.shape[0]) if transa else A.shape
    K2, N = (B.shape[1], B.shape[0]) if transb else B.shape
    if K != K2:
        raise RuntimeError("gemm: inner dimensions differ (%d vs %d)" % (K, K2))
    e = lambda v: (v + 1) // 2 * 2
    Me, Ne, Ke = e(M), e(N), e(K)
    Ap = _even_pad(A, *((Ke, Me) if transa else (Me, Ke)))
    Bp = _even_pad(B, *((Ne, Ke) if transb else (Ke, Ne)))
    Cp = torch.zeros(Me, Ne, dtype=torch.float64, device=A.device)
    if C is not None and beta != 0.0:
        Cp[:M, :N] = C
    check(_lib.load().bcbf_gemm(int(transa), int(transb), Me, Ne, Ke, float(alpha), _ptr(Ap), Ap.stride(0), _ptr(Bp),
                                Bp.stride(0), float(beta), _ptr(Cp), Cp.stride(0), _stream()))
    return Cp[:M, :N]


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


def potrf_(A, N, jitter=None, jitter_scale=1e-5, check_pd=True, info_out=None):
    """In-place lower Cholesky of the padded matrix A (Npad,Npad) + jitter_scale*diag(jitter).
    Returns (A, dinv).  Raises NotPositiveDefiniteError (a RuntimeError) when a pivot is not positive.
    check_pd=False skips the (synchronising) status read: the caller reads `info_out` (a 1-element int32 CUDA tensor that
    receives 0 or 1 + the index of the failing pivot) when it chooses — what a CUDA-graph capture needs."""
    _req(A, jitter)
    Npad = A.shape[0]
    lib = _lib.load()
    dinv = torch.empty(lib.bcbf_dinv_elems(Npad), dtype=torch.float64, device=A.device)
    info = torch.zeros(1, dtype=torch.int32, device=A.device) if info_out is None else info_out
    check(lib.bcbf_potrf(_ptr(A), A.stride(0), Npad, N, _ptr(jitter), float(jitter_scale), _ptr(dinv), _ptr(info),
                         _stream()))
    if check_pd:
        check(lib.bcbf_check_info(_ptr(info), _stream()))
    return A, dinv


def check_info(info):
    """Read a factorisation status written by potrf_(..., check_pd=False, info_out=info) (synchronises the stream);
    raises NotPositiveDefiniteError when a pivot was not positive."""
    check(_lib.load().bcbf_check_info(_ptr(info), _stream()))


def trtri(L, dinv):
    _req(L, dinv)
    Npad = L.shape[0]
    Linv = torch.empty_like(L)
    scratch = torch.empty_like(L)
    check(_lib.load().bcbf_trtri(_ptr(L), _ptr(dinv), _ptr(Linv), _ptr(scratch), L.stride(0), Npad, _stream()))
    return Linv


def trmm_lower(A, Bm, trans=False, alpha=1.0):
    """C = alpha * op(A) @ Bm with A (Npad,Npad) lower triangular; Bm (Npad, ncols)."""
    _req(A)
    _req(Bm, contiguous=False)
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


def oz_split_factor(Linv, ndigits=7):
    """Digits of L^-1 for the int8 tensor-core covariance path (csrc/ozaki.cu): (digits uint8 blob, rowscale (Npad)).
    ndigits: 7 (default) or 6 (the opt-in 21-product mode; pass the same count to posterior_blocks_i8)."""
    _req(Linv)
    Npad = Linv.shape[0]
    lib = _lib.load()
    digits = torch.empty(lib.bcbf_oz_factor_bytes_d(Npad, int(ndigits)), dtype=torch.int8, device=Linv.device)
    rowscale = torch.empty(Npad, dtype=torch.float64, device=Linv.device)
    check(lib.bcbf_oz_split_factor_d(_ptr(Linv), Linv.stride(0), Npad, _ptr(digits), _ptr(rowscale), int(ndigits),
                                     _stream()))
    return digits, rowscale


def posterior_var_i8(digits, rowscale, Kstar, G, Bmat, kss, p, Q):
    """B_k (Q,p,p) with the N^2 p contraction on the int8 tensor cores; same result as posterior_blocks' B_k."""
    _req(rowscale, Kstar, G, Bmat)
    Npad = rowscale.shape[0]
    Bk = torch.empty(Q, p, p, dtype=torch.float64, device=Kstar.device)
    check(_lib.load().bcbf_posterior_var_i8(_ptr(digits), _ptr(rowscale), Npad, _ptr(Kstar), Kstar.stride(0), _ptr(G),
                                            _ptr(Bmat), float(kss), p, Q, _ptr(Bk), _stream()))
    return Bk


def posterior_blocks_i8(digits, rowscale, Kstar, G, W, Bmat, Ct, kss, n, p, Q, want_mean=True, want_cov=True, ndigits=7):
    """posterior_blocks with the covariance contraction on the int8 tensor cores and the mean fused into the pass over
    K* that finds the column scales (bcbf_posterior_blocks_i8)."""
    _req(rowscale, Kstar, G, W, Bmat, Ct)
    Npad = rowscale.shape[0]
    dev = Kstar.device
    Mk = torch.empty(Q, n, p, dtype=torch.float64, device=dev) if want_mean else None
    Bk = torch.empty(Q, p, p, dtype=torch.float64, device=dev) if want_cov else None
    check(_lib.load().bcbf_posterior_blocks_i8_d(_ptr(digits), _ptr(rowscale), Npad, _ptr(Kstar), Kstar.stride(0), _ptr(G),
                                                 _ptr(W), _ptr(Bmat), _ptr(Ct), float(kss), n, p, Q, _ptr(Mk), _ptr(Bk),
                                                 int(ndigits), _stream()))
    return Mk, Bk


def oz_gemm(A, B, alpha=1.0, tri=0, out=None):
    """C = alpha A B on the int8 tensor cores, FP64-accurate (bcbf_oz_gemm).  A (M,K), B (K,N) row-major views (last
    stride 1); M % 128 == 0, N % 64 == 0, K % 32 == 0.  tri: 0, 1 (A lower triangular), 2 (B lower triangular)."""
    assert A.is_cuda and B.is_cuda and A.dtype is torch.float64 and B.dtype is torch.float64
    assert A.stride(1) == 1 and B.stride(1) == 1
    M, K = A.shape
    N = B.shape[1]
    C = torch.empty(M, N, dtype=torch.float64, device=A.device) if out is None else out
    check(_lib.load().bcbf_oz_gemm(M, N, K, float(alpha), _ptr(A), A.stride(0), _ptr(B), B.stride(0), _ptr(C), C.stride(0),
                                   int(tri), _stream()))
    return C


def oz_gemm_tn(A, B, alpha=1.0, lower=False):
    """C = alpha A^T B on the int8 tensor cores (bcbf_oz_gemm_tn); lower: A and B square lower triangular."""
    assert A.is_cuda and B.is_cuda and A.stride(1) == 1 and B.stride(1) == 1
    K, M = A.shape
    N = B.shape[1]
    C = torch.empty(M, N, dtype=torch.float64, device=A.device)
    check(_lib.load().bcbf_oz_gemm_tn(M, N, K, float(alpha), _ptr(A), A.stride(0), _ptr(B), B.stride(0), _ptr(C), C.stride(0),
                                      1 if lower else 0, _stream()))
    return C


def oz_update_(C, PA, PB, alpha=-1.0, lower=False):
    """C += alpha PA PB^T in place on the int8 tensor cores (bcbf_oz_update); lower: only tiles touching the lower triangle."""
    assert C.is_cuda and PA.stride(1) == 1 and PB.stride(1) == 1 and C.stride(1) == 1
    M, K = PA.shape
    N = PB.shape[0]
    check(_lib.load().bcbf_oz_update(M, N, K, float(alpha), _ptr(PA), PA.stride(0), _ptr(PB), PB.stride(0), _ptr(C),
                                     C.stride(0), 1 if lower else 0, _stream()))
    return C


def oz_max_npad():
    return _lib.load().bcbf_oz_max_npad()


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


def gram_train_backward(X, UH, B, lengthscale, outputscale, Pinv, alphaAi, alpha):
    """sum_ij Gbar_ij dKb_ij/dtheta with Gbar = 1/2 (alphaAi alpha^T - nout P): returns (g_outputscale (),
    g_lengthscale (n,), g_B (p,p)).  Pinv (>=N, ldp) is Kb^-1; alphaAi, alpha (N, nout) contiguous."""
    _req(X, UH, B, lengthscale, Pinv, alphaAi, alpha)
    N, n = X.shape
    p = UH.shape[1]
    nout = alpha.shape[1]
    lib = _lib.load()
    import ctypes
    oe, mn, mp = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    lib.bcbf_gram_backward_layout(ctypes.byref(oe), ctypes.byref(mn), ctypes.byref(mp))
    nblk = ((N + 63) // 64) ** 2
    partial = torch.empty(nblk * oe.value, dtype=torch.float64, device=X.device)
    out = torch.empty(oe.value, dtype=torch.float64, device=X.device)
    check(lib.bcbf_gram_train_backward(_ptr(X), _ptr(UH), _ptr(B), _ptr(lengthscale), float(outputscale), N, n, p,
                                       _ptr(Pinv), Pinv.stride(0), _ptr(alphaAi), _ptr(alpha), alpha.stride(0), nout,
                                       _ptr(partial), partial.numel(), _ptr(out), _stream()))
    gB = out[1 + mn.value:1 + mn.value + mp.value * mp.value].reshape(mp.value, mp.value)[:p, :p]
    return out[0], out[1:1 + n], gB


def socp_factor(Asq, reg=0.0):
    """Batched Asq (Q,p,p) = Ls Ls^T -> (A_socp (Q,p,m), bfb (Q,p), status (Q,))."""
    _req(Asq)
    Q, p, _ = Asq.shape
    f64 = dict(dtype=torch.float64, device=Asq.device)
    A_socp = torch.empty(Q, p, p - 1, **f64)
    bfb = torch.empty(Q, p, **f64)
    status = torch.empty(Q, dtype=torch.int32, device=Asq.device)
    check(_lib.load().bcbf_socp_factor(_ptr(Asq), p, Q, float(reg), _ptr(A_socp), _ptr(bfb), _ptr(status), _stream()))
    return A_socp, bfb, status


def socp_solve(w, c, d, A, b, rho, r=None, tol=1e-9, q=None):
    """Batched tiny SOCPs (bcbf_socp_solve[_lin]): minimise sum_i w_i (y_i - r_i)^2 + q^T y
    s.t. c_k^T y + d_k >= rho ||A_k y + b_k||.
    c (Q,K,nv), d (Q,K), A (Q,K,pc,nv), b (Q,K,pc); w (nv,) or (Q,nv); r, q (Q,nv) or None.
    Returns y (Q,nv) [NaN where infeasible], status (Q,) int32 [0 optimal, 1 infeasible], iters (Q,) int32."""
    _req(w, c, d, A, b, r, q)
    Q, K, nv = c.shape
    pc = A.shape[2]
    dev = c.device
    y = torch.empty(Q, nv, dtype=torch.float64, device=dev)
    status = torch.empty(Q, dtype=torch.int32, device=dev)
    iters = torch.empty(Q, dtype=torch.int32, device=dev)
    check(_lib.load().bcbf_socp_solve_lin(Q, nv, K, pc, float(rho), _ptr(w), int(w.ndim == 2), _ptr(r), _ptr(q), _ptr(c),
                                          _ptr(d), _ptr(A), _ptr(b), float(tol), _ptr(y), _ptr(status), _ptr(iters),
                                          _stream()))
    return y, status, iters
