"""MVGPModel — thin Python owner of a `bcbf_model` handle (include/bcbf.h): one fitted MVGP resident on one GPU.

fit  = Gram + jittered Cholesky + L^{-1} + alpha at fixed hyper-parameters (what the reference's
       `_perturbed_cholesky` + the alpha solve do on first use, control_affine_model.py:366-385, 545);
query = posterior mean / covariance of F(x)[1;u] for many states (control_affine_model.py:931-961, 983-1096).
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import Hyper, check


def make_hyper(n, p, lengthscale, outputscale, A, B, C):
    h = Hyper()
    h.n, h.p, h.outputscale = int(n), int(p), float(outputscale)
    ls = np.asarray(lengthscale, dtype=np.float64).reshape(-1)
    A = np.asarray(A, dtype=np.float64).reshape(n, n)
    B = np.asarray(B, dtype=np.float64).reshape(p, p)
    C = np.asarray(C, dtype=np.float64).reshape(p, n)
    for i in range(n):
        h.lengthscale[i] = ls[i]
    for i, v in enumerate(A.reshape(-1)):
        h.A[i] = v
    for i, v in enumerate(B.reshape(-1)):
        h.B[i] = v
    for i, v in enumerate(C.reshape(-1)):
        h.C[i] = v
    return h


class _DevArray:
    """__cuda_array_interface__ view of a raw device pointer owned by the handle."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = dict(shape=tuple(shape), typestr='<f8', data=(int(ptr), False), version=3,
                                             strides=None)


def _hptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


class MVGPModel:
    def __init__(self, device=0):
        self._lib = _lib.load()
        self._h = ctypes.c_void_p()
        self.device = int(device)
        check(self._lib.bcbf_model_create(ctypes.byref(self._h), self.device))
        self.hyper = None
        self.N = 0

    def close(self):
        if self._h:
            self._lib.bcbf_model_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ fit
    def fit(self, hyper, X, U, Xdot, jitter=None, jitter_scale=1e-5):
        """Host numpy inputs (float64).  Raises NotPositiveDefiniteError (RuntimeError) on a bad pivot."""
        X = np.ascontiguousarray(X, dtype=np.float64)
        U = np.ascontiguousarray(U, dtype=np.float64)
        Xdot = np.ascontiguousarray(Xdot, dtype=np.float64)
        jit = None if jitter is None else np.ascontiguousarray(jitter, dtype=np.float64)
        self.hyper = hyper
        self.N = X.shape[0]
        check(self._lib.bcbf_model_fit(self._h, ctypes.byref(hyper), _hptr(X), _hptr(U), _hptr(Xdot), self.N,
                                       _hptr(jit), float(jitter_scale)))
        return self

    def fit_timing_ms(self):
        ms = (ctypes.c_double * 5)()
        check(self._lib.bcbf_model_fit_timing(self._h, ctypes.byref(ms)))
        return dict(gram=ms[0], potrf=ms[1], trtri=ms[2], alpha=ms[3], total=ms[4],
                    oz_split=self._lib.bcbf_model_oz_split_ms(self._h))

    # ------------------------------------------------------------------ which kernel computes B_k
    VAR_PATHS = {'dmma': 0, 'int8': 1}

    def set_var_path(self, path):
        """'dmma': FP64 tensor pipe (post_var_kernel);  'int8': tcgen05 int8 tensor cores with error-free digit
        splitting (oz_var_kernel), FP64-accurate, Npad <= 18432."""
        check(self._lib.bcbf_model_set_var_path(self._h, self.VAR_PATHS[path] if isinstance(path, str) else int(path)))
        return self

    def set_oz_digits(self, digits):
        """Digits per operand of the int8 path: 7 (default, FP64 rounding level) or 6 (opt-in: 21 instead of 28 digit
        products, B_k to ~3e-11 of the prior scale)."""
        check(self._lib.bcbf_model_set_oz_digits(self._h, int(digits)))
        return self

    @property
    def var_path(self):
        return {v: k for k, v in self.VAR_PATHS.items()}[self._lib.bcbf_model_get_var_path(self._h)]

    # ------------------------------------------------------------------ state (multi-GPU broadcast)
    def alloc_state(self, hyper, N):
        self.hyper = hyper
        self.N = int(N)
        check(self._lib.bcbf_model_alloc_state(self._h, ctypes.byref(hyper), self.N))

    def adopt_state(self):
        """Receiving ranks: the buffers of state_tensors() now hold a broadcast fit (bcbf_model_adopt)."""
        check(self._lib.bcbf_model_adopt(self._h))
        return self

    def state_tensors(self):
        """torch views (no copy) of the fitted state, in the order they are broadcast: Linv, L, alpha, G, W, X."""
        N, Npad = ctypes.c_int(), ctypes.c_int()
        ptrs = [ctypes.c_void_p() for _ in range(6)]
        check(self._lib.bcbf_model_state(self._h, ctypes.byref(N), ctypes.byref(Npad), *[ctypes.byref(p) for p in ptrs]))
        L, Linv, alpha, G, W, X = [p.value for p in ptrs]
        n, p = self.hyper.n, self.hyper.p
        Np = Npad.value
        dev = torch.device('cuda', self.device)
        mk = lambda ptr, shape: torch.as_tensor(_DevArray(ptr, shape), device=dev)
        ldy = (n + 1) // 2 * 2
        return dict(Linv=mk(Linv, (Np, Np)), L=mk(L, (Np, Np)), alpha=mk(alpha, (Np, ldy)), G=mk(G, (Np, p)),
                    W=mk(W, (Np, n * p)), X=mk(X, (N.value, n)))

    # ------------------------------------------------------------------ query
    def query(self, Xq, Uq=None, want=('mean', 'svar', 'Mk', 'Bk')):
        """HOST numpy in / out; the host<->device copies happen inside the call (this is the e2e path)."""
        Xq = np.ascontiguousarray(Xq, dtype=np.float64)
        Q = Xq.shape[0]
        n, p = self.hyper.n, self.hyper.p
        Uq = None if Uq is None else np.ascontiguousarray(Uq, dtype=np.float64)
        out = {}
        if 'mean' in want:
            out['mean'] = np.empty((Q, n))
        if 'svar' in want:
            out['svar'] = np.empty(Q)
        if 'Mk' in want:
            out['Mk'] = np.empty((Q, n, p))
        if 'Bk' in want:
            out['Bk'] = np.empty((Q, p, p))
        check(self._lib.bcbf_model_query(self._h, _hptr(Xq), _hptr(Uq), Q, _hptr(out.get('mean')),
                                         _hptr(out.get('svar')), _hptr(out.get('Mk')), _hptr(out.get('Bk'))))
        return out

    def query_into(self, Xq, Uq, mean, svar, Mk, Bk):
        """Host buffers supplied by the caller (e.g. pinned torch tensors' numpy views)."""
        check(self._lib.bcbf_model_query(self._h, _hptr(Xq), _hptr(Uq), Xq.shape[0], _hptr(mean), _hptr(svar),
                                         _hptr(Mk), _hptr(Bk)))

    def query_device(self, Xq, Uq=None, want=('mean', 'svar', 'Mk', 'Bk'), stream=None):
        """CUDA tensors in / out, asynchronous on the current torch stream."""
        assert Xq.is_cuda and Xq.dtype is torch.float64 and Xq.is_contiguous()
        Q = Xq.shape[0]
        n, p = self.hyper.n, self.hyper.p
        f = dict(dtype=torch.float64, device=Xq.device)
        out = {}
        if 'mean' in want:
            out['mean'] = torch.empty(Q, n, **f)
        if 'svar' in want:
            out['svar'] = torch.empty(Q, **f)
        if 'Mk' in want:
            out['Mk'] = torch.empty(Q, n, p, **f)
        if 'Bk' in want:
            out['Bk'] = torch.empty(Q, p, p, **f)
        ptr = lambda t: None if t is None else t.data_ptr()
        s = torch.cuda.current_stream().cuda_stream if stream is None else stream
        check(self._lib.bcbf_model_query_device(self._h, ptr(Xq), ptr(Uq), Q, ptr(out.get('mean')), ptr(out.get('svar')),
                                                ptr(out.get('Mk')), ptr(out.get('Bk')), s))
        return out
