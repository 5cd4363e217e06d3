"""Log marginal likelihood of the MVGP and its hyper-parameter gradients on the GPU — what `fit` maximises.

The reference hands `MultivariateNormal(M(XU), Kb (x) A)` to gpytorch's `ExactMarginalLogLikelihood`
(control_affine_model.py:309-321) and back-propagates through gpytorch's lazy-tensor algebra.  With the Kronecker
identity (verified equal to the (N n)-dimensional density in oracle/mvgp_oracle.py:mll_dense)

    log N(vec Xdot; vec(UH C), Kb (x) A) = -1/2 [ tr(A^-1 Y^T Kb^-1 Y) + n logdet Kb + N logdet A + N n log 2 pi ]

and the closed-form adjoints

    d/dKb = 1/2 (alpha A^-1 alpha^T - n Kb^-1),   d/dA = 1/2 (A^-1 Y^T alpha A^-1 - N A^-1),   d/dC = UH^T alpha A^-1

(alpha = Kb^-1 Y), one value+gradient evaluation is: fused Gram -> blocked Cholesky -> triangular inverse ->
Kb^-1 = L^-T L^-1 (DMMA GEMM) -> one fused reduction over the N x N adjoint (`bcbf_gram_train_backward`).
The n x n / p x p pieces are host-side glue.  Parity: unpinned in the reference (gpytorch-internal);
tests compare against torch autograd of the dense density.
"""
import math

import torch

from . import ops
from ._lib import NotPositiveDefiniteError


class _MVGPLogMarginal(torch.autograd.Function):
    # set by `capturable(info)` around a CUDA-graph capture: the Cholesky status goes to this tensor instead of being read
    # back (a read synchronises, which a capture forbids); the caller checks it after the replay
    deferred_info = None

    @staticmethod
    def forward(ctx, ls, s, A, B, C, X, UH, Xdot):
        N, n = X.shape
        p = UH.shape[1]
        nout = Xdot.shape[1]
        dev = X.device
        # the outputscale is folded into B (Kb = k1 o (UH (s B) UH^T) with a unit-scale kernel k1): every hyper-parameter
        # stays on the device and the Gram kernels read them there — no device->host read in the iteration
        ls_d = ls.detach().contiguous()
        s_d = s.detach().reshape(())
        B_d = (s_d * B.detach()).contiguous()
        s_f = 1.0
        Y = (Xdot - UH @ C.detach()).contiguous()
        ones = torch.ones(N, dtype=torch.float64, device=dev)
        jitter = 0.0
        L = dinv = None
        for attempt in range(7):          # psd-safe escalation like gpytorch's psd_safe_cholesky (1e-8 * 10^t)
            Kb = ops.gram_train_lower(X, UH, B_d, ls_d, s_f)
            if _MVGPLogMarginal.deferred_info is not None:      # no read-back: a failed factor leaves NaNs and a status
                L, dinv = ops.potrf_(Kb, N, None, 0.0, check_pd=False, info_out=_MVGPLogMarginal.deferred_info)
                break
            try:
                L, dinv = ops.potrf_(Kb, N, ones if jitter > 0 else None, jitter)
                break
            except NotPositiveDefiniteError:
                if attempt == 6:
                    raise
                jitter = 1e-8 if jitter == 0.0 else jitter * 10
        Npad = L.shape[0]
        Linv = ops.trtri(L, dinv)
        Ypad = torch.zeros(Npad, nout, dtype=torch.float64, device=dev)
        Ypad[:N] = Y
        z = ops.trmm_lower(Linv, Ypad)                              # L^-1 Y
        alpha = ops.trmm_lower(Linv, z.contiguous(), trans=True)    # Kb^-1 Y
        alpha = alpha[:N].contiguous()
        # n x n glue on the device (2x2 / 3x3 matrices, elementwise column-by-column factorisation: no library solver, no
        # host round trip)
        from .ensemble import _small_cholesky, _small_lower_inverse
        La = _small_cholesky(A.detach().unsqueeze(0))
        Lai = _small_lower_inverse(La)[0]
        Ai = Lai.transpose(0, 1) @ Lai
        YtA = ops.gemm(z[:N].contiguous(), z[:N].contiguous(), transa=True)   # Y^T Kb^-1 Y  (n x n)
        quad = torch.trace(Ai @ YtA)
        logdetK = 2.0 * torch.log(torch.diagonal(L)[:N]).sum()
        logdetA = 2.0 * torch.log(torch.diagonal(La[0])).sum()
        value = -0.5 * (quad + nout * logdetK + N * logdetA + N * nout * math.log(2 * math.pi))
        # ---- gradients (always needed by fit; computed eagerly) -----------------------------------------------
        if Linv.shape[0] >= 2048 and Linv.shape[0] <= ops.oz_max_npad():
            Pinv = ops.oz_gemm_tn(Linv, Linv, lower=True)               # the same on the int8 tensor cores (exact digits)
        else:
            Pinv = ops.gemm(Linv, Linv, transa=True)                 # Kb^-1 = L^-T L^-1  (Npad, Npad)
        alphaAi = (alpha @ Ai).contiguous()
        g_s1, g_ls, g_Beff = ops.gram_train_backward(X, UH, B_d, ls_d, s_f, Pinv.contiguous(), alphaAi, alpha)
        # chain rule of the folding B_eff = s B, s_kernel = 1:  d/ds = (d/ds_kernel) / s  (Kb is linear in both),
        # d/dB = s d/dB_eff
        g_s = g_s1 / s_d
        g_B = g_Beff * s_d
        g_A = 0.5 * (Ai @ YtA @ Ai - N * Ai)
        g_C = ops.gemm(UH, alphaAi, transa=True)
        ctx.save_for_backward(g_ls.clone(), g_s.clone(), g_A, g_B.clone(), g_C)
        ctx.jitter = jitter
        return value

    @staticmethod
    def backward(ctx, g):
        g_ls, g_s, g_A, g_B, g_C = ctx.saved_tensors
        return g * g_ls, g * g_s, g * g_A, g * g_B, g * g_C, None, None, None


class capturable:
    """Context manager: inside it `mvgp_log_marginal` makes no device->host read (the Cholesky status is written to `info`,
    a 1-element int32 CUDA tensor), so that value + gradients can be captured in a CUDA graph."""

    def __init__(self, info):
        self.info = info

    def __enter__(self):
        _MVGPLogMarginal.deferred_info = self.info
        return self

    def __exit__(self, *a):
        _MVGPLogMarginal.deferred_info = None
        return False


def _need_cuda(*tensors):
    for t in tensors:
        if not t.is_cuda:
            raise RuntimeError("mvgp_log_marginal runs on a CUDA device only (no CPU fallback)")


def mvgp_log_marginal(lengthscale, outputscale, A, B, C, X, UH, Xdot):
    """log N(vec Xdot; vec(UH C), Kb (x) A); float64 CUDA tensors; differentiable w.r.t. the five hyper-parameters."""
    _need_cuda(lengthscale, A, B, C, X, UH, Xdot)
    return _MVGPLogMarginal.apply(lengthscale.reshape(-1), outputscale.reshape(()), A, B, C, X.contiguous(),
                                  UH.contiguous(), Xdot.contiguous())


class _DenseLogMarginal(torch.autograd.Function):
    """log N(r; 0, K) for a dense SPD K (M x M) — the CoGP comparator's marginal likelihood, whose (N n) x (N n)
    covariance has no Kronecker structure (reference control_affine_model.py:1106-1127 hands it to gpytorch's
    ExactMarginalLogLikelihood).  Forward: blocked Cholesky (psd-safe jitter escalation 1e-8 * 10^t like gpytorch),
    triangular inverse, alpha = K^-1 r, all on the CUDA kernels; backward: dK = 1/2 (alpha alpha^T - K^-1), dr = -alpha.
    K itself is assembled by the caller with differentiable glue (HetergeneousCoregionalizationKernel)."""

    @staticmethod
    def forward(ctx, K, r):
        M = K.shape[0]
        Mpad = ops.padded(M)
        dev = K.device
        jitter = 0.0
        ones = torch.ones(M, dtype=torch.float64, device=dev)
        for attempt in range(7):
            buf = torch.eye(Mpad, dtype=torch.float64, device=dev)
            buf[:M, :M] = K.detach()
            try:
                L, dinv = ops.potrf_(buf, M, ones if jitter > 0 else None, jitter)
                break
            except NotPositiveDefiniteError:
                if attempt == 6:
                    raise
                jitter = 1e-8 if jitter == 0.0 else jitter * 10
        Linv = ops.trtri(L, dinv)
        rp = torch.zeros(Mpad, 2, dtype=torch.float64, device=dev)
        rp[:M, 0] = r.detach()
        z = ops.trmm_lower(Linv, rp)
        alpha = ops.trmm_lower(Linv, z.contiguous(), trans=True)[:M, 0].contiguous()
        quad = (z[:M, 0] * z[:M, 0]).sum()
        logdet = 2.0 * torch.log(torch.diagonal(L)[:M]).sum()
        if Mpad >= 2048 and Mpad <= ops.oz_max_npad():
            Kinv = ops.oz_gemm_tn(Linv, Linv, lower=True)
        else:
            Kinv = ops.gemm(Linv, Linv, transa=True)
        ctx.save_for_backward(alpha, Kinv[:M, :M])
        return -0.5 * (quad + logdet + M * math.log(2 * math.pi))

    @staticmethod
    def backward(ctx, g):
        alpha, Kinv = ctx.saved_tensors
        gK = 0.5 * g * (torch.outer(alpha, alpha) - Kinv)
        return gK, -g * alpha


def dense_log_marginal(K, r):
    """log N(r; 0, K), float64 CUDA tensors, differentiable w.r.t. K and r."""
    _need_cuda(K, r)
    return _DenseLogMarginal.apply(K, r)
