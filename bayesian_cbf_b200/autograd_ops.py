"""Differentiable wrappers over the CUDA ops, for the small-batch API paths that the reference differentiates
through with autograd (`get_affine_terms`, `get_quadratic_terms`, `GradientGP`, `t_hessian`;
bayes_cbf/misc.py:236-285, bayes_cbf/gp_algebra.py:319-405).

The reference back-propagates through gpytorch's RBF kernel and `torch.cholesky_solve`.  Here:
  * `rbf_kernel`   forward = CUDA `bcbf_rbf_blocks` K; backward / double backward = the closed-form derivative blocks
                    dK, d2K emitted by the same kernel (control_affine_model.py:465-477, tests/test_gp_algebra.py:117-127);
  * `linv_mm`      y = op(L^-1) M on the triangular DMMA GEMM; linear, so backward is the same op transposed;
  * `mm_tn`        A^T B over the N-sized axis on the DMMA GEMM.
Third derivatives are not provided (nothing on the path needs them).
"""
import torch

from . import ops


def _c(t):
    return t.contiguous()


class _RbfD2K(torch.autograd.Function):
    @staticmethod
    def forward(ctx, X1, X2, ls, s):
        _, _, d2K = ops.rbf_blocks(_c(X1.detach()), _c(X2.detach()), _c(ls.detach()), float(s), False, True)
        return d2K

    @staticmethod
    def backward(ctx, g):
        raise RuntimeError("bayesian_cbf_b200: third derivatives of the RBF kernel are not implemented")


class _RbfDK(torch.autograd.Function):
    """dK[i,j,:] = d k(x1_i, x2_j) / d x1_i."""

    @staticmethod
    def forward(ctx, X1, X2, ls, s):
        ctx.save_for_backward(X1, X2, ls, s)
        _, dK, _ = ops.rbf_blocks(_c(X1.detach()), _c(X2.detach()), _c(ls.detach()), float(s), True, False)
        return dK

    @staticmethod
    def backward(ctx, g):
        X1, X2, ls, s = ctx.saved_tensors
        d2K = _RbfD2K.apply(X1, X2, ls, s)                     # d^2 k / dx1 dx2^T  (a,c,n,n)
        # d(dK_d)/dx1_e = -d2K_de ;  d(dK_d)/dx2_e = +d2K_de
        t = torch.einsum('ijd,ijde->ije', g, d2K)
        gX1 = -t.sum(1) if ctx.needs_input_grad[0] else None
        gX2 = t.sum(0) if ctx.needs_input_grad[1] else None
        return gX1, gX2, None, None


class _RbfK(torch.autograd.Function):
    @staticmethod
    def forward(ctx, X1, X2, ls, s):
        ctx.save_for_backward(X1, X2, ls, s)
        K, _, _ = ops.rbf_blocks(_c(X1.detach()), _c(X2.detach()), _c(ls.detach()), float(s), False, False)
        return K

    @staticmethod
    def backward(ctx, g):
        X1, X2, ls, s = ctx.saved_tensors
        dK = _RbfDK.apply(X1, X2, ls, s)                       # (a,c,n), differentiable again
        gX1 = torch.einsum('ij,ijd->id', g, dK) if ctx.needs_input_grad[0] else None
        gX2 = -torch.einsum('ij,ijd->jd', g, dK) if ctx.needs_input_grad[1] else None
        # hyper-parameter gradients flow through the marginal-likelihood op (fit), not through predictions
        return gX1, gX2, None, None


def rbf_kernel(X1, X2, lengthscale, outputscale):
    """k(X1, X2) (a, c); differentiable (twice) w.r.t. X1 and X2.  lengthscale (n,), outputscale 0-d tensor."""
    return _RbfK.apply(X1, X2, lengthscale.reshape(-1), outputscale.reshape(()))


class _LinvMM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, Linv, M, trans):
        ctx.save_for_backward(Linv)
        ctx.trans = trans
        return ops.trmm_lower(Linv, _c(M.detach()), trans=trans)

    @staticmethod
    def backward(ctx, g):
        (Linv,) = ctx.saved_tensors
        return None, _LinvMM.apply(Linv, g, not ctx.trans), None


def linv_mm(Linv, M, trans=False):
    """op(Linv) @ M with Linv (Npad, Npad) lower triangular; differentiable w.r.t. M to any order."""
    return _LinvMM.apply(Linv, M, bool(trans))


class _MMtn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, A, B):
        ctx.save_for_backward(A, B)
        return ops.gemm(_c(A.detach()), _c(B.detach()), transa=True)

    @staticmethod
    def backward(ctx, g):
        A, B = ctx.saved_tensors
        # the remaining products contract over the small (query / output) axis: plain broadcasting glue
        gA = B @ g.transpose(0, 1) if ctx.needs_input_grad[0] else None
        gB = A @ g if ctx.needs_input_grad[1] else None
        return gA, gB


def mm_tn(A, B):
    """A^T @ B contracting the leading (N-sized) axis on the FP64 tensor-core GEMM; differentiable."""
    return _MMtn.apply(A, B)
