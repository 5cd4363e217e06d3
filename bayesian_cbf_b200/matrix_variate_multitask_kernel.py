"""MVGP kernel classes with the reference's names and call signatures
(bayes_cbf/matrix_variate_multitask_kernel.py:18-49, 99-204), evaluated densely on the GPU.

Covariance of the heterogeneous observation model (train rows, mask 1, observe F(x)[1;u]: n outputs; test rows,
mask 0, observe F(x): p*n outputs), rows sorted train-first, output index order (point, q in p, r in n), r fastest:

    [ (K11 o UH1 B UH2^T) (x) A          ((K12 (x) 1_p^T) o (UH1 B)) (x) A ]
    [            (.)^T                       (K22 (x) B) (x) A             ]

Every block is one launch of the fused control-affine Gram kernel (`bcbf_gram_ca`) followed by the Kronecker
expansion with A.
"""
import operator
from functools import reduce

import torch

from . import ops
from .gp_modules import Evaluated, IndexKernel, Kernel
from .misc import torch_kron


def prod(L):
    return reduce(operator.mul, L, 1)


class MatrixVariateIndexKernel(Kernel):
    """covar_matrix = V (x) U for U (n,n) row covariance and V (p,p) column covariance (reference :18-49)."""

    def __init__(self, U: IndexKernel, V: IndexKernel):
        super().__init__()
        self.U = U
        self.V = V
        self.matshape = (self.U.raw_var.shape[-1], self.V.raw_var.shape[-1])

    @property
    def covar_matrix(self):
        return Evaluated(torch_kron(self.V.covar_matrix.evaluate(), self.U.covar_matrix.evaluate(), batch_dims=0))

    def forward(self, i1, i2, **params):
        assert i1.dtype in (torch.int64, torch.int32) and i2.dtype in (torch.int64, torch.int32)
        C = self.covar_matrix.evaluate()
        return C[i1.reshape(-1)][:, i2.reshape(-1)]


class MatrixVariateKernel(Kernel):
    @property
    def num_tasks(self):
        return prod(self.task_covar_module.matshape)

    def __init__(self, task_covar_module, data_covar_module, decoder, **kwargs):
        super().__init__()
        self.task_covar_module = task_covar_module
        self.data_covar_module = data_covar_module
        self.decoder = decoder


def _train_end(M1s):
    idxs = torch.nonzero(M1s - torch.ones_like(M1s))
    return int(torch.min(idxs).item()) if idxs.numel() else M1s.size(-1)


def _dense(t):
    """Whatever a (plug-in) kernel module hands back -> dense torch tensor (gpytorch lazy results expose .evaluate())."""
    if hasattr(t, 'evaluate'):
        t = t.evaluate()
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(t)
    return t


class HetergeneousMatrixVariateKernel(MatrixVariateKernel):
    """Covariance of the heterogeneous observation model (reference :99-204).  The class evaluates the modules it is
    given, like the reference: `data_covar_module.forward(X1, X2)` for the data kernel, `task_covar_module.U / .V
    .covar_matrix` for A and B.  When the data kernel is the stock `ScaleKernel(RBFKernel)` / `RBFKernel` of
    `gp_modules`, every block is ONE launch of the fused control-affine Gram kernel (bcbf_gram_ca) instead of a dense
    K followed by the weighting; any other module takes the plug-in path: its dense K, then bcbf_ca_weight."""

    def num_outputs_per_input(self, mxu1, mxu2):
        M1, X1, _ = self.decoder.decode(mxu1)
        M1s = M1[..., 0]
        end = _train_end(M1s)
        train_size = end * X1.shape[-1]
        test_size = (M1s.size(-1) - end) * prod(self.task_covar_module.matshape)
        return (train_size + test_size) / M1s.size(-1)

    # ---- which evaluation path ------------------------------------------------------------------------------------
    def _stock_rbf(self):
        """(rbf module, outputscale float) if data_covar_module is the stock RBF-ARD (x scale) kernel, else None."""
        from .gp_modules import RBFKernel, ScaleKernel
        dk = self.data_covar_module
        if type(dk) is ScaleKernel and type(dk.base_kernel) is RBFKernel:
            return dk.base_kernel, float(dk.outputscale.detach())
        if type(dk) is RBFKernel:
            return dk, 1.0
        return None

    def _AB(self):
        return _dense(self.task_covar_module.U.covar_matrix), _dense(self.task_covar_module.V.covar_matrix)

    def _hyper(self, ref):
        rbf, s = self._stock_rbf()
        n = ref.shape[-1]
        ls = rbf.lengthscale.reshape(-1).double().expand(n).contiguous().detach()
        A, B = self._AB()
        return ls, s, A, B

    @staticmethod
    def _onehot_cols(X, p):
        """(X repeated p times point-major, identity rows) so that gram_ca yields frakB columns (point, q)."""
        Xr = X.repeat_interleave(p, dim=0).contiguous()
        E = torch.eye(p, dtype=X.dtype, device=X.device).repeat(X.shape[0], 1).contiguous()
        return Xr, E

    # ---- the three block types (reference kernel1 / kernel2 / correlation_kernel_12, :112-134) -------------------
    # Fused forms (stock data kernel): operands are the points themselves.
    def kernel1(self, X1, UH1, X2, UH2):
        ls, s, A, B = self._hyper(X1)
        Kb = ops.gram_ca(X1, X2, ls, s, UH1, UH2, B.double().contiguous())
        return torch_kron(Kb.to(A.dtype), A, batch_dims=0)

    def kernel2(self, X1, X2):
        ls, s, A, B = self._hyper(X1)
        K = ops.gram_ca(X1, X2, ls, s)
        return torch_kron(torch_kron(K.to(A.dtype), B, batch_dims=0), A, batch_dims=0)

    def correlation_kernel_12(self, X1, UH1, X2):
        ls, s, A, B = self._hyper(X1)
        p = B.shape[0]
        X2r, E = self._onehot_cols(X2, p)
        K12 = ops.gram_ca(X1, X2r, ls, s, UH1, E, B.double().contiguous())
        return torch_kron(K12.to(A.dtype), A, batch_dims=0)

    # Plug-in forms (any data kernel): operands are blocks of the dense K it returned.
    def _kernel1_dense(self, K11, UH1, UH2, A, B):
        Kb = ops.ca_weight(K11.double().contiguous(), UH1, UH2, B.double().contiguous())        # H1 (K (x) B) H2^T
        return torch_kron(Kb.to(A.dtype), A, batch_dims=0)

    def _kernel2_dense(self, K22, A, B):
        return torch_kron(torch_kron(K22.to(A.dtype), B, batch_dims=0), A, batch_dims=0)

    def _correlation_kernel_12_dense(self, K12, UH1, A, B):
        p = B.shape[0]
        E = torch.eye(p, dtype=torch.float64, device=K12.device).repeat(K12.shape[1], 1).contiguous()
        K12r = K12.double().repeat_interleave(p, dim=1).contiguous()                             # columns (point, q)
        return torch_kron(ops.ca_weight(K12r, UH1, E, B.double().contiguous()).to(A.dtype), A, batch_dims=0)   # H1 (K12 (x) B)

    def mask_dependent_covar(self, M1s, U1, M2s, U2, X1, X2, covar_xx=None):
        e1, e2 = _train_end(M1s), _train_end(M2s)
        assert (M1s[e1:] == 0).all() and (M2s[e2:] == 0).all(), "rows must be sorted train-first"
        d = lambda t: t.double().contiguous()
        UH1a, UH2a = d(U1[:e1]), d(U2[:e2])
        n11 = bool(e1 and e2)
        n22 = bool((X1.shape[0] - e1) and (X2.shape[0] - e2))
        if covar_xx is None:            # stock data kernel: fused blocks
            X1a, X1b, X2a, X2b = d(X1[:e1]), d(X1[e1:]), d(X2[:e2]), d(X2[e2:])
            k11 = self.kernel1(X1a, UH1a, X2a, UH2a) if n11 else None
            k22 = self.kernel2(X1b, X2b) if n22 else None
            if n11 and n22:
                k12 = self.correlation_kernel_12(X1a, UH1a, X2b)
                k21 = self.correlation_kernel_12(X2a, UH2a, X1b).transpose(0, 1)
        else:                           # plug-in data kernel: blocks of its dense matrix
            A, B = self._AB()
            k11 = self._kernel1_dense(covar_xx[:e1, :e2], UH1a, UH2a, A, B) if n11 else None
            k22 = self._kernel2_dense(covar_xx[e1:, e2:], A, B) if n22 else None
            if n11 and n22:
                k12 = self._correlation_kernel_12_dense(covar_xx[:e1, e2:], UH1a, A, B)
                k21 = self._correlation_kernel_12_dense(covar_xx[e1:, :e2].transpose(0, 1), UH2a, A, B).transpose(0, 1)
        if n11 and n22:
            return torch.cat([torch.cat([k11, k12], dim=1), torch.cat([k21, k22], dim=1)], dim=0)
        return k22 if k11 is None else k11

    def forward(self, mxu1, mxu2, diag=False, last_dim_is_batch=False, **params):
        assert not torch.isnan(mxu1).any() and not torch.isnan(mxu2).any()
        if last_dim_is_batch:
            raise RuntimeError("HetergeneousMatrixVariateKernel does not accept the last dimension to be treated as a batch dimension.")
        M1, X1, U1 = self.decoder.decode(mxu1)
        M2, X2, U2 = self.decoder.decode(mxu2)
        covar_xx = None
        if self._stock_rbf() is None:
            covar_xx = _dense(self.data_covar_module.forward(X1, X2, **params)).to(mxu1.device)
            for name, value in self.data_covar_module.named_parameters():
                assert not torch.isnan(value).any()
        res = self.mask_dependent_covar(M1[..., 0], U1, M2[..., 0], U2, X1, X2, covar_xx).to(mxu1.dtype)
        return torch.diagonal(res) if diag else res


class HetergeneousCoregionalizationKernel(MatrixVariateKernel):
    """The CoGP comparator's covariance (reference :207-316): vec F(x) ~ GP with ONE (p n) x (p n) coregionalisation matrix
    Sigma = task_covar_module.covar_matrix instead of the Kronecker pair (A, B); rows sorted train-first as above:

        [ (H1 (x) I_n)(K11 (x) Sigma)(H2^T (x) I_n)      (H1 (x) I_n)(K12 (x) Sigma) ]
        [                (.)^T                                  K22 (x) Sigma         ]

    H = blockdiag(uh_i^T).  Sigma's index is (q in p, r in n), r fastest.  The data kernel is whatever module was handed
    in, evaluated densely; the contractions over the small (p, n) indices are einsum glue and stay differentiable (the
    comparator's fit back-propagates through them)."""

    def num_outputs_per_input(self, mxu1, mxu2):
        M1, X1, _ = self.decoder.decode(mxu1)
        M1s = M1[..., 0]
        end = _train_end(M1s)
        test_size = (M1s.size(-1) - end) * self.task_covar_module.covar_matrix.shape[-1]
        return (end * X1.shape[-1] + test_size) / M1s.size(-1)

    def _sigma4(self, dtype):
        """Sigma in the dtype the covariance is assembled in (the inputs' dtype: a float32 model evaluated on float64
        inputs — what the comparator's fit does — gets a float64 covariance of its float32 parameters)."""
        _, n, p = self.decoder.sizes
        Sigma = _dense(self.task_covar_module.covar_matrix).to(dtype)
        return Sigma, Sigma.reshape(p, n, p, n), n, p

    def kernel1(self, Kxx, UH1, UH2):
        """(H1 (x) I)(K (x) Sigma)(H2^T (x) I): entry [(i,r),(j,s)] = K_ij sum_qq' uh1_i[q] Sigma[(q,r),(q',s)] uh2_j[q']."""
        _, S4, n, _ = self._sigma4(Kxx.dtype)
        uSu = torch.einsum('iq,qrts,jt->irjs', UH1.to(S4.dtype), S4, UH2.to(S4.dtype))
        return (Kxx.unsqueeze(1).unsqueeze(-1) * uSu).reshape(UH1.shape[0] * n, UH2.shape[0] * n)

    def kernel2(self, Kxx):
        Sigma, _, _, _ = self._sigma4(Kxx.dtype)
        return torch_kron(Kxx, Sigma, batch_dims=0)

    def correlation_kernel_12(self, Kxx, UH1):
        """(H1 (x) I)(K12 (x) Sigma): entry [(i,r),(j,q',s)] = K_ij sum_q uh1_i[q] Sigma[(q,r),(q',s)]."""
        _, S4, n, p = self._sigma4(Kxx.dtype)
        uS = torch.einsum('iq,qrts->irts', UH1.to(S4.dtype), S4)                       # (N1, n, p, n)
        return (Kxx.reshape(Kxx.shape[0], 1, Kxx.shape[1], 1, 1) * uS.unsqueeze(2)).reshape(
            UH1.shape[0] * n, Kxx.shape[1] * p * n)

    def mask_dependent_covar(self, M1s, U1, M2s, U2, covar_xx):
        e1, e2 = _train_end(M1s), _train_end(M2s)
        assert (M1s[e1:] == 0).all() and (M2s[e2:] == 0).all(), "rows must be sorted train-first"
        n11 = bool(e1 and e2)
        n22 = bool((covar_xx.shape[0] - e1) and (covar_xx.shape[1] - e2))
        k11 = self.kernel1(covar_xx[:e1, :e2], U1[:e1], U2[:e2]) if n11 else None
        k22 = self.kernel2(covar_xx[e1:, e2:]) if n22 else None
        if n11 and n22:
            k12 = self.correlation_kernel_12(covar_xx[:e1, e2:], U1[:e1])
            k21 = self.correlation_kernel_12(covar_xx[e1:, :e2].transpose(0, 1), U2[:e2]).transpose(0, 1)
            return torch.cat([torch.cat([k11, k12], dim=1), torch.cat([k21, k22], dim=1)], dim=0)
        return k22 if k11 is None else k11

    def forward(self, mxu1, mxu2, diag=False, last_dim_is_batch=False, **params):
        assert not torch.isnan(mxu1).any() and not torch.isnan(mxu2).any()
        if last_dim_is_batch:
            raise RuntimeError("HetergeneousCoregionalizationKernel does not accept the last_dim_is_batch argument.")
        M1, X1, U1 = self.decoder.decode(mxu1)
        M2, X2, U2 = self.decoder.decode(mxu2)
        covar_x = _dense(self.data_covar_module.forward(X1, X2, **params)).to(device=mxu1.device, dtype=mxu1.dtype)
        for name, value in self.data_covar_module.named_parameters():
            assert not torch.isnan(value).any()
        res = self.mask_dependent_covar(M1[..., 0], U1, M2[..., 0], U2, covar_x)
        return torch.diagonal(res) if diag else res
