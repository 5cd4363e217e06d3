"""Parameter-holding modules with the gpytorch names the reference's model is built from
(control_affine_model.py:139-177): ScaleKernel(RBFKernel(ard)), IndexKernel, ConstantMean, MultitaskMean, GammaPrior.
gpytorch itself (a fork, requirements.txt:3) is not a dependency: these classes keep its parameter names and
constraints (softplus-positive `raw_*` parameters) so that state_dicts line up, and evaluate on the GPU through
`bayesian_cbf_b200.ops` — there is no CPU evaluation path.
"""
import copy
import math

import torch
import torch.nn.functional as F
from torch import nn

from . import autograd_ops, ops


def inv_softplus(x):
    return x + torch.log(-torch.expm1(-x))


class Evaluated:
    """Stand-in for gpytorch's LazyTensor results: `.evaluate()`, `.diag()`, shape/indexing/matmul passthrough."""

    def __init__(self, t):
        self.tensor = t

    def evaluate(self):
        return self.tensor

    def diag(self):
        return torch.diagonal(self.tensor, dim1=-2, dim2=-1)

    def t(self):
        return Evaluated(self.tensor.transpose(-2, -1))

    @property
    def shape(self):
        return self.tensor.shape

    @property
    def dtype(self):
        return self.tensor.dtype

    @property
    def device(self):
        return self.tensor.device

    def size(self, *a):
        return self.tensor.size(*a)

    def numel(self):
        return self.tensor.numel()

    def __getitem__(self, idx):
        return Evaluated(self.tensor[idx])

    def __matmul__(self, other):
        o = other.tensor if isinstance(other, Evaluated) else other
        return Evaluated(self.tensor @ o)

    def __rmatmul__(self, other):
        return Evaluated(other @ self.tensor)


def _need_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError("bayesian_cbf_b200.%s evaluates on a CUDA device only (no CPU fallback); got %s" % (what, t.device))


class Kernel(nn.Module):
    """Minimal `gpytorch.kernels.Kernel` surface: __call__(x1, x2=None, diag=False) -> Evaluated."""

    def __call__(self, x1, x2=None, diag=False, **params):
        x1_ = x1.unsqueeze(-1) if x1.ndim == 1 else x1
        x2_ = x1_ if x2 is None else (x2.unsqueeze(-1) if x2.ndim == 1 else x2)
        res = self.forward(x1_, x2_, diag=diag, **params)
        return res if isinstance(res, Evaluated) else Evaluated(res)


class GammaPrior:
    def __init__(self, concentration, rate):
        self.concentration = float(concentration)
        self.rate = float(rate)

    def log_prob(self, x):
        a, b = self.concentration, self.rate
        return (a * math.log(b) + (a - 1) * torch.log(x) - b * x - math.lgamma(a)).sum()


class RBFKernel(Kernel):
    """exp(-1/2 |(x - x') / lengthscale|^2), ARD; `raw_lengthscale` (1, d), lengthscale = softplus(raw)."""

    def __init__(self, ard_num_dims=None, lengthscale_prior=None):
        super().__init__()
        self.ard_num_dims = ard_num_dims
        self.raw_lengthscale = nn.Parameter(torch.zeros(1, ard_num_dims or 1))
        self.lengthscale_prior = lengthscale_prior

    @property
    def lengthscale(self):
        return F.softplus(self.raw_lengthscale)

    @lengthscale.setter
    def lengthscale(self, value):
        v = torch.as_tensor(value, dtype=self.raw_lengthscale.dtype, device=self.raw_lengthscale.device)
        with torch.no_grad():
            self.raw_lengthscale.copy_(inv_softplus(v.expand_as(self.raw_lengthscale)))

    def forward(self, x1, x2, diag=False, outputscale=None, **params):
        _need_cuda(x1, 'RBFKernel')
        s = torch.ones((), dtype=torch.float64, device=x1.device) if outputscale is None else outputscale.double()
        ls = self.lengthscale.reshape(-1).double().expand(x1.shape[-1]).contiguous()
        a, c = x1.double(), x2.double()
        if torch.is_grad_enabled() and self.training and (self.raw_lengthscale.requires_grad or s.requires_grad):
            # hyper-parameter fit of a model that evaluates this module directly (the CoGP comparator): the squared
            # distances come from the fused kernel (log of the unit-lengthscale Gram), the lengthscale / scale
            # dependence is elementwise and differentiable
            one = torch.ones(a.shape[-1], dtype=torch.float64, device=a.device)
            if self.raw_lengthscale.numel() == 1:
                D2 = -2.0 * torch.log(ops.gram_ca(a.contiguous(), c.contiguous(), one, 1.0).clamp_min(1e-300))
                K = s * torch.exp(-0.5 * D2 / (self.lengthscale.double().reshape(()) ** 2))
            else:
                d = (a.unsqueeze(1) - c.unsqueeze(0)) / ls
                K = s * torch.exp(-0.5 * (d * d).sum(-1))
        elif a.requires_grad or c.requires_grad:
            K = autograd_ops.rbf_kernel(a, c, ls.detach(), s.detach())
        else:
            K = ops.gram_ca(a.contiguous(), c.contiguous(), ls.detach(), float(s.detach()))
        K = K.to(x1.dtype)
        return torch.diagonal(K) if diag else K


class ScaleKernel(Kernel):
    def __init__(self, base_kernel):
        super().__init__()
        self.base_kernel = base_kernel
        self.raw_outputscale = nn.Parameter(torch.zeros(()))

    @property
    def outputscale(self):
        return F.softplus(self.raw_outputscale)

    @outputscale.setter
    def outputscale(self, value):
        v = torch.as_tensor(value, dtype=self.raw_outputscale.dtype, device=self.raw_outputscale.device)
        with torch.no_grad():
            self.raw_outputscale.copy_(inv_softplus(v.reshape(())))

    def forward(self, x1, x2, diag=False, **params):
        return self.base_kernel.forward(x1, x2, diag=diag, outputscale=self.outputscale, **params)


class IndexKernel(Kernel):
    """covar_matrix = F F^T + diag(softplus(raw_var)); F ~ randn(num_tasks, rank), raw_var ~ randn(num_tasks)."""

    def __init__(self, num_tasks, rank=1):
        super().__init__()
        self.covar_factor = nn.Parameter(torch.randn(num_tasks, rank))
        self.raw_var = nn.Parameter(torch.randn(num_tasks))

    @property
    def var(self):
        return F.softplus(self.raw_var)

    @property
    def covar_matrix(self):
        return Evaluated(self.covar_factor @ self.covar_factor.transpose(-1, -2) + torch.diag(self.var))

    def forward(self, i1, i2, **params):
        C = self.covar_matrix.evaluate()
        return C[i1.reshape(-1)][:, i2.reshape(-1)]


class ConstantMean(nn.Module):
    def __init__(self):
        super().__init__()
        self.constant = nn.Parameter(torch.zeros(1))

    def forward(self, x):
        return self.constant.expand(x.shape[:-1])


class MultitaskMean(nn.Module):
    def __init__(self, base_means, num_tasks):
        super().__init__()
        self.base_means = nn.ModuleList([base_means] + [copy.deepcopy(base_means) for _ in range(num_tasks - 1)])
        self.num_tasks = num_tasks


class MultivariateNormalResult:
    """What `ControlAffineExactGP.forward` returns: mean vector + dense covariance."""

    def __init__(self, mean, covar):
        self.mean = mean
        self._covar = covar.evaluate() if isinstance(covar, Evaluated) else covar

    @property
    def covariance_matrix(self):
        return self._covar

    @property
    def lazy_covariance_matrix(self):
        return Evaluated(self._covar)


class LinearKernel(Kernel):
    """variance * x1 x2^T (gpytorch LinearKernel; `raw_variance` (1,1), softplus-positive)."""

    def __init__(self):
        super().__init__()
        self.raw_variance = nn.Parameter(torch.zeros(1, 1))

    @property
    def variance(self):
        return F.softplus(self.raw_variance)

    def forward(self, x1, x2, diag=False, outputscale=None, **params):
        _need_cuda(x1, 'LinearKernel')
        if torch.is_grad_enabled() and self.training and self.raw_variance.requires_grad:
            G = ops.gemm(x1.double(), x2.double(), transb=True)
            sc = self.variance.double().reshape(()) * (1.0 if outputscale is None else outputscale.double())
            K = (sc * G).to(x1.dtype)
            return torch.diagonal(K) if diag else K
        s = 1.0 if outputscale is None else float(outputscale.detach())
        K = ops.gemm(x1.double(), x2.double(), transb=True, alpha=s * float(self.variance.detach()))
        K = K.to(x1.dtype)
        return torch.diagonal(K) if diag else K


class AdditiveKernel(Kernel):
    def __init__(self, *kernels):
        super().__init__()
        self.kernels = nn.ModuleList(kernels)

    def forward(self, x1, x2, diag=False, **params):
        out = None
        for k in self.kernels:
            r = k.forward(x1, x2, diag=diag, **params)
            out = r if out is None else out + r
        return out


def _kernel_add(self, other):
    return AdditiveKernel(self, other)


Kernel.__add__ = _kernel_add
