"""TEST INFRASTRUCTURE ONLY — dense stand-in for the gpytorch surface the reference touches.

The reference (wecacuee/Bayesian_CBF) depends on a gpytorch fork that is not vendored
(requirements.txt:3, `wecacuee/gpytorch@fractional-outputs-per-input`, egg gpytorch==0.3.7fopi1) and
on matplotlib / kwplus, none of which exist in this image.  ``install()`` registers minimal *dense*
replacements in ``sys.modules`` so that the reference's own source files

    bayes_cbf/control_affine_model.py, matrix_variate_multitask_kernel.py,
    matrix_variate_multitask_model.py, gp_algebra.py, cbc1.py, cbc2.py, misc.py

can be imported UNMODIFIED from /root/reference and executed on CPU.  It is used only by
``oracle/gen_golden.py`` (golden-vector generation, run in the build container where /root/reference
exists) and by tests that validate the oracle restatement.  Nothing in the product package imports it.

The arithmetic restated here follows gpytorch 0.3.x definitions:
  * RBFKernel:   exp(-0.5 * sqdist(x1/l, x2/l)), sqdist by the mean-centred |a|^2+|b|^2-2ab form
                 (gpytorch/kernels/kernel.py `Distance._sq_dist`), l = softplus(raw_lengthscale)
  * ScaleKernel: softplus(raw_outputscale) * base
  * IndexKernel: F F^T + diag(softplus(raw_var)), F ~ randn(T, rank), raw_var ~ randn(T)
  * ConstantMean: constant (zeros(1)) broadcast
  * MultivariateNormal.log_prob: dense Cholesky with psd-safe jitter escalation
"""
import copy
import importlib.machinery
import math
import sys
import types
from contextlib import contextmanager
from unittest import mock

import torch
from torch import nn
from torch.nn.functional import softplus


# ----------------------------------------------------------------------------- lazy tensors
class LazyTensor:
    """Dense-backed object exposing the LazyTensor methods the reference calls."""

    def __init__(self, t):
        if isinstance(t, LazyTensor):
            t = t.tensor
        if not isinstance(t, torch.Tensor):
            t = torch.as_tensor(t)
        self.tensor = t

    # basic info
    @property
    def shape(self):
        return self.tensor.shape

    @property
    def dtype(self):
        return self.tensor.dtype

    @property
    def device(self):
        return self.tensor.device

    @property
    def ndim(self):
        return self.tensor.ndim

    def size(self, *a):
        return self.tensor.size(*a)

    def dim(self):
        return self.tensor.dim()

    def numel(self):
        return self.tensor.numel()

    def evaluate(self):
        return self.tensor

    def t(self):
        return LazyTensor(self.tensor.transpose(-1, -2))

    def transpose(self, a, b):
        return LazyTensor(self.tensor.transpose(a, b))

    def diag(self):
        return torch.diagonal(self.tensor, dim1=-2, dim2=-1)

    def detach(self):
        return LazyTensor(self.tensor.detach())

    def add_jitter(self, jitter_val=1e-3):
        eye = torch.eye(self.tensor.shape[-1], dtype=self.dtype, device=self.device)
        return LazyTensor(self.tensor + jitter_val * eye)

    def __getitem__(self, idx):
        return LazyTensor(self.tensor[idx])

    def __matmul__(self, other):
        o = other.tensor if isinstance(other, LazyTensor) else other
        res = self.tensor @ o
        return LazyTensor(res) if isinstance(other, LazyTensor) else res

    def __rmatmul__(self, other):
        return other @ self.tensor

    def matmul(self, other):
        return self.__matmul__(other)

    def __add__(self, other):
        o = other.tensor if isinstance(other, LazyTensor) else other
        return LazyTensor(self.tensor + o)

    def __mul__(self, other):
        o = other.tensor if isinstance(other, LazyTensor) else other
        return LazyTensor(self.tensor * o)

    def mul(self, other):
        return self.__mul__(other)


class NonLazyTensor(LazyTensor):
    pass


def _dense(x):
    return x.tensor if isinstance(x, LazyTensor) else x


def _kron2(a, b):
    # dense Kronecker by broadcasting (torch.kron rejects some non-contiguous views)
    ra, ca = a.shape[-2:]
    rb, cb = b.shape[-2:]
    return (a[..., :, None, :, None] * b[..., None, :, None, :]).reshape(*a.shape[:-2], ra * rb, ca * cb)


class KroneckerProductLazyTensor(LazyTensor):
    def __init__(self, *ts):
        res = _dense(ts[0])
        for t in ts[1:]:
            res = _kron2(res, _dense(t))
        super().__init__(res)


class BlockDiagLazyTensor(LazyTensor):
    def __init__(self, base, block_dim=-3):
        blocks = _dense(base)  # (..., N, r, c)
        N, r, c = blocks.shape[-3:]
        out = blocks.new_zeros(*blocks.shape[:-3], N * r, N * c)
        for i in range(N):
            out[..., i * r:(i + 1) * r, i * c:(i + 1) * c] = blocks[..., i, :, :]
        super().__init__(out)


class InterpolatedLazyTensor(LazyTensor):
    def __init__(self, base_lazy_tensor, left_interp_indices=None, right_interp_indices=None, **kw):
        base = _dense(base_lazy_tensor)
        li = left_interp_indices.reshape(-1)
        ri = right_interp_indices.reshape(-1)
        super().__init__(base[li][:, ri])


def lazify(x):
    return x if isinstance(x, LazyTensor) else NonLazyTensor(x)


def delazify(x):
    return _dense(x)


def lazycat(inputs, dim=0, output_device=None):
    return LazyTensor(torch.cat([_dense(i) for i in inputs], dim=dim))


# ----------------------------------------------------------------------------- settings
class _Flag:
    def __init__(self, *a, **k):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


# ----------------------------------------------------------------------------- kernels
class Kernel(nn.Module):
    has_lengthscale = False

    def __init__(self, ard_num_dims=None, lengthscale_prior=None, batch_shape=torch.Size([]), **kwargs):
        super().__init__()
        self.ard_num_dims = ard_num_dims
        self._priors = []
        if self.has_lengthscale:
            d = 1 if ard_num_dims is None else ard_num_dims
            self.raw_lengthscale = nn.Parameter(torch.zeros(1, d))
            if lengthscale_prior is not None:
                self._priors.append((lengthscale_prior, lambda: self.lengthscale))

    @property
    def lengthscale(self):
        return softplus(self.raw_lengthscale)

    def named_priors_shim(self):
        out = list(self._priors)
        for m in self.children():
            if isinstance(m, Kernel):
                out.extend(m.named_priors_shim())
        return out

    def __call__(self, x1, x2=None, diag=False, last_dim_is_batch=False, **params):
        if isinstance(x1, torch.Tensor) and x1.ndim == 1:
            x1 = x1.unsqueeze(1)
        if x2 is not None and isinstance(x2, torch.Tensor) and x2.ndim == 1:
            x2 = x2.unsqueeze(1)
        if x2 is None:
            x2 = x1
        kw = dict(params)
        if diag:
            kw['diag'] = True
        res = nn.Module.__call__(self, x1, x2, **kw)
        return res if diag else lazify(res)

    def __add__(self, other):
        return AdditiveKernel(self, other)


class AdditiveKernel(Kernel):
    def __init__(self, *kernels):
        super().__init__()
        self.kernels = nn.ModuleList(kernels)

    def forward(self, x1, x2, diag=False, **params):
        res = None
        for k in self.kernels:
            t = _dense(k.forward(x1, x2, diag=diag, **params) if diag else k.forward(x1, x2, **params))
            res = t if res is None else res + t
        return res


def _sq_dist(x1, x2, x1_eq_x2=False):
    adjustment = x1.mean(-2, keepdim=True)
    x1 = x1 - adjustment
    x2 = x2 - adjustment
    x1_norm = x1.pow(2).sum(dim=-1, keepdim=True)
    x1_pad = torch.ones_like(x1_norm)
    # gpytorch only takes the x1==x2 shortcut when no input gradient is required; without that guard the
    # reference's GradientGP Hessians at x == x' (gp_algebra.py:352-393) would vanish.
    x1_eq_x2 = x1_eq_x2 and not x1.requires_grad and not x2.requires_grad
    if x1_eq_x2:
        x2_norm, x2_pad = x1_norm, x1_pad
    else:
        x2_norm = x2.pow(2).sum(dim=-1, keepdim=True)
        x2_pad = torch.ones_like(x2_norm)
    x1_ = torch.cat([-2.0 * x1, x1_norm, x1_pad], dim=-1)
    x2_ = torch.cat([x2, x2_pad, x2_norm], dim=-1)
    res = x1_.matmul(x2_.transpose(-2, -1))
    if x1_eq_x2:
        res = res - torch.diag_embed(torch.diagonal(res, dim1=-2, dim2=-1))
    return res.clamp_min(0)


class Distance(nn.Module):
    def _sq_dist(self, x1, x2, postprocess=None, x1_eq_x2=False):
        r = _sq_dist(x1, x2, x1_eq_x2)
        return postprocess(r) if postprocess else r


def default_postprocess_script(x):
    return x


class RBFKernel(Kernel):
    has_lengthscale = True

    def forward(self, x1, x2, diag=False, **params):
        x1_ = x1.div(self.lengthscale)
        x2_ = x2.div(self.lengthscale)
        if diag:
            return (x1_ - x2_).pow(2).sum(-1).div(-2).exp()
        x1_eq_x2 = (x1_.shape == x2_.shape) and bool(torch.equal(x1_, x2_))
        return _sq_dist(x1_, x2_, x1_eq_x2).div(-2).exp()


class LinearKernel(Kernel):
    def __init__(self, **kw):
        super().__init__(**kw)
        self.raw_variance = nn.Parameter(torch.zeros(1, 1))

    @property
    def variance(self):
        return softplus(self.raw_variance)

    def forward(self, x1, x2, diag=False, **params):
        x1_ = x1 * self.variance.sqrt()
        x2_ = x2 * self.variance.sqrt()
        if diag:
            return (x1_ * x2_).sum(-1)
        return x1_ @ x2_.transpose(-2, -1)


class ScaleKernel(Kernel):
    def __init__(self, base_kernel, outputscale_prior=None, **kw):
        super().__init__(**kw)
        self.base_kernel = base_kernel
        self.raw_outputscale = nn.Parameter(torch.zeros(()))

    @property
    def outputscale(self):
        return softplus(self.raw_outputscale)

    def forward(self, x1, x2, diag=False, **params):
        orig = _dense(self.base_kernel.forward(x1, x2, diag=diag, **params) if diag
                      else self.base_kernel.forward(x1, x2, **params))
        return orig * self.outputscale


class IndexKernel(Kernel):
    def __init__(self, num_tasks, rank=1, prior=None, **kw):
        super().__init__(**kw)
        self.covar_factor = nn.Parameter(torch.randn(num_tasks, rank))
        self.raw_var = nn.Parameter(torch.randn(num_tasks))

    @property
    def var(self):
        return softplus(self.raw_var)

    @property
    def covar_matrix(self):
        return LazyTensor(self.covar_factor @ self.covar_factor.transpose(-1, -2) + torch.diag(self.var))

    def forward(self, i1, i2, **params):
        cm = self.covar_matrix.tensor
        return cm[i1.reshape(-1)][:, i2.reshape(-1)]


class MultitaskKernel(Kernel):
    pass


# ----------------------------------------------------------------------------- means
class ConstantMean(nn.Module):
    def __init__(self, prior=None, batch_shape=torch.Size(), **kw):
        super().__init__()
        self.constant = nn.Parameter(torch.zeros(*batch_shape, 1))

    def forward(self, x):
        return self.constant.expand(x.shape[:-1])


class MultitaskMean(nn.Module):
    def __init__(self, base_means, num_tasks):
        super().__init__()
        if isinstance(base_means, nn.Module):
            base_means = [base_means] + [copy.deepcopy(base_means) for _ in range(num_tasks - 1)]
        self.base_means = nn.ModuleList(base_means)
        self.num_tasks = num_tasks

    def forward(self, x):
        return torch.cat([m(x).unsqueeze(-1) for m in self.base_means], dim=-1)


# ----------------------------------------------------------------------------- distributions
def psd_safe_cholesky(A, max_tries=6):
    try:
        return torch.linalg.cholesky(A)
    except RuntimeError:
        jitter = 1e-6 if A.dtype == torch.float32 else 1e-8
        for i in range(max_tries):
            try:
                return torch.linalg.cholesky(A + jitter * (10 ** i) * torch.eye(A.shape[-1], dtype=A.dtype))
            except RuntimeError:
                continue
        raise


class MultivariateNormal:
    def __init__(self, mean, covariance_matrix, validate_args=False):
        self.loc = mean
        self._covar = lazify(covariance_matrix)

    @property
    def mean(self):
        return self.loc

    @property
    def covariance_matrix(self):
        return self._covar.evaluate()

    @property
    def lazy_covariance_matrix(self):
        return self._covar

    @property
    def variance(self):
        return self._covar.diag()

    def log_prob(self, value):
        diff = value - self.loc
        L = psd_safe_cholesky(self._covar.evaluate())
        z = torch.linalg.solve_triangular(L, diff.unsqueeze(-1), upper=False)
        inv_quad = (z * z).sum()
        logdet = 2 * torch.diagonal(L).log().sum()
        return -0.5 * (inv_quad + logdet + diff.size(-1) * math.log(2 * math.pi))


# ----------------------------------------------------------------------------- likelihoods / models / mlls
class FixedGaussianNoise(nn.Module):
    def __init__(self, noise):
        super().__init__()
        self.noise = noise


class _GaussianLikelihoodBase(nn.Module):
    def __init__(self, noise_covar=None, **kw):
        nn.Module.__init__(self)
        self.noise_covar = noise_covar

    def marginal(self, function_dist, *a, **k):
        return function_dist

    def __call__(self, inp, *a, **k):
        if isinstance(inp, MultivariateNormal):
            return self.marginal(inp, *a, **k)
        return self.forward(inp, *a, **k)


class GaussianLikelihood(_GaussianLikelihoodBase):
    pass


class ExactGP(nn.Module):
    def __init__(self, train_inputs, train_targets, likelihood):
        super().__init__()
        if train_inputs is not None and isinstance(train_inputs, torch.Tensor):
            train_inputs = (train_inputs,)
        self.train_inputs = train_inputs
        self.train_targets = train_targets
        self.likelihood = likelihood

    def set_train_data(self, inputs=None, targets=None, strict=True):
        if inputs is not None:
            if isinstance(inputs, torch.Tensor):
                inputs = (inputs,)
            self.train_inputs = tuple(i.unsqueeze(-1) if i.ndim == 1 else i for i in inputs)
        if targets is not None:
            self.train_targets = targets

    def __call__(self, *args, **kwargs):
        if self.training or self.train_inputs is None:
            return self.forward(*args, **kwargs)
        # eval mode: condition the joint prior on the training targets (dense; parity-unpinned)
        train = self.train_inputs[0]
        test = args[0]
        ntr = None
        full = torch.cat([train, test], dim=-2)
        joint = self.forward(full)
        mean, cov = joint.mean, joint.covariance_matrix
        ntr = self.train_targets.shape[-1]
        m1, m2 = mean[:ntr], mean[ntr:]
        K11, K12, K22 = cov[:ntr, :ntr], cov[:ntr, ntr:], cov[ntr:, ntr:]
        L = psd_safe_cholesky(K11)
        alpha = torch.cholesky_solve((self.train_targets - m1).unsqueeze(-1), L).squeeze(-1)
        pm = m2 + K12.transpose(-1, -2) @ alpha
        pc = K22 - K12.transpose(-1, -2) @ torch.cholesky_solve(K12, L)
        return MultivariateNormal(pm, pc)


class ExactMarginalLogLikelihood(nn.Module):
    def __init__(self, likelihood, model):
        super().__init__()
        self.likelihood = likelihood
        self.model = model

    def forward(self, output, target, *params):
        output = self.likelihood(output, *params)
        res = output.log_prob(target)
        for mod in self.model.modules():
            if isinstance(mod, Kernel):
                for prior, closure in mod._priors:
                    res = res + prior.log_prob(closure()).sum()
        return res / target.size(-1)


class GammaPrior:
    def __init__(self, concentration, rate, **kw):
        self.d = torch.distributions.Gamma(torch.as_tensor(float(concentration)), torch.as_tensor(float(rate)))

    def log_prob(self, x):
        return self.d.log_prob(x)


def cached(*a, **k):
    def deco(f):
        return f
    if len(a) == 1 and callable(a[0]) and not k:
        return a[0]
    return deco


# ----------------------------------------------------------------------------- install
class _MockModule(types.ModuleType):
    """Importable stand-in for an absent package: every attribute is a MagicMock."""

    def __init__(self, name):
        super().__init__(name)
        self.__path__ = []
        self.__spec__ = importlib.machinery.ModuleSpec(name, None, is_package=True)

    def __getattr__(self, item):
        if item.startswith('__') and item.endswith('__'):
            raise AttributeError(item)
        val = mock.MagicMock(name=self.__name__ + '.' + item)
        setattr(self, item, val)
        return val


def _mod(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def install(reference_root='/root/reference'):
    """Register the shim + matplotlib/kwplus stubs and put the reference on sys.path."""
    if 'gpytorch' in sys.modules and getattr(sys.modules['gpytorch'], '_bcbf_shim', False):
        return
    settings = _mod('gpytorch.settings', lazily_evaluate_kernels=_Flag, max_cg_iterations=_Flag, debug=_Flag,
                    fast_computations=_Flag, fast_pred_var=_Flag)
    lazy = _mod('gpytorch.lazy', LazyTensor=LazyTensor, NonLazyTensor=NonLazyTensor,
                KroneckerProductLazyTensor=KroneckerProductLazyTensor, BlockDiagLazyTensor=BlockDiagLazyTensor,
                InterpolatedLazyTensor=InterpolatedLazyTensor, lazify=lazify, delazify=delazify, cat=lazycat)
    kernel_mod = _mod('gpytorch.kernels.kernel', Kernel=Kernel, Distance=Distance,
                      default_postprocess_script=default_postprocess_script)
    kernels = _mod('gpytorch.kernels', Kernel=Kernel, RBFKernel=RBFKernel, ScaleKernel=ScaleKernel,
                   IndexKernel=IndexKernel, LinearKernel=LinearKernel, MultitaskKernel=MultitaskKernel,
                   AdditiveKernel=AdditiveKernel, kernel=kernel_mod)
    means = _mod('gpytorch.means', ConstantMean=ConstantMean, MultitaskMean=MultitaskMean)
    dists = _mod('gpytorch.distributions', MultivariateNormal=MultivariateNormal,
                 base_distributions=torch.distributions)
    noise_models = _mod('gpytorch.likelihoods.noise_models', FixedGaussianNoise=FixedGaussianNoise)
    likelihoods = _mod('gpytorch.likelihoods', _GaussianLikelihoodBase=_GaussianLikelihoodBase,
                       GaussianLikelihood=GaussianLikelihood, noise_models=noise_models)
    models = _mod('gpytorch.models', ExactGP=ExactGP)
    mlls = _mod('gpytorch.mlls', ExactMarginalLogLikelihood=ExactMarginalLogLikelihood)
    priors = _mod('gpytorch.priors', GammaPrior=GammaPrior)
    memoize = _mod('gpytorch.utils.memoize', cached=cached)
    utils = _mod('gpytorch.utils', memoize=memoize)
    g = _mod('gpytorch', settings=settings, lazy=lazy, kernels=kernels, means=means, distributions=dists,
             likelihoods=likelihoods, models=models, mlls=mlls, priors=priors, utils=utils)
    g._bcbf_shim = True
    # matplotlib / kwplus / other absent packages -> permissive mock modules
    for name in ['matplotlib', 'matplotlib.pyplot', 'matplotlib.transforms', 'matplotlib.patches',
                 'matplotlib.colors', 'matplotlib.cm', 'matplotlib.ticker', 'matplotlib.animation',
                 'matplotlib.lines', 'matplotlib.axes', 'matplotlib.collections', 'matplotlib.gridspec',
                 'matplotlib.figure', 'mpl_toolkits', 'mpl_toolkits.mplot3d', 'kwplus', 'kwplus.functools',
                 'kwplus.variations', 'cvxopt', 'cvxpy', 'bdlqr', 'bdlqr.full', 'mpc', 'mpc.mpc']:
        if name not in sys.modules:
            sys.modules[name] = _MockModule(name)
    sys.modules['matplotlib.pyplot'].subplots = mock.MagicMock(return_value=(mock.MagicMock(), mock.MagicMock()))
    # removed torch APIs still used by the reference (gp_algebra.py:385,389)
    if not getattr(torch, '_bcbf_eig_compat', False):
        def _eig(A, eigenvectors=False):
            w, v = torch.linalg.eig(A)
            return torch.stack([w.real, w.imag], dim=-1), v.real
        torch.eig = _eig
        torch.symeig = lambda A, eigenvectors=False, upper=True: torch.linalg.eigh(A, UPLO='U' if upper else 'L')
        torch._bcbf_eig_compat = True
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
