"""Algebra of (vector-valued) Gaussian processes given as closures — the API of the reference's
bayes_cbf/gp_algebra.py (classes at :12, :70, :109, :133, :170, :201, :225, :258, :319), consumed by cbc1 / cbc2 and
the controllers.  An expression tree over leaves (`GaussianProcess`, `DeterministicGP`) propagates

    mean(x)            E[Z(x)]
    knl(x, x')         cov(Z(x), Z(x'))
    covar(Y, x, x')    cov(Z(x), Y(x'))

through `+`, scalar `*`, `det.t() @ gp`, `gp.t() @ gp`, `.t()` and the gradient operator `GradientGP`.
Everything here is shape bookkeeping on n- / m-sized tensors plus autograd drivers; the numbers come from the
regressor closures (CUDA kernels, differentiable through bayesian_cbf_b200.autograd_ops).

Conscious deviations from the reference: `torch.eig` (removed from torch) in the PSD clean-up of
`GradientGP.knl` (:384-393) is replaced by `torch.linalg.eigh` with the reconstruction V diag(w) V^T.
"""
from abc import ABC, abstractmethod

import torch
from torch.distributions import MultivariateNormal

from .misc import t_hessian, t_jac, variable_required_grad


class GaussianProcessBase(ABC):
    @property
    @abstractmethod
    def shape(self):
        ...

    @abstractmethod
    def mean(self, x):
        ...

    @abstractmethod
    def knl(self, x, xp):
        ...

    @abstractmethod
    def covar(self, Z, x, xp):
        ...

    def sample(self, x, sample_shape=torch.Size([])):
        return MultivariateNormal(self.mean(x), self.knl(x, x)).sample(sample_shape)

    def __add__(self, Y):
        return GaussianProcessAddExpr(self, Y)

    def __mul__(self, a):
        return GaussianProcessMulExpr(self, a)

    def __truediv__(self, a):
        return GaussianProcessMulExpr(self, 1 / a)

    __div__ = __truediv__

    def __matmul__(self, Y):
        if isinstance(self, DeterministicGP):
            return GaussianProcessDetMatmulExpr(self, Y)
        return GaussianProcessMatmulExpr(self, Y)

    def t(self):
        return GaussianProcessTranspose(self)


class GaussianProcessLeaf(GaussianProcessBase):
    @classmethod
    def isleaf(cls):
        return True

    def name(self):
        return self._name

    def __str__(self):
        return "GaussianProcessLeaf(name={})".format(self._name)


class GaussianProcessExpr(GaussianProcessBase):
    @classmethod
    def isleaf(cls):
        return False


class DeterministicGP(GaussianProcessLeaf):
    """A deterministic function seen as a GP with zero kernel."""

    def __init__(self, mean, shape, name="{mean}"):
        assert len(shape) <= 2
        assert len(shape) == 1 or min(shape) == 1
        self._mean = mean
        self._shape = shape
        self._name = name.format(mean=mean)

    @property
    def shape(self):
        return self._shape

    def mean(self, x):
        return self._mean(x)

    def knl(self, x, xp):
        k = max(self._shape)
        return x.new_zeros(k, k)

    def covar(self, Z, x, xp):
        assert isinstance(Z, GaussianProcessBase)
        if isinstance(Z, DeterministicGP):
            return x.new_zeros(max(self._shape), max(Z.shape))
        return Z.covar(self, x, xp).t()

    def sample(self, x, sample_shape=torch.Size([])):
        return self.mean(x).expand(*sample_shape, -1)

    def __repr__(self):
        return "DeterministicGP(mean={})".format(self._mean)

    def __str__(self):
        return "DeterministicGP(name={})".format(self._name)


class GaussianProcessAddExpr(GaussianProcessExpr):
    def __init__(self, X, Y):
        assert isinstance(X, GaussianProcessBase) and isinstance(Y, GaussianProcessBase)
        assert X.shape == Y.shape
        self.lhs, self.rhs = X, Y

    @property
    def shape(self):
        return self.lhs.shape

    def mean(self, x):
        return self.lhs.mean(x) + self.rhs.mean(x)

    def knl(self, x, xp):
        X, Y = self.lhs, self.rhs
        return X.knl(x, xp) + Y.knl(x, xp) + Y.covar(X, x, xp) + X.covar(Y, x, xp)

    def covar(self, Z, x, xp):
        assert isinstance(Z, GaussianProcessBase)
        return self.lhs.covar(Z, x, xp) + self.rhs.covar(Z, x, xp)


class GaussianProcessMatmulExpr(GaussianProcessExpr):
    """X^T Y for two jointly Gaussian vector GPs (second-order moment matching, reference :133-168)."""

    def __init__(self, X, Y):
        assert isinstance(X, GaussianProcessBase) and isinstance(Y, GaussianProcessBase)
        assert X.shape[-1] == Y.shape[0]
        self.lhs = X.t()
        self.rhs = Y

    @property
    def shape(self):
        return (1,)

    def mean(self, x):
        X, Y = self.lhs, self.rhs
        return X.mean(x).t() @ Y.mean(x) + 0.5 * X.covar(Y, x, x).trace() + 0.5 * Y.covar(X, x, x).trace()

    def knl(self, x, xp):
        X, Y = self.lhs, self.rhs
        # the last term follows the reference as written (it carries a FIXME there, :158-159)
        return (2 * X.covar(Y, x, xp).trace() ** 2
                + Y.mean(x).t() @ X.knl(x, xp) @ Y.mean(xp)
                + X.mean(x).t() @ Y.knl(x, xp) @ X.mean(xp)
                + 2 * Y.mean(x).t() @ Y.covar(X, x, xp) @ X.mean(xp))

    def covar(self, Z, x, xp):
        X, Y = self.lhs, self.rhs
        assert isinstance(Z, GaussianProcessBase)
        return X.mean(x).t() @ Y.covar(Z, x, xp) + Y.mean(x).t() @ X.covar(Z, x, xp)

    def __str__(self):
        return "{!s} @ {!s}".format(self.lhs, self.rhs)


class GaussianProcessDetMatmulExpr(GaussianProcessExpr):
    """d(x)^T Y(x) for deterministic d (reference :170-199)."""

    def __init__(self, X, Y):
        assert isinstance(X, DeterministicGP) and isinstance(Y, GaussianProcessBase)
        assert X.t().shape == Y.shape
        self.lhs = X.t()
        self.rhs = Y

    @property
    def shape(self):
        return (1,)

    def mean(self, x):
        return self.lhs.mean(x).t() @ self.rhs.mean(x)

    def knl(self, x, xp):
        X, Y = self.lhs, self.rhs
        return X.mean(x).t() @ Y.knl(x, xp) @ X.mean(xp)

    def covar(self, Z, x, xp):
        assert isinstance(Z, GaussianProcessBase)
        return self.lhs.mean(x).t() @ self.rhs.covar(Z, x, xp)

    def __str__(self):
        return "{!s} @ {!s}".format(self.lhs, self.rhs)


class GaussianProcessMulExpr(GaussianProcessExpr):
    def __init__(self, X, a):
        assert isinstance(X, GaussianProcessBase)
        assert isinstance(a, (float, int, torch.Tensor))
        self.rhs = X
        self.α = a

    @property
    def shape(self):
        return self.rhs.shape

    def mean(self, x):
        return self.α * self.rhs.mean(x)

    def knl(self, x, xp):
        return (self.α ** 2) * self.rhs.knl(x, xp)

    def covar(self, Z, x, xp):
        assert isinstance(Z, GaussianProcessBase)
        return self.α * self.rhs.covar(Z, x, xp)

    def __str__(self):
        return "{!s} * {!s}".format(self.α, self.rhs)


class GaussianProcessTranspose(GaussianProcessExpr):
    def __init__(self, gp):
        assert isinstance(gp, GaussianProcessBase)
        assert 1 <= len(gp.shape) <= 2
        self.gp = gp

    @property
    def shape(self):
        if len(self.gp.shape) == 2:
            assert self.gp.shape[0] == 1
            return (self.gp.shape[1],)
        return (1, self.gp.shape[0])

    def mean(self, x):
        return self.gp.mean(x).t()

    def knl(self, x, xp):
        return self.gp.knl(x, xp)

    def covar(self, Y, x, xp):
        assert isinstance(Y, GaussianProcessBase)
        return self.gp.covar(Y, x, xp).t()

    def t(self):
        return self.gp

    def __str__(self):
        return "{!s}.t()".format(self.gp)


class GaussianProcess(GaussianProcessLeaf):
    """Leaf GP given by closures mean(x) (k,), knl(x, x') (k,k); cross-covariances with other leaves are registered."""

    def __init__(self, mean, knl, shape, assume_independence=False, name="{mean}"):
        self._mean = mean
        self._knl = knl
        self._shape = shape
        self._covars = dict()
        self.register_covar(self, self.knl)
        self.assume_independence = assume_independence
        self._name = name.format(mean=mean)

    @property
    def shape(self):
        return self._shape

    @property
    def dtype(self):
        owner = getattr(self._mean, '__self__', None)
        if owner is None and hasattr(self._mean, 'func'):
            owner = getattr(self._mean.func, '__self__', None)
        return getattr(owner, 'dtype', None)

    def to(self, dtype):
        for f in (self._mean, self._knl):
            owner = getattr(f, '__self__', None)
            if owner is None and hasattr(f, 'func'):
                owner = getattr(f.func, '__self__', None)
            if owner is not None and hasattr(owner, 'to'):
                owner.to(dtype=dtype)

    def mean(self, x):
        return self._mean(x)

    def knl(self, x, xp):
        return self._knl(x, xp)

    def covar(self, Z, x, xp):
        assert isinstance(Z, GaussianProcessBase)
        if isinstance(Z, GaussianProcess):
            if id(Z) in self._covars:
                return self._covars[id(Z)](x, xp)
            if self.assume_independence:
                return x.new_zeros(max(self.shape), max(Z.shape))
            raise ValueError("No covariance registered among two leaf GaussianProcesses: {!s} and {!s}".format(self, Z))
        if isinstance(Z, DeterministicGP):
            return x.new_zeros(max(self.shape), max(Z.shape))
        return Z.covar(self, x, xp).t()

    def register_covar(self, gp, covar_func):
        assert isinstance(gp, GaussianProcess)
        self._covars[id(gp)] = covar_func
        gp._covars[id(self)] = covar_func

    def __repr__(self):
        return "GaussianProcess(mean={}, knl={}, shape={})".format(self._mean, self._knl, self.shape)

    def __str__(self):
        return "GaussianProcess(name={})".format(self._name)


EPS = 2e-3


class GradientGP(GaussianProcessExpr):
    """The gradient process of a scalar GP f: mean = grad f.mean, knl = d^2 f.knl / dx dx'^T, covar = Jacobian."""

    def __init__(self, f, x_shape, grad_check=False, analytical_hessian=True):
        self.gp = f
        self.x_shape = x_shape
        self.grad_check = grad_check
        self.analytical_hessian = analytical_hessian

    @property
    def shape(self):
        return self.x_shape

    @property
    def dtype(self):
        return self.gp.dtype

    def to(self, dtype):
        self.gp.to(dtype)

    def mean(self, x):
        with variable_required_grad(x) as xg:
            return torch.autograd.grad(self.gp.mean(xg), xg)[0]

    def knl(self, x, xp, eigeps=EPS):
        f = self.gp
        if xp is x:
            xp = xp.detach().clone()
        if self.analytical_hessian:
            Hxx_k = t_hessian(f.knl, x, xp)
        else:
            raise NotImplementedError("numerical Hessians (analytical_hessian=False) are a debugging aid of the "
                                      "reference; the closed-form double backward is always available here")
        if torch.allclose(x, xp):
            Hs = 0.5 * (Hxx_k + Hxx_k.t())
            w, V = torch.linalg.eigh(Hs)
            assert (w > -eigeps).all(), " Hessian must be positive definite"
            if ((w > -eigeps) & (w < 0)).any():
                Hxx_k = V @ torch.diag(w.clamp_min(0)) @ V.t()
        return Hxx_k

    def covar(self, G, x, xp):
        """cov(grad f, g) given cov(f, g)."""
        with variable_required_grad(x) as xg:
            J = t_jac(self.gp.covar(G, xg, xp), xg)
        return J.t()

    def __str__(self):
        return "∇ {!s}".format(self.gp)
