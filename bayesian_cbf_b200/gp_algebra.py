"""Algebra of (vector-valued) Gaussian processes given as closures — the public surface of the reference's
bayes_cbf/gp_algebra.py (`GaussianProcess`, `DeterministicGP`, `GradientGP`, operators `+`, scalar `*` and `/`, `@`,
`.t()`, methods `mean(x)`, `knl(x, x')`, `covar(Z, x, x')`, `sample`, `register_covar`), consumed by cbc1 / cbc2 and the
controllers.

Design.  Every expression that is *linear* in its GP operands — a sum, a scalar multiple, a deterministic row vector
dotted with a GP — is one node type here, `LinearCombination`: a list of (coefficient, operand) pairs with

    mean(x)          = sum_i c_i(x) . E[Z_i(x)]
    knl(x, x')       = sum_i c_i(x) K_i(x,x') c_i(x')^T  +  sum_{i != j} c_i(x) cov(Z_i(x), Z_j(x')) c_j(x')^T
    covar(W, x, x')  = sum_i c_i(x) cov(Z_i(x), W(x'))

where a coefficient is a number or a deterministic function of x (a `DeterministicGP`, applied as d(x)^T).  The two
non-linear constructions keep their own nodes: `ProductOfGPs` (X^T Y for jointly Gaussian X, Y, second-order moment
matching — reference :133-168) and `GradientGP` (:319-405).  Operand cross-covariances are resolved exactly like the
reference does (leaf registry; a leaf asked about an expression defers to the expression and transposes), because the
reference's numbers — pinned by tests/golden/ref_cbc2_pendulum_f64.npz — depend on that order of evaluation.

Everything here is shape bookkeeping on n- / m-sized tensors plus autograd drivers; the numbers come from the
regressor closures (CUDA kernels, differentiable through bayesian_cbf_b200.autograd_ops).

Conscious deviation: the PSD clean-up of `GradientGP.knl` used `torch.eig` (removed from torch,
reference :384-393); here `torch.linalg.eigh` on the symmetrised Hessian with the reconstruction V diag(w) V^T.
"""
from abc import ABC, abstractmethod
from numbers import Number

import torch
from torch.distributions import MultivariateNormal

from .misc import t_hessian, t_jac, variable_required_grad


def _width(shape):
    return max(shape)


class GaussianProcessBase(ABC):
    """Operator overloading shared by leaves and expression nodes."""

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

    def __add__(self, other):
        assert isinstance(other, GaussianProcessBase) and self.shape == other.shape
        return LinearCombination([(1, self), (1, other)], self.shape, symbol='+')

    def __mul__(self, a):
        assert isinstance(a, (Number, torch.Tensor))
        return LinearCombination([(a, self)], self.shape, symbol='*')

    def __truediv__(self, a):
        return self * (1 / a)

    __div__ = __truediv__

    def __matmul__(self, other):
        assert isinstance(other, GaussianProcessBase)
        if isinstance(self, DeterministicGP):
            row = self.t()                      # `d.t() @ Y` arrives here as (row vector) @ Y: undo the transpose
            assert row.shape == other.shape
            return LinearCombination([(row, other)], (1,), symbol='@')
        return ProductOfGPs(self, other)

    def t(self):
        return Transposed(self)


# ------------------------------------------------------------------------------------------------------------ leaves
class _Leaf(GaussianProcessBase):
    @classmethod
    def isleaf(cls):
        return True

    def name(self):
        return self._name


class DeterministicGP(_Leaf):
    """A deterministic function of x seen as a GP with zero kernel (also used as a coefficient)."""

    def __init__(self, mean, shape, name="{mean}"):
        assert len(shape) <= 2 and (len(shape) == 1 or min(shape) == 1)
        self._mean, self._shape = mean, shape
        self._name = name.format(mean=mean)

    @property
    def shape(self):
        return self._shape

    def mean(self, x):
        return self._mean(x)

    def knl(self, x, xp):
        return x.new_zeros(_width(self._shape), _width(self._shape))

    def covar(self, Z, x, xp):
        assert isinstance(Z, GaussianProcessBase)
        if isinstance(Z, DeterministicGP):
            return x.new_zeros(_width(self._shape), _width(Z.shape))
        return Z.covar(self, x, xp).t()

    def sample(self, x, sample_shape=torch.Size([])):
        return self.mean(x).expand(*sample_shape, -1)

    def __repr__(self):
        return "DeterministicGP(mean={})".format(self._mean)

    __str__ = lambda self: "DeterministicGP(name={})".format(self._name)


class GaussianProcess(_Leaf):
    """Leaf GP given by closures mean(x) (k,), knl(x, x') (k,k).  Cross-covariances with other leaves are looked up in
    a registry (symmetric registration, as in the reference :305-308); unregistered pairs raise unless
    `assume_independence`."""

    def __init__(self, mean, knl, shape, assume_independence=False, name="{mean}"):
        self._mean, self._knl, self._shape = mean, knl, shape
        self._covars = {}
        self.assume_independence = assume_independence
        self._name = name.format(mean=mean)
        self.register_covar(self, self.knl)      # covariance with itself is the kernel

    @property
    def shape(self):
        return self._shape

    @staticmethod
    def _owner(f):
        f = getattr(f, 'func', f)                # functools.partial of a bound method
        return getattr(f, '__self__', None)

    @property
    def dtype(self):
        return getattr(self._owner(self._mean), 'dtype', None)

    def to(self, dtype):
        for f in (self._mean, self._knl):
            owner = self._owner(f)
            if owner is not None and hasattr(owner, 'to'):
                owner.to(dtype=dtype)

    def mean(self, x):
        return self._mean(x)

    def knl(self, x, xp):
        return self._knl(x, xp)

    def register_covar(self, gp, covar_func):
        assert isinstance(gp, GaussianProcess)
        self._covars[id(gp)] = covar_func
        gp._covars[id(self)] = covar_func

    def covar(self, Z, x, xp):
        assert isinstance(Z, GaussianProcessBase)
        if isinstance(Z, DeterministicGP):
            return x.new_zeros(_width(self.shape), _width(Z.shape))
        if not isinstance(Z, GaussianProcess):
            return Z.covar(self, x, xp).t()      # an expression knows how to expand itself against a leaf
        fn = self._covars.get(id(Z))
        if fn is not None:
            return fn(x, xp)
        if self.assume_independence:
            return x.new_zeros(_width(self.shape), _width(Z.shape))
        raise ValueError("No covariance registered among two leaf GaussianProcesses: {!s} and {!s}".format(self, Z))

    def __repr__(self):
        return "GaussianProcess(mean={}, knl={}, shape={})".format(self._mean, self._knl, self.shape)

    __str__ = lambda self: "GaussianProcess(name={})".format(self._name)


# ------------------------------------------------------------------------------------------------------------- nodes
class _Node(GaussianProcessBase):
    @classmethod
    def isleaf(cls):
        return False


class LinearCombination(_Node):
    """sum_i c_i . Z_i with c_i a number or a deterministic row (DeterministicGP whose mean(x) is applied as c(x)^T)."""

    def __init__(self, terms, shape, symbol='+'):
        self.terms, self._shape, self.symbol = list(terms), shape, symbol

    @property
    def shape(self):
        return self._shape

    # --- coefficient application ---------------------------------------------------------------------------------------
    @staticmethod
    def _left(c, x, M):
        """c(x) applied from the left."""
        if isinstance(c, DeterministicGP):
            return c.mean(x).t() @ M
        return M if (isinstance(c, Number) and c == 1) else c * M

    @staticmethod
    def _right(c, xp, M):
        """c(x')^T applied from the right."""
        if isinstance(c, DeterministicGP):
            return M @ c.mean(xp)
        return M if (isinstance(c, Number) and c == 1) else M * c

    def mean(self, x):
        out = None
        for c, Z in self.terms:
            term = self._left(c, x, Z.mean(x))
            out = term if out is None else out + term
        return out

    def knl(self, x, xp):
        out = None
        for c, Z in self.terms:                                    # own kernels first ...
            term = self._right(c, xp, self._left(c, x, Z.knl(x, xp)))
            out = term if out is None else out + term
        for i in reversed(range(len(self.terms))):                  # ... then the cross terms (Y with X before X with Y)
            for j in range(len(self.terms)):
                if i == j:
                    continue
                (ci, Zi), (cj, Zj) = self.terms[i], self.terms[j]
                out = out + self._right(cj, xp, self._left(ci, x, Zi.covar(Zj, x, xp)))
        return out

    def covar(self, W, x, xp):
        assert isinstance(W, GaussianProcessBase)
        out = None
        for c, Z in self.terms:
            term = self._left(c, x, Z.covar(W, x, xp))
            out = term if out is None else out + term
        return out

    def __str__(self):
        if self.symbol == '+':
            return " + ".join(str(Z) for _, Z in self.terms)
        c, Z = self.terms[0]
        return "{!s} {} {!s}".format(c, self.symbol, Z)

    # the reference exposes the operands of its binary nodes; keep those handles
    @property
    def lhs(self):
        return self.terms[0][0] if self.symbol == '@' else self.terms[0][1]

    @property
    def rhs(self):
        return self.terms[-1][1]


class ProductOfGPs(_Node):
    """X^T Y for two jointly Gaussian vector GPs: mean and kernel by second-order moment matching (reference :133-168;
    its last kernel term carries a FIXME there and is reproduced as written)."""

    def __init__(self, X, Y):
        assert X.shape[-1] == Y.shape[0]
        self.lhs, self.rhs = X.t(), Y

    @property
    def shape(self):
        return (1,)

    def mean(self, x):
        X, Y = self.lhs, self.rhs
        return X.mean(x).t() @ Y.mean(x) + 0.5 * X.covar(Y, x, x).trace() + 0.5 * Y.covar(X, x, x).trace()

    def knl(self, x, xp):
        X, Y = self.lhs, self.rhs
        mx, mxp, my, myp = X.mean(x), X.mean(xp), Y.mean(x), Y.mean(xp)
        return (2 * X.covar(Y, x, xp).trace() ** 2
                + my.t() @ X.knl(x, xp) @ myp
                + mx.t() @ Y.knl(x, xp) @ mxp
                + 2 * my.t() @ Y.covar(X, x, xp) @ mxp)

    def covar(self, Z, x, xp):
        assert isinstance(Z, GaussianProcessBase)
        X, Y = self.lhs, self.rhs
        return X.mean(x).t() @ Y.covar(Z, x, xp) + Y.mean(x).t() @ X.covar(Z, x, xp)

    def __str__(self):
        return "{!s} @ {!s}".format(self.lhs, self.rhs)


class Transposed(_Node):
    """Shape bookkeeping: (k,) <-> (1,k).  Means transpose, kernels stay, cross-covariances transpose."""

    def __init__(self, gp):
        assert isinstance(gp, GaussianProcessBase) and 1 <= len(gp.shape) <= 2
        self.gp = gp

    @property
    def shape(self):
        s = self.gp.shape
        if len(s) == 2:
            assert s[0] == 1
            return (s[1],)
        return (1, s[0])

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


# `d.t()` of a DeterministicGP must stay usable as a coefficient: give it the leaf's interface
def _det_t(self):
    if len(self._shape) == 2:
        assert self._shape[0] == 1
        shape = (self._shape[1],)
    else:
        shape = (1, self._shape[0])
    return DeterministicGP(lambda x: self._mean(x).t(), shape, name=self._name + ".t()")


DeterministicGP.t = _det_t

EPS = 2e-3


class GradientGP(_Node):
    """The gradient process of a scalar GP f: mean = grad f.mean, knl = d^2 f.knl / dx dx'^T, covar = Jacobian of
    f.covar — all by autograd through the closures (closed-form derivative kernels underneath)."""

    def __init__(self, f, x_shape, grad_check=False, analytical_hessian=True):
        self.gp, self.x_shape = f, x_shape
        self.grad_check, self.analytical_hessian = grad_check, analytical_hessian

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
        if not self.analytical_hessian:
            raise NotImplementedError("numerical Hessians (analytical_hessian=False) are a debugging aid of the "
                                      "reference; the closed-form double backward is always available here")
        if xp is x:
            xp = xp.detach().clone()
        H = t_hessian(self.gp.knl, x, xp)
        if torch.allclose(x, xp):
            w, V = torch.linalg.eigh(0.5 * (H + H.t()))
            assert (w > -eigeps).all(), " Hessian must be positive definite"
            if (w < 0).any():
                H = V @ torch.diag(w.clamp_min(0)) @ V.t()
        return H

    def covar(self, G, x, xp):
        with variable_required_grad(x) as xg:
            return t_jac(self.gp.covar(G, xg, xp), xg).t()

    def __str__(self):
        return "∇ {!s}".format(self.gp)


# names used by the reference for its node classes
GaussianProcessExpr = _Node
GaussianProcessLeaf = _Leaf
GaussianProcessMatmulExpr = ProductOfGPs
GaussianProcessTranspose = Transposed
