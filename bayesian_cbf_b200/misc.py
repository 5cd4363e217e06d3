"""Host-side numeric helpers the MVGP path needs (the counterpart of the reference's bayes_cbf/misc.py,
restricted to what sits on the hot path: SURVEY §2 "misc.py numeric helpers").

    torch_kron                 bayes_cbf/misc.py:80-106
    t_jac / t_hessian          bayes_cbf/misc.py:47-53, 236-245
    variable_required_grad     bayes_cbf/misc.py:219-233
    get_affine_terms / get_quadratic_terms   bayes_cbf/misc.py:268-285
    DynamicsModel / BayesianDynamicsModel / ZeroDynamicsModel   bayes_cbf/misc.py:109-213

These are tiny tensor-shape utilities and autograd drivers; the arithmetic they differentiate through runs in
the CUDA ops of `bayesian_cbf_b200.autograd_ops`.
"""
import math
from abc import ABC, abstractmethod
from contextlib import contextmanager

import torch


def to_numpy(x):
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().double().numpy()
    return x


def torch_kron(A, B, batch_dims=1):
    """Kronecker product over the trailing dims, broadcasting over `batch_dims` leading dims.

    >>> B = torch.rand(5, 3, 3); A = torch.rand(5, 2, 2)
    >>> bool(torch.allclose(torch_kron(A, B)[1, :3, :3], A[1, 0, 0] * B[1]))
    True
    >>> torch_kron(torch.rand(2, 3), torch.rand(3, 2), batch_dims=0).shape
    torch.Size([6, 6])
    """
    assert A.ndim == B.ndim
    ta, tb = A.shape[batch_dims:], B.shape[batch_dims:]
    # interleave: A gets (s, 1) per trailing axis, B gets (1, s)
    A_ = A.reshape(*A.shape[:batch_dims], *[d for s in ta for d in (s, 1)])
    B_ = B.reshape(*B.shape[:batch_dims], *[d for s in tb for d in (1, s)])
    out = A_ * B_
    lead = out.shape[:batch_dims]
    return out.reshape(*lead, *[x * y for x, y in zip(ta, tb)])


def isleaf(x):
    return x.grad_fn is None


@contextmanager
def variable_required_grad(x):
    """Context in which `x` (or a detached clone of it, when x is not a leaf) requires grad."""
    was = x.requires_grad
    leaf = x if isleaf(x) else x.detach().clone()
    try:
        yield leaf.requires_grad_(True)
    finally:
        if isleaf(x):
            x.requires_grad_(was)


def t_jac(f_x, x, retain_graph=False, **kw):
    """Jacobian of a vector (row per output) or gradient of a scalar w.r.t. x."""
    if f_x.ndim:
        rows = [torch.autograd.grad(f_x[i], x, retain_graph=True, **kw)[0].unsqueeze(0) for i in range(f_x.shape[0])]
        return torch.cat(rows, dim=0)
    return torch.autograd.grad(f_x, x, retain_graph=retain_graph, **kw)[0]


def t_hessian(f, x, xp, grad_check=True):
    """Mixed second derivative d^2 f(x, xp) / dx dxp^T of a scalar-valued f."""
    with variable_required_grad(x):
        with variable_required_grad(xp):
            g = torch.autograd.grad(f(x, xp), x, create_graph=True)[0]
            return t_jac(g, xp)


def get_affine_terms(func, x):
    """func(x) = linear @ x + const around x (exact when func is affine)."""
    with variable_required_grad(x):
        f_x = func(x)
        linear = torch.autograd.grad(f_x, x, create_graph=True)[0]
    with torch.no_grad():
        const = f_x - linear @ x
    return linear, const


def get_quadratic_terms(func, x):
    """func(x) = x^T quad x + linear @ x + const around x (exact when func is quadratic)."""
    with variable_required_grad(x):
        f_x = func(x)
        linear_more = torch.autograd.grad(f_x, x, create_graph=True)[0]
        quad = t_jac(linear_more, x) / 2
    with torch.no_grad():
        linear = linear_more - 2 * quad @ x
        const = f_x - x @ quad @ x - linear @ x
    return quad, linear, const


def random_psd(m):
    M = torch.rand(m, m)
    return M @ M.T


def normalize_radians(theta):
    return (theta + math.pi) % (2 * math.pi) - math.pi


class DynamicsModel(ABC):
    """xdot = f(x) + g(x) u"""

    def __init__(self):
        self._state = None

    @property
    @abstractmethod
    def ctrl_size(self):
        """dimension of u"""

    @property
    @abstractmethod
    def state_size(self):
        """dimension of x"""

    @abstractmethod
    def f_func(self, X):
        """f(X) for X (d, n) or (n,)"""

    @abstractmethod
    def g_func(self, X):
        """g(X): (d, n, m) or (n, m)"""

    def normalize_state(self, X_in):
        return X_in

    def forward(self, x, u):
        X_b = x.unsqueeze(0) if x.ndim == 1 else x
        if u.ndim == 1:
            U_b = u.unsqueeze(0).unsqueeze(-1)
        elif u.ndim == 2:
            U_b = u.unsqueeze(0)
        else:
            U_b = u
        Xdot_b = self.f_func(X_b) + self.g_func(X_b).bmm(U_b).squeeze(-1)
        return Xdot_b.squeeze(0) if x.ndim == 1 else Xdot_b

    def step(self, u, dt):
        x = self._state
        xdot = self.forward(x, u)
        xtp1 = self.normalize_state(x + xdot * dt)
        self._state = xtp1
        return dict(x=xtp1, xdot=xdot)

    def set_init_state(self, x0):
        self._state = x0.clone()

    def F_func(self, X):
        return torch.cat([self.f_func(X).unsqueeze(-1), self.g_func(X)], dim=-1)


class BayesianDynamicsModel(DynamicsModel):
    @abstractmethod
    def fu_func_gp(self, U):
        """GaussianProcessBase of F(.)[1;u]"""


class ZeroDynamicsModel(DynamicsModel):
    def __init__(self, m, n):
        super().__init__()
        self.m = m
        self.n = n

    @property
    def ctrl_size(self):
        return self.m

    @property
    def state_size(self):
        return self.n

    def f_func(self, X):
        return (torch.zeros((self.n,)) if X.dim() <= 1 else torch.zeros(X.shape)) * X

    def g_func(self, X):
        return torch.zeros((*X.shape, self.m)) * X.unsqueeze(-1)
