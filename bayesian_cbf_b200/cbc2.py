"""Relative-degree-2 control barrier condition and the extraction of its affine / quadratic terms in u
(the functions of the reference's bayes_cbf/cbc2.py:7-63, same names and return conventions).

    cbc2(x; u) = L_{f+gu} (L_f h)(x) + k_0 h(x) + k_1 L_f h(x)

as a GP in x.  Its mean is affine in u and its kernel quadratic in u; `cbc2_quadratic_terms` recovers those
coefficients from one first- and one second-order Taylor expansion in u around an arbitrary point (exact because the
functions are affine / quadratic)."""
import math
from abc import ABC, abstractmethod

import torch

from .gp_algebra import DeterministicGP, GradientGP
from .misc import t_jac, variable_required_grad


def _taylor(func, u, order):
    """value, gradient [, Hessian] of a scalar function of u at u (gradient graph kept for the Hessian)."""
    with variable_required_grad(u) as ug:
        val = func(ug)
        grad = torch.autograd.grad(val, ug, create_graph=True)[0]
        hess = t_jac(grad, ug) if order == 2 else None
    return val.detach(), grad.detach(), (None if hess is None else hess.detach())


def cbc2_quadratic_terms(cbc2, x, u):
    """((A, b), (Q, p, r), mean(u), var(u)) with  cbc2(u).mean(x) = A u + b  and  cbc2(u).knl(x,x) = u^T Q u + p^T u + r."""
    mean_of = lambda up: cbc2(up).mean(x)
    var_of = lambda up: cbc2(up).knl(x, x)
    m0, A, _ = _taylor(mean_of, u, 1)
    b = m0 - A @ u
    v0, g, H = _taylor(var_of, u, 2)
    Q = H / 2
    p = g - H @ u                      # gradient of the quadratic at 0
    r = v0 - u @ Q @ u - p @ u
    for name, t in (('mean_A', A), ('mean_b', b), ('k_Q', Q), ('k_p', p), ('k_r', r)):
        assert not torch.isnan(t).any(), name
    return (A, b), (Q, p, r), mean_of(u), var_of(u)


def cbc2_gp(h, grad_h, learned_model, utest, k_α):
    n = learned_model.state_size
    f_gp, fu_gp = learned_model.f_func_gp(), learned_model.fu_func_gp(utest)
    h_gp = DeterministicGP(h, shape=(1,), name="h(x)")
    L1h = DeterministicGP(grad_h, shape=(n,), name="∇ h(x)").t() @ f_gp          # L_f h
    L2h = GradientGP(L1h, x_shape=(n,)).t() @ fu_gp                               # L_{f+gu} L_f h
    return L2h + h_gp * k_α[0] + L1h * k_α[1]


def cbc2_safety_factor(δ):
    assert δ < 0.5            # ask for more than 50 % safety
    factor = math.sqrt((1 - δ) / δ)
    assert factor > 1
    return factor


class RelDeg2Safety(ABC):
    """Mixin for a relative-degree-2 barrier: provide k_alpha, model, max_unsafe_prob, cbf, grad_cbf."""

    @property
    @abstractmethod
    def k_alpha(self):
        ...

    @property
    @abstractmethod
    def model(self):
        ...

    @property
    @abstractmethod
    def max_unsafe_prob(self):
        ...

    @abstractmethod
    def cbf(self, x):
        ...

    @abstractmethod
    def grad_cbf(self, x):
        ...

    def cbc(self, u0):
        return cbc2_gp(self.cbf, self.grad_cbf, self.model, u0, self.k_alpha)

    def safety_factor(self):
        return cbc2_safety_factor(self.max_unsafe_prob)
