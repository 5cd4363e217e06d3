"""Relative-degree-1 control barrier condition as a GP in x parameterised by u (reference bayes_cbf/cbc1.py:10-46):

    cbc(x; u) = grad_h(x)^T F(x)[1;u] + gamma h(x)

For the MVGP posterior this is affine in u in the mean and quadratic in u in the variance,
    mean = grad_h^T (Fbar + M_k) [1;u] + gamma h,      var = ([1;u]^T B_k [1;u]) (grad_h^T A grad_h),
which `cbc1_terms_batched` evaluates for many states at once on the GPU (bcbf_cbc1_terms) without autograd."""
import math
from abc import ABC, abstractmethod

from scipy.special import erfinv

from . import ops
from .gp_algebra import DeterministicGP


def cbc1_safety_factor(δ):
    assert δ < 0.5  # ask for more than 50% safety
    factor = math.sqrt(2) * erfinv(1 - 2 * δ)
    assert factor > 1
    return factor


class RelDeg1Safety(ABC):
    @property
    @abstractmethod
    def gamma(self):
        ...

    @property
    @abstractmethod
    def model(self):
        ...

    @abstractmethod
    def cbf(self, x):
        ...

    @abstractmethod
    def grad_cbf(self, x):
        ...

    @property
    @abstractmethod
    def max_unsafe_prob(self):
        ...

    def cbc(self, u0):
        h_gp = DeterministicGP(lambda x: self.gamma * self.cbf(x), shape=(1,), name="h(x)")
        grad_h_gp = DeterministicGP(self.grad_cbf, shape=(self.model.state_size,), name="∇ h(x)")
        fu_gp = self.model.fu_func_gp(u0)
        return grad_h_gp.t() @ fu_gp + h_gp

    def safety_factor(self):
        return cbc1_safety_factor(self.max_unsafe_prob)


def cbc1_terms_batched(Mk, Bk, A, grad_h, h, gamma, Fbar=None):
    """Closed-form CBC terms for Q states (CUDA): returns dict(bfe (Q,m), e (Q,), Asq (Q,p,p), A_socp (Q,p,m),
    bfb (Q,p), status (Q,)) with  mean = bfe^T u + e  and  ||A_socp u + bfb||^2 = var(u)."""
    bfe, e, Asq, A_socp, bfb, status = ops.cbc1_terms(Mk.contiguous(), Bk.contiguous(), A.contiguous(),
                                                      grad_h.contiguous(), h.contiguous(), gamma,
                                                      None if Fbar is None else Fbar.contiguous())
    return dict(bfe=bfe, e=e, Asq=Asq, A_socp=A_socp, bfb=bfb, status=status)
