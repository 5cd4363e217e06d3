"""TEST INFRASTRUCTURE (run as a subprocess by tests/test_reference_known_answers.py, only where /root/reference exists).

The drop-in claim checked from the consumer side: the REFERENCE's own gp_algebra.py, cbc1.py, cbc2.py and
controllers.SOCPController.convert_cbc_terms_to_socp_terms (imported unmodified from /root/reference; matplotlib / cvxpy
stubbed by oracle/gpytorch_shim.install) consume the closures of `bayesian_cbf_b200.ControlAffineRegressor`
(f_func_mean / f_func_knl / fu_func_mean / fu_func_knl / covar_fu_f, wired into the reference's GaussianProcess leaves
exactly as the reference's regressor wires its own, control_affine_model.py:707-744, 793-798) and must reproduce the
goldens that the reference's regressor produced (tests/golden/ref_predict_*.npz: cbc1_*, socp_*; ref_cbc2_pendulum_f64).

    python tests/ref_consumers_check.py cpu|cuda
On 'cpu' the regressor runs over the torch stand-ins of tests/fake_ops.py (host logic); on 'cuda' over libbcbf.so."""
import os
import sys
from functools import partial

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import gpytorch_shim  # noqa: E402

gpytorch_shim.install()
import bayes_cbf.gp_algebra as rga  # noqa: E402  (the reference's)
import bayes_cbf.cbc1 as rcbc1  # noqa: E402
import bayes_cbf.cbc2 as rcbc2  # noqa: E402
import bayes_cbf.controllers as rctl  # noqa: E402

from tests.golden_util import PREDICT_CASES, T, load  # noqa: E402


class _ContextPatch:
    def setattr(self, obj, name, value):
        self._undo.append((obj, name, getattr(obj, name)))
        setattr(obj, name, value)

    def __init__(self):
        self._undo = []


class RefFacing:
    """The reference-side view of our regressor: reference GaussianProcess leaves over OUR closures."""

    def __init__(self, reg):
        self.reg = reg
        self._f = rga.GaussianProcess(reg.f_func_mean, reg.f_func_knl, (reg.x_dim,), name="f")

    @property
    def state_size(self):
        return self.reg.x_dim

    def f_func_gp(self):
        return self._f

    def fu_func_gp(self, u):
        gp = rga.GaussianProcess(mean=partial(self.reg.fu_func_mean, u), knl=partial(self.reg.fu_func_knl, u),
                                 shape=(self.reg.x_dim,), name="F(.)u")
        gp.register_covar(self._f, partial(self.reg.covar_fu_f, u))
        return gp


def close(got, want, rel, what, scale=None):
    got = got.detach().cpu().double().numpy().reshape(np.asarray(want).shape)
    want = np.asarray(want, dtype=np.float64)
    scale = max(np.abs(want).max(), 1e-300) if scale is None else scale
    err = np.abs(got - want).max() / scale
    assert err < rel, '%s: %.3e >= %.1e' % (what, err, rel)
    return err


def main(dev):
    if dev == 'cpu':
        from tests import fake_ops
        fake_ops.installed(_ContextPatch()).__enter__()
    from bayesian_cbf_b200.control_affine_model import ControlAffineRegressor
    worst = 0.0
    for case in PREDICT_CASES:
        d = load(case)
        f64 = str(d['dtype']) == 'float64'
        tol = 1e-9 if f64 else 1e-4
        dt = torch.float64 if f64 else torch.float32
        n, m = int(d['n']), int(d['m'])
        reg = ControlAffineRegressor(n, m, device=dev)
        if f64:
            reg.model.double()
        reg.set_hyperparameters(d['hb_lengthscale'], d['hb_outputscale'], d['hb_A'], d['hb_B'], d['hb_C'])
        reg.fit(T(d['X']), T(d['U']), T(d['Xdot']), training_iter=0)
        reg.set_jitter_source([T(d['base_first_rand%d' % i]) for i in range(int(d['base_first_nrand']))])
        x0, u0 = T(d['cbc1_x']).to(dt).to(dev), T(d['cbc1_u0']).to(dt).to(dev)
        gamma = float(d['cbc1_gamma'])

        facing, gam = RefFacing(reg), gamma

        class Safety(rcbc1.RelDeg1Safety):               # the reference's ABC, our model behind it
            @property
            def gamma(self):
                return gam

            @property
            def model(self):
                return facing

            @property
            def max_unsafe_prob(self):
                return 0.01

            def cbf(self, x):
                return (x * x).sum() - 0.3

            def grad_cbf(self, x):
                return 2 * x

        safety = Safety()
        (bfe, e), (V, bfv, v), mean, var = rcbc2.cbc2_quadratic_terms(safety.cbc, x0, u0)
        for name, got in dict(bfe=bfe, e=e, mean=mean).items():
            worst = max(worst, close(got, d['cbc1_' + name], tol, case + ' ' + name))
        vscale = float(np.abs(d['cbc1_V']).max())
        for name, got in dict(V=V, bfv=bfv, v=v, var=var).items():
            worst = max(worst, close(got, d['cbc1_' + name], tol, case + ' ' + name, vscale))
        A_s, bfb, bfc, dd = rctl.SOCPController.convert_cbc_terms_to_socp_terms(
            bfe.float().cpu(), e.float().reshape(()).cpu(), V.float().cpu(), bfv.float().cpu(), v.float().reshape(()).cpu(), 1)
        for name, got in dict(A=A_s, bfb=bfb, bfc=bfc, d=dd).items():
            close(got, d['socp_' + name], 1e-4, case + ' socp ' + name)
        assert abs(safety.safety_factor() - rcbc1.cbc1_safety_factor(0.01)) == 0
    # relative degree two: the reference's cbc2_gp (GradientGP with double backward through OUR closures)
    d = load('ref_cbc2_pendulum_f64')
    reg = ControlAffineRegressor(2, 1, device=dev)
    reg.model.double()
    reg.set_hyperparameters(d['h_lengthscale'], d['h_outputscale'], d['h_A'], d['h_B'], d['h_C'])
    reg.fit(T(d['X']), T(d['U']), T(d['Xdot']), training_iter=0)
    reg.set_jitter_source([T(d['jitter'])])
    h = lambda x: (x[0] - 0.2) ** 2 + 0.5 * x[1] ** 2 - 0.1
    grad_h = lambda x: torch.stack([2 * (x[0] - 0.2), x[1]])
    k_alpha, x0, u0 = T(d['k_alpha']).to(dev), T(d['x0']).to(dev), T(d['u0']).to(dev)
    model = RefFacing(reg)
    cbc2 = rcbc2.cbc2_gp(h, grad_h, model, u0, k_alpha)
    worst = max(worst, close(cbc2.mean(x0), d['cbc2_mean'], 1e-8, 'cbc2 mean'))
    worst = max(worst, close(cbc2.knl(x0, x0), d['cbc2_knl'], 1e-8, 'cbc2 knl'))
    (bfe, e), (V, bfv, v), mean, var = rcbc2.cbc2_quadratic_terms(
        lambda u: rcbc2.cbc2_gp(h, grad_h, model, u, k_alpha), x0, u0)
    for name, got in dict(bfe=bfe, e=e, V=V, bfv=bfv, v=v, mean=mean, var=var).items():
        close(got, d['q_' + name], 1e-7, 'q_' + name)
    print('REF_CONSUMERS_OK device=%s worst_rel=%.2e' % (dev, worst))


if __name__ == '__main__':
    main(sys.argv[1] if len(sys.argv) > 1 else 'cpu')
