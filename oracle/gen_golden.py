"""TEST INFRASTRUCTURE ONLY — generate golden vectors by running the UNMODIFIED reference.

Run in the build container (where /root/reference exists):

    python oracle/gen_golden.py            # writes tests/golden/*.npz

The reference source files are imported from /root/reference over the dense gpytorch stand-in of
`oracle/gpytorch_shim.py`; nothing is copied.  Every `torch.rand` draw made inside the reference during a
recorded call (the Cholesky jitter of `make_psd`, control_affine_model.py:907-910 and :1089) is captured
and stored next to the outputs so that the oracle restatement and the CUDA path can be fed the same jitter.
"""
import os
import sys
from functools import partial

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import gpytorch_shim  # noqa: E402

gpytorch_shim.install()

from bayes_cbf.control_affine_model import (ControlAffineRegressor, ControlAffineRegressorExact,  # noqa: E402
                                            ControlAffineExactGP)
from bayes_cbf.matrix_variate_multitask_kernel import (HetergeneousMatrixVariateKernel,  # noqa: E402
                                                       MatrixVariateIndexKernel)
from bayes_cbf.gp_algebra import DeterministicGP, GradientGP  # noqa: E402
from bayes_cbf.cbc2 import cbc2_quadratic_terms, cbc2_gp  # noqa: E402
import bayes_cbf.controllers as ref_controllers  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')


class RandRecorder:
    """Context manager recording every torch.rand draw (in order)."""

    def __enter__(self):
        self.draws = []
        self._orig = torch.rand

        def rec(*a, **k):
            r = self._orig(*a, **k)
            self.draws.append(r.detach().clone())
            return r
        torch.rand = rec
        return self

    def __exit__(self, *a):
        torch.rand = self._orig
        return False


def np64(t):
    return t.detach().cpu().double().numpy()


def randomise_hyper(reg, gen):
    """Give the model non-trivial hyper-parameters (fit() is parity-unpinned, so we set them directly)."""
    with torch.no_grad():
        for name, prm in reg.model.named_parameters():
            prm.copy_(0.5 * torch.randn(prm.shape, generator=gen, dtype=prm.dtype))


def extract_hyper(reg):
    m = reg.model
    p, n = m.matshape
    return dict(
        lengthscale=np64(m.input_covar.base_kernel.lengthscale.reshape(-1)),
        outputscale=np64(m.input_covar.outputscale.reshape(())),
        A=np64(m.task_covar.U.covar_matrix.evaluate()),
        B=np64(m.task_covar.V.covar_matrix.evaluate()),
        C=np64(torch.stack([bm.constant.reshape(()) for bm in m.mean_module.base_means]).reshape(p, n)),
    )


def make_data(n, m, N, b, gen, dtype, scale=1.0):
    X = scale * (2 * torch.rand(N, n, generator=gen, dtype=dtype) - 1)
    U = 2 * torch.rand(N, m, generator=gen, dtype=dtype) - 1
    Xdot = torch.randn(N, n, generator=gen, dtype=dtype)
    Xt = scale * (2 * torch.rand(b, n, generator=gen, dtype=dtype) - 1)
    Ut = 2 * torch.rand(b, m, generator=gen, dtype=dtype) - 1
    Xtp = scale * (2 * torch.rand(b, n, generator=gen, dtype=dtype) - 1)
    Utp = 2 * torch.rand(b, m, generator=gen, dtype=dtype) - 1
    return X, U, Xdot, Xt, Ut, Xtp, Utp


def case_predict(name, n, m, N, b, seed, dtype=torch.float64, fit_iters=0, rank_one=False, scale=1.0):
    torch.set_default_dtype(dtype)
    gen = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    X, U, Xdot, Xt, Ut, Xtp, Utp = make_data(n, m, N, b, gen, dtype, scale)
    out = dict(X=np64(X), U=np64(U), Xdot=np64(Xdot), Xt=np64(Xt), Ut=np64(Ut), Xtp=np64(Xtp), Utp=np64(Utp),
               n=n, m=m, dtype=str(dtype).split('.')[-1])
    mc = partial(ControlAffineExactGP, rank=1) if rank_one else ControlAffineExactGP
    base = ControlAffineRegressor(n, m, device='cpu', model_class=mc)
    exact = ControlAffineRegressorExact(n, m, device='cpu', model_class=mc)
    if dtype is torch.float64:
        base.model.double()
        exact.model.double()
    randomise_hyper(base, gen)
    exact.model.load_state_dict = None  # never used; keep nn.Module API untouched
    with torch.no_grad():
        for (_, p_src), (_, p_dst) in zip(base.model.named_parameters(), exact.model.named_parameters()):
            p_dst.copy_(p_src)
    # prior-only predictions (no train data): control_affine_model.py:495-506, 1024-1026
    pm, pc = base.custom_predict(Xt, Ut, Xtestp_in=Xtp, Utestp_in=Utp)
    out['prior_base_mean'], out['prior_base_cov'] = np64(pm), np64(pc)
    pm, pA, pB = exact._custom_predict_matrix(Xt, Xtp)
    out['prior_exact_Mk'], out['prior_exact_BkXX'] = np64(pm), np64(pB)
    out['hyper_prior'] = extract_hyper(base)
    # "fit": sets the train data and runs fit_iters Adam steps of the (shimmed, parity-unpinned) MLL
    base.fit(X, U, Xdot, training_iter=fit_iters)
    exact.fit(X, U, Xdot, training_iter=fit_iters)
    hb, he = extract_hyper(base), extract_hyper(exact)
    for k, v in hb.items():
        out['hb_' + k] = v
    for k, v in he.items():
        out['he_' + k] = v
    for k, v in out.pop('hyper_prior').items():
        out['hp_' + k] = v

    def record(tag, fn):
        with RandRecorder() as rr:
            res = fn()
        res = res if isinstance(res, tuple) else (res,)
        for i, r in enumerate(res):
            out['%s_out%d' % (tag, i)] = np64(r)
        for i, d in enumerate(rr.draws):
            out['%s_rand%d' % (tag, i)] = np64(d)
        out['%s_nrand' % tag] = len(rr.draws)

    # base class: first call draws the factor jitter (cached afterwards, :379-385)
    record('base_first', lambda: base.custom_predict(Xt, Ut))
    record('base_cross', lambda: base.custom_predict(Xt, Ut, Xtestp_in=Xtp, Utestp_in=Utp))
    record('base_nocov', lambda: base.custom_predict(Xt, Ut, compute_cov=False))
    record('base_uhfill0', lambda: base.custom_predict(Xt, Ut, UHfill=0))
    record('base_noU', lambda: base.custom_predict(Xt))
    record('base_svar', lambda: base.custom_predict(Xt, Ut, scalar_var_only=True))
    record('base_f_func_mean', lambda: base.f_func_mean(Xt))
    record('base_fu_func_mean', lambda: base.fu_func_mean(Ut, Xt))
    record('base_fu_func_knl', lambda: base.fu_func_knl(Ut[0], Xt[0], Xtp[0]))
    record('base_covar_fu_f', lambda: base.covar_fu_f(Ut[0], Xt[0], Xtp[0]))
    record('base_f_func_knl', lambda: base.f_func_knl(Xt[0], Xtp[0]))
    # Exact
    record('exact_matrix', lambda: exact._custom_predict_matrix(Xt))
    record('exact_predict', lambda: exact.custom_predict(Xt, Ut))
    record('exact_predict_cross', lambda: exact.custom_predict(Xt, Ut, Xtestp_in=Xtp, Utestp_in=Utp))
    record('exact_fullmat', lambda: exact.custom_predict_fullmat(Xt))
    record('exact_nocov', lambda: exact.custom_predict(Xt, Ut, compute_cov=False))
    record('exact_b1', lambda: exact._custom_predict_matrix(Xt[:1]))
    # the factor itself
    out['base_L'] = np64(base._cache['perturbed_cholesky'])
    out['exact_L'] = np64(exact._cache['perturbed_cholesky'])

    # relative-degree-1 CBC through the reference's gp_algebra + autograd (cbc1.py:38-46, cbc2.py:7-23)
    x0 = Xt[0].clone()
    u0 = Ut[0].clone()
    gamma = 0.7
    cbf = lambda x: (x * x).sum() - 0.3
    grad_cbf = lambda x: 2 * x

    def cbc(model, u):
        h_gp = DeterministicGP(lambda x: gamma * cbf(x), shape=(1,), name="h(x)")
        grad_h_gp = DeterministicGP(grad_cbf, shape=(n,), name="grad h(x)")
        return grad_h_gp.t() @ model.fu_func_gp(u) + h_gp

    (bfe, e), (V, bfv, v), mean, var = cbc2_quadratic_terms(partial(cbc, base), x0, u0)
    out['cbc1_x'], out['cbc1_u0'], out['cbc1_gamma'] = np64(x0), np64(u0), gamma
    for k_, v_ in dict(bfe=bfe, e=e, V=V, bfv=bfv, v=v, mean=mean, var=var).items():
        out['cbc1_' + k_] = np64(v_)
    A_s, bfb, bfc, d = ref_controllers.SOCPController.convert_cbc_terms_to_socp_terms(
        bfe.float(), e.float().reshape(()), V.float(), bfv.float(), v.float().reshape(()), 1)
    out['socp_A'], out['socp_bfb'], out['socp_bfc'], out['socp_d'] = map(np64, (A_s, bfb, bfc, d))
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)
    print('wrote', name, len(out), 'arrays')
    return base, exact


def case_kernel(name, n, m, Ntr, Nte, seed):
    """HetergeneousMatrixVariateKernel / mean on train-only, test-only and mixed inputs with the real
    RBF + IndexKernel modules (matrix_variate_multitask_kernel.py:187-204)."""
    torch.set_default_dtype(torch.float64)
    gen = torch.Generator().manual_seed(seed)
    reg = ControlAffineRegressor(n, m, device='cpu')
    reg.model.double()
    randomise_hyper(reg, gen)
    X = torch.rand(Ntr, n, generator=gen)
    U = torch.rand(Ntr, m, generator=gen)
    Xte = torch.rand(Nte, n, generator=gen)
    _, MXU = reg.model.encode_from_XU(X, U, 1)
    _, MXUte = reg.model.encode_from_XU(Xte)
    both = torch.cat([MXU, MXUte], dim=0)
    out = dict(X=np64(X), U=np64(U), Xte=np64(Xte), MXU=np64(MXU), MXUte=np64(MXUte), n=n, m=m)
    for k, v in extract_hyper(reg).items():
        out['h_' + k] = v
    cm = reg.model.covar_module
    out['K_train'] = np64(cm(MXU, MXU).evaluate())
    out['K_test'] = np64(cm(MXUte, MXUte).evaluate())
    out['K_mixed'] = np64(cm(both, both).evaluate())
    out['K_mixed_diag'] = np64(cm(both, both, diag=True))
    out['mean_train'] = np64(reg.model.mean_module(MXU))
    out['mean_test'] = np64(reg.model.mean_module(MXUte))
    out['mean_mixed'] = np64(reg.model.mean_module(both))
    out['nopi_mixed'] = float(cm.num_outputs_per_input(both, both))
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)
    print('wrote', name)


def case_cbc2(name, seed):
    """Relative-degree-2 CBC on a pendulum-shaped model (n=2, m=1): GradientGP / MatmulExpr goldens
    (gp_algebra.py:133-168, 319-405; cbc2.py:26-33)."""
    torch.set_default_dtype(torch.float64)
    gen = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    n, m, N = 2, 1, 24
    X, U, Xdot, Xt, Ut, _, _ = make_data(n, m, N, 3, gen, torch.float64)
    reg = ControlAffineRegressor(n, m, device='cpu')
    reg.model.double()
    randomise_hyper(reg, gen)
    with RandRecorder() as rr:
        reg.fit(X, U, Xdot, training_iter=0)
        reg.custom_predict(Xt, Ut)  # draws the factor jitter once; cached afterwards
    out = dict(X=np64(X), U=np64(U), Xdot=np64(Xdot), Xt=np64(Xt), Ut=np64(Ut), jitter=np64(rr.draws[0]))
    for k, v in extract_hyper(reg).items():
        out['h_' + k] = v
    h = lambda x: (x[0] - 0.2) ** 2 + 0.5 * x[1] ** 2 - 0.1
    grad_h = lambda x: torch.stack([2 * (x[0] - 0.2), x[1]])
    k_alpha = torch.tensor([0.9, 1.7])
    x0, u0 = Xt[0].clone(), Ut[0].clone()
    f_gp = reg.f_func_gp()
    grad_h_gp = DeterministicGP(grad_h, shape=(n,), name="grad h")
    L1h = grad_h_gp.t() @ f_gp
    gL1h = GradientGP(L1h, x_shape=(n,))
    out['x0'], out['u0'], out['k_alpha'] = np64(x0), np64(u0), np64(k_alpha)
    out['L1h_mean'] = np64(L1h.mean(x0))
    out['L1h_knl'] = np64(L1h.knl(x0, x0))
    out['gL1h_mean'] = np64(gL1h.mean(x0))
    out['gL1h_knl'] = np64(gL1h.knl(x0, x0))
    cbc2 = cbc2_gp(h, grad_h, reg, u0, k_alpha)
    out['cbc2_mean'] = np64(cbc2.mean(x0))
    out['cbc2_knl'] = np64(cbc2.knl(x0, x0))
    (bfe, e), (V, bfv, v), mean, var = cbc2_quadratic_terms(
        lambda u: cbc2_gp(h, grad_h, reg, u, k_alpha), x0, u0)
    for k_, v_ in dict(bfe=bfe, e=e, V=V, bfv=bfv, v=v, mean=mean, var=var).items():
        out['q_' + k_] = np64(v_)
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)
    print('wrote', name)


if __name__ == '__main__':
    os.makedirs(OUT, exist_ok=True)
    case_predict('ref_predict_unicycle_f64', n=3, m=2, N=48, b=5, seed=11)
    case_predict('ref_predict_pendulum_f64', n=2, m=1, N=30, b=6, seed=12)
    # float32 as in the pendulum recipes; a wide state box keeps Kb well conditioned so that the reference's own
    # float32 round-off stays below the 1e-4 parity tolerance
    case_predict('ref_predict_pendulum_f32', n=2, m=1, N=30, b=6, seed=13, dtype=torch.float32, scale=6.0)
    case_predict('ref_predict_unicycle_rank1_fit_f64', n=3, m=2, N=40, b=4, seed=14, fit_iters=15, rank_one=True)
    case_kernel('ref_kernel_f64', n=2, m=2, Ntr=6, Nte=3, seed=21)
    case_cbc2('ref_cbc2_pendulum_f64', seed=31)
