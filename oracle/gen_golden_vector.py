"""TEST INFRASTRUCTURE ONLY — goldens for the CoGP comparator `ControlAffineRegressorVector`
(bayes_cbf/control_affine_model.py:1106-1331, the baseline series of the reference's speed test) from the UNMODIFIED
reference over the dense gpytorch stand-in.  Run where /root/reference exists:  python oracle/gen_golden_vector.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import gpytorch_shim  # noqa: E402

gpytorch_shim.install()
from bayes_cbf.control_affine_model import ControlAffineRegressorVector  # noqa: E402
from oracle.gen_golden import RandRecorder, make_data, np64, randomise_hyper  # noqa: E402


def main():
    torch.set_default_dtype(torch.float64)
    n, m, N, b, seed = 2, 1, 20, 4, 41
    gen = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    X, U, Xdot, Xt, Ut, Xtp, Utp = make_data(n, m, N, b, gen, torch.float64)
    reg = ControlAffineRegressorVector(n, m, device='cpu')
    reg.model.double()
    randomise_hyper(reg, gen)
    mdl = reg.model
    p = m + 1
    out = dict(X=np64(X), U=np64(U), Xdot=np64(Xdot), Xt=np64(Xt), Ut=np64(Ut), Xtp=np64(Xtp), Utp=np64(Utp), n=n, m=m,
               Sigma=np64(mdl.task_covar.covar_matrix.evaluate()),
               outputscale=np64(mdl.input_covar.outputscale.reshape(())),
               C=np64(torch.stack([bm.constant.reshape(()) for bm in mdl.mean_module.base_means]).reshape(p, n)))
    add = mdl.input_covar.base_kernel          # AdditiveKernel(RBF, Linear)
    rbf, lin = add.kernels[0], add.kernels[1]
    out['lengthscale'] = np64(rbf.lengthscale.reshape(-1))
    out['linear_variance'] = np64(lin.variance.reshape(()))
    out['K_data'] = np64(mdl.input_covar(Xt, X).evaluate())
    pm, pk = reg._custom_predict_matrix(Xt, Xtp)
    out['prior_Mk'], out['prior_Kk'] = np64(pm), np64(pk)
    reg.fit(X, U, Xdot, training_iter=0)

    def record(tag, fn):
        with RandRecorder() as rr:
            res = fn()
        for i, r in enumerate(res if isinstance(res, tuple) else (res,)):
            out['%s_out%d' % (tag, i)] = np64(r)
        for i, d in enumerate(rr.draws):
            out['%s_rand%d' % (tag, i)] = np64(d)
        out['%s_nrand' % tag] = len(rr.draws)

    record('matrix', lambda: reg._custom_predict_matrix(Xt))
    record('predict', lambda: reg.custom_predict(Xt, Ut))
    record('fullmat', lambda: reg.custom_predict_fullmat(Xt))
    record('nocov', lambda: reg.custom_predict(Xt, Ut, compute_cov=False))
    out['L'] = np64(reg._cache['perturbed_cholesky'])
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'ref_vector_cogp_f64.npz'), **out)
    print('wrote ref_vector_cogp_f64.npz', {k: np.shape(v) for k, v in out.items()})


if __name__ == '__main__':
    main()
