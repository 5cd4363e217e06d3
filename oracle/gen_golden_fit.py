"""TEST INFRASTRUCTURE ONLY — golden TRAJECTORIES of the reference's hyper-parameter fit.

Runs the UNMODIFIED `ControlAffineRegressor._fit_with_warnings` (bayes_cbf/control_affine_model.py:274-335: train-data
set-up, Adam + MultiStepLR, the fresh 1e-6 multiplicative target noise of every iteration, loss = -mll) from
/root/reference over the dense gpytorch stand-in (oracle/gpytorch_shim.py: ExactMarginalLogLikelihood =
[log N(y; M, K) + sum of lengthscale priors] / num_data evaluated densely) and records

  * the raw parameters before the first step (by parameter name),
  * every `torch.rand_like` target-noise draw, in order,
  * the loss of every iteration,
  * the raw parameters and the constrained hyper-parameters after the last step.

tests/test_fit_mll.py replays the same initial parameters and noise through bayesian_cbf_b200's fit (CUDA log marginal
with closed-form adjoints) and must reproduce the loss sequence and the final parameters.  The marginal likelihood is
gpytorch-internal in the reference ("parity unpinned" against real gpytorch); this pins everything the
reference's OWN code contributes to the trajectory, and the density itself is checked separately against
torch.distributions.MultivariateNormal.

    python oracle/gen_golden_fit.py        # writes tests/golden/ref_fit_*.npz
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

from bayes_cbf.control_affine_model import (ControlAffineRegressor, ControlAffineExactGP,  # noqa: E402
                                            ControlAffineRegressorVector)
import gpytorch  # noqa: E402  (the shim)

OUT = os.path.join(ROOT, 'tests', 'golden')
np64 = lambda t: t.detach().cpu().double().numpy().copy()      # copy: .numpy() aliases the (later updated) parameter


def case_fit(name, n, m, N, seed, iters, rank=None, prior=None, lr=0.1, vector=False):
    torch.set_default_dtype(torch.float64)
    gen = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    X = 2 * torch.rand(N, n, generator=gen) - 1
    U = 2 * torch.rand(N, m, generator=gen) - 1
    Wt = torch.randn(n, n, generator=gen)
    Xdot = torch.sin(X @ Wt) + (torch.cos(X) * U[:, :1]) + 0.05 * torch.randn(N, n, generator=gen)
    mc = partial(ControlAffineExactGP, rank=rank, gamma_length_scale_prior=prior) if (rank is not None or prior) \
        else ControlAffineExactGP
    reg = ControlAffineRegressorVector(n, m, device='cpu') if vector else ControlAffineRegressor(n, m, device='cpu', model_class=mc)
    reg.model.double()
    with torch.no_grad():                                  # non-trivial, reproducible start
        for _, prm in reg.model.named_parameters():
            prm.copy_(0.3 * torch.randn(prm.shape, generator=gen, dtype=prm.dtype))
    out = dict(X=np64(X), U=np64(U), Xdot=np64(Xdot), n=n, m=m, iters=iters, lr=lr,
               rank=-1 if rank is None else rank, prior=np.asarray(prior if prior else [], dtype=np.float64))
    for pname, prm in reg.model.named_parameters():
        out['init/' + pname] = np64(prm)
    noises, losses = [], []
    orig_rand_like = torch.rand_like
    orig_fwd = gpytorch.mlls.ExactMarginalLogLikelihood.forward

    def rec_rand_like(t, *a, **k):
        r = orig_rand_like(t, *a, **k)
        noises.append(r.detach().clone())
        return r

    def rec_fwd(self, output, target, *params):
        v = orig_fwd(self, output, target, *params)
        losses.append(-float(v))
        return v

    torch.rand_like = rec_rand_like
    gpytorch.mlls.ExactMarginalLogLikelihood.forward = rec_fwd
    try:
        reg.fit(X, U, Xdot, training_iter=iters, lr=lr)
    finally:
        torch.rand_like = orig_rand_like
        gpytorch.mlls.ExactMarginalLogLikelihood.forward = orig_fwd
    assert len(noises) == iters and len(losses) == iters
    out['noise'] = np.stack([np64(z) for z in noises])           # (iters, N*n)
    out['loss'] = np.asarray(losses)
    for pname, prm in reg.model.named_parameters():
        out['final/' + pname] = np64(prm)
    mdl = reg.model
    p_, n_ = mdl.matshape
    out['final_outputscale'] = np64(mdl.input_covar.outputscale.reshape(()))
    if vector:                                             # CoGP comparator: scalar RBF lengthscale + linear kernel, one Sigma
        out['final_lengthscale'] = np64(mdl.input_covar.base_kernel.kernels[0].lengthscale.reshape(-1))
        out['final_linear_variance'] = np64(mdl.input_covar.base_kernel.kernels[1].variance.reshape(()))
        out['final_Sigma'] = np64(mdl.task_covar.covar_matrix.evaluate())
    else:
        out['final_lengthscale'] = np64(mdl.input_covar.base_kernel.lengthscale.reshape(-1))
        out['final_A'] = np64(mdl.task_covar.U.covar_matrix.evaluate())
        out['final_B'] = np64(mdl.task_covar.V.covar_matrix.evaluate())
    out['final_C'] = np64(torch.stack([bm.constant.reshape(()) for bm in mdl.mean_module.base_means]).reshape(p_, n_))
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)
    print('wrote', name, 'loss %.6f -> %.6f' % (losses[0], losses[-1]))


if __name__ == '__main__':
    case_fit('ref_fit_unicycle_f64', n=3, m=2, N=40, seed=41, iters=20)
    case_fit('ref_fit_pendulum_rank1_prior_f64', n=2, m=1, N=36, seed=42, iters=50, rank=1, prior=(1e-3, 1e-3))
    case_fit('ref_fit_pendulum_cogp_vector_f64', n=2, m=1, N=30, seed=43, iters=25, vector=True)
