#!/usr/bin/env python
"""BASELINE configs[0] — pendulum `learn_dynamics_matrix_vector` (reference bayes_cbf/pendulum.py:1053-1246) through the
drop-in API on the GPU: simulate a pendulum trajectory under ControlRandom, fit the MVGP (Adam on the fused GPU log
marginal, 50 iterations as in the reference) on N = 200 samples, predict F(x) = [f(x) | g(x)] on the 20x20 grid and
compare with the true dynamics.  Prints one JSON line (errors relative to the true F, posterior std, timings)."""
import json
import math
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.speed_test_matrix_vector import grid_from_Xtrain, pendulum_trajectory  # noqa: E402


def true_F(X, mass=1.0, gravity=10.0, length=1.0):
    """[f | g] of the pendulum (pendulum.py:82-130): f = [omega, -(g/l) sin theta], g = [0, 1/(m l)]."""
    F = np.zeros((X.shape[0], 2, 2))
    F[:, 0, 0] = X[:, 1]
    F[:, 1, 0] = -(gravity / length) * np.sin(X[:, 0])
    F[:, 1, 1] = 1.0 / (mass * length)
    return F


def main():
    from bayesian_cbf_b200.control_affine_model import ControlAffineRegressorExact
    torch.manual_seed(0)
    np.random.seed(0)
    dX, X, U = pendulum_trajectory(1001)
    idx = torch.randint(1000, (200,)).numpy()          # pendulum.py:350-354: sampled with replacement
    Xtr, Utr, dXtr = (torch.from_numpy(M[idx]) for M in (X, U, dX))
    dgp = ControlAffineRegressorExact(2, 1, device='cuda')
    dgp.model.double()
    t0 = time.perf_counter()
    dgp.fit(Xtr, Utr, dXtr, training_iter=50)
    torch.cuda.synchronize()
    fit_s = time.perf_counter() - t0
    grid = grid_from_Xtrain(X[idx])
    Xtest = torch.from_numpy(grid).cuda()
    t0 = time.perf_counter()
    mean, cov = dgp.custom_predict_fullmat(Xtest)
    torch.cuda.synchronize()
    pred_s = time.perf_counter() - t0
    b = grid.shape[0]
    Fhat = mean.reshape(b, 2, 2).transpose(1, 2).cpu().numpy()      # (b, p, n) -> (b, n, p)
    Ftrue = true_F(grid)
    err = np.abs(Fhat - Ftrue)
    std = torch.sqrt(torch.diagonal(cov).clamp_min(0)).reshape(b, 2, 2).transpose(1, 2).cpu().numpy()
    ls, s = dgp.get_kernel_param('lengthscale').detach().cpu().numpy().tolist(), float(dgp.get_kernel_param('scalefactor'))
    print(json.dumps(dict(config='pendulum learn_dynamics_matrix_vector (BASELINE configs[0]): N=200, 50 Adam iterations, 20x20 grid',
                          fit_s=fit_s, predict_fullmat_s=pred_s,
                          rms_error_f=float(np.sqrt((err[:, :, 0] ** 2).mean())), rms_error_g=float(np.sqrt((err[:, :, 1] ** 2).mean())),
                          rms_true_f=float(np.sqrt((Ftrue[:, :, 0] ** 2).mean())),
                          mean_posterior_std=float(std.mean()), frac_within_3std=float((err <= 3 * std + 1e-9).mean()),
                          lengthscale=ls, outputscale=s)))


if __name__ == '__main__':
    main()
