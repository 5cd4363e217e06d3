#!/usr/bin/env python
"""The reference's own speed test (BASELINE configs[1], bayes_cbf/pendulum.py:1305-1394) through the drop-in API:

    timed statement:  dgp.custom_predict_fullmat(Xtest); dgp.clear_cache()      (pendulum.py:1367-1372)

i.e. Gram + jittered Cholesky (+ L^-1) + posterior mean and full (b p n)^2 covariance of F(x) on a 20x20 grid, with the
factor cache cleared inside the timed statement.  Pendulum n=2, m=1; data: Euler-simulated trajectory with the
reference's ControlRandom (pendulum.py:69-78, 200-252), theta wrapped to [-pi, pi).  Hyper-parameters: fitted with
`training_iter` Adam steps when --fit is given (as the reference does, :1366), else the seeded defaults.
min over `repeat` of `number` calls, divided by `number` — the reference logs exactly that (`elapsed/ntimes`, :1388).
BASELINE.md section 1 holds the reference's published numbers for N = 256..512 (unnamed CUDA GPU, float32).
Prints one JSON line per N and a final summary line."""
import argparse
import json
import math
import os
import sys
import timeit

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

PUBLISHED = {   # BASELINE.md section 1, */elapsed [s] at N = 256, 320, 384, 512
    'matrix': {256: 0.04359, 320: 0.04527, 384: 0.05029, 512: 0.07753},
    'vector': {256: 0.06428, 320: 0.08648, 384: 0.11682, 512: 0.19145},
    'matrixdiag': {256: 0.03314, 320: 0.03629, 384: 0.04174, 512: 0.05108},
    'vectordiag': {256: 0.05899, 320: 0.08177, 384: 0.11233, 512: 0.17858},
}


def pendulum_trajectory(steps, tau=0.01, x0=(5 * math.pi / 6, -0.01), mass=1.0, gravity=10.0, length=1.0, seed=0):
    g = torch.Generator().manual_seed(seed)
    theta, omega = x0
    X = np.zeros((steps, 2))
    U = np.zeros((steps, 1))
    for t in range(steps):
        X[t] = (theta, omega)
        u = mass * gravity * math.sin(theta) * (float(torch.rand(1, generator=g)) * 0.8 + 0.6)
        U[t, 0] = u
        omega_n = omega + (-(gravity / length) * math.sin(theta) + u / (mass * length)) * tau
        theta_n = theta + omega * tau
        theta = ((theta_n + math.pi) % (2 * math.pi)) - math.pi
        omega = omega_n
    dX = (X[1:] - X[:-1]) / tau
    return dX, X[:-1], U[:-1]


def grid_from_Xtrain(Xtrain, k=20):
    th = np.arange(Xtrain[:, 0].min(), Xtrain[:, 0].max(), (Xtrain[:, 0].max() - Xtrain[:, 0].min()) / k)[:k]
    om = np.arange(Xtrain[:, 1].min(), Xtrain[:, 1].max(), (Xtrain[:, 1].max() - Xtrain[:, 1].min()) / k)[:k]
    T, O_ = np.meshgrid(th, om, indexing='ij')
    return np.stack([T.reshape(-1), O_.reshape(-1)], axis=1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--sizes', type=int, nargs='*', default=[256, 320, 384, 512, 1024, 2048, 4096])
    ap.add_argument('--repeat', type=int, default=5)
    ap.add_argument('--number', type=int, default=50)
    ap.add_argument('--fit', type=int, default=50, help='Adam iterations before timing (the reference: 50, pendulum.py:1366)')
    ap.add_argument('--dtype', default='float32', choices=['float32', 'float64'])
    ap.add_argument('--series', nargs='*', default=['matrix', 'vector', 'matrixdiag', 'vectordiag'])
    a = ap.parse_args()
    from bayesian_cbf_b200 import control_affine_model as cam
    classes = dict(matrix=cam.ControlAffineRegressorExact, vector=cam.ControlAffineRegressorVector,
                   matrixdiag=cam.ControlAffineRegMatrixDiag, vectordiag=cam.ControlAffineRegVectorDiag)
    dt = torch.float32 if a.dtype == 'float32' else torch.float64
    steps = max(2001, max(a.sizes) + 2)
    dX, X, U = pendulum_trajectory(steps)
    order = np.random.RandomState(0).permutation(X.shape[0])
    out = []
    for series in a.series:
      for N in a.sizes:
        if series.startswith('vector') and N > 2048:
            continue   # (N n)^2 factor: the reference itself runs out of memory here
        idx = order[:N]
        Xtr, Utr, dXtr = (torch.from_numpy(M[idx]).to(dt) for M in (X, U, dX))
        Xtest = torch.from_numpy(grid_from_Xtrain(X[idx])).to(dt)
        torch.manual_seed(0)
        dgp = classes[series](2, 1, device='cuda')
        if dt is torch.float64:
            dgp.model.double()
        dgp.fit(Xtr, Utr, dXtr, training_iter=a.fit)     # all four series, like the reference (pendulum.py:1366)
        Xtest_d = Xtest.cuda()

        def stmt():
            dgp.custom_predict_fullmat(Xtest_d)
            dgp.clear_cache()
        stmt()
        torch.cuda.synchronize()
        number = a.number if N <= 1024 else max(5, a.number // 5)

        def timed():
            stmt()
            torch.cuda.synchronize()   # the result is consumed on the host in the reference (plots / logs)
        elapsed = min(timeit.repeat(timed, repeat=a.repeat, number=number)) / number
        mean, cov = dgp.custom_predict_fullmat(Xtest_d)
        rec = dict(series=series, N=N, b=int(Xtest.shape[0]), seconds_per_call=elapsed, queries_per_s=Xtest.shape[0] / elapsed,
                   published_reference_s=PUBLISHED[series].get(N),
                   speedup_vs_published=(PUBLISHED[series][N] / elapsed) if N in PUBLISHED[series] else None,
                   cov_shape=list(cov.shape), dtype=a.dtype, fit_iters=a.fit)
        print(json.dumps(rec), flush=True)
        out.append(rec)
    print(json.dumps(dict(summary='pendulum speed_test_matrix_vector (MVGP full), custom_predict_fullmat + clear_cache',
                          results=out)))


if __name__ == '__main__':
    main()
