#!/usr/bin/env python
"""Full-size (N=16384) comparison of the two covariance kernels on the bench workload: int8 tensor cores with digit
splitting (oz_var_kernel) against the FP64 DMMA kernel (post_var_kernel), same fitted model, same queries."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from bayesian_cbf_b200.model import MVGPModel, make_hyper


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    Q = int(sys.argv[2]) if len(sys.argv) > 2 else 4200
    X, U, Xdot, hyp, jitter = bench.make_workload(N)
    h = make_hyper(3, 3, hyp['lengthscale'].numpy(), float(hyp['outputscale']), hyp['A'].numpy(), hyp['B'].numpy(),
                   hyp['C'].numpy())
    model = MVGPModel(0).set_var_path('int8')
    model.fit(h, X.numpy(), U.numpy(), Xdot.numpy(), jitter.numpy(), 1e-5)
    Xq, Uq = bench.make_queries(Q, 5)
    a = model.query(Xq.numpy(), Uq.numpy())
    model.set_var_path('dmma')
    b = model.query(Xq.numpy(), Uq.numpy())
    prior = float(hyp['outputscale'] * torch.linalg.matrix_norm(hyp['B'], 2))
    out = dict(N=N, Q=Q, fit_ms=model.fit_timing_ms(),
               Bk_diff_over_prior=float(np.abs(a['Bk'] - b['Bk']).max() / prior),
               svar_diff_over_prior=float(np.abs(a['svar'] - b['svar']).max() / prior),
               svar_min=float(a['svar'].min()), svar_max=float(a['svar'].max()),
               svar_rel_diff_max=float((np.abs(a['svar'] - b['svar']) / np.abs(b['svar'])).max()))
    print(json.dumps(out))


if __name__ == '__main__':
    main()
