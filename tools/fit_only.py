#!/usr/bin/env python
"""One fit of the bench workload (Gram + jittered Cholesky + L^-1 + alpha + digits of L^-1) and nothing else: the target
of the per-kernel ncu pass of the factorisation (profiles/*fit_kernels*)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from bayesian_cbf_b200.model import MVGPModel, make_hyper


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    X, U, Xdot, hyp, jitter = bench.make_workload(N)
    h = make_hyper(3, 3, hyp['lengthscale'].numpy(), float(hyp['outputscale']), hyp['A'].numpy(), hyp['B'].numpy(),
                   hyp['C'].numpy())
    model = MVGPModel(0).set_var_path('int8')
    for _ in range(2):
        model.fit(h, X.numpy(), U.numpy(), Xdot.numpy(), jitter.numpy(), 1e-5)
    print(json.dumps(dict(N=N, fit_ms=model.fit_timing_ms())))


if __name__ == '__main__':
    main()
