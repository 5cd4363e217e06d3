#!/usr/bin/env python
"""Where the host time of one small-N Adam iteration goes (cProfile of ControlAffineRegressor.fit at N = 200)."""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from bayesian_cbf_b200.control_affine_model import ControlAffineRegressorExact
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    g = torch.Generator().manual_seed(0)
    X = 4 * torch.rand(N, 3, generator=g, dtype=torch.float64) - 2
    U = 2 * torch.rand(N, 2, generator=g, dtype=torch.float64) - 1
    Xdot = torch.sin(X) * (1 + U[:, :1]) + 0.01 * torch.randn(N, 3, generator=g, dtype=torch.float64)
    reg = ControlAffineRegressorExact(3, 2, device='cuda')
    reg.model.double()
    reg.fit(X, U, Xdot, training_iter=5, lr=0.01)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reg.fit(X, U, Xdot, training_iter=50, lr=0.01)
    torch.cuda.synchronize()
    print('ms per iteration: %.3f' % (1e3 * (time.perf_counter() - t0) / 50))
    pr = cProfile.Profile()
    pr.enable()
    reg.fit(X, U, Xdot, training_iter=50, lr=0.01)
    torch.cuda.synchronize()
    pr.disable()
    pstats.Stats(pr).sort_stats('cumulative').print_stats(45)


if __name__ == '__main__':
    main()
