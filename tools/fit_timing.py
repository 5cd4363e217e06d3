#!/usr/bin/env python
"""Per-iteration time of ControlAffineRegressor.fit (Adam on the fused GPU log marginal, bayesian_cbf_b200/mll.py) —
SURVEY 8a row 6/13.  CPU reference points (dense Kronecker MLL + autograd, 8 threads):
N = 200 / 512 / 1024 / 4096 / 8192 -> 5.8 ms / 12 ms / 68 ms / 2.25 s / 12.5 s per iteration."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CPU_REF_MS = {200: 5.8, 512: 12.0, 1024: 68.0, 4096: 2250.0, 8192: 12500.0}


def main():
    from bayesian_cbf_b200.control_affine_model import ControlAffineRegressorExact
    sizes = [int(a) for a in sys.argv[1:]] or [200, 512, 1024, 4096, 8192]
    g = torch.Generator().manual_seed(0)
    out = []
    for N in sizes:
        n, m = 3, 2
        X = 4 * torch.rand(N, n, generator=g, dtype=torch.float64) - 2
        U = 2 * torch.rand(N, m, generator=g, dtype=torch.float64) - 1
        Xdot = torch.sin(X) * (1 + U[:, :1]) + 0.01 * torch.randn(N, n, generator=g, dtype=torch.float64)
        reg = ControlAffineRegressorExact(n, m, device='cuda')
        reg.model.double()
        reg.fit(X, U, Xdot, training_iter=3, lr=0.01)
        torch.cuda.synchronize()
        iters = 50 if N <= 1024 else 5      # 50 = the reference's default training_iter (control_affine_model.py:274)

        def timed(**kw):
            t0 = time.perf_counter()
            reg.fit(X, U, Xdot, training_iter=iters, lr=0.01, **kw)
            torch.cuda.synchronize()
            return 1e3 * (time.perf_counter() - t0) / iters
        ms = timed()                        # default policy: captured iteration when Npad <= 1024 (capture cost included)
        rec = dict(N=N, iterations=iters, ms_per_adam_iteration=ms, cpu_reference_ms=CPU_REF_MS.get(N),
                   speedup_vs_cpu_reference=(CPU_REF_MS[N] / ms) if N in CPU_REF_MS else None)
        if N <= 1024:
            rec['ms_per_adam_iteration_eager'] = timed(cuda_graph=False)
        print(json.dumps(rec), flush=True)
        out.append(rec)


if __name__ == '__main__':
    main()
