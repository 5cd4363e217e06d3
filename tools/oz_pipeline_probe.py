#!/usr/bin/env python
"""In-kernel pipeline counters of oz_var_kernel (bcbf_oz_debug_counters) on the bench workload: how many cycles per K step
the MMA issue thread needs, and how much of that it waits for operand stages / for the epilogue."""
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from bayesian_cbf_b200 import _lib
from bayesian_cbf_b200.model import MVGPModel, make_hyper


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    Q = int(sys.argv[2]) if len(sys.argv) > 2 else 18648
    lib = _lib.load()
    X, U, Xdot, hyp, jitter = bench.make_workload(N)
    h = make_hyper(3, 3, hyp['lengthscale'].numpy(), float(hyp['outputscale']), hyp['A'].numpy(), hyp['B'].numpy(),
                   hyp['C'].numpy())
    model = MVGPModel(0).set_var_path('int8')
    model.fit(h, X.numpy(), U.numpy(), Xdot.numpy(), jitter.numpy(), 1e-5)
    Xq, Uq = bench.make_queries(Q, 5)
    Xq, Uq = Xq.cuda(), Uq.cuda()
    for _ in range(3):
        model.query_device(Xq, Uq)
    torch.cuda.synchronize()
    out = (ctypes.c_ulonglong * 8)()
    lib.bcbf_oz_debug_counters(1, None)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    model.query_device(Xq, Uq)
    e1.record()
    torch.cuda.synchronize()
    lib.bcbf_oz_debug_counters(0, ctypes.byref(out))
    c = list(out)
    ksteps = max(c[4], 1)
    print(json.dumps(dict(N=N, Q=Q, step_ms=e0.elapsed_time(e1), ctas=148, ksteps=c[4],
                          cycles_per_kstep=c[0] / ksteps, wait_operands_per_kstep=c[1] / ksteps,
                          wait_epilogue_per_kstep=c[2] / ksteps, producer_wait_free_stage_per_kstep=c[3] / ksteps,
                          issue_thread_mcycles_per_cta=c[0] / 148e6)))


if __name__ == '__main__':
    main()
