#!/usr/bin/env python
"""Where the time of the reference's speed-test statement goes at small N (pendulum n=2, m=1, b=400):
`dgp.custom_predict_fullmat(Xtest); dgp.clear_cache()`.  Prints (1) wall time per call, (2) the kernels of one call with
their device time (torch.profiler), device-busy total and launch count, (3) wall time of the stages when each is
followed by a synchronize."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.speed_test_matrix_vector import grid_from_Xtrain, pendulum_trajectory


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--N', type=int, default=256)
    ap.add_argument('--series', default='matrix')
    ap.add_argument('--dtype', default='float32')
    ap.add_argument('--number', type=int, default=100)
    ap.add_argument('--cprofile', action='store_true', help='host-side profile of 200 calls instead of the device timeline')
    a = ap.parse_args()
    from bayesian_cbf_b200 import control_affine_model as cam
    classes = dict(matrix=cam.ControlAffineRegressorExact, vector=cam.ControlAffineRegressorVector)
    dt = torch.float32 if a.dtype == 'float32' else torch.float64
    dX, X, U = pendulum_trajectory(2001)
    idx = np.random.RandomState(0).permutation(X.shape[0])[:a.N]
    Xtr, Utr, dXtr = (torch.from_numpy(M[idx]).to(dt) for M in (X, U, dX))
    Xtest = torch.from_numpy(grid_from_Xtrain(X[idx])).to(dt).cuda()
    torch.manual_seed(0)
    dgp = classes[a.series](2, 1, device='cuda')
    dgp.fit(Xtr, Utr, dXtr, training_iter=10)

    def stmt():
        dgp.custom_predict_fullmat(Xtest)
        dgp.clear_cache()
    for _ in range(5):
        stmt()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(a.number):
        stmt()
        torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / a.number
    print(json.dumps(dict(series=a.series, N=a.N, wall_ms_per_call=wall * 1e3)))
    if a.cprofile:
        import cProfile
        import pstats
        pr = cProfile.Profile()
        pr.enable()
        for _ in range(200):
            stmt()
        torch.cuda.synchronize()
        pr.disable()
        st = pstats.Stats(pr)
        st.sort_stats('tottime').print_stats(45)
        st.sort_stats('cumulative').print_stats(45)
        return
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        stmt()
        torch.cuda.synchronize()
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    ev.sort(key=lambda e: e.time_range.start)
    busy = sum(e.time_range.elapsed_us() for e in ev)
    span = (ev[-1].time_range.end - ev[0].time_range.start) if ev else 0
    print(json.dumps(dict(device_events=len(ev), device_busy_us=busy, device_span_us=span)))
    agg = {}
    for e in ev:
        k = e.name[:70]
        c = agg.setdefault(k, [0, 0.0])
        c[0] += 1
        c[1] += e.time_range.elapsed_us()
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print('%-72s %4d %9.1f us' % (k, c, t))
    print('--- in launch order ---')
    for e in ev[:200]:
        print('%9.1f +%8.1f  %s' % (e.time_range.start - ev[0].time_range.start, e.time_range.elapsed_us(), e.name[:90]))


if __name__ == '__main__':
    main()
