#!/usr/bin/env python
"""Rollout-ensemble regime (BASELINE configs[4]): R independent MVGPs with N training points each, one posterior query
per rollout and step.  Reports rollout-steps/s and the HBM roofline of ens_posterior_kernel (algorithmic bytes =
each rollout's lower-triangular L^-1 + its X, G, W rows), plus the batched fit time.  One JSON line."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--rollouts', type=int, default=4096)
    ap.add_argument('--n-train', type=int, default=200)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=10)
    a = ap.parse_args()
    from bayesian_cbf_b200.ensemble import MVGPEnsemble
    R, N, n, m = a.rollouts, a.n_train, 3, 2
    p = m + 1
    g = torch.Generator().manual_seed(0)
    f = dict(generator=g, dtype=torch.float64)
    X = 4 * torch.rand(R, N, n, **f) - 2
    U = 2 * torch.rand(R, N, m, **f) - 1
    Xdot = torch.sin(X) + 0.01 * torch.randn(R, N, n, **f)
    ls = torch.tensor([0.7, 0.9, 1.1], dtype=torch.float64).repeat(R, 1) * (0.9 + 0.2 * torch.rand(R, 1, **f))
    s = 1.3 * (0.9 + 0.2 * torch.rand(R, **f))
    Ra, Rb = torch.randn(R, n, n, **f), torch.randn(R, p, p, **f)
    A = Ra @ Ra.transpose(1, 2) + torch.eye(n, dtype=torch.float64)
    B = Rb @ Rb.transpose(1, 2) + torch.eye(p, dtype=torch.float64)
    C = torch.zeros(R, p, n, dtype=torch.float64)
    ens = MVGPEnsemble(n, m)
    args = [t.cuda() for t in (X, U, Xdot, ls, s, A, B, C)]
    for _ in range(3):                  # warm-up: first launches, and the caching allocator reaches its steady state
        ens.fit(*args)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ens.fit(*args)
    e1.record()
    torch.cuda.synchronize()
    fit_ms = e0.elapsed_time(e1)
    xq = (4 * torch.rand(a.steps + a.warmup, R, n, **f) - 2).cuda()
    out = ens.posterior(xq[0])
    for i in range(a.warmup):
        ens.posterior(xq[i], out)
    torch.cuda.synchronize()
    e0.record()
    for i in range(a.warmup, a.warmup + a.steps):
        ens.posterior(xq[i], out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))
    nbytes = ens.posterior_bytes()
    achieved = nbytes / (ms * 1e-3) * 1e-9
    print(json.dumps(dict(metric='rollout posterior steps/sec (M_k, B_k per rollout)', value=R / (ms * 1e-3), unit='rollout-steps/s',
                          config=dict(workload='ensemble of %d independent unicycle MVGPs, N=%d each, 1 query per rollout per step '
                                               '(BASELINE configs[4])' % (R, N), rollouts=R, n_train=N,
                                      l2='factors total %.0f MB per step, larger than L2' % (nbytes / 1e6)),
                          ms_per_step=ms, steps=a.steps, warmup=a.warmup, dtype='f64', data='synthetic',
                          fit_ms=fit_ms, fit_note='batched Gram + Cholesky + L^-1 + alpha for all rollouts (second call)',
                          roofline=dict(bound='hbm', achieved=achieved, peak=peaks['hbm_gbs'], unit='GB/s',
                                        frac=achieved / peaks['hbm_gbs'], traffic=None, kernel='ens_posterior_kernel',
                                        algorithmic_bytes_per_launch=nbytes, peak_source='MEASURED_PEAKS.json hbm_gbs (of measured)'))))


if __name__ == '__main__':
    main()
