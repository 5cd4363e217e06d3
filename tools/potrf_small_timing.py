#!/usr/bin/env python
"""Device time of the in-CTA 128x128 Cholesky + inverse kernel (potf2_inv_kernel): R = 148 independent blocks in one
batched launch (one CTA per SM, one wave) timed with CUDA events — the latency of the serial link of the blocked
factorisation, free of host launch overhead."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesian_cbf_b200 import _lib  # noqa: E402

lib = _lib.load()
g = torch.Generator().manual_seed(0)
for n, R in ((128, 148), (128, 1), (256, 148)):
    M = torch.randn(R, n, n, generator=g, dtype=torch.float64)
    A0 = (M @ M.transpose(1, 2) + n * torch.eye(n, dtype=torch.float64)).cuda()
    dinv = torch.empty(R, lib.bcbf_dinv_elems(n), dtype=torch.float64, device='cuda')
    info = torch.zeros(R, dtype=torch.int32, device='cuda')
    ts = []
    for rep in range(6):
        A = A0.clone()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.bcbf_potrf_batched(A.data_ptr(), n, n, n, None, 0.0, dinv.data_ptr(), info.data_ptr(), R, 0))
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    assert int(info.abs().max()) == 0
    print('potrf_batched n=%d R=%d: min %.1f us' % (n, R, 1e3 * min(ts)))
