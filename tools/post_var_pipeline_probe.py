#!/usr/bin/env python
"""Development probe: where does post_var_kernel lose time?  Runs a few bench-shaped steps with the in-kernel pipeline
counters enabled (bcbf_debug_counters) and prints barrier-wait fractions."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from bayesian_cbf_b200 import _lib  # noqa: E402
from bayesian_cbf_b200.model import MVGPModel, make_hyper  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
QS = 18944
lib = _lib.load()
X, U, Xdot, hyp, jitter = bench.make_workload(N)
model = MVGPModel(0)
model.fit(make_hyper(3, 3, hyp['lengthscale'].numpy(), float(hyp['outputscale']), hyp['A'].numpy(), hyp['B'].numpy(),
                     hyp['C'].numpy()), X.numpy(), U.numpy(), Xdot.numpy(), jitter.numpy(), 1e-5)
Xq, Uq = bench.make_queries(QS, 0)
Xq, Uq = Xq.cuda(), Uq.cuda()
for _ in range(2):
    model.query_device(Xq, Uq)
torch.cuda.synchronize()
out = (ctypes.c_ulonglong * 8)()
lib.bcbf_debug_counters(1, None)
model.query_device(Xq, Uq, want=('Bk',))
lib.bcbf_debug_counters(0, ctypes.byref(out))
c = list(out)
print('consumer: wait cycles/stage %.1f, blocked stages %.2f%%' % (c[0] / max(c[1], 1), 100.0 * c[2] / max(c[1], 1)))
print('producer: empty-wait cycles/stage %.1f, issue cycles/stage %.1f' % (c[3] / max(c[4], 1), c[5] / max(c[4], 1)))
print('CTA lifetime cycles avg %.0f, stages per CTA %.0f -> cycles per stage %.1f (ideal 6144)' %
      (c[6] / max(c[7], 1), c[4] / max(c[7], 1), c[6] / max(c[4], 1)))
print('consumer wait share of CTA time: %.2f%%' % (100.0 * (c[0] / 8.0) / max(c[6], 1)))
