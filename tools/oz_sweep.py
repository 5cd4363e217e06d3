#!/usr/bin/env python
"""Within-one-box A/B of oz_var_kernel's scheduling knobs on the bench workload (N=16384, 18648 queries per launch):
cluster size (L^-1 digits multicast to 1 / 2 / 4 CTAs: bcbf_oz_set_cluster) x row blocks per scheduling group
(bcbf_oz_set_group).  Box-to-box spread of this kernel is +-5 % (power cap), so configurations are only comparable inside
one process: every configuration runs `reps` back-to-back launches (sustained, power cap active) in each of `rounds`
interleaved rounds; the B_k of every configuration must equal the first one's bit for bit.
Prints one JSON line per configuration and round."""
import argparse
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
    ap = argparse.ArgumentParser()
    ap.add_argument('--n-train', type=int, default=16384)
    ap.add_argument('--queries', type=int, default=18648)
    ap.add_argument('--clusters', type=int, nargs='*', default=[1, 2, 4])
    ap.add_argument('--groups', type=int, nargs='*', default=[4, 8])
    ap.add_argument('--reps', type=int, default=6)
    ap.add_argument('--rounds', type=int, default=2)
    ap.add_argument('--skips', type=int, nargs='*', default=[0],
                    help='bcbf_oz_debug_skip_loads masks to time (1: no A copies, 2: no B copies, 3: none; results void)')
    a = ap.parse_args()
    lib = _lib.load()
    X, U, Xdot, hyp, jitter = bench.make_workload(a.n_train)
    h = make_hyper(3, 3, hyp['lengthscale'].numpy(), float(hyp['outputscale']), hyp['A'].numpy(), hyp['B'].numpy(),
                   hyp['C'].numpy())
    model = MVGPModel(0).set_var_path('int8')
    model.fit(h, X.numpy(), U.numpy(), Xdot.numpy(), jitter.numpy(), 1e-5)
    Xq, Uq = bench.make_queries(a.queries, 5)
    Xq, Uq = Xq.cuda(), Uq.cuda()
    want = None
    for rnd in range(a.rounds):
        for cl, g, skip in [(c, gg, sk) for c in a.clusters for gg in a.groups for sk in a.skips]:
            if True:
                _lib.check(lib.bcbf_oz_set_cluster(cl))
                _lib.check(lib.bcbf_oz_set_group(g))
                _lib.check(lib.bcbf_oz_debug_skip_loads(skip if cl == 1 else 0))
                out = model.query_device(Xq, Uq)          # warm-up of this configuration (function attributes, workspaces)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(a.reps):
                    out = model.query_device(Xq, Uq)
                e1.record()
                torch.cuda.synchronize()
                Bk = out[1] if isinstance(out, (tuple, list)) else out['Bk']
                if want is None and skip == 0:
                    want = Bk.clone()
                same = bool(torch.equal(Bk, want)) if (want is not None and skip == 0) else None
                ms = e0.elapsed_time(e1) / a.reps
                # one more launch with the in-kernel counters on: SM cycles per K step as the MMA issue thread sees them
                # (independent of the clock the power cap settles at), and how long it waits for operands / the epilogue
                cnt = (ctypes.c_ulonglong * 8)()
                _lib.check(lib.bcbf_oz_debug_counters(1, None))
                model.query_device(Xq, Uq)
                torch.cuda.synchronize()
                _lib.check(lib.bcbf_oz_debug_counters(0, ctypes.byref(cnt)))
                ks = max(cnt[4], 1)
                print(json.dumps(dict(round=rnd, cluster=cl, group=g, skip_loads=skip, ms_per_launch=ms, queries_per_s=a.queries / ms * 1e3,
                                      bit_identical_to_first=same, cycles_per_kstep=cnt[0] / ks,
                                      wait_operands_per_kstep=cnt[1] / ks, wait_epilogue_per_kstep=cnt[2] / ks,
                                      producer_wait_free_stage_per_kstep=cnt[3] / ks)), flush=True)
    _lib.check(lib.bcbf_oz_set_cluster(1))
    _lib.check(lib.bcbf_oz_set_group(4))
    _lib.check(lib.bcbf_oz_debug_skip_loads(0))


if __name__ == '__main__':
    main()
