#!/usr/bin/env python
"""BASELINE configs[4] shape, R rollouts per GPU (run under torchrun for several GPUs: the rollouts are independent, every
rank takes its own R with its own seeds, no data-path collective; rank 0 prints the aggregate): R independent `learning_helps_avoid_getting_stuck` rollouts
(reference unicycle_move_to_pose.py:1948-1969: true Ackermann L=1, prior L=12 with kernel_diag_A=[1,1,1], learning on,
2 obstacles, PiecewiseLinearPlanner, dt=0.001), each with its own MVGP refitted every `--train-every` steps on at most
200 of its own samples.  Per control step and rollout: ensemble posterior (HBM-bound kernel) -> CBC/CLC cone terms ->
batched SOCP.  --adam-iters K runs K Adam steps on every rollout's log marginal at each refit (reference: 100).  Reports rollout-steps/s; one JSON line."""
import argparse
import json
import math
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--rollouts', type=int, default=512)
    ap.add_argument('--steps', type=int, default=600)
    ap.add_argument('--train-every', type=int, default=200)
    ap.add_argument('--max-train', type=int, default=200)
    ap.add_argument('--eager', action='store_true', help='no CUDA graph')
    ap.add_argument('--adam-iters', type=int, default=0, help='Adam steps per refit on every rollout (reference: 100)')
    a = ap.parse_args()
    rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        sys.stdout.flush()
        saved_stdout = os.dup(1)          # NCCL's printf banner (NCCL_DEBUG=VERSION) goes to stderr, not into the JSON
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
            dist.all_reduce(torch.zeros(1, device='cuda'))
            torch.cuda.synchronize()
        finally:
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    from bayesian_cbf_b200 import unicycle as U
    R, dt, numSteps = a.rollouts, 0.001, 2000
    x0 = [-3.0, -1.0, -math.pi / 4]
    xg = [0.0, 0.0, math.pi / 4]
    planner = U.PiecewiseLinearPlanner(x0, xg, numSteps, dt, frac_time_to_reach_goal=0.95)
    cbfs = U.obstacles_at_mid_from_start_and_goal(x0, xg, term_weights=(0.7, 0.3))
    learner = U.EnsembleLearner(R, dt, model_L=12.0, max_train=a.max_train, train_every_n_steps=a.train_every,
                                lengthscale=(1.0, 1.0, 0.7), outputscale=1.0, adam_iters=a.adam_iters, seed=rank)
    ctrl = U.BayesCBFController(planner, U.CLFCartesian(Kp=(0.9, 1.5, 0.0)), cbfs, [5.0, 5.0], model_L=12.0,
                                clf_gamma=10.0, max_risk=0.01, posterior=learner.posterior)
    g = torch.Generator().manual_seed(rank)          # rollout r of rank k: its own start perturbation and jitter stream
    X0 = (torch.tensor(x0, dtype=torch.float64).repeat(R, 1)
          + 0.05 * (torch.rand(R, 3, generator=g, dtype=torch.float64) - 0.5)).cuda()
    U.rollout(ctrl, X0, 5, dt, true_L=1.0)          # warm-up launches
    learner.Xs, learner.Us = [], []
    if a.eager:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = U.rollout(ctrl, X0, a.steps, dt, true_L=1.0, on_step=learner.record)
    else:
        # the graph is captured with the prior-only posterior and RE-captured after every refit (the posterior switches
        # from the analytic prior to the ensemble kernel, and the factor buffers are re-allocated)
        gr = U.GraphedRollout(ctrl, X0, dt, true_L=1.0).capture()
        state = dict(refits=0)

        def on_step(t, X, u, xdot, ok):
            learner.record(t, X, u, xdot, ok)
            if learner.refits != state['refits']:
                state['refits'] = learner.refits
                gr.capture()
                gr._set_plan(t)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = gr.run(a.steps, on_step=on_step, record=False)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    alive = int(out['alive'].sum())
    if world > 1:
        t = torch.tensor([wall], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)                 # slowest rank
        al = torch.tensor([alive], dtype=torch.int64, device='cuda')
        dist.all_reduce(al, op=dist.ReduceOp.SUM)
        wall, alive = float(t.item()), int(al.item())
        dist.destroy_process_group()
        if rank != 0:
            return
    R = R * world
    print(json.dumps(dict(metric='controlled rollout steps/sec (posterior + CBC terms + SOCP per step)',
                          value=R * a.steps / wall, unit='rollout-steps/s', rollouts=R, n_gpus=world, steps=a.steps,
                          ms_per_step=1e3 * wall / a.steps, refits=learner.refits, cuda_graph=not a.eager, adam_iters_per_refit=a.adam_iters, alive_at_end=alive,
                          n_train_last=getattr(learner.ens, 'N', 0),
                          config=dict(workload='ensemble of %d unicycle learning rollouts (BASELINE configs[4] shape), '
                                               'refit every %d steps, max_train %d' % (R, a.train_every, a.max_train)),
                          dtype='f64', data='synthetic')))


if __name__ == '__main__':
    main()
