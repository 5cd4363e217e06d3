"""TEST INFRASTRUCTURE ONLY — the per-step record of BASELINE configs[2], `unicycle_bayes_cbf_safe_obstacle`
(reference bayes_cbf/unicycle_move_to_pose.py:1889-1905, 1925-1928: start (-3,-1,-pi/4), goal (0,0,pi/4), dt = 0.001,
2000 steps, true AckermannDrive(L=12), prior mean AckermannDrive(L=1, kernel_diag_A=1e-2), learning off, 2 ObstacleCBFs,
PiecewiseLinearPlanner(0.95), CLF Kp=[0.9,1.5,0], max_risk 0.01), produced by the CPU RESTATEMENTS of the kernels on the
path (tests/fake_ops.py: closed-form CBC terms, oracle/socp_oracle.py barrier solver).  The reference itself cannot
produce it here (its per-step program goes to cvxpy + GUROBI); its cone terms at 12 states of this recipe ARE pinned to
the reference (tests/golden/ref_controller_f64.npz).  The CUDA rollout must reproduce this record: identical
feasibility decision at every one of the 2000 steps, states / controls / cone terms to 1e-7.

    python oracle/gen_rollout_record.py   -> tests/golden/rollout_safe_obstacle_2000.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)


class _Patch:
    def setattr(self, obj, name, value):
        setattr(obj, name, value)


def build(dev):
    from bayesian_cbf_b200 import unicycle as U
    from tests.golden_util import load
    d = load('ref_controller_f64')
    planner = U.PiecewiseLinearPlanner(d['x0'], d['xg'], int(d['numSteps']), float(d['dt']), frac_time_to_reach_goal=0.95)
    cbfs = U.obstacles_at_mid_from_start_and_goal(d['x0'], d['xg'], term_weights=(0.7, 0.3))
    ctrl = U.BayesCBFController(planner, U.CLFCartesian(Kp=(0.9, 1.5, 0.0)), cbfs, [5.0, 5.0], model_L=1.0,
                                kernel_diag_A=(1e-2, 1e-2, 1e-2), clf_gamma=10.0, max_risk=0.01)
    X0 = torch.from_numpy(d['x0']).reshape(1, 3).to(dev)
    return U, ctrl, X0, float(d['dt']), int(d['numSteps'])


def run(dev, steps=None, term_every=100):
    U, ctrl, X0, dt, numSteps = build(dev)
    steps = numSteps if steps is None else steps
    terms = {}

    def on_step(t, X, u, xdot, ok):
        if t % term_every == 0:
            c, d, A, b = ctrl.constraint_terms(X, t)
            terms[t] = np.concatenate([v[0].detach().cpu().numpy().reshape(-1) for v in (c, d, A, b)])
    out = U.rollout(ctrl, X0, steps, dt, true_L=12.0, on_step=on_step)
    ts = sorted(terms)
    return dict(X=out['X'][:, 0].cpu().numpy(), U=out['U'][:, 0].cpu().numpy(),
                feasible=out['feasible'][:, 0].cpu().numpy(), term_steps=np.array(ts),
                terms=np.stack([terms[t] for t in ts]))


def main():
    from tests import fake_ops
    fake_ops.installed(_Patch()).__enter__()
    rec = run('cpu')
    keep = np.arange(0, rec['X'].shape[0], 25)
    out = dict(steps=rec['U'].shape[0], sample_every=25, X=rec['X'][keep], U=rec['U'][keep[:-1]], X_final=rec['X'][-1],
               feasible=np.packbits(rec['feasible']), n_feasible=int(rec['feasible'].sum()),
               term_steps=rec['term_steps'], terms=rec['terms'])
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'rollout_safe_obstacle_2000.npz'), **out)
    print('wrote rollout_safe_obstacle_2000.npz: %d steps, %d feasible, final state %s' %
          (out['steps'], out['n_feasible'], rec['X'][-1]))


if __name__ == '__main__':
    main()
