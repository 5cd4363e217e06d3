"""TEST INFRASTRUCTURE ONLY — golden vectors for the consumer side of the path, produced by the UNMODIFIED reference:
`ControllerCLFBayesian._clc_terms / _cbc_terms` (bayes_cbf/unicycle_move_to_pose.py:880-920) with the recipe of
`unicycle_bayes_cbf_safe_obstacle` (:1889-1905, 1925-1928: prior-only AckermannDrive(L=1, kernel_diag_A=1e-2),
CLFCartesian(Kp=[0.9,1.5,0]), 2 ObstacleCBFs with weights [0.7,0.3], PiecewiseLinearPlanner(frac=0.95)), i.e. the
(A, bfb, bfc, d) cone terms of every constraint at a set of states / time steps.  cvxpy / GUROBI are not needed for
these.  Run where /root/reference exists:  python oracle/gen_golden_controller.py  -> tests/golden/ref_controller_f64.npz
"""
import functools
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import gpytorch_shim  # noqa: E402

gpytorch_shim.install()
kw = types.ModuleType('kwplus')
kw.__path__ = []
kw.default_kw = lambda f: {}
kw.recpartial = lambda f, d=None, **k: functools.partial(f)
sys.modules['kwplus'] = kw
fm = types.ModuleType('kwplus.functools')
fm.recpartial = kw.recpartial
sys.modules['kwplus.functools'] = fm
vm = types.ModuleType('kwplus.variations')
vm.kwvariations = lambda *a, **k: []
vm.expand_variations = lambda *a, **k: []
sys.modules['kwplus.variations'] = vm

import bayes_cbf.unicycle_move_to_pose as U  # noqa: E402  (sets the default dtype to float64, :50)


def main():
    torch.manual_seed(0)
    x0 = torch.tensor([-3.0, -1.0, -np.pi / 4])
    xg = torch.tensor([0.0, 0.0, np.pi / 4])
    numSteps, dt = 2000, 0.001
    planner = U.PiecewiseLinearPlanner(x0, xg, numSteps, dt, frac_time_to_reach_goal=0.95)
    cbfs = U.obstacles_at_mid_from_start_and_goal(x0, xg, term_weights=[0.7, 0.3])
    dyn = U.LearnedShiftInvariantDynamics(dt=dt, mean_dynamics=U.AckermannDrive(L=1.0, kernel_diag_A=[1e-2, 1e-2, 1e-2]),
                                          enable_learning=False)
    ctrl = U.ControllerCLFBayesian(planner, dynamics=dyn, clf=U.CLFCartesian(Kp=torch.tensor([0.9, 1.5, 0.])),
                                   clf_gamma=10., cbfs=cbfs, cbf_gammas=[5., 5.], max_risk=0.01)
    g = torch.Generator().manual_seed(3)
    states = [x0.clone()]
    for _ in range(11):
        states.append(torch.tensor([-3.0, -1.0, 0.0]) + torch.tensor([3.5, 2.0, 1.5]) * torch.rand(3, generator=g))
    ts = [0, 1, 7, 100, 400, 800, 1200, 1500, 1700, 1850, 1899, 1950]
    out = dict(x0=x0.numpy(), xg=xg.numpy(), numSteps=numSteps, dt=dt, states=np.stack([s.numpy() for s in states]),
               ts=np.array(ts), rho=ctrl._factor(),
               obstacle_centers=np.stack([c.center.numpy() for c in cbfs]),
               obstacle_radii=np.array([float(c.radius) for c in cbfs]))
    plans, dplans, clc, cbc = [], [], [], []
    for x, t in zip(states, ts):
        goal = planner.plan(t)
        plans.append(goal.numpy())
        dplans.append(planner.dot_plan(t).numpy())
        A, bfb, bfc, d = ctrl._clc_terms(x, goal, t)
        clc.append(np.concatenate([A.reshape(-1), bfb.reshape(-1), bfc.reshape(-1), np.reshape(d, -1)]))
        row = []
        for (A, bfb, bfc, d) in ctrl._cbcs(x, t):
            row.append(np.concatenate([A.reshape(-1), bfb.reshape(-1), bfc.reshape(-1), np.reshape(d, -1)]))
        cbc.append(np.stack(row))
    out.update(plans=np.stack(plans), dot_plans=np.stack(dplans), clc=np.stack(clc), cbc=np.stack(cbc))
    # pieces, for finer-grained checks
    out['clf'] = np.array([float(ctrl.clf.clf_terms(x, planner.plan(t)).sum()) for x, t in zip(states, ts)])
    out['grad_clf'] = np.stack([ctrl.clf.grad_clf(x, planner.plan(t)).numpy() for x, t in zip(states, ts)])
    out['grad_clf_goal'] = np.stack([ctrl.clf.grad_clf_wrt_goal(x, planner.plan(t)).numpy() for x, t in zip(states, ts)])
    out['cbf'] = np.array([[float(c.cbf(x)) for c in cbfs] for x in states])
    out['grad_cbf'] = np.stack([np.stack([c.grad_cbf(x).numpy() for c in cbfs]) for x in states])
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'ref_controller_f64.npz'), **out)
    print('wrote ref_controller_f64.npz', {k: np.shape(v) for k, v in out.items()})


if __name__ == '__main__':
    main()
