"""TEST INFRASTRUCTURE ONLY — goldens for the LEARNING rollouts (BASELINE configs[4]) from the UNMODIFIED reference:
`LearnedShiftInvariantDynamics` (bayes_cbf/unicycle_move_to_pose.py:295-428) around
`ControlAffineRegressorExactRankOne`, and `ControllerCLFBayesian._clc_terms / _cbcs` (:880-920) on top of it.

A short seeded rollout under a FIXED control law (the reference's own controller needs cvxpy + GUROBI, absent here) with
the recipe of `unicycle_learning_helps_avoid_getting_stuck` (:1948-1969: true AckermannDrive(L=1), prior mean
AckermannDrive(L=12, kernel_diag_A=[1,1,1]), learning on), scaled down (train every 20 steps, max_train 30, 5 Adam steps)
records, at every refit:

  * the training set handed to the regressor: shift-invariant states [0, 0, theta], controls, the finite-difference
    Xdot minus the prior model's prediction, after the random subsampling (:340-386) — plus the subsample indices;
  * the hyper-parameters after the fit;

and, after the last refit, at a few raw states (the reference queries the learned GP at the RAW state, :388-397):

  * the factor jitter and output jitter drawn (make_psd, control_affine_model.py:907-910, :1089) and the posterior
    blocks M_k, A, B_k of the learned part (`_custom_predict_matrix`);
  * the cone terms (A, bfb, bfc, d) of the CLC and both CBCs from the reference controller (each evaluation of the GP
    inside the autograd term extraction draws a fresh output jitter: variance terms are pinned to jitter level only).

    python oracle/gen_golden_learned_dynamics.py   -> tests/golden/ref_learned_dynamics_f64.npz
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

import bayes_cbf.unicycle_move_to_pose as U  # noqa: E402  (default dtype -> float64, :50)

np64 = lambda t: np.array(t.detach().cpu().double().numpy(), copy=True)


def main():
    torch.manual_seed(0)
    np.random.seed(0)
    dt, steps, every, max_train, adam = 0.01, 61, 20, 30, 5
    x0 = torch.tensor([-3.0, -1.0, -np.pi / 4])
    xg = torch.tensor([0.0, 0.0, np.pi / 4])
    true = U.AckermannDrive(L=1.0)
    dyn = U.LearnedShiftInvariantDynamics(dt=dt, mean_dynamics=U.AckermannDrive(L=12.0, kernel_diag_A=[1.0, 1.0, 1.0]),
                                          max_train=max_train, training_iter=adam, train_every_n_steps=every,
                                          enable_learning=True)
    reg = dyn.learned_dynamics
    fits, picks = [], []
    orig_fit = reg.fit
    orig_shuffle = np.random.shuffle

    def rec_fit(X, Uc, Xdot, training_iter=50, **k):
        r = orig_fit(X, Uc, Xdot, training_iter=adam, **k)       # (the reference passes its default 100 here, :386)
        m = reg.model
        p_, n_ = m.matshape
        fits.append(dict(X=np64(X), U=np64(Uc), Xdot=np64(Xdot),
                         lengthscale=np64(m.input_covar.base_kernel.lengthscale.reshape(-1)),
                         outputscale=np64(m.input_covar.outputscale.reshape(())),
                         A=np64(m.task_covar.U.covar_matrix.evaluate()), B=np64(m.task_covar.V.covar_matrix.evaluate()),
                         C=np64(torch.stack([bm.constant.reshape(()) for bm in m.mean_module.base_means]).reshape(p_, n_))))
        return r

    def rec_shuffle(a):
        orig_shuffle(a)
        picks.append(np.array(a, copy=True))

    reg.fit = rec_fit
    np.random.shuffle = rec_shuffle
    xs, us = [], []
    x = x0.clone()
    try:
        for t in range(steps):
            u = torch.tensor([1.0 + 0.3 * np.sin(0.05 * t), 0.5 * np.cos(0.03 * t)])
            dyn.train(x, u)                                        # trains every `every` steps, then records (x, u)
            xs.append(x.clone())
            us.append(u.clone())
            x = x + (true.f_func(x) + true.g_func(x) @ u) * dt     # sampling.py:49-75 with the true model
    finally:
        reg.fit = orig_fit
        np.random.shuffle = orig_shuffle
    out = dict(dt=dt, every=every, max_train=max_train, states=np.stack([np64(v) for v in xs]),
               controls=np.stack([np64(v) for v in us]), nfits=len(fits), x0=np64(x0), xg=np64(xg))
    for i, f in enumerate(fits):
        for k_, v_ in f.items():
            out['fit%d_%s' % (i, k_)] = v_
    for i, pk in enumerate(picks):
        out['pick%d' % i] = pk
    out['npicks'] = len(picks)
    # ---- posterior of the learned part at raw states, jitter recorded --------------------------------------------
    g = torch.Generator().manual_seed(7)
    Xq = torch.stack([xs[-1], xs[30] + 0.1 * torch.rand(3, generator=g), torch.tensor([-2.0, -0.5, 0.3])])
    draws = []
    orig_rand = torch.rand

    def rec_rand(*a, **k):
        r = orig_rand(*a, **k)
        draws.append(r.detach().clone())
        return r

    torch.rand = rec_rand
    try:
        reg.clear_cache()
        blocks = []
        for q in range(Xq.shape[0]):
            Mk, A, Bk = reg._custom_predict_matrix(Xq[q:q + 1], compute_cov=True)
            blocks.append((np64(Mk[0]), np64(A), np64(Bk[0, 0])))
    finally:
        torch.rand = orig_rand
    out['Xq'] = np64(Xq)
    out['factor_jitter'] = np64(draws[0])                           # first draw: make_psd of Kb (cached afterwards)
    out['out_jitter'] = np.stack([np64(d) for d in draws[1:]])      # one (p,) draw per _custom_predict_matrix call
    out['Mk'] = np.stack([b[0] for b in blocks])
    out['Amat'] = blocks[0][1]
    out['Bk_with_jitter'] = np.stack([b[2] for b in blocks])
    # ---- cone terms from the reference controller on the learned dynamics -------------------------------------------
    planner = U.PiecewiseLinearPlanner(x0, xg, 2000, 0.001, frac_time_to_reach_goal=0.95)
    cbfs = U.obstacles_at_mid_from_start_and_goal(x0, xg, term_weights=[0.7, 0.3])
    ctrl = U.ControllerCLFBayesian(planner, dynamics=dyn, clf=U.CLFCartesian(Kp=torch.tensor([0.9, 1.5, 0.])),
                                   clf_gamma=10., cbfs=cbfs, cbf_gammas=[5., 5.], max_risk=0.01)
    ts = [100, 700, 1500]
    clc, cbc = [], []
    for q, t in enumerate(ts):
        xq = Xq[q]
        A_, bfb, bfc, d = ctrl._clc_terms(xq, planner.plan(t), t)
        clc.append(np.concatenate([A_.reshape(-1), bfb.reshape(-1), bfc.reshape(-1), np.reshape(d, -1)]))
        row = []
        for (A_, bfb, bfc, d) in ctrl._cbcs(xq, t):
            row.append(np.concatenate([A_.reshape(-1), bfb.reshape(-1), bfc.reshape(-1), np.reshape(d, -1)]))
        cbc.append(np.stack(row))
    out.update(ts=np.array(ts), clc=np.stack(clc), cbc=np.stack(cbc))
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'ref_learned_dynamics_f64.npz'), **out)
    print('wrote ref_learned_dynamics_f64.npz: %d fits, %d subsamples, train sizes %s' %
          (len(fits), len(picks), [f['X'].shape[0] for f in fits]))


if __name__ == '__main__':
    main()
