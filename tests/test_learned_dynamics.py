"""The learning rollouts of BASELINE configs[4]: `EnsembleLearner` (bayesian_cbf_b200/unicycle.py) against the UNMODIFIED
reference's `LearnedShiftInvariantDynamics` + `ControlAffineRegressorExactRankOne` + `ControllerCLFBayesian` on a short
seeded rollout (oracle/gen_golden_learned_dynamics.py -> tests/golden/ref_learned_dynamics_f64.npz):

  * training-set construction at every refit (shift-invariant states, finite-difference Xdot, residual against the
    prior model, subsampling) — host logic, runs without a GPU;
  * posterior M_k / B_k of the learned part after the last refit, at RAW query states, with the reference's factor
    jitter (CUDA: batched Gram / Cholesky / inverse / ens_posterior_kernel);
  * cone terms (A, bfb, bfc, d) of the CLC and both CBCs through `BayesCBFController` (bcbf_cbc1_terms).
"""
import numpy as np
import pytest
import torch

from tests.golden_util import load


def _feed(learner, d, R, upto):
    X = torch.from_numpy(d['states'])
    U = torch.from_numpy(d['controls'])
    dev = learner.device
    for t in range(upto):
        learner.Xs.append(X[t].repeat(R, 1).to(dev))
        learner.Us.append(U[t].repeat(R, 1).to(dev))


def test_training_sets_match_the_reference():
    from bayesian_cbf_b200.unicycle import EnsembleLearner
    d = load('ref_learned_dynamics_f64')
    R, every = 2, int(d['every'])
    picks = [d['pick%d' % i] for i in range(int(d['npicks']))]
    learner = EnsembleLearner(R, float(d['dt']), model_L=12.0, max_train=int(d['max_train']), train_every_n_steps=every,
                              device='cpu')
    learner.set_subsample_source(picks)
    assert int(d['nfits']) == 3
    for i in range(int(d['nfits'])):
        learner.Xs, learner.Us = [], []
        _feed(learner, d, R, every * (i + 1))          # what LearnedShiftInvariantDynamics.train has seen at refit i
        Xtr, Utr, err = learner.training_set()
        for r in range(R):
            assert np.array_equal(Xtr[r].numpy(), d['fit%d_X' % i])                    # [0, 0, theta]: exact
            assert np.array_equal(Utr[r].numpy(), d['fit%d_U' % i])
            assert np.abs(err[r].numpy() - d['fit%d_Xdot' % i]).max() < 1e-12 * max(1.0, np.abs(d['fit%d_Xdot' % i]).max())
        assert Xtr.shape[1] == min(every * (i + 1) - 1, int(d['max_train']))


def test_record_triggers_refits_like_the_reference(monkeypatch):
    """`record` = LearnedShiftInvariantDynamics.train (:340-354): a refit exactly when len(recorded) is a positive
    multiple of train_every_n_steps, BEFORE the current pair is appended."""
    from bayesian_cbf_b200.unicycle import EnsembleLearner
    d = load('ref_learned_dynamics_f64')
    learner = EnsembleLearner(1, float(d['dt']), train_every_n_steps=20, device='cpu')
    calls = []
    monkeypatch.setattr(learner, 'fit', lambda: calls.append(len(learner.Xs)))
    X, U = torch.from_numpy(d['states']), torch.from_numpy(d['controls'])
    for t in range(61):
        learner.record(t, X[t:t + 1], U[t:t + 1], None, None)
    assert calls == [20, 40, 60]


@pytest.mark.gpu
def test_posterior_and_cone_terms_after_refit_match_the_reference():
    from bayesian_cbf_b200 import unicycle as Un
    d = load('ref_learned_dynamics_f64')
    R, every, last = 3, int(d['every']), int(d['nfits']) - 1
    h = {k: torch.from_numpy(np.asarray(d['fit%d_%s' % (last, k)])) for k in ('lengthscale', 'outputscale', 'A', 'B', 'C')}
    learner = Un.EnsembleLearner(R, float(d['dt']), model_L=12.0, max_train=int(d['max_train']),
                                 train_every_n_steps=every, device='cuda')
    f64 = dict(dtype=torch.float64, device='cuda')
    learner.ls = h['lengthscale'].to(**f64).expand(R, 3).contiguous()
    learner.s = h['outputscale'].to(**f64).expand(R).contiguous()
    learner.A = h['A'].to(**f64).expand(R, 3, 3).contiguous()
    learner.B = h['B'].to(**f64).expand(R, 3, 3).contiguous()
    learner.C = h['C'].to(**f64).expand(R, 3, 3).contiguous()
    learner.set_subsample_source([d['pick%d' % (int(d['npicks']) - 1)]])
    learner.set_jitter_source([np.tile(d['factor_jitter'], (R, 1))])
    _feed(learner, d, R, every * (last + 1))
    learner.fit()
    assert learner.fitted and learner.ens.N == int(d['max_train'])
    prior = float(d['fit%d_outputscale' % last]) * np.linalg.norm(d['fit%d_B' % last], 2)
    Xq = torch.from_numpy(d['Xq']).cuda()
    assert Xq.shape[0] == R                                    # rollout r is asked at raw state r
    Mk, Bk, A = learner.posterior(Xq)
    assert np.abs(Mk.cpu().numpy() - d['Mk']).max() < 1e-9 * max(1.0, np.abs(d['Mk']).max())
    want_Bk = d['Bk_with_jitter'] - 1e-5 * np.stack([np.diag(j) for j in d['out_jitter']])   # the reference ADDED this (:1089)
    assert np.abs(Bk.cpu().numpy() - want_Bk).max() < 1e-9 * prior
    assert np.abs(A[0].cpu().numpy() - d['Amat']).max() < 1e-12
    # shift-invariant querying is the OTHER behaviour: it must differ for states with x, y != 0
    learner.query_raw_state = False
    Mk_si, _, _ = learner.posterior(Xq)
    assert (Mk_si - Mk).abs().max() > 1e-6
    learner.query_raw_state = True
    # cone terms of the reference controller built on the learned dynamics
    planner = Un.PiecewiseLinearPlanner(d['x0'], d['xg'], 2000, 0.001, frac_time_to_reach_goal=0.95)
    cbfs = Un.obstacles_at_mid_from_start_and_goal(d['x0'], d['xg'], term_weights=(0.7, 0.3))
    ctrl = Un.BayesCBFController(planner, Un.CLFCartesian(Kp=(0.9, 1.5, 0.0)), cbfs, [5.0, 5.0], model_L=12.0,
                                 clf_gamma=10.0, max_risk=0.01, posterior=learner.posterior)
    for q, t in enumerate(d['ts']):
        Xr = Xq[q].repeat(R, 1)                               # every rollout at the same state: row q is the golden's
        c, dd, Acone, b = ctrl.constraint_terms(Xr, int(t))
        assert int(ctrl.last_notpd.max()) == 0
        rows = [d['clc'][q]] + [d['cbc'][q, k] for k in range(len(cbfs))]
        for k, want in enumerate(rows):
            A_w, b_w, c_w, d_w = want[:6].reshape(3, 2), want[6:9], want[9:11], want[11]      # A (3,2) | bfb | bfc | d
            # mean terms: exact to 1e-9
            assert np.abs(c[0, k, 1:].cpu().numpy() - c_w).max() < 1e-9 * max(1.0, np.abs(c_w).max()), (q, k)
            assert abs(float(dd[0, k]) - d_w) < 1e-9 * max(1.0, abs(d_w)), (q, k)
            # variance terms: the reference's term extraction evaluates the GP several times, each with a fresh 1e-5
            # output jitter on B_k (:1089) — pinned to that level (relative to the size of the factor)
            got = np.concatenate([Acone[0, k, :, 1:].cpu().numpy().reshape(-1), b[0, k].cpu().numpy()])
            wantv = np.concatenate([A_w.reshape(-1), b_w])
            assert np.abs(got - wantv).max() < 3e-4 * max(np.abs(wantv).max(), 1e-3), (q, k, got, wantv)
