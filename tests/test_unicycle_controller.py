"""Consumer side of the path: planner, CLF/CBF algebra and the cone terms of every constraint of
ControllerCLFBayesian against vectors produced by the UNMODIFIED reference (oracle/gen_golden_controller.py ->
tests/golden/ref_controller_f64.npz), and fixed-seed rollouts whose per-step feasibility decisions must be identical
between the CUDA kernels and their CPU restatements."""
import numpy as np
import pytest
import torch

from tests import fake_ops
from tests.golden_util import load


@pytest.fixture(params=['cpu-fakeops', pytest.param('cuda', marks=pytest.mark.gpu)])
def dev(request, monkeypatch):
    if request.param == 'cuda':
        yield 'cuda'
    else:
        with fake_ops.installed(monkeypatch):
            yield 'cpu'


def _setup(d, dev):
    from bayesian_cbf_b200 import unicycle as U
    planner = U.PiecewiseLinearPlanner(d['x0'], d['xg'], int(d['numSteps']), float(d['dt']), frac_time_to_reach_goal=0.95)
    cbfs = U.obstacles_at_mid_from_start_and_goal(d['x0'], d['xg'], term_weights=(0.7, 0.3))
    ctrl = U.BayesCBFController(planner, U.CLFCartesian(Kp=(0.9, 1.5, 0.0)), cbfs, [5.0, 5.0], model_L=1.0,
                                kernel_diag_A=(1e-2, 1e-2, 1e-2), clf_gamma=10.0, max_risk=0.01)
    return U, planner, cbfs, ctrl


def test_terms_match_reference_controller(dev):
    d = load('ref_controller_f64')
    U, planner, cbfs, ctrl = _setup(d, dev)
    assert abs(ctrl.rho - float(d['rho'])) < 1e-14
    for i, c in enumerate(cbfs):
        assert np.allclose(c.center, d['obstacle_centers'][i]) and abs(c.radius - d['obstacle_radii'][i]) < 1e-12
    X = torch.from_numpy(d['states']).to(dev)
    for i, t in enumerate(d['ts']):
        t = int(t)
        assert np.allclose(planner.plan(t), d['plans'][i], rtol=1e-12, atol=1e-12)
        assert np.allclose(planner.dot_plan(t), d['dot_plans'][i], rtol=1e-12, atol=1e-12, equal_nan=True)
        x = X[i:i + 1]
        goal = torch.tensor(planner.plan(t), dtype=torch.float64, device=dev).reshape(1, 3)
        assert abs(ctrl.clf.clf(x, goal).item() - d['clf'][i]) < 1e-12 * max(1, abs(d['clf'][i]))
        assert np.allclose(ctrl.clf.grad_clf(x, goal).cpu().numpy()[0], d['grad_clf'][i], rtol=1e-11, atol=1e-12)
        assert np.allclose(ctrl.clf.grad_clf_wrt_goal(x, goal).cpu().numpy()[0], d['grad_clf_goal'][i], rtol=1e-11, atol=1e-12)
        for k, cbf in enumerate(cbfs):
            assert abs(cbf.cbf(x).item() - d['cbf'][i, k]) < 1e-12 * max(1, abs(d['cbf'][i, k]))
            assert np.allclose(cbf.grad_cbf(x).cpu().numpy()[0], d['grad_cbf'][i, k], rtol=1e-11, atol=1e-12)
        c, dd, A, b = ctrl.constraint_terms(x, t)
        rows = [d['clc'][i]] + [d['cbc'][i, k] for k in range(len(cbfs))]
        for k, want in enumerate(rows):
            got = np.concatenate([A[0, k, :, 1:].cpu().numpy().reshape(-1), b[0, k].cpu().numpy(),
                                  c[0, k, 1:].cpu().numpy(), dd[0, k].cpu().numpy().reshape(-1)])
            # the reference extracts these with autograd at a random linearisation point (cbc2.py:7-23): 1e-9
            assert np.abs(got - want).max() < 1e-9 * max(1.0, np.abs(want).max()), (i, k, got, want)
            assert float(c[0, k, 0]) == (1.0 if k == 0 else 0.0)
            assert float(A[0, k, :, 0].abs().max()) == 0.0


def _run(dev, steps=120, R=5):
    d = load('ref_controller_f64')
    U, planner, cbfs, ctrl = _setup(d, dev)
    g = torch.Generator().manual_seed(5)
    X0 = torch.from_numpy(d['x0']).repeat(R, 1) + 0.2 * (torch.rand(R, 3, generator=g, dtype=torch.float64) - 0.5)
    X0[0] = torch.from_numpy(d['x0'])
    # one start well inside an obstacle's unsafe set: its program must be infeasible from the first step
    X0[R - 1] = torch.tensor([cbfs[0].center[0], cbfs[0].center[1] + 0.05, -1.5], dtype=torch.float64)
    return U.rollout(ctrl, X0.to(dev), steps, float(d['dt']), true_L=12.0)


def test_rollout_runs_and_detects_infeasibility(dev):
    out = _run(dev, steps=15)
    feas = out['feasible'].cpu().numpy()
    assert feas[:, 0].all()               # the recipe's own start stays feasible
    assert not feas[:, -1].any()          # inside the obstacle: infeasible from step 0
    assert out['X'].shape == (16, 5, 3) and torch.isfinite(out['X']).all()


@pytest.mark.gpu
def test_rollout_decisions_identical_cuda_vs_restatement(monkeypatch):
    """Fixed-seed rollouts: the CUDA path (bcbf_cbc1_terms + bcbf_socp_solve) and the CPU restatements make the same
    feasibility decision at every step, and the trajectories agree to 1e-7."""
    got = _run('cuda')
    with fake_ops.installed(monkeypatch):
        want = _run('cpu')
    assert torch.equal(got['feasible'].cpu(), want['feasible'])
    assert (got['X'].cpu() - want['X']).abs().max() < 1e-7
    assert (got['U'].cpu() - want['U']).abs().max() < 1e-6


@pytest.mark.gpu
def test_graphed_rollout_equals_eager_rollout():
    """The CUDA-graph replay of the control step reproduces the eager rollout bit for bit (same kernels, same order)."""
    d = load('ref_controller_f64')
    U, planner, cbfs, ctrl = _setup(d, 'cuda')
    g = torch.Generator().manual_seed(5)
    R = 6
    X0 = (torch.from_numpy(d['x0']).repeat(R, 1) + 0.2 * (torch.rand(R, 3, generator=g, dtype=torch.float64) - 0.5)).cuda()
    X0[R - 1] = torch.tensor([cbfs[0].center[0], cbfs[0].center[1] + 0.05, -1.5], dtype=torch.float64)
    eager = U.rollout(ctrl, X0, 40, float(d['dt']), true_L=12.0)
    graphed = U.GraphedRollout(ctrl, X0, float(d['dt']), true_L=12.0).capture().run(40)
    assert torch.equal(eager['feasible'], graphed['feasible'])
    assert torch.equal(eager['X'], graphed['X'])
    assert torch.equal(eager['U'], graphed['U'])


@pytest.mark.gpu
def test_full_2000_step_recipe_against_the_recorded_run():
    """BASELINE configs[2], all 2000 steps of `unicycle_bayes_cbf_safe_obstacle`: the CUDA rollout (captured control
    step: bcbf_cbc1_terms + bcbf_socp_solve) against the per-step record of the CPU restatements
    (oracle/gen_rollout_record.py -> tests/golden/rollout_safe_obstacle_2000.npz): the same feasibility decision at
    every step, states / controls / cone terms to 1e-7."""
    from oracle import gen_rollout_record as G
    d = load('rollout_safe_obstacle_2000')
    steps = int(d['steps'])
    rec = G.run('cuda', steps=steps)
    feas = np.unpackbits(d['feasible'])[:steps].astype(bool)
    assert np.array_equal(rec['feasible'], feas) and int(d['n_feasible']) == int(feas.sum())
    keep = np.arange(0, steps + 1, int(d['sample_every']))
    assert np.abs(rec['X'][keep] - d['X']).max() < 1e-7
    assert np.abs(rec['U'][keep[:-1]] - d['U']).max() < 1e-6
    assert np.abs(rec['X'][-1] - d['X_final']).max() < 1e-7
    assert np.array_equal(rec['term_steps'], d['term_steps'])
    assert np.abs(rec['terms'] - d['terms']).max() < 1e-7 * max(1.0, np.abs(d['terms']).max())
    # and the CUDA-graph replay of the same 2000 steps is bit-identical to the eager CUDA rollout
    U, ctrl, X0, dt, _ = G.build('cuda')
    graphed = U.GraphedRollout(ctrl, X0, dt, true_L=12.0).capture().run(steps)
    assert np.array_equal(graphed['X'][:, 0].cpu().numpy(), rec['X'])
    assert np.array_equal(graphed['feasible'][:, 0].cpu().numpy(), rec['feasible'])
