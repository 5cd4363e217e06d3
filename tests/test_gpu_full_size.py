"""Parity at BASELINE.json's full size (N = 16384, the bench workload): against the CPU oracle on 2048 of the bench's own
queries (test_full_size_against_the_oracle: ~1 min of host time for the oracle's factorisations and the measured parity
floor of the mean), and through size-independent identities the exact result must satisfy:

  * factor:   L (L^T z) = (Kb + jitter) z  and  L^-1 (L z) = z  for random probe vectors (a checksum of the N^2 entries);
  * posterior at the training inputs: the variance of F(x_i)[1;u_i] collapses to the jitter level and the mean
    reproduces the targets to the same level;
  * B_k is symmetric positive semi-definite and bounded by the prior; the fold-in form (one column per query,
    bcbf_posterior_fu) equals the [1;u] contraction of the matrix form (bcbf_posterior_blocks);
  * query batching: a permuted query set gives the permuted answers.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def fitted():
    import bench
    from bayesian_cbf_b200.model import MVGPModel, make_hyper
    N = 16384
    X, U, Xdot, hyp, jitter = bench.make_workload(N)
    model = MVGPModel(0)
    model.fit(make_hyper(3, 3, hyp['lengthscale'].numpy(), float(hyp['outputscale']), hyp['A'].numpy(), hyp['B'].numpy(),
                         hyp['C'].numpy()), X.numpy(), U.numpy(), Xdot.numpy(), jitter.numpy(), 1e-5)
    yield model, X, U, Xdot, hyp, jitter
    model.close()


def test_factor_identities_at_full_size(fitted):
    from bayesian_cbf_b200 import ops
    model, X, U, Xdot, hyp, jitter = fitted
    st = model.state_tensors()
    L, Linv = st['L'], st['Linv']
    N = X.shape[0]
    g = torch.Generator().manual_seed(1)
    z = torch.randn(N, 4, generator=g, dtype=torch.float64).cuda()
    Ltz = ops.trmm_lower(L, z, trans=True)
    LLtz = ops.trmm_lower(L, Ltz.contiguous())
    UH = torch.cat([torch.ones(N, 1, dtype=torch.float64), U], dim=1).cuda()
    Kb = ops.gram_train(X.cuda(), UH, hyp['B'].cuda(), hyp['lengthscale'].cuda(), float(hyp['outputscale']))
    Kz = ops.gemm(Kb, z) + 1e-5 * jitter.cuda().unsqueeze(1) * z
    assert ((LLtz - Kz).abs().max() / Kz.abs().max()).item() < 1e-12        # backward error of the factorisation
    back = ops.trmm_lower(Linv, ops.trmm_lower(L, z).contiguous())
    # forward error of the explicit inverse is conditioning-limited: eps * cond(L) ~ 1e-16 * 1e5
    assert ((back - z).abs().max() / z.abs().max()).item() < 1e-8
    del Kb


def test_posterior_properties_at_full_size(fitted):
    model, X, U, Xdot, hyp, jitter = fitted
    prior = float(hyp['outputscale'] * torch.linalg.matrix_norm(hyp['B'], 2))
    # (1) at training inputs
    idx = torch.arange(0, X.shape[0], 37)[:400]
    out = model.query(X[idx].numpy(), U[idx].numpy())
    UHi = torch.cat([torch.ones(len(idx), 1, dtype=torch.float64), U[idx]], dim=1)
    ubu = torch.einsum('qa,ab,qb->q', UHi, hyp['B'], UHi).numpy() * float(hyp['outputscale'])
    assert (out['svar'] > -1e-9 * prior).all()
    assert (out['svar'] < 1e-4 * ubu + 1e-9 * prior).all()                   # collapsed to the 1e-5 jitter level
    assert np.abs(out['mean'] - Xdot[idx].numpy()).max() < 5e-2               # targets carry 0.01 noise; jitter-level fit
    # (2) random queries: symmetry, PSD, bounded by the prior, fold-in == contraction
    g = torch.Generator().manual_seed(3)
    Q = 2000
    Xq = (4 * torch.rand(Q, 3, generator=g, dtype=torch.float64) - 2)
    Uq = (2 * torch.rand(Q, 2, generator=g, dtype=torch.float64) - 1)
    out = model.query(Xq.numpy(), Uq.numpy())
    Bk = torch.from_numpy(out['Bk'])
    assert (Bk - Bk.transpose(1, 2)).abs().max().item() == 0.0
    ev = torch.linalg.eigvalsh(Bk)
    assert ev.min().item() > -1e-9 * prior
    assert ev.max().item() < prior * (1 + 1e-9)
    UHq = torch.cat([torch.ones(Q, 1, dtype=torch.float64), Uq], dim=1)
    svar_c = torch.einsum('qa,qab,qb->q', UHq, Bk, UHq).numpy()
    assert np.abs(out['svar'] - svar_c).max() < 1e-12 * prior * 9
    from bayesian_cbf_b200 import ops
    st = model.state_tensors()
    Ks = ops.cross_gram(st['X'], Xq.cuda(), hyp['lengthscale'].cuda(), float(hyp['outputscale']), Npad=st['Linv'].shape[0])
    sv_fold = ops.posterior_fu_var(st['Linv'], Ks, st['G'], hyp['B'].cuda(), UHq.cuda(), float(hyp['outputscale']), 3, 3)
    assert np.abs(sv_fold.cpu().numpy() - out['svar']).max() < 1e-10 * prior
    # (3) permutation equivariance through the batching loop
    perm = torch.randperm(Q, generator=g)
    outp = model.query(Xq[perm].numpy(), Uq[perm].numpy(), want=('mean', 'svar'))
    assert np.abs(outp['svar'] - out['svar'][perm.numpy()]).max() < 1e-12 * prior
    assert np.abs(outp['mean'] - out['mean'][perm.numpy()]).max() < 1e-11 * max(1.0, np.abs(out['mean']).max())


def test_int8_tensor_core_path_at_full_size(fitted):
    """The int8 (tcgen05, digit-splitting) covariance kernel on the same fitted N=16384 model: it must satisfy the same
    properties and agree with the FP64 DMMA kernel to rounding level (both contract the same L^-1, K*, G)."""
    model, X, U, Xdot, hyp, jitter = fitted
    prior = float(hyp['outputscale'] * torch.linalg.matrix_norm(hyp['B'], 2))
    g = torch.Generator().manual_seed(5)
    Q = 2500                                                                   # 120 column tiles, last one partial
    Xq = (4 * torch.rand(Q, 3, generator=g, dtype=torch.float64) - 2)
    Uq = (2 * torch.rand(Q, 2, generator=g, dtype=torch.float64) - 1)
    idx = torch.arange(0, X.shape[0], 41)[:400]
    Xq[:400], Uq[:400] = X[idx], U[idx]                                        # 400 training inputs among the queries
    ref = model.query(Xq.numpy(), Uq.numpy())
    model.set_var_path('int8')
    try:
        out = model.query(Xq.numpy(), Uq.numpy())
        perm = torch.randperm(Q, generator=g)
        outp = model.query(Xq[perm].numpy(), Uq[perm].numpy(), want=('svar', 'Bk'))
    finally:
        model.set_var_path('dmma')
    assert np.abs(out['Bk'] - ref['Bk']).max() < 1e-11 * prior               # measured 2e-13
    assert np.abs(out['svar'] - ref['svar']).max() < 1e-11 * prior
    Bk = torch.from_numpy(out['Bk'])
    assert (Bk - Bk.transpose(1, 2)).abs().max().item() == 0.0
    ev = torch.linalg.eigvalsh(Bk)
    assert ev.min().item() > -1e-9 * prior and ev.max().item() < prior * (1 + 1e-9)
    UHi = torch.cat([torch.ones(400, 1, dtype=torch.float64), U[idx]], dim=1)
    ubu = torch.einsum('qa,ab,qb->q', UHi, hyp['B'], UHi).numpy() * float(hyp['outputscale'])
    assert (out['svar'][:400] < 1e-4 * ubu + 1e-9 * prior).all()              # collapsed to the jitter level
    # integer arithmetic is exact and the tile a query lands in only changes which other columns share its MMAs:
    # permuting the queries permutes the answers bit for bit
    assert np.array_equal(outp['Bk'], out['Bk'][perm.numpy()])


def test_alpha_refinement_has_converged_at_full_size(fitted):
    """bcbf_model_fit runs three compensated refinement steps on alpha; a fourth must not move the posterior mean."""
    from bayesian_cbf_b200 import ops
    import bench
    model, X, U, Xdot, hyp, jitter = fitted
    st = model.state_tensors()
    N, Npad = X.shape[0], st['Linv'].shape[0]
    UH = torch.cat([torch.ones(N, 1, dtype=torch.float64), U], dim=1).cuda()
    Y = torch.zeros(Npad, 3, dtype=torch.float64, device='cuda')
    Y[:N] = (Xdot - UH.cpu() @ hyp['C']).cuda()
    args = (X.cuda(), UH, hyp['B'].cuda(), hyp['lengthscale'].cuda(), float(hyp['outputscale']), st['Linv'], Y,
            jitter.cuda(), 1e-5)
    a2 = ops.alpha_refine(*args, iters=3)
    a3 = ops.alpha_refine(*args, iters=4)
    a0 = ops.alpha_refine(*args, iters=0)
    assert torch.equal(a2[:N], st['alpha'][:N, :3])                 # what the fit stored
    Xq, Uq = bench.make_queries(1024, 0)
    UHq = torch.cat([torch.ones(1024, 1, dtype=torch.float64), Uq], dim=1).cuda()
    kb = ops.gram_ca(X.cuda(), Xq.cuda(), hyp['lengthscale'].cuda(), float(hyp['outputscale']), UH, UHq, hyp['B'].cuda())
    m0, m2, m3 = [(kb.T @ a[:N]) for a in (a0, a2, a3)]
    sc = m3.abs().max()
    d23, d03 = float((m2 - m3).abs().max() / sc), float((m0 - m3).abs().max() / sc)
    print('mean: explicit inverse only vs converged %.2e, 3 steps vs 4 steps %.2e' % (d03, d23))
    assert d23 < 1e-10


def test_full_size_against_the_oracle(fitted):
    """CUDA (int8 tensor-core and FP64 DMMA covariance kernels) against oracle.posterior_blocks at N = 16384 on 2048 bench
    queries.  Tolerances: B_k / svar 1e-9 of the prior scale; mean / M_k 1e-9 relative or the measured floor of the
    reference's own arithmetic where that is larger (tests/parity_util.py) — against the exact solution of the reference's
    linear system AND against its LAPACK evaluation."""
    import bench
    from oracle import mvgp_oracle as O
    from tests.parity_util import mean_reference, rel
    model, X, U, Xdot, hyp_d, jitter = fitted
    hyp = bench.oracle_hyper(hyp_d)
    Q = 2048
    Xq, Uq = bench.make_queries(Q, 0)
    torch.set_num_threads(max(1, __import__('os').cpu_count() or 1))
    L = O.perturbed_cholesky(hyp, X, O.homogeneous(U), [jitter], direct=True)
    Mk_o, Bk_o, mean_o, svar_o = O.posterior_blocks(hyp, X, U, Xdot, L, Xq, Uq, direct=True, chunk=1024)
    del L
    ref = mean_reference(hyp, X, U, Xdot, jitter, Xq, Uq)
    prior = float(hyp.outputscale * torch.linalg.matrix_norm(hyp.B, 2))
    report = {}
    for path in ('int8', 'dmma'):
        model.set_var_path(path)
        try:
            out = model.query(Xq.numpy(), Uq.numpy())
        finally:
            model.set_var_path('dmma')
        r = dict(Bk=float(np.abs(out['Bk'] - Bk_o.numpy()).max() / prior),
                 svar=float(np.abs(out['svar'] - svar_o.numpy()).max() / prior),
                 mean_vs_exact=rel(out['mean'], ref['mean_exact']), Mk_vs_exact=rel(out['Mk'], ref['Mk_exact']),
                 mean_vs_lapack=rel(out['mean'], mean_o), Mk_vs_lapack=rel(out['Mk'], Mk_o))
        report[path] = r
        print(path, r, {k: ref[k] for k in ('ulp_sensitivity', 'lapack_vs_exact', 'tol_exact', 'tol_lapack')})
        assert r['Bk'] < 1e-9 and r['svar'] < 1e-9, r
        assert r['mean_vs_exact'] < ref['tol_exact'] and r['Mk_vs_exact'] < ref['tol_exact'], (r, ref['tol_exact'])
        assert r['mean_vs_lapack'] < ref['tol_lapack'] and r['Mk_vs_lapack'] < ref['tol_lapack'], (r, ref['tol_lapack'])
