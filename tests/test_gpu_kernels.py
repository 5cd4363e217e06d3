"""GPU parity tests: every CUDA kernel (called through the C ABI) against the CPU oracle on the same
seeded inputs.  float64; tolerances are norm-wise relative, written next to each check."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import mvgp_oracle as O
from tests.parity_util import mean_reference, rel

pytestmark = pytest.mark.gpu


def _mk(seed, N, n, m, Q, box=2.0):
    g = torch.Generator().manual_seed(seed)
    p = m + 1
    X = box * (2 * torch.rand(N, n, generator=g, dtype=torch.float64) - 1)
    U = 2 * torch.rand(N, m, generator=g, dtype=torch.float64) - 1
    Xdot = torch.sin(X @ torch.randn(n, n, generator=g, dtype=torch.float64)) \
        + 0.01 * torch.randn(N, n, generator=g, dtype=torch.float64)
    Ra = torch.randn(n, n, generator=g, dtype=torch.float64)
    Rb = torch.randn(p, p, generator=g, dtype=torch.float64)
    hyp = O.Hyper(lengthscale=0.6 + 0.5 * torch.rand(n, generator=g, dtype=torch.float64),
                  outputscale=torch.tensor(1.3, dtype=torch.float64),
                  A=Ra @ Ra.T + torch.eye(n, dtype=torch.float64),
                  B=Rb @ Rb.T + torch.eye(p, dtype=torch.float64),
                  C=0.3 * torch.randn(p, n, generator=g, dtype=torch.float64))
    jit = torch.rand(N, generator=g, dtype=torch.float64)
    Xq = box * (2 * torch.rand(Q, n, generator=g, dtype=torch.float64) - 1)
    Uq = 2 * torch.rand(Q, m, generator=g, dtype=torch.float64) - 1
    return X, U, Xdot, hyp, jit, Xq, Uq


def _d(t):
    return t.contiguous().cuda()


def _fma_diag(diag, jit, scale):
    """diag + scale * jit rounded once, as the fused multiply-add of potf2_inv_kernel / gram_resid_kernel does (long
    double would round twice: with |alpha| ~ 1e4 a last-bit difference on the diagonal shows up in the residual)."""
    from fractions import Fraction
    s = Fraction(scale)
    N = len(jit)          # the pad diagonal (if any) carries no jitter
    out = [float(Fraction(a) + s * Fraction(b)) for a, b in zip(diag[:N].tolist(), jit.tolist())] + diag[N:].tolist()
    return torch.tensor(out, dtype=torch.float64)


def _relerr(got, want):
    want = want if isinstance(want, torch.Tensor) else torch.as_tensor(want)
    return ((got.cpu() - want).abs().max() / want.abs().max().clamp_min(1e-300)).item()


@pytest.mark.parametrize('N,n,m', [(300, 3, 2), (129, 2, 1), (64, 1, 3)])
def test_gram_train(N, n, m):
    from bayesian_cbf_b200 import ops
    X, U, _, hyp, _, _, _ = _mk(1, N, n, m, 4)
    UH = O.homogeneous(U)
    Kb = ops.gram_train(_d(X), _d(UH), _d(hyp.B), _d(hyp.lengthscale), float(hyp.outputscale))
    Npad = ops.padded(N)
    assert Kb.shape == (Npad, Npad)
    ref = O.gram_train(hyp, X, UH, direct=True)
    assert _relerr(Kb[:N, :N], ref) < 1e-14          # same formula, FMA-level differences only
    ref_g = O.gram_train(hyp, X, UH, direct=False)    # gpytorch's expanded-distance form
    assert _relerr(Kb[:N, :N], ref_g) < 1e-13
    pad = Kb.cpu()[N:, N:]
    assert torch.equal(pad, torch.eye(Npad - N, dtype=torch.float64))
    assert Kb.cpu()[N:, :N].abs().max() == 0 and Kb.cpu()[:N, N:].abs().max() == 0


@pytest.mark.parametrize('N,n,m', [(300, 3, 2), (129, 2, 1), (200, 1, 3), (1000, 3, 2)])
def test_gram_train_lower_has_the_bits_of_the_full_matrix(N, n, m):
    """The fit path builds only the 64x64 tiles on/below the diagonal (bcbf_gram_train_lower); entry (i,j), j <= i, is
    bit-identical to the full matrix, which is exactly symmetric (the upper triangle mirrors the lower one)."""
    from bayesian_cbf_b200 import ops
    X, U, _, hyp, _, _, _ = _mk(1, N, n, m, 4)
    UH = O.homogeneous(U)
    args = (_d(X), _d(UH), _d(hyp.B), _d(hyp.lengthscale), float(hyp.outputscale))
    full = ops.gram_train(*args).cpu()
    low = ops.gram_train_lower(*args).cpu()
    assert torch.equal(full, full.T)
    assert torch.equal(torch.tril(low), torch.tril(full))
    # storage above the lower TILES is left untouched: poison the buffer and look
    Npad = full.shape[0]
    buf = torch.full((Npad, Npad), 7.0, dtype=torch.float64, device='cuda')
    from bayesian_cbf_b200 import _lib
    _lib.check(_lib.load().bcbf_gram_train_lower(args[0].data_ptr(), args[1].data_ptr(), args[2].data_ptr(),
                                                 args[3].data_ptr(), args[4], N, n, m + 1, buf.data_ptr(), Npad, Npad,
                                                 torch.cuda.current_stream().cuda_stream))
    buf = buf.cpu()
    tile_r = torch.arange(Npad).unsqueeze(1) // 64
    tile_c = torch.arange(Npad).unsqueeze(0) // 64
    assert (buf[tile_c > tile_r] == 7.0).all() and torch.equal(buf[tile_c <= tile_r], full[tile_c <= tile_r])


@pytest.mark.parametrize('N,n,m', [(700, 3, 2), (333, 2, 1), (260, 4, 3), (1500, 3, 2)])
def test_gram_resid_is_the_exact_residual(N, n, m):
    """bcbf_gram_resid: Y - (Kb + jitter) alpha with Kb re-evaluated on the fly and Dot2 accumulation, against the
    oracle's error-free residual of the SAME matrix (the GPU's own Gram entries, downloaded)."""
    from bayesian_cbf_b200 import ops
    X, U, Xdot, hyp, jit, _, _ = _mk(13, N, n, m, 4, box=2.0)
    UH = O.homogeneous(U)
    args = (_d(X), _d(UH), _d(hyp.B), _d(hyp.lengthscale), float(hyp.outputscale))
    Kb = ops.gram_train(*args).cpu()[:N, :N]
    Kbp = Kb.clone()
    Kbp.diagonal().copy_(_fma_diag(Kb.diagonal(), jit, 1e-5))               # fma(jscale, jitter, Kb_ii): ONE rounding
    Y = O.residual_targets(hyp, UH, Xdot)
    alpha = torch.cholesky_solve(Y, torch.linalg.cholesky(Kbp))
    R = ops.gram_resid(*args, _d(alpha), _d(Y), _d(jit), 1e-5).cpu()
    R_exact = O.residual_exact(Kbp, alpha, Y)
    scale = (Kbp.abs() @ alpha.abs())
    # Dot2: error <= eps |result| + O(N eps^2) |Kb||alpha|; the oracle itself truncates at 2^-76 of scale per term
    # (4 x 19-bit slices, ~1e-23 N of scale).  Plain float64 would be ~1e-16 * scale, orders of magnitude larger.
    assert ((R - R_exact).abs() / scale).max().item() < 1e-19 + 3e-16 * (R_exact.abs() / scale).max().item()
    assert ((Y - Kbp @ alpha - R_exact).abs() / scale).max().item() > 1e-18      # (the float64 residual is not)


@pytest.mark.parametrize('N,n,m', [(300, 3, 2), (1000, 2, 1), (2111, 3, 2)])
def test_stored_residual_and_refinement_equal_the_on_the_fly_ones_bit_for_bit(N, n, m):
    """bcbf_gram_resid_stored reads the lower triangle bcbf_gram_train_lower wrote (garbage in the upper one) and must
    return the bits of bcbf_gram_resid; bcbf_alpha_refine_ws with and without the Kb workspace likewise."""
    from bayesian_cbf_b200 import ops
    X, U, Xdot, hyp, jit, _, _ = _mk(17, N, n, m, 4, box=2.0)
    UH = O.homogeneous(U)
    args = (_d(X), _d(UH), _d(hyp.B), _d(hyp.lengthscale), float(hyp.outputscale))
    Y = _d(O.residual_targets(hyp, UH, Xdot))
    g = torch.Generator().manual_seed(3)
    alpha = _d(torch.randn(N, n, generator=g, dtype=torch.float64))
    Npad = ops.padded(N)
    Kb = torch.full((Npad, Npad), float('nan'), dtype=torch.float64, device='cuda')
    Kb = torch.triu(Kb, diagonal=1)                                # NaN strictly above the diagonal, zeros below
    Kb = Kb + torch.tril(ops.gram_train_lower(*args))              # the lower triangle as the factorisation sees it
    assert torch.isnan(Kb[0, 1])
    R1 = ops.gram_resid(*args, alpha, Y, _d(jit), 1e-5)
    R2 = ops.gram_resid_stored(Kb, alpha, Y, _d(jit), 1e-5)
    assert torch.equal(R1, R2)
    L, dinv = ops.potrf_(ops.gram_train_lower(*args), N, _d(jit), 1e-5)
    Linv = ops.trtri(L, dinv)
    Yp = torch.zeros(Npad, n, dtype=torch.float64, device='cuda')
    Yp[:N] = Y
    a1 = ops.alpha_refine(*args, Linv, Yp, _d(jit), 1e-5, iters=3, store_kb=False)
    a2 = ops.alpha_refine(*args, Linv, Yp, _d(jit), 1e-5, iters=3, store_kb=True)
    assert torch.equal(a1, a2)


def test_small_tile_gemm_variant_returns_the_same_bits():
    """The 32 x 128 row-tile variant of the FP64 GEMM (picked for problems of a few tiles) against the 128 x 128 one:
    general products in the four storage orders with ragged extents, the triangular products, and a whole blocked
    Cholesky + triangular inverse (in-place panel solves, SYRK updates, k-range masks) — bit for bit."""
    from bayesian_cbf_b200 import _lib, ops
    lib = _lib.load()
    g = torch.Generator().manual_seed(5)
    rnd = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64).cuda()

    def run(policy):
        _lib.check(lib.bcbf_set_gemm_tile_policy(policy))
        out = []
        try:
            for (M, N, K) in ((2, 802, 256), (300, 130, 260), (97, 33, 515)):
                for ta in (False, True):
                    for tb in (False, True):
                        A = rnd(*((K, M) if ta else (M, K)))
                        B = rnd(*((N, K) if tb else (K, N)))
                        C = rnd(M, N)
                        out.append(ops.gemm(A, B, transa=ta, transb=tb, alpha=-0.5, beta=1.0, C=C))
            n = 640
            Mx = rnd(n, n)
            S = Mx @ Mx.t() + n * torch.eye(n, dtype=torch.float64, device='cuda')
            L, dinv = ops.potrf_(S.clone(), n, None, 0.0)
            Linv = ops.trtri(L, dinv)
            Bm = rnd(n, 34)
            out += [L.clone(), Linv, ops.trmm_lower(Linv, Bm), ops.trmm_lower(Linv, Bm, trans=True)]
        finally:
            _lib.check(lib.bcbf_set_gemm_tile_policy(0))
        return out

    g.manual_seed(5)
    big = run(1)
    g.manual_seed(5)
    small = run(2)
    assert len(big) == len(small) == 16
    for a, b in zip(big, small):
        assert torch.equal(a, b)
    L, Linv = big[12], big[13]
    assert (torch.tril(L) @ Linv - torch.eye(640, dtype=torch.float64, device='cuda')).abs().max().item() < 1e-10


def test_alpha_refine_reaches_the_exact_solution_of_the_factorised_matrix():
    """N = 2048 of the bench workload (cond ~ 4e9): bcbf_alpha_refine (explicit inverse + 2 compensated refinement steps)
    returns the float64 rounding of the exact solution of the system the GPU factorised — closer to it than LAPACK's
    cholesky_solve of the same matrix, which is what the reference runs (control_affine_model.py:545)."""
    import bench
    from bayesian_cbf_b200 import ops
    N = 2048
    X, U, Xdot, hyp_d, jit = bench.make_workload(N)
    hyp = bench.oracle_hyper(hyp_d)
    Xq, Uq = bench.make_queries(512, 0)
    UH = O.homogeneous(U)
    args = (_d(X), _d(UH), _d(hyp.B), _d(hyp.lengthscale), float(hyp.outputscale))
    Kb = ops.gram_train(*args).cpu()
    Kbp = Kb.clone()
    Kbp.diagonal().copy_(_fma_diag(Kb.diagonal(), jit, 1e-5))               # fma(jscale, jitter, Kb_ii): ONE rounding
    L, dinv = ops.potrf_(ops.gram_train_lower(*args), N, _d(jit), 1e-5)
    Linv = ops.trtri(L, dinv)
    Y = O.residual_targets(hyp, UH, Xdot)
    a0 = ops.alpha_refine(*args, Linv, _d(Y), _d(jit), 1e-5, iters=0).cpu()
    a2 = ops.alpha_refine(*args, Linv, _d(Y), _d(jit), 1e-5, iters=2).cpu()
    a3 = ops.alpha_refine(*args, Linv, _d(Y), _d(jit), 1e-5, iters=3).cpu()
    Lc = torch.linalg.cholesky(Kbp)
    a_exact, last = O.solve_exact(Kbp, Lc, Y)
    a_lapack = torch.cholesky_solve(Y, Lc)
    assert last < 1e-15
    kb = O.rbf_ard(X, Xq, hyp.lengthscale, hyp.outputscale, direct=True) * ((UH @ hyp.B) @ O.homogeneous(Uq).t())
    m_exact = kb.t() @ a_exact
    err = lambda a: float((kb.t() @ a - m_exact).abs().max() / m_exact.abs().max())
    e0, e2, e3, el = err(a0), err(a2), err(a3), err(a_lapack)
    print('mean error vs exact solution of the same matrix: inverse only %.2e, 2 steps %.2e, 3 steps %.2e, LAPACK %.2e'
          % (e0, e2, e3, el))
    assert e2 < 5e-10 and e3 == e2              # converged: only the float64 rounding of alpha (1e6-fold cancellation) is left
    assert e2 < 0.1 * el                        # an order of magnitude closer than the reference's own solve
    assert float((a2 - a_exact).abs().max() / a_exact.abs().max()) < 1e-13


def test_cross_gram_and_rbf_blocks():
    from bayesian_cbf_b200 import ops
    X, _, _, hyp, _, Xq, _ = _mk(2, 200, 3, 2, 70)
    Ks = ops.cross_gram(_d(X), _d(Xq), _d(hyp.lengthscale), float(hyp.outputscale))
    assert Ks.shape == (256, 192)
    ref = O.rbf_ard(X, Xq, hyp.lengthscale, hyp.outputscale, direct=True)
    assert _relerr(Ks[:200, :70], ref) < 1e-14
    assert Ks.cpu()[200:].abs().max() == 0 and Ks.cpu()[:, 70:].abs().max() == 0
    K, dK, d2K = ops.rbf_blocks(_d(Xq[:5]), _d(X[:7]), _d(hyp.lengthscale), float(hyp.outputscale), True, True)
    for i in range(5):
        for j in range(7):
            k, g, H = O.rbf_grad_hess(Xq[i], X[j], hyp.lengthscale, hyp.outputscale)
            assert abs(K[i, j].item() - k.item()) < 1e-14
            assert (dK[i, j].cpu() - g).abs().max() < 1e-13
            assert (d2K[i, j].cpu() - H).abs().max() < 1e-13


@pytest.mark.parametrize('N', [100, 128, 300, 640, 1100, 2100])
def test_potrf_trtri(N):
    from bayesian_cbf_b200 import ops
    X, U, _, hyp, jit, _, _ = _mk(3, N, 3, 2, 4, box=3.0)
    UH = O.homogeneous(U)
    Kb = ops.gram_train(_d(X), _d(UH), _d(hyp.B), _d(hyp.lengthscale), float(hyp.outputscale))
    Kb_host = Kb.cpu().clone()
    Npad = Kb.shape[0]
    L, dinv = ops.potrf_(Kb, N, _d(jit), 1e-5)
    Lh = L.cpu()
    jpad = torch.zeros(Npad, dtype=torch.float64)
    jpad[:N] = 1e-5 * jit
    Kbp = Kb_host + torch.diag(jpad)
    assert torch.equal(Lh, torch.tril(Lh))                               # upper triangle zeroed
    resid = (Lh @ Lh.T - Kbp).abs().max() / Kbp.abs().max()
    assert resid < 1e-14, resid                                          # backward error of the factorisation
    Lref = torch.linalg.cholesky(Kbp)
    # forward agreement with LAPACK is conditioning-limited: eps * cond(Kb) — bound it loosely
    assert _relerr(L, Lref) < 1e-7
    Linv = ops.trtri(L, dinv).cpu()
    assert torch.equal(Linv, torch.tril(Linv))
    I = torch.eye(Npad, dtype=torch.float64)
    condL = torch.linalg.cond(Lref).item()
    assert (Linv @ Lh - I).abs().max() < 1e-15 * condL * 50 + 1e-12


def test_potrf_not_pd_raises_runtimeerror():
    from bayesian_cbf_b200 import ops
    A = torch.eye(256, dtype=torch.float64)
    A[200, 200] = -1.0
    with pytest.raises(RuntimeError, match="not positive-definite"):
        ops.potrf_(_d(A), 256, None, 0.0)


def test_trmm_lower():
    from bayesian_cbf_b200 import ops
    g = torch.Generator().manual_seed(5)
    Npad = 384
    A = torch.tril(torch.randn(Npad, Npad, generator=g, dtype=torch.float64))
    Bm = torch.randn(Npad, 7, generator=g, dtype=torch.float64)
    C = ops.trmm_lower(_d(A), _d(Bm))
    assert _relerr(C, A @ Bm) < 1e-13
    Ct = ops.trmm_lower(_d(A), _d(Bm), trans=True, alpha=-2.0)
    assert _relerr(Ct, -2.0 * A.T @ Bm) < 1e-13


def _fit_on_gpu(X, U, Xdot, hyp, jit):
    from bayesian_cbf_b200 import ops
    N = X.shape[0]
    UH = O.homogeneous(U)
    args = (_d(X), _d(UH), _d(hyp.B), _d(hyp.lengthscale), float(hyp.outputscale))
    Kb = ops.gram_train_lower(*args)
    L, dinv = ops.potrf_(Kb, N, _d(jit), 1e-5)
    Linv = ops.trtri(L, dinv)
    Npad = L.shape[0]
    G = torch.zeros(Npad, hyp.p, dtype=torch.float64)
    G[:N] = UH @ hyp.B
    Y = torch.zeros(Npad, hyp.n, dtype=torch.float64)
    Y[:N] = O.residual_targets(hyp, UH, Xdot)
    alpha = ops.alpha_refine(*args, Linv, _d(Y), _d(jit), 1e-5, iters=3).contiguous()
    W = (alpha.unsqueeze(-1) * _d(G).unsqueeze(1)).reshape(Npad, -1).contiguous()
    return L, Linv, _d(G), alpha, W


@pytest.mark.parametrize('N,n,m,Q', [(500, 3, 2, 200), (1000, 2, 1, 333), (260, 3, 3, 50), (150, 2, 0, 40)])
def test_posterior_blocks(N, n, m, Q):
    from bayesian_cbf_b200 import ops
    X, U, Xdot, hyp, jit, Xq, Uq = _mk(7, N, n, m, Q, box=3.0)
    p = m + 1
    L, Linv, G, alpha, W = _fit_on_gpu(X, U, Xdot, hyp, jit)
    Ks = ops.cross_gram(_d(X), _d(Xq), _d(hyp.lengthscale), float(hyp.outputscale))
    Mk, Bk = ops.posterior_blocks(Linv, Ks, G, W, _d(hyp.B), _d(hyp.C.t()), float(hyp.outputscale), n, p, Q)
    # oracle on the SAME jitter (CPU LAPACK factor)
    Lref = O.perturbed_cholesky(hyp, X, O.homogeneous(U), [jit], direct=True)
    Mk_o, Bk_o, mean_o, svar_o = O.posterior_blocks(hyp, X, U, Xdot, Lref, Xq, Uq, direct=True)
    prior = (hyp.outputscale * torch.linalg.matrix_norm(hyp.B, 2)).item()
    # covariance: norm-wise relative to the prior scale s*|B| (difference of nearly equal numbers)
    assert (Bk.cpu() - Bk_o).abs().max() / prior < 1e-9
    # mean: 1e-9 relative, or the measured floor of the reference's own arithmetic where that is larger (parity_util)
    ref = mean_reference(hyp, X, U, Xdot, jit, Xq, Uq)
    assert rel(Mk, ref['Mk_exact']) < ref['tol_exact'], (rel(Mk, ref['Mk_exact']), ref)
    assert rel(Mk, Mk_o) < ref['tol_lapack'], (rel(Mk, Mk_o), ref)
    # u-contraction
    UHq = O.homogeneous(Uq)
    mean, svar = ops.contract_u(Mk, Bk, _d(UHq))
    assert rel(mean, ref['mean_exact']) < ref['tol_exact'], (rel(mean, ref['mean_exact']), ref)
    assert rel(mean, mean_o) < ref['tol_lapack']
    assert (svar.cpu() - svar_o).abs().max() / prior < 1e-8
    # fold-in form
    sv2 = ops.posterior_fu_var(Linv, Ks, G, _d(hyp.B), _d(UHq), float(hyp.outputscale), n, p)
    assert (sv2.cpu() - svar_o).abs().max() / prior < 1e-8
    assert (svar_o > -1e-9 * prior).all()


def test_cbc1_terms():
    from bayesian_cbf_b200 import ops
    g = torch.Generator().manual_seed(9)
    Q, n, p = 37, 3, 3
    Mk = torch.randn(Q, n, p, generator=g, dtype=torch.float64)
    R = torch.randn(Q, p, p, generator=g, dtype=torch.float64)
    Bk = R @ R.transpose(1, 2) + 0.1 * torch.eye(p, dtype=torch.float64)
    Ra = torch.randn(n, n, generator=g, dtype=torch.float64)
    A = Ra @ Ra.T + torch.eye(n, dtype=torch.float64)
    gh = torch.randn(Q, n, generator=g, dtype=torch.float64)
    h = torch.randn(Q, generator=g, dtype=torch.float64)
    Fbar = torch.randn(Q, n, p, generator=g, dtype=torch.float64)
    bfe, e, Asq, A_socp, bfb, status = ops.cbc1_terms(_d(Mk), _d(Bk), _d(A), _d(gh), _d(h), 0.7, _d(Fbar))
    assert (status.cpu() == 0).all()
    for q in range(Q):
        bfe_o, e_o, V, bfv, v = O.cbc1_terms_closed_form(Mk[q], Bk[q], A, gh[q], h[q], 0.7, Fbar[q])
        A_o, bfb_o, bfc_o, d_o = O.convert_cbc_terms_to_socp_terms(bfe_o, e_o, V, bfv, v, 0)
        assert (bfe[q].cpu() - bfe_o).abs().max() < 1e-12
        assert abs(e[q].item() - e_o.item()) < 1e-12
        assert (A_socp[q].cpu() - A_o).abs().max() < 1e-11
        assert (bfb[q].cpu() - bfb_o).abs().max() < 1e-11


def test_model_handle_host_pointers():
    """bcbf_model_* entry points with HOST buffers (what bench.py's e2e leg and a non-torch caller bind)."""
    from bayesian_cbf_b200 import _lib
    lib = _lib.load()
    N, n, m, Q = 700, 3, 2, 1000
    p = m + 1
    X, U, Xdot, hyp, jit, Xq, Uq = _mk(11, N, n, m, Q, box=3.0)
    h = _lib.Hyper()
    h.n, h.p, h.outputscale = n, p, float(hyp.outputscale)
    for i, v in enumerate(hyp.lengthscale.tolist()):
        h.lengthscale[i] = v
    for i, v in enumerate(hyp.A.reshape(-1).tolist()):
        h.A[i] = v
    for i, v in enumerate(hyp.B.reshape(-1).tolist()):
        h.B[i] = v
    for i, v in enumerate(hyp.C.reshape(-1).tolist()):
        h.C[i] = v
    model = ctypes.c_void_p()
    _lib.check(lib.bcbf_model_create(ctypes.byref(model), 0))
    try:
        arrs = [a.contiguous().numpy() for a in (X, U, Xdot, jit, Xq, Uq)]
        ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        _lib.check(lib.bcbf_model_fit(model, ctypes.byref(h), ptr(arrs[0]), ptr(arrs[1]), ptr(arrs[2]), N,
                                      ptr(arrs[3]), 1e-5))
        mean = np.empty((Q, n)); svar = np.empty(Q); Mk = np.empty((Q, n, p)); Bk = np.empty((Q, p, p))
        _lib.check(lib.bcbf_model_query(model, ptr(arrs[4]), ptr(arrs[5]), Q, ptr(mean), ptr(svar), ptr(Mk), ptr(Bk)))
        ms = (ctypes.c_double * 5)()
        _lib.check(lib.bcbf_model_fit_timing(model, ctypes.byref(ms)))
        assert ms[4] > 0
    finally:
        lib.bcbf_model_destroy(model)
    Lref = O.perturbed_cholesky(hyp, X, O.homogeneous(U), [jit], direct=True)
    Mk_o, Bk_o, mean_o, svar_o = O.posterior_blocks(hyp, X, U, Xdot, Lref, Xq, Uq, direct=True)
    prior = (hyp.outputscale * torch.linalg.matrix_norm(hyp.B, 2)).item()
    assert np.abs(Bk - Bk_o.numpy()).max() / prior < 1e-9
    assert np.abs(svar - svar_o.numpy()).max() / prior < 1e-8
    ref = mean_reference(hyp, X, U, Xdot, jit, Xq, Uq)
    assert rel(Mk, ref['Mk_exact']) < ref['tol_exact'] and rel(mean, ref['mean_exact']) < ref['tol_exact'], ref
    assert rel(Mk, Mk_o) < ref['tol_lapack'] and rel(mean, mean_o) < ref['tol_lapack'], ref


def test_model_handle_batches_large_query_sets_and_reports_errors():
    """bcbf_model_query splits query sets larger than one device batch (148 * 32 * 4 queries at p = 3) and the pieces
    agree with a single small call; error paths return codes, not crashes."""
    from bayesian_cbf_b200 import _lib
    from bayesian_cbf_b200.model import MVGPModel, make_hyper
    N, n, m = 260, 3, 2
    X, U, Xdot, hyp, jit, _, _ = _mk(21, N, n, m, 4)
    model = MVGPModel(0)
    # querying before fitting is an error, not a crash
    with pytest.raises(_lib.BcbfError, match="not fitted"):
        model.hyper = make_hyper(n, m + 1, hyp.lengthscale.numpy(), 1.3, hyp.A.numpy(), hyp.B.numpy(), hyp.C.numpy())
        model.query(np.zeros((2, n)), np.zeros((2, m)))
    model.fit(make_hyper(n, m + 1, hyp.lengthscale.numpy(), float(hyp.outputscale), hyp.A.numpy(), hyp.B.numpy(),
                         hyp.C.numpy()), X.numpy(), U.numpy(), Xdot.numpy(), jit.numpy(), 1e-5)
    g = torch.Generator().manual_seed(2)
    Q = 148 * 32 * 4 + 777            # one full device batch plus a ragged tail
    Xq = (4 * torch.rand(Q, n, generator=g, dtype=torch.float64) - 2).numpy()
    Uq = (2 * torch.rand(Q, m, generator=g, dtype=torch.float64) - 1).numpy()
    big = model.query(Xq, Uq)
    idx = np.array([0, 1, 18943, 18944, 18945, Q - 1])
    small = model.query(Xq[idx], Uq[idx])
    prior = float(hyp.outputscale * torch.linalg.matrix_norm(hyp.B, 2)) * 3
    for k in ('mean', 'svar', 'Mk', 'Bk'):
        # same kernels; only the split of the row-block reduction differs with the number of query tiles
        assert np.abs(big[k][idx] - small[k]).max() < 1e-12 * max(prior, np.abs(small[k]).max()), k
    # a non-PD Gram (negative outputscale) reports NOT_PD as a RuntimeError subclass: the caller's jitter-retry loop
    bad = make_hyper(n, m + 1, hyp.lengthscale.numpy(), -1.0, hyp.A.numpy(), hyp.B.numpy(), hyp.C.numpy())
    with pytest.raises(RuntimeError, match="not positive-definite"):
        model.fit(bad, X.numpy(), U.numpy(), Xdot.numpy(), jit.numpy(), 1e-5)
    model.close()


def test_ops_reject_wrong_inputs():
    from bayesian_cbf_b200 import ops
    with pytest.raises(RuntimeError, match="float64"):
        ops.gram_train(torch.zeros(4, 2, device='cuda'), torch.zeros(4, 2, device='cuda'), torch.eye(2, device='cuda'),
                       torch.ones(2, device='cuda'), 1.0)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.cross_gram(torch.zeros(4, 2, dtype=torch.float64), torch.zeros(3, 2, dtype=torch.float64),
                       torch.ones(2, dtype=torch.float64), 1.0)
    with pytest.raises(RuntimeError, match="libbcbf error -1"):     # bcbf_potrf: Npad must be a multiple of 128
        A = torch.eye(100, dtype=torch.float64, device='cuda')
        ops.potrf_(A, 100, None, 0.0)


def test_pack_unpack_lower_on_the_device():
    """bcbf_pack_lower / bcbf_unpack_lower (the one-collective factor broadcast) against the torch indexing form."""
    from bayesian_cbf_b200.sharding import pack_lower, packed_lower_elems, unpack_lower
    g = torch.Generator().manual_seed(0)
    M = torch.tril(torch.randn(640, 640, generator=g, dtype=torch.float64))
    buf = pack_lower(M.cuda())
    assert buf.numel() == packed_lower_elems(640)
    assert torch.equal(buf.cpu(), pack_lower(M))                       # CPU path: per-block-row strided copies
    out = torch.full((640, 640), 7.0, dtype=torch.float64, device='cuda')
    assert torch.equal(unpack_lower(buf, out).cpu(), M)                # strictly-upper blocks zeroed, lower restored
