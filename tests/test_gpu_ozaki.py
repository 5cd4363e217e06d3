"""GPU parity of the int8 tensor-core covariance path (csrc/ozaki.cu: tcgen05 kind::i8, error-free digit splitting) against
the CPU oracle and against the FP64 DMMA path, through the C ABI.  Integer arithmetic is exact, so the digit arrays are
checked bit-for-bit against a host reconstruction and B_k to FP64 rounding level (tolerances written at each check)."""
import numpy as np
import pytest
import torch

from oracle import mvgp_oracle as O
from tests.test_gpu_kernels import _d, _fit_on_gpu, _mk

pytestmark = pytest.mark.gpu

S, TM, TN, KSTEP = 7, 128, 64, 32
A_STEP = S * TM * KSTEP


def _decode_factor_digits(blob, rowscale, Npad):
    """Host reconstruction of L^-1 from the digit blob: [row block I][K step][slice][row group][k chunk][row][16]."""
    nb = Npad // TM
    blob = blob.cpu().numpy().astype(np.int64)
    out = np.zeros((Npad, Npad))
    w = 256.0 ** -(np.arange(S) + 1.0)
    for I in range(nb):
        for ks in range(4 * (I + 1)):
            off = (2 * I * (I + 1) + ks) * A_STEP
            t = blob[off:off + A_STEP].reshape(S, TM // 8, 2, 8, 16)          # slice, row group, k chunk, row, k
            t = t.transpose(0, 1, 3, 2, 4).reshape(S, TM, KSTEP)
            out[I * TM:(I + 1) * TM, ks * KSTEP:(ks + 1) * KSTEP] = np.tensordot(w, t.astype(np.float64), axes=1)
    return out * rowscale.cpu().numpy()[:, None]


def test_factor_digits_reconstruct_the_factor():
    from bayesian_cbf_b200 import ops
    X, U, Xdot, hyp, jit, Xq, Uq = _mk(3, 300, 3, 2, 10)
    L, Linv, G, alpha, W = _fit_on_gpu(X, U, Xdot, hyp, jit)
    digits, rowscale = ops.oz_split_factor(Linv)
    Npad = Linv.shape[0]
    rs = rowscale.cpu().numpy()
    assert np.all(np.log2(rs) == np.round(np.log2(rs)))                        # powers of two
    Lh = Linv.cpu().numpy()
    assert np.all(np.abs(Lh).max(1) / rs <= 0.498) and np.all(np.abs(Lh).max(1) / rs > 0.124)
    rec = _decode_factor_digits(digits, rowscale, Npad)
    # 56 bits below the row scale: |error| <= 2^-57 * rowscale (round to nearest of the last digit)
    assert np.all(np.abs(rec - Lh) <= 2.0 ** -57 * rs[:, None] * (1 + 1e-12))
    d = digits.cpu().numpy()
    assert d.min() >= -128 and d.max() <= 127


@pytest.mark.parametrize('N,n,m,Q', [(500, 3, 2, 200), (1000, 2, 1, 333), (260, 3, 3, 50), (150, 2, 0, 40), (700, 3, 2, 21)])
def test_posterior_var_i8(N, n, m, Q):
    from bayesian_cbf_b200 import ops
    X, U, Xdot, hyp, jit, Xq, Uq = _mk(7, N, n, m, Q, box=3.0)
    p = m + 1
    L, Linv, G, alpha, W = _fit_on_gpu(X, U, Xdot, hyp, jit)
    Ks = ops.cross_gram(_d(X), _d(Xq), _d(hyp.lengthscale), float(hyp.outputscale))
    _, Bk_dmma = ops.posterior_blocks(Linv, Ks, G, W, _d(hyp.B), _d(hyp.C.t()), float(hyp.outputscale), n, p, Q,
                                      want_mean=False)
    digits, rowscale = ops.oz_split_factor(Linv)
    Bk = ops.posterior_var_i8(digits, rowscale, Ks, G, _d(hyp.B), float(hyp.outputscale), p, Q)
    torch.cuda.synchronize()
    prior = (hyp.outputscale * torch.linalg.matrix_norm(hyp.B, 2)).item()
    # same inputs (L^-1, K*, G) through two exact-to-rounding contractions: FP64 rounding level apart
    d_paths = ((Bk - Bk_dmma).abs().max() / prior).item()
    assert d_paths < 1e-12, d_paths
    assert (Bk - Bk.transpose(1, 2)).abs().max().item() == 0.0
    Lref = O.perturbed_cholesky(hyp, X, O.homogeneous(U), [jit], direct=True)
    Mk_o, Bk_o, mean_o, svar_o = O.posterior_blocks(hyp, X, U, Xdot, Lref, Xq, Uq, direct=True)
    d_oracle = ((Bk.cpu() - Bk_o).abs().max() / prior).item()
    assert d_oracle < 1e-9, d_oracle                                            # north-star tolerance (rel. 1e-9, float64)


def test_model_handle_int8_path_matches_dmma_path():
    from bayesian_cbf_b200.model import MVGPModel, make_hyper
    X, U, Xdot, hyp, jit, Xq, Uq = _mk(11, 900, 3, 2, 5000, box=2.0)
    h = make_hyper(3, 3, hyp.lengthscale.numpy(), float(hyp.outputscale), hyp.A.numpy(), hyp.B.numpy(), hyp.C.numpy())
    outs = {}
    for path in ('dmma', 'int8'):
        model = MVGPModel(0).set_var_path(path)
        assert model.var_path == path
        model.fit(h, X.numpy(), U.numpy(), Xdot.numpy(), jit.numpy(), 1e-5)
        outs[path] = model.query(Xq.numpy(), Uq.numpy())
        if path == 'int8':
            assert model.fit_timing_ms()['oz_split'] > 0.0
        model.close()
    prior = (hyp.outputscale * torch.linalg.matrix_norm(hyp.B, 2)).item()
    assert np.abs(outs['int8']['Bk'] - outs['dmma']['Bk']).max() / prior < 1e-12
    assert np.abs(outs['int8']['svar'] - outs['dmma']['svar']).max() / prior < 1e-11
    # the mean rides on a different pass over K* (other summation order of a cancelling sum): rounding level of sum |K* W|
    assert np.abs(outs['int8']['Mk'] - outs['dmma']['Mk']).max() < 1e-8 * np.abs(outs['dmma']['Mk']).max()
    assert np.abs(outs['int8']['mean'] - outs['dmma']['mean']).max() < 1e-8 * np.abs(outs['dmma']['mean']).max()


def test_host_class_picks_the_int8_kernel_for_large_batches():
    """ControlAffineRegressorExact.custom_predict_blocks: 'auto' switches to the int8 kernel at N >= 1024, >= 512 queries;
    both kernels give the same blocks."""
    from bayesian_cbf_b200.control_affine_model import ControlAffineRegressorExact
    torch.manual_seed(0)
    X, U, Xdot, hyp, jit, Xq, Uq = _mk(21, 1100, 3, 2, 700, box=2.0)
    reg = ControlAffineRegressorExact(3, 2, device='cuda')
    reg.double_()
    reg.fit(X, U, Xdot, training_iter=0)
    out = {}
    for kern in ('dmma', 'int8', 'auto'):
        reg.covariance_kernel = kern
        Mk, Bk, mean, svar = reg.custom_predict_blocks(Xq, Uq)
        out[kern] = (Mk.cpu(), Bk.cpu(), mean.cpu(), svar.cpu())
        assert ('_oz' in reg._cache) == (kern != 'dmma')
    scale = out['dmma'][1].abs().max().item()
    assert (out['int8'][1] - out['dmma'][1]).abs().max().item() < 1e-12 * scale
    assert torch.equal(out['auto'][1], out['int8'][1]) and torch.equal(out['auto'][0], out['int8'][0])
    assert (out['int8'][0] - out['dmma'][0]).abs().max().item() < 1e-8 * out['dmma'][0].abs().max().item()
    reg.clear_cache()
    assert '_oz' not in reg._cache


def test_cluster_multicast_variant_is_bit_identical():
    """oz_var_kernel as single CTAs (the default) and as clusters of two CTAs (multicast of the L^-1 digits): the same
    integers are accumulated, so B_k agrees bit for bit; odd and even numbers of column tiles."""
    from bayesian_cbf_b200 import _lib, ops
    lib = _lib.load()
    X, U, Xdot, hyp, jit, Xq, Uq = _mk(9, 640, 3, 2, 21 * 5 + 3, box=2.5)
    L, Linv, G, alpha, W = _fit_on_gpu(X, U, Xdot, hyp, jit)
    digits, rowscale = ops.oz_split_factor(Linv)
    try:
        for Q in (21 * 5 + 3, 21 * 4, 1):
            Ks = ops.cross_gram(_d(X), _d(Xq[:Q]), _d(hyp.lengthscale), float(hyp.outputscale))
            res = []
            for ctas in (1, 2):
                assert lib.bcbf_oz_set_cluster(ctas) == 0
                res.append(ops.posterior_var_i8(digits, rowscale, Ks, G, _d(hyp.B), float(hyp.outputscale), 3, Q).cpu())
            assert torch.equal(res[0], res[1]), Q
    finally:
        lib.bcbf_oz_set_cluster(1)


def test_int8_entry_points_reject_bad_arguments():
    from bayesian_cbf_b200 import _lib
    lib = _lib.load()
    big = lib.bcbf_oz_max_npad() + 128
    assert lib.bcbf_oz_max_npad() == 18432 and lib.bcbf_oz_factor_bytes(256) == 2 * 2 * 3 * 7 * 128 * 32
    assert lib.bcbf_oz_factor_bytes(100) == 0
    x = torch.zeros(16, dtype=torch.float64, device='cuda')
    ptr = x.data_ptr()
    # factor larger than the exact-int32 limit, Npad not a multiple of 128, ldks < Q, unsupported p, null pointers
    assert lib.bcbf_oz_split_factor(ptr, big, big, ptr, ptr, None) == _lib.BCBF_ERR_INVALID
    assert lib.bcbf_posterior_var_i8(ptr, ptr, big, ptr, 64, ptr, ptr, 1.0, 3, 8, ptr, None) == _lib.BCBF_ERR_INVALID
    assert lib.bcbf_posterior_var_i8(ptr, ptr, 200, ptr, 64, ptr, ptr, 1.0, 3, 8, ptr, None) == _lib.BCBF_ERR_INVALID
    assert lib.bcbf_posterior_var_i8(ptr, ptr, 256, ptr, 4, ptr, ptr, 1.0, 3, 8, ptr, None) == _lib.BCBF_ERR_INVALID
    assert lib.bcbf_posterior_var_i8(ptr, ptr, 256, ptr, 64, ptr, ptr, 1.0, 5, 8, ptr, None) == _lib.BCBF_ERR_INVALID
    assert lib.bcbf_posterior_var_i8(None, ptr, 256, ptr, 64, ptr, ptr, 1.0, 3, 8, ptr, None) == _lib.BCBF_ERR_INVALID
    assert b'bcbf_posterior' in lib.bcbf_last_error()
    assert lib.bcbf_oz_set_cluster(3) == _lib.BCBF_ERR_INVALID


@pytest.mark.parametrize('n,m', [(2, 0), (2, 1), (3, 3)])
def test_model_handle_int8_other_control_dimensions(n, m):
    """p = 1, 2, 4 (64, 32, 16 queries per tile) through the model handle, against the oracle."""
    from bayesian_cbf_b200.model import MVGPModel, make_hyper
    p = m + 1
    X, U, Xdot, hyp, jit, Xq, Uq = _mk(31 + p, 400, n, m, 333, box=2.0)
    h = make_hyper(n, p, hyp.lengthscale.numpy(), float(hyp.outputscale), hyp.A.numpy(), hyp.B.numpy(), hyp.C.numpy())
    model = MVGPModel(0).set_var_path('int8')
    model.fit(h, X.numpy(), U.numpy().reshape(400, m), Xdot.numpy(), jit.numpy(), 1e-5)
    out = model.query(Xq.numpy(), Uq.numpy().reshape(333, m) if m else None)
    model.close()
    Lref = O.perturbed_cholesky(hyp, X, O.homogeneous(U), [jit], direct=True)
    Mk_o, Bk_o, mean_o, svar_o = O.posterior_blocks(hyp, X, U, Xdot, Lref, Xq, Uq, direct=True)
    prior = (hyp.outputscale * torch.linalg.matrix_norm(hyp.B, 2)).item()
    assert np.abs(out['Bk'] - Bk_o.numpy()).max() / prior < 1e-9
    assert np.abs(out['svar'] - svar_o.numpy()).max() / prior < 1e-8
    from tests.parity_util import mean_reference, rel
    ref = mean_reference(hyp, X, U, Xdot, jit, Xq, Uq if m else None)
    assert rel(out['Mk'], ref['Mk_exact']) < ref['tol_exact'], (rel(out['Mk'], ref['Mk_exact']), ref)
    assert rel(out['Mk'], Mk_o) < ref['tol_lapack']


@pytest.mark.parametrize('M,N,K,tri', [(256, 192, 160, 0), (384, 128, 384, 1), (128, 256, 256, 2), (1024, 1024, 1024, 0),
                                       (2048, 2048, 2048, 1), (2048, 2048, 2048, 2)])
def test_oz_gemm_matches_float64_matmul(M, N, K, tri):
    """bcbf_oz_gemm against torch float64 matmul on operands with a wide dynamic range inside rows / columns; error
    measured against |A| |B| (the quantity FP64 rounding errors scale with)."""
    from bayesian_cbf_b200 import ops
    g = torch.Generator().manual_seed(M + N + K + tri)
    A = torch.randn(M, K, generator=g, dtype=torch.float64) * torch.exp2(torch.randint(-12, 4, (M, K), generator=g).double())
    B = torch.randn(K, N, generator=g, dtype=torch.float64) * torch.exp2(torch.randint(-12, 4, (K, N), generator=g).double())
    junk = 1e300                                         # the unread triangle may hold anything
    if tri == 1:
        A = torch.tril(A) + torch.triu(torch.full_like(A, junk), 1)
    if tri == 2:
        B = torch.tril(B) + torch.triu(torch.full_like(B, junk), 1)
    C = ops.oz_gemm(A.cuda(), B.cuda(), alpha=-0.75, tri=tri).cpu()
    Ar = torch.tril(A) if tri == 1 else A
    Br = torch.tril(B) if tri == 2 else B
    ref = -0.75 * (Ar @ Br)
    assert torch.isfinite(C).all()
    # 56-bit digits below a per-row / per-column power-of-two scale: the error scales with rowmax_i * colmax_j (not with
    # |A||B| entry by entry); worst case 10 K 2^-58 of scale_i scale_j <= K 2^-50 rowmax colmax
    bound = Ar.abs().amax(1, keepdim=True) * Br.abs().amax(0, keepdim=True)
    assert ((C - ref).abs() / bound).max().item() < K * 2.0 ** -50
    if tri == 0:  # full rows and columns: also within a few hundred ulps of what FP64 rounding errors scale with
        assert ((C - ref).abs() / (Ar.abs() @ Br.abs()).clamp_min(1e-300)).max().item() < 1e-13
    # embedded in a larger buffer (leading dimensions > extents, as bcbf_trtri calls it)
    big = torch.zeros(M + 128, N + 64, dtype=torch.float64, device='cuda')
    Ab = torch.zeros(M, K + 32, dtype=torch.float64, device='cuda')
    Ab[:, :K] = A.cuda()
    ops.oz_gemm(Ab[:, :K], B.cuda(), alpha=-0.75, tri=tri, out=big[128:, 64:])
    assert torch.equal(big[128:, 64:].cpu(), C) and big[:128].abs().max().item() == 0.0


def test_trtri_large_levels_on_int8_match_the_fp64_pipe():
    """bcbf_trtri runs its levels with half-size >= 2048 through bcbf_oz_gemm: same inverse as the all-DMMA run to the
    conditioning-limited level, and the same residual |L^-1 L - I|."""
    from bayesian_cbf_b200 import _lib, ops
    lib = _lib.load()
    N = 4200                                            # Npad = 4224: one full 2048 pair, a clipped 4096 pair
    X, U, Xdot, hyp, jit, Xq, Uq = _mk(17, N, 3, 2, 8, box=3.0)
    UH = O.homogeneous(U)
    res = {}
    try:
        for on in (0, 1):
            assert lib.bcbf_set_trtri_i8(on) == 0
            Kb = ops.gram_train(_d(X), _d(UH), _d(hyp.B), _d(hyp.lengthscale), float(hyp.outputscale))
            L, dinv = ops.potrf_(Kb, N, _d(jit), 1e-5)
            res[on] = (L, ops.trtri(L, dinv))
    finally:
        lib.bcbf_set_trtri_i8(1)
    L, Linv0 = res[0]
    Linv1 = res[1][1]
    scale = Linv0.abs().max().item()
    assert (Linv1 - Linv0).abs().max().item() < 1e-9 * scale
    assert torch.equal(torch.triu(Linv1, 1), torch.zeros_like(Linv1))
    Npad = L.shape[0]
    I = torch.eye(Npad, dtype=torch.float64, device='cuda')
    r0 = (Linv0 @ L - I).abs().max().item()
    r1 = (Linv1 @ L - I).abs().max().item()
    # measured 6.1e-11 against 1.5e-11 (without the inner-dimension balancing of bcbf_oz_gemm: 6.4e-9)
    assert r1 < 10 * r0 + 1e-12, (r0, r1)


@pytest.mark.parametrize('M,N,K,lower', [(384, 128, 96, False), (1024, 512, 512, False), (640, 640, 512, True),
                                         (2048, 2048, 512, True)])
def test_oz_update_rank_k(M, N, K, lower):
    """bcbf_oz_update (the Cholesky trailing update on the int8 tensor cores) against float64: C += alpha PA PB^T on every
    tile, or only on the 128 x 64 tiles that touch the lower triangle."""
    from bayesian_cbf_b200 import ops
    g = torch.Generator().manual_seed(M + N + K)
    PA = torch.randn(M, K, generator=g, dtype=torch.float64) * torch.exp2(torch.randint(-6, 3, (M, 1), generator=g).double())
    PB = PA if lower else torch.randn(N, K, generator=g, dtype=torch.float64)
    C0 = torch.randn(M, N, generator=g, dtype=torch.float64)
    big = torch.zeros(M + 128, N + 64, dtype=torch.float64, device='cuda')
    big[128:, 64:] = C0.cuda()
    PAd = PA.cuda()
    C = ops.oz_update_(big[128:, 64:], PAd, PAd if lower else PB.cuda(), alpha=-1.0, lower=lower).cpu()
    ref = C0 - PA @ PB.T
    touched = torch.ones(M, N, dtype=torch.bool)
    if lower:
        i = torch.arange(M).unsqueeze(1) // 128
        j = torch.arange(N).unsqueeze(0) // 64
        touched = j <= 2 * i + 1
    bound = PA.abs().amax(1, keepdim=True) * PB.abs().amax(1, keepdim=True).T
    err = ((C - ref).abs() / bound)[touched].max().item()
    assert err < K * 2.0 ** -50 + 2e-16, err
    assert torch.equal(C[~touched], C0[~touched])
    assert big[:128].abs().max().item() == 0.0 and big[:, :64].abs().max().item() == 0.0


def test_potrf_trailing_updates_on_int8_match_the_fp64_pipe():
    """bcbf_potrf with its large trailing updates on bcbf_oz_update: same backward error |L L^T - (Kb + jitter)| as the
    all-DMMA factorisation, factors equal to the conditioning-limited level."""
    from bayesian_cbf_b200 import _lib, ops
    lib = _lib.load()
    N = 4200
    X, U, Xdot, hyp, jit, Xq, Uq = _mk(23, N, 3, 2, 8, box=3.0)
    UH = O.homogeneous(U)
    res = {}
    try:
        for on in (0, 1):
            assert lib.bcbf_set_potrf_i8(on) == 0
            Kb = ops.gram_train(_d(X), _d(UH), _d(hyp.B), _d(hyp.lengthscale), float(hyp.outputscale))
            Kfull = Kb.clone()
            L, dinv = ops.potrf_(Kb, N, _d(jit), 1e-5)
            res[on] = L
    finally:
        lib.bcbf_set_potrf_i8(1)
    Npad = Kfull.shape[0]
    jpad = torch.zeros(Npad, dtype=torch.float64, device='cuda')
    jpad[:N] = 1e-5 * _d(jit)
    target = Kfull + torch.diag(jpad)
    back = [((res[on] @ res[on].T - target).abs().max() / target.abs().max()).item() for on in (0, 1)]
    assert back[1] < 3 * back[0] + 1e-15 and back[1] < 1e-13, back
    assert ((res[1] - res[0]).abs().max() / res[0].abs().max()).item() < 1e-7


@pytest.mark.parametrize('M,N,K,tri', [(256, 192, 160, 0), (384, 128, 384, 1), (128, 256, 256, 2)])
def test_oz_gemm_is_bit_identical_to_the_cpu_restatement(M, N, K, tri):
    """Integer digit products are exact and every FP64 step of the kernel is a single rounding in a fixed order:
    bcbf_oz_gemm must equal oracle/ozaki_oracle.py:gemm bit for bit (scales, balancing, digits, recombination)."""
    from bayesian_cbf_b200 import ops
    from oracle import ozaki_oracle as Z
    g = torch.Generator().manual_seed(5 * M + N + K + tri)
    A = torch.randn(M, K, generator=g, dtype=torch.float64) * torch.exp2(torch.randint(-20, 6, (M, K), generator=g).double())
    B = torch.randn(K, N, generator=g, dtype=torch.float64) * torch.exp2(torch.randint(-20, 6, (K, N), generator=g).double())
    A[3] = 0.0                                           # an all-zero row and column: unit scale, zero digits
    B[:, 5] = 0.0
    C = ops.oz_gemm(A.cuda(), B.cuda(), alpha=-0.75, tri=tri).cpu().numpy()
    ref = Z.gemm(A.numpy(), B.numpy(), alpha=-0.75, tri=tri)
    assert np.array_equal(C, ref)


@pytest.mark.parametrize('lower', [False, True])
def test_oz_update_is_bit_identical_to_the_cpu_restatement(lower):
    from bayesian_cbf_b200 import ops
    from oracle import ozaki_oracle as Z
    g = torch.Generator().manual_seed(77 + lower)
    M, K = 384, 96
    PA = torch.randn(M, K, generator=g, dtype=torch.float64) * torch.exp2(torch.randint(-9, 4, (M, K), generator=g).double())
    PB = PA if lower else torch.randn(128, K, generator=g, dtype=torch.float64)
    C0 = torch.randn(M, PB.shape[0], generator=g, dtype=torch.float64)
    C = C0.clone().cuda()
    PAd = PA.cuda()
    ops.oz_update_(C, PAd, PAd if lower else PB.cuda(), alpha=-1.0, lower=lower)
    ref = Z.update(C0.numpy(), PA.numpy(), PB.numpy(), alpha=-1.0)
    touched = np.ones(ref.shape, dtype=bool)
    if lower:
        touched = (np.arange(ref.shape[1])[None, :] // 64) <= 2 * (np.arange(M)[:, None] // 128) + 1
    got = C.cpu().numpy()
    assert np.array_equal(got[touched], ref[touched])
    assert np.array_equal(got[~touched], C0.numpy()[~touched])


@pytest.mark.parametrize('N,n,m,Q', [(300, 3, 2, 50), (200, 2, 1, 70)])
def test_posterior_var_i8_is_bit_identical_to_the_cpu_restatement(N, n, m, Q):
    """oz_var_kernel + its split / finalize kernels against oracle/ozaki_oracle.py:posterior_bk on the same L^-1, K*, G
    (downloaded from the device): scales, digits, integer accumulation, recombination, warp-butterfly Gram reduction
    and the fixed-order sum over row blocks are all reproduced, so B_k agrees bit for bit."""
    from bayesian_cbf_b200 import ops
    from oracle import ozaki_oracle as Z
    X, U, Xdot, hyp, jit, Xq, Uq = _mk(41, N, n, m, Q, box=2.0)
    p = m + 1
    L, Linv, G, alpha, W = _fit_on_gpu(X, U, Xdot, hyp, jit)
    Ks = ops.cross_gram(_d(X), _d(Xq), _d(hyp.lengthscale), float(hyp.outputscale))
    digits, rowscale = ops.oz_split_factor(Linv)
    Bk = ops.posterior_var_i8(digits, rowscale, Ks, G, _d(hyp.B), float(hyp.outputscale), p, Q).cpu().numpy()
    ref = Z.posterior_bk(Linv.cpu().numpy(), Ks.cpu().numpy()[:, :Q], G.cpu().numpy(), hyp.B.numpy(), float(hyp.outputscale))
    assert np.array_equal(Bk, ref)


def test_oz_gemm_tn_lower_is_bit_identical_and_accurate():
    """Kb^-1 = L^-T L^-1 through bcbf_oz_gemm_tn (column-split operands, K steps below max(i, j) skipped): bit-identical to
    the CPU restatement, symmetric, and at FP64 level against the float64 product."""
    from bayesian_cbf_b200 import ops
    from oracle import ozaki_oracle as Z
    X, U, Xdot, hyp, jit, Xq, Uq = _mk(51, 300, 3, 2, 4, box=2.0)
    L, Linv, G, alpha, W = _fit_on_gpu(X, U, Xdot, hyp, jit)
    P = ops.oz_gemm_tn(Linv, Linv, lower=True).cpu().numpy()
    Lh = Linv.cpu().numpy()
    ref = Z.gemm(Lh.T.copy(), Lh, balance=False)
    assert np.array_equal(P, ref)
    assert np.array_equal(P, P.T)
    exact = (Lh.T.astype(np.longdouble) @ Lh.astype(np.longdouble)).astype(np.float64)
    bound = np.abs(Lh).max(0)[:, None] * np.abs(Lh).max(0)[None, :]
    assert (np.abs(P - exact) / bound).max() < Lh.shape[0] * 2.0 ** -50


@pytest.mark.parametrize('N,n,m,Q', [(300, 3, 2, 50), (640, 2, 1, 200)])
def test_six_digit_mode_is_bit_identical_to_its_restatement_and_inside_the_parity_tolerance(N, n, m, Q):
    """The opt-in 6-digit mode of oz_var_kernel (21 digit products instead of 28): bit for bit the arithmetic of
    oracle/ozaki_oracle.py:posterior_bk(digits=6); against the CPU oracle of the path its B_k stays inside the 1e-9
    parity tolerance (measured ~1e-11 of the prior scale) but is, as designed, less exact than the 7-digit default."""
    from bayesian_cbf_b200 import ops
    from oracle import ozaki_oracle as Z
    X, U, Xdot, hyp, jit, Xq, Uq = _mk(43, N, n, m, Q, box=2.0)
    p = m + 1
    L, Linv, G, alpha, W = _fit_on_gpu(X, U, Xdot, hyp, jit)
    Ks = ops.cross_gram(_d(X), _d(Xq), _d(hyp.lengthscale), float(hyp.outputscale))
    out = {}
    for nd in (6, 7):
        digits, rowscale = ops.oz_split_factor(Linv, ndigits=nd)
        _, Bk = ops.posterior_blocks_i8(digits, rowscale, Ks, G, W, _d(hyp.B), _d(hyp.C.t()), float(hyp.outputscale), n, p, Q,
                                        want_mean=False, ndigits=nd)
        out[nd] = Bk.cpu().numpy()
    ref6 = Z.posterior_bk(Linv.cpu().numpy(), Ks.cpu().numpy()[:, :Q], G.cpu().numpy(), hyp.B.numpy(),
                          float(hyp.outputscale), digits=6)
    assert np.array_equal(out[6], ref6)
    Lref = O.perturbed_cholesky(hyp, X, O.homogeneous(U), [jit], direct=True)
    _, Bk_o = O.posterior_blocks(hyp, X, U, Xdot, Lref, Xq, direct=True)
    prior = (hyp.outputscale * torch.linalg.matrix_norm(hyp.B, 2)).item()
    e6 = np.abs(out[6] - Bk_o.numpy()).max() / prior
    e7 = np.abs(out[7] - Bk_o.numpy()).max() / prior
    print('B_k error / prior scale: 6 digits %.2e, 7 digits %.2e' % (e6, e7))
    assert e6 < 1e-9 and e7 < 1e-9
    assert np.abs(out[6] - out[7]).max() / prior < 1e-9 and not np.array_equal(out[6], out[7])


def test_model_handle_six_digit_option():
    from bayesian_cbf_b200.model import MVGPModel, make_hyper
    X, U, Xdot, hyp, jit, Xq, Uq = _mk(12, 900, 3, 2, 3000, box=2.0)
    h = make_hyper(3, 3, hyp.lengthscale.numpy(), float(hyp.outputscale), hyp.A.numpy(), hyp.B.numpy(), hyp.C.numpy())
    model = MVGPModel(0).set_var_path('int8')
    model.fit(h, X.numpy(), U.numpy(), Xdot.numpy(), jit.numpy(), 1e-5)
    o7 = model.query(Xq.numpy(), Uq.numpy())
    o6 = model.set_oz_digits(6).query(Xq.numpy(), Uq.numpy())        # re-splits L^-1 with six digits
    o7b = model.set_oz_digits(7).query(Xq.numpy(), Uq.numpy())
    model.close()
    prior = (hyp.outputscale * torch.linalg.matrix_norm(hyp.B, 2)).item()
    assert np.array_equal(o7['Bk'], o7b['Bk']) and np.array_equal(o7['mean'], o6['mean'])
    d = np.abs(o6['Bk'] - o7['Bk']).max() / prior
    assert 0 < d < 1e-9, d
    with pytest.raises(Exception):
        model2 = MVGPModel(0)
        model2.set_oz_digits(5)


def test_scratch_using_entry_points_are_safe_across_streams():
    """Two torch streams call the int8 posterior (which shares per-device scratch: frakB digits, partial sums) back to back
    without any host synchronisation between the calls; every result must equal the one computed alone (bit for bit: the
    kernels are deterministic).  Round 1 documented this as a trap; ScratchScope (csrc/common.cuh) orders the calls."""
    from bayesian_cbf_b200 import ops
    X, U, Xdot, hyp, jit, Xq, Uq = _mk(51, 900, 3, 2, 3000, box=2.0)
    L, Linv, G, alpha, W = _fit_on_gpu(X, U, Xdot, hyp, jit)
    digits, rowscale = ops.oz_split_factor(Linv)
    Bd, Ct = _d(hyp.B), _d(hyp.C.t())
    Xd = _d(X)
    ls = _d(hyp.lengthscale)
    chunks = [(0, 1100), (1100, 2300), (2300, 3000), (500, 2900)]
    Ks = [ops.cross_gram(Xd, _d(Xq[a:b]), ls, float(hyp.outputscale)) for a, b in chunks]
    alone = []
    for K, (a, b) in zip(Ks, chunks):
        alone.append(ops.posterior_blocks_i8(digits, rowscale, K, G, W, Bd, Ct, float(hyp.outputscale), 3, 3, b - a))
        torch.cuda.synchronize()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    torch.cuda.synchronize()
    got = [None] * len(chunks)
    for rep in range(3):
        for i, (K, (a, b)) in enumerate(zip(Ks, chunks)):
            with torch.cuda.stream(streams[i % 2]):
                got[i] = ops.posterior_blocks_i8(digits, rowscale, K, G, W, Bd, Ct, float(hyp.outputscale), 3, 3, b - a)
        torch.cuda.synchronize()
        for i in range(len(chunks)):
            assert torch.equal(got[i][1], alone[i][1]) and torch.equal(got[i][0], alone[i][0]), (rep, i)
