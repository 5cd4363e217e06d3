"""CPU checks of the digit-splitting arithmetic (oracle/ozaki_oracle.py) that csrc/ozaki.cu runs on the int8 tensor cores:
exactness of the digit representation, the error bound of the product against extended precision, and what the
inner-dimension balancing buys on factor-like operands.  The CUDA kernels are compared with these functions bit for bit in
tests/test_gpu_ozaki.py."""
import numpy as np
import pytest

from oracle import ozaki_oracle as Z


def test_digits_represent_56_bits_exactly():
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-0.498, 0.498, 20000), rng.uniform(-1, 1, 2000) * 2.0 ** rng.integers(-60, -1, 2000),
                        [0.0, 0.498, -0.498, 2.0 ** -57, -2.0 ** -57, 2.0 ** -58, 127.0 / 256, -0.5 + 2.0 ** -9]])
    d = Z.digits_of(x)
    assert d.min() >= -128 and d.max() <= 127
    back = Z.undigits(d)
    assert np.abs(back - x).max() <= 2.0 ** -57                   # round to nearest of the 56th bit
    exact = np.abs(x) * 2.0 ** 56 == np.rint(np.abs(x) * 2.0 ** 56)
    assert np.array_equal(back[exact], x[exact])                   # values with <= 56 bits below the scale are exact


def test_scale_of_keeps_the_row_inside_the_digit_range():
    rng = np.random.default_rng(1)
    mx = np.concatenate([rng.uniform(0, 1, 1000) * 2.0 ** rng.integers(-40, 40, 1000), [1.0, 0.5, 0.996, 0.9961, 2.0 ** -30]])
    sc = Z.scale_of(mx)
    assert np.all(np.log2(sc) == np.rint(np.log2(sc)))
    r = mx / sc
    assert r.max() <= 0.498 and r.min() > 0.124
    assert Z.scale_of(0.0) == 1.0


def _graded(M, K, N, rng):
    """Operands shaped like a Cholesky factor's rows and its inverse's columns: magnitudes fall / rise steeply along k."""
    g = 2.0 ** (-np.linspace(0, 30, K))
    A = rng.standard_normal((M, K)) * g[None, :]
    B = rng.standard_normal((K, N)) / g[:, None]
    return A, B


@pytest.mark.parametrize('tri', [0, 1, 2])
def test_gemm_error_bound_against_extended_precision(tri):
    rng = np.random.default_rng(2 + tri)
    M = K = N = 96
    A = rng.standard_normal((M, K)) * 2.0 ** rng.integers(-12, 4, (M, K))
    B = rng.standard_normal((K, N)) * 2.0 ** rng.integers(-12, 4, (K, N))
    C = Z.gemm(A, B, alpha=-0.75, tri=tri)
    Am, Bm = Z._masked(A, B, tri)
    ref = (-0.75 * (Am.astype(np.longdouble) @ Bm.astype(np.longdouble))).astype(np.float64)
    D = Z.inner_scale(A, B, tri)
    bound = np.abs(Am * D).max(1)[:, None] * np.abs(Bm / D[:, None]).max(0)[None, :]
    assert (np.abs(C - ref) / bound).max() < K * 2.0 ** -50


def test_inner_balancing_recovers_the_bits_of_graded_operands():
    rng = np.random.default_rng(5)
    A, B = _graded(64, 128, 64, rng)
    ref = (A.astype(np.longdouble) @ B.astype(np.longdouble)).astype(np.float64)
    scale = np.abs(A) @ np.abs(B)
    err_plain = (np.abs(Z.gemm(A, B, balance=False) - ref) / scale).max()
    err_bal = (np.abs(Z.gemm(A, B, balance=True) - ref) / scale).max()
    assert err_bal < 1e-14 and err_plain > 100 * err_bal            # measured: ~1e-16 against ~1e-9


def test_update_matches_float64_rank_k_update():
    rng = np.random.default_rng(7)
    P = rng.standard_normal((64, 96)) * 2.0 ** rng.integers(-6, 3, (64, 1))
    C0 = rng.standard_normal((64, 64))
    C = Z.update(C0, P, P, alpha=-1.0)
    ref = (C0.astype(np.longdouble) - P.astype(np.longdouble) @ P.astype(np.longdouble).T).astype(np.float64)
    bound = np.abs(P).max(1)[:, None] * np.abs(P).max(1)[None, :]
    assert (np.abs(C - ref) / bound).max() < 96 * 2.0 ** -50 + 2.0 ** -52


def test_accumulator_bound_is_enforced():
    with pytest.raises(AssertionError):
        Z.gemm(np.ones((8, 18752)), np.ones((18752, 8)))
