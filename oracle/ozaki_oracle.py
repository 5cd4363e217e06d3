"""CPU restatement (numpy, exact integer arithmetic) of the digit-splitting GEMM that csrc/ozaki.cu runs on the int8 tensor
cores.  TEST INFRASTRUCTURE ONLY: imported by tests/ (never by the product package).

This is not an algorithm of the reference: it is the arithmetic by which this repo evaluates the reference's FP64 products
(V = L^-1 frakB of control_affine_model.py:1051-1088, the Cholesky trailing update of :907, the triangular inverse) on
hardware without a fast FP64 pipe.  Every step is exact or a single correctly rounded FP64 operation in a fixed order, so
the CUDA kernels must reproduce these functions BIT FOR BIT (tests/test_gpu_ozaki.py), and the functions themselves are
checked against extended-precision products in tests/test_ozaki_oracle.py:

  scale_of(mx)        power of two 2^e with |x| / 2^e <= 0.498 for |x| <= mx
  digits_of(x)        seven signed base-256 digits of round(x 2^56), x = sum_j d_j 256^-(j+1)  (|x| <= 0.498)
  gemm(A, B, ...)     C = alpha (A D)(D^-1 B): inner-dimension balancing D, per-row / per-column scales, the 28 exact
                      integer digit products with digit sum <= 6 accumulated per diagonal, FP64 recombination from the
                      smallest diagonal up, scales applied last
"""
import numpy as np

S = 7


def scale_of(mx):
    """2^e with max |x| / 2^e in (0.124, 0.498]; 1.0 for an all-zero row (csrc/ozaki.cu: scale_of)."""
    mx = np.asarray(mx, dtype=np.float64)
    f, ex = np.frexp(mx)                       # mx = f 2^ex, f in [0.5, 1)
    e = ex + 1 + (f * 0.5 >= 0.498)
    return np.where(mx > 0.0, np.ldexp(1.0, e), 1.0)


def digits_of(x, S=S):
    """x (|x| <= 0.498) -> int64 array (..., S) of digits in [-128, 127] of round(x 2^(8 S)) (csrc/ozaki.cu: digits_of;
    S = 7 everywhere except the opt-in 6-digit mode of the covariance kernel)."""
    I = np.rint(np.asarray(x, dtype=np.float64) * 2.0 ** (8 * S)).astype(np.int64)   # exact: power-of-two scaling, one rint
    d = np.zeros(I.shape + (S,), dtype=np.int64)
    for j in range(S - 1, 0, -1):
        b = ((I & 0xFF) ^ 0x80) - 0x80         # low byte read as signed
        d[..., j] = b
        I = (I - b) >> 8
    d[..., 0] = np.clip(I, -128, 127)
    return d


def undigits(d):
    """sum_j d_j 256^-(j+1), exact in float64 as long as the digits came from digits_of (56 bits)."""
    return sum(d[..., j].astype(np.float64) * 256.0 ** -(j + 1) for j in range(S - 1, -1, -1))


def inner_scale(A, B, tri=0):
    """D_k = 2^round-ish(0.5 log2(rowmax_k(B) / colmax_k(A))) exactly as kscale_kernel computes it from frexp exponents."""
    A, B = _masked(A, B, tri)
    ca, rb = np.abs(A).max(axis=0), np.abs(B).max(axis=1)
    _, ea = np.frexp(ca)
    _, eb = np.frexp(rb)
    e = eb.astype(np.int64) - ea.astype(np.int64)
    e = np.where(e >= 0, e + 1, e)
    e = np.where(e >= 0, e // 2, -((-e) // 2))                      # C integer division truncates toward zero
    return np.where((ca > 0) & (rb > 0), np.ldexp(1.0, e.astype(np.int64)), 1.0)


def _masked(A, B, tri):
    A, B = np.asarray(A, dtype=np.float64), np.asarray(B, dtype=np.float64)
    if tri == 1:
        A = np.tril(A)
    if tri == 2:
        B = np.tril(B)
    return A, B


def gemm(A, B, alpha=1.0, tri=0, balance=True):
    """C = alpha A B as bcbf_oz_gemm computes it (tri: 0, 1 = A lower triangular, 2 = B lower triangular)."""
    A, B = _masked(A, B, tri)
    K = A.shape[1]
    assert 7 * K * 2 ** 14 < 2 ** 31, "int32 accumulators would not be exact"
    D = inner_scale(A, B) if balance else np.ones(K)
    As, Bs = A * D[None, :], B / D[:, None]                         # exact (powers of two)
    rs, cs = scale_of(np.abs(As).max(axis=1)), scale_of(np.abs(Bs).max(axis=0))
    da, db = digits_of(As / rs[:, None]), digits_of(Bs / cs[None, :])
    v = np.zeros((A.shape[0], B.shape[1]))
    for d in range(S - 1, -1, -1):                                  # smallest diagonal first, as the kernel epilogue
        acc = np.zeros(v.shape, dtype=np.int64)
        for a in range(d + 1):
            acc += da[..., a] @ db[..., d - a]
        assert np.abs(acc).max() < 2 ** 31
        v = acc.astype(np.float64) * 2.0 ** (-8 * (d + 2)) + v      # product exact -> one rounding, like the kernel's fma
    return v * ((alpha * rs)[:, None] * cs[None, :])


def update(C, PA, PB, alpha=-1.0):
    """C + alpha PA PB^T as bcbf_oz_update computes it (no inner balancing: the operands share their K profile)."""
    PA, PB = np.asarray(PA, dtype=np.float64), np.asarray(PB, dtype=np.float64)
    K = PA.shape[1]
    assert 7 * K * 2 ** 14 < 2 ** 31
    rsa, rsb = scale_of(np.abs(PA).max(axis=1)), scale_of(np.abs(PB).max(axis=1))
    da, db = digits_of(PA / rsa[:, None]), digits_of(PB / rsb[:, None])
    v = np.zeros((PA.shape[0], PB.shape[0]))
    for d in range(S - 1, -1, -1):
        acc = np.zeros(v.shape, dtype=np.int64)
        for a in range(d + 1):
            acc += da[..., a] @ db[..., d - a].T
        v = acc.astype(np.float64) * 2.0 ** (-8 * (d + 2)) + v
    # the kernel: fma(v, (alpha rs_i) rs_j, c_old).  For |alpha| a power of two the product is exact and the fma is the
    # plain sum with one rounding, which float64 numpy reproduces bit for bit.
    assert np.frexp(abs(alpha))[0] == 0.5, "bit-exact emulation of the fused multiply-add needs |alpha| = 2^k"
    scale = (alpha * rsa)[:, None] * rsb[None, :]
    return v * scale + np.asarray(C, dtype=np.float64)


def posterior_bk(Linv, Kstar, G, Bmat, kss, digits=S):
    """B_k (Q,p,p) = kss B - V^T V, V = L^-1 frakB, exactly as oz_var_kernel + finalize_kernel compute it:
    Linv (Npad,Npad) lower triangular, Kstar (Npad,Q), G (Npad,p).  Row scales from the lower triangle of L^-1, column
    scales from max_i |K*[i,q]| |G[i,t]|, digits of (K* G) / scale, recombination as in gemm(); per 128-row block the p(p+1)/2
    Gram products of every query are summed over 32 rows with the warp butterfly (xor 16, 8, 4, 2, 1), the four warps as
    (w0 + w1) + (w2 + w3), and the row blocks in increasing order."""
    Linv = np.tril(np.asarray(Linv, dtype=np.float64))
    Kstar, G = np.asarray(Kstar, dtype=np.float64), np.asarray(G, dtype=np.float64)
    Npad, Q = Kstar.shape
    p = G.shape[1]
    assert Npad % 128 == 0 and 7 * Npad * 2 ** 14 < 2 ** 31
    rs = scale_of(np.abs(Linv).max(axis=1))
    frak = (Kstar[:, :, None] * G[:, None, :]).reshape(Npad, Q * p)            # column = p q + t, one rounding per entry
    cmax = (np.abs(Kstar)[:, :, None] * np.abs(G)[:, None, :]).reshape(Npad, Q * p).max(axis=0)
    cs = scale_of(cmax)
    da, db = digits_of(Linv / rs[:, None], digits), digits_of(frak / cs[None, :], digits)
    v = np.zeros((Npad, Q * p))
    for d in range(digits - 1, -1, -1):                 # digits = 6: the 21 products with digit sum <= 5
        acc = np.zeros(v.shape, dtype=np.int64)
        for a in range(d + 1):
            acc += da[..., a] @ db[..., d - a]
        v = acc.astype(np.float64) * 2.0 ** (-8 * (d + 2)) + v
    V = (v * (rs[:, None] * cs[None, :])).reshape(Npad, Q, p)
    pairs = [(x, y) for x in range(p) for y in range(x, p)]
    prod = np.stack([V[:, :, x] * V[:, :, y] for x, y in pairs], axis=-1)      # (Npad, Q, npair)
    s = prod.reshape(Npad // 128, 4, 32, Q, len(pairs))                        # row block, warp, lane
    lanes = np.arange(32)
    for o in (16, 8, 4, 2, 1):
        s = s + s[:, :, lanes ^ o]
    w = s[:, :, 0]                                                             # lane 0 of each warp
    blk = (w[:, 0] + w[:, 1]) + (w[:, 2] + w[:, 3])                            # (row block, Q, npair)
    tot = np.zeros(blk.shape[1:])
    for sp in range(blk.shape[0]):
        tot = tot + blk[sp]
    Bk = np.zeros((Q, p, p))
    for e, (x, y) in enumerate(pairs):
        Bk[:, x, y] = Bk[:, y, x] = kss * np.asarray(Bmat, dtype=np.float64)[x, y] - tot[:, e]
    return Bk
