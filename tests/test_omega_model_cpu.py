"""NumPy models of the host-side schedules of the Omega build (csrc/omega_refine.cu, csrc/ozaki_int8.cu), run on the CPU:

  * the triangular-aware strips of the Cholesky inverse (trtri_lower_gemm, vtv_upper_strips) -- which sub-blocks each
    DGEMM reads and writes -- must reproduce inv(L) and upper(V'V) while skipping only exact zeros;
  * the upper-strip schedule of ozaki_product_upper must cover the upper triangle;
  * the digit-plane product (7-bit planes, per-column powers of two, one GEMM per diagonal s + t = d, pairs with
    s + t >= S dropped) must reproduce an exact product to S 2^(-7 S) of rowmax * colmax per term, and the Newton-Schulz
    step built from it must square the residual of a start with cond ~ 1e8.

The loops below restate the C++ loops index for index (the C++ passes the same offsets to cuBLAS); they are models of the
schedule, not of the kernels -- those are covered by the `-m gpu` parity tests (test_tps_matrix_vs_oracle,
test_omega_column_blocks_match_single_build, the C3 / C4 jobs)."""
from fractions import Fraction

import numpy as np
import pytest


# ---- csrc/omega_refine.cu: trtri_lower_gemm / vtv_upper_strips ---------------------------------------------------
def trtri_lower_strips(L, strip, leaf, log):
    """In-place V = L^-1 of a lower-triangular L (strict upper part zero), recursion and strips as in the C++."""
    n = L.shape[0]
    if n <= leaf:
        L[:, :] = np.linalg.solve(L, np.eye(n))                     # cublasDtrsm on an identity
        return
    h = ((n // 2) + 63) // 64 * 64                                  # as in the C++ (its leaves are 512 rows: h < n)
    assert 0 < h < n
    r = n - h
    L11, L21, L22 = L[:h, :h], L[h:, :h], L[h:, h:]
    trtri_lower_strips(L11, strip, leaf, log)
    trtri_lower_strips(L22, strip, leaf, log)
    T = np.empty((r, h))
    for j0 in range(0, h, strip):                                   # T = L21 V11, column strips of V11
        nb = min(strip, h - j0)
        T[:, j0:j0 + nb] = L21[:, j0:h] @ L11[j0:h, j0:j0 + nb]
        log.append(2.0 * r * nb * (h - j0))
    for i0 in range(0, r, strip):                                   # V21 = -V22 T, row strips of V22
        nb = min(strip, r - i0)
        L21[i0:i0 + nb, :] = -(L22[i0:i0 + nb, :i0 + nb] @ T[:i0 + nb, :])
        log.append(2.0 * nb * h * (i0 + nb))


def vtv_upper_strips(V, strip, log):
    n = V.shape[0]
    C = np.full((n, n), np.nan)                                     # the part below the diagonal blocks is never written
    for j0 in range(0, n, strip):
        nb = min(strip, n - j0)
        C[:j0 + nb, j0:j0 + nb] = V[j0:, :j0 + nb].T @ V[j0:, j0:j0 + nb]
        log.append(2.0 * (j0 + nb) * nb * (n - j0))
    return C


@pytest.mark.parametrize("n,strip,leaf", [(300, 64, 64), (517, 128, 64), (1000, 256, 128), (96, 1024, 512)])
def test_cholesky_inverse_strips(n, strip, leaf):
    rng = np.random.default_rng(n)
    G = rng.normal(size=(n, n + 20))
    B = G @ G.T + n * np.eye(n)
    L = np.linalg.cholesky(B)
    V = L.copy()
    flops = []
    trtri_lower_strips(V, strip, leaf, flops)
    assert np.max(np.abs(V @ L - np.eye(n))) <= 1e-12
    assert np.all(np.triu(V, 1) == 0.0)                             # the strips never write above the diagonal
    log2 = []
    C = vtv_upper_strips(V, strip, log2)
    iu = np.triu_indices(n)
    want = np.linalg.inv(B)
    assert np.max(np.abs(C[iu] - want[iu])) / np.max(np.abs(want)) <= 1e-12
    # skipping the zero part of the triangular operand: fewer flops than the dense products, never fewer than the ideal
    dense = 2.0 * n ** 3
    assert n ** 3 / 3.0 <= sum(log2) <= (dense / 2 if n > strip else dense)
    if n > 2 * strip:
        assert sum(log2) < 0.75 * dense / 2 + 2.0 * n * n * strip


# ---- csrc/ozaki_int8.cu: ozaki_product_upper --------------------------------------------------------------------
@pytest.mark.parametrize("m,n,strip", [(10003, 10003, 2048), (10000, 10000, 2048), (2503, 2503, 2048), (100, 100, 16),
                                         (33, 33, 32), (5, 5, 2048)])
def test_upper_strip_schedule_covers_upper_triangle(m, n, strip):
    covered_to = np.zeros(n, dtype=np.int64)                        # rows [0, covered_to[j]) of column j are computed
    products = 0
    for j0 in range(0, n, strip):
        nb = min(strip, n - j0)
        mr = min(m, j0 + nb)
        assert (j0 % 16) == 0 and (mr + 15) // 16 * 16 <= (m + 15) // 16 * 16
        covered_to[j0:j0 + nb] = mr
        products += mr * nb
    assert np.all(covered_to >= np.minimum(np.arange(n) + 1, m))    # the diagonal and everything above it
    assert products <= m * n
    if n >= 4 * strip:
        assert products <= 0.65 * m * n                             # 5 strips: 0.6 of the full product


# ---- csrc/ozaki_int8.cu: digit planes ----------------------------------------------------------------------------
def expo(mx):
    e = np.zeros(mx.shape, dtype=np.int64)
    nz = mx > 0
    e[nz] = np.frexp(mx[nz])[1]                                     # mx = f 2^e, f in [0.5, 1)
    return e


def planes(M, e_cols, S):
    y = np.ldexp(M, (6 - e_cols)[None, :].astype(np.int32))         # |y| < 64, exact
    out = []
    for _ in range(S):
        d = np.rint(y)
        assert np.max(np.abs(d)) <= 64
        out.append(d.astype(np.int64))
        y = (y - d) * 128.0                                         # exact
    return out


def plane_product(At, ea, B, eb, S):
    """At' B from S planes per operand: left planes forward, right planes reversed, one product per diagonal over the
    concatenated contraction axis (ozaki_product), accumulated from the smallest weight up."""
    k = At.shape[0]
    left = np.concatenate(planes(At, ea, S), axis=0)                # (S k) x m, plane s at rows [s k, (s+1) k)
    right = np.concatenate(planes(B, eb, S)[::-1], axis=0)          # plane t at rows [(S-1-t) k, (S-t) k)
    acc = np.zeros((At.shape[1], B.shape[1]))
    for d in range(S - 1, -1, -1):
        a = left[: (d + 1) * k]                                     # planes 0 .. d
        b = right[(S - 1 - d) * k:]                                 # planes d .. 0
        C = a.T @ b                                                 # int64: exact; the device holds it in int32
        assert np.max(np.abs(C)) < 2 ** 31
        acc += np.ldexp(C.astype(np.float64), -7 * d)
    return np.ldexp(acc, (ea[:, None] + eb[None, :] - 12).astype(np.int32))


@pytest.mark.parametrize("S", [4, 7, 12])
def test_digit_plane_product_error_bound(S):
    rng = np.random.default_rng(S)
    k, m, n = 24, 7, 5
    A = rng.normal(size=(k, m)) * np.exp(rng.uniform(-8, 8, size=(1, m)))     # wide per-column dynamic range
    B = rng.normal(size=(k, n)) * np.exp(rng.uniform(-8, 8, size=(1, n)))
    A[3, 2] = 0.0
    ea, eb = expo(np.abs(A).max(axis=0)), expo(np.abs(B).max(axis=0))
    got = plane_product(A, ea, B, eb, S)
    for i in range(m):
        for j in range(n):
            exact = sum(Fraction(A[t, i]) * Fraction(B[t, j]) for t in range(k))
            err = abs(Fraction(got[i, j]) - exact)
            # per term: the dropped pairs s + t >= S weigh < (S + 1) 2^(-7 S) of 2^(ea + eb); plus FP64 accumulation
            bound = k * (S + 2) * Fraction(2) ** int(ea[i] + eb[j] - 7 * S) + Fraction(2) ** int(ea[i] + eb[j] - 45)
            assert err <= bound, (i, j, float(err), float(bound))


def test_newton_schulz_step_on_digit_planes_squares_the_residual():
    """The refinement of thinPlateSplineMatrix's bordered system (src/RcppSpatPCA.cpp:58-69) at a size the CPU model
    handles: a start with |I - A X0| ~ 1e-7, one step with a 12-plane residual and a 5-plane correction."""
    rng = np.random.default_rng(5)
    p = 120
    x = rng.uniform(size=(p, 2))
    r = np.sqrt(((x[:, None, :] - x[None, :, :]) ** 2).sum(-1))
    with np.errstate(divide="ignore", invalid="ignore"):
        E = np.where(r > 0, r * r * np.log(r) / (8 * np.pi), 0.0)
    T = np.hstack([np.ones((p, 1)), x])
    q = p + 3
    A = np.zeros((q, q))
    A[:p, :p] = E + 1e-8 * np.eye(p)
    A[:p, p:] = T
    A[p:, :p] = T.T
    A[p:, p:] = 1e-8 * np.eye(3)
    Al = A.astype(np.longdouble)
    X0 = np.linalg.inv(A)
    X0 = X0 * (1.0 + 1e-9 * rng.normal(size=X0.shape))              # a start a few digits short of FP64
    X0 = np.triu(X0) + np.triu(X0, 1).T
    I = np.eye(q, dtype=np.longdouble)
    r0 = float(np.max(np.abs(I - Al @ X0.astype(np.longdouble))))
    ea = expo(np.abs(A).max(axis=0))
    R = np.eye(q) - plane_product(A, ea, X0, expo(np.abs(X0).max(axis=0)), 12)          # rows of A = its columns
    Rex = (I - Al @ X0.astype(np.longdouble)).astype(np.float64)
    assert np.max(np.abs(R - Rex)) <= 1e-4 * r0                                        # the residual itself is accurate
    X1 = X0 + plane_product(X0, expo(np.abs(X0).max(axis=0)), R, expo(np.abs(R).max(axis=0)), 5)
    r1 = float(np.max(np.abs(I - Al @ X1.astype(np.longdouble))))
    assert r0 > 1e-9 and r1 <= max(50 * r0 * r0, 1e-3 * r0), (r0, r1)
