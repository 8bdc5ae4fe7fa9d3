"""CPU checks of the factored-solve geometry (no GPU): the block partition and the stream phases that the ADMM
kernel executes per iteration are host-side decisions (csrc/spd_inverse.cu: spd_blocking*, csrc/admm_kernels.cu:
make_geom).  Here a NumPy model builds the block U'DU factor F exactly as spd_factor_blocked does (right-looking,
explicit inverses of the diagonal Schur complements, -D^-1 R panels, mirrored) and applies the library's phase
plan to it; the result must solve the system.  Together with the GPU parity tests this pins the algorithm: the CUDA
kernels only have to compute the same rectangles' products."""
import numpy as np
import pytest

from spatpca_b200 import engine as eng

FUSED, FWD, BWD = 0, 1, 2


def build_factor(A, off):
    """F = block U'DU factor of SPD A over the partition `off` (full symmetric storage)."""
    p = A.shape[0]
    F = np.triu(A).copy()
    m = len(off) - 1
    for j in range(m):
        o, e = off[j], off[j + 1]
        D = np.triu(F[o:e, o:e])
        D = D + np.triu(D, 1).T
        Dinv = np.linalg.inv(D)
        F[o:e, o:e] = Dinv
        if e < p:
            R = F[o:e, e:].copy()
            Wm = -Dinv @ R
            F[e:, e:] = np.triu(F[e:, e:] + R.T @ Wm)
            F[o:e, e:] = Wm
    return np.triu(F) + np.triu(F, 1).T


def apply_plan(F, plan, b):
    """Run the phases: z starts as b, w ends as A^-1 b.  Also returns how often each row was declared final."""
    p = F.shape[0]
    z, w = b.copy(), np.zeros_like(b)
    final_count = np.zeros(p, dtype=int)
    for ph in plan:
        r0, r1 = ph["row0"], ph["row0"] + ph["nrows"]
        c0, c1 = ph["col0"], ph["col0"] + ph["ncols"]
        if ph["kind"] == FUSED:
            w[r0:r1] = F[r0:r1, c0:c1] @ z[c0:c1]
        elif ph["kind"] == FWD:
            prod = F[r0:r1, c0:c1] @ z[c0:c1]
            nd = c1 - c0                       # FWD: row0 == col0, the first ncols rows are the diagonal block
            assert r0 == c0
            w[r0:r0 + nd] = prod[:nd]
            z[r0 + nd:r1] += prod[nd:]
        else:
            w[r0:r1] += F[r0:r1, c0:c1] @ w[c0:c1]
        final_count[ph["fin0"]:ph["fin1"]] += 1
    return w, final_count


@pytest.mark.parametrize("p,K,blocks", [(700, 3, 1), (1300, 5, 3), (1537, 5, 4), (2048, 2, 4), (2300, 12, 3),
                                        (4100, 5, 8), (5000, 5, 0), (6200, 20, 12)])
def test_phase_plan_solves_the_system(p, K, blocks):
    rng = np.random.default_rng(p + K)
    off = eng.block_offsets(p, blocks)
    plan = eng.admm_phase_plan(p, K, blocks)
    m = len(off) - 1
    assert off[0] == 0 and off[-1] == p and all(a < b for a, b in zip(off, off[1:]))
    assert all(o % 512 == 0 for o in off[:-1])                 # block starts are row-tile aligned
    assert len(plan) == 2 * m - 1
    # the rectangles tile the matrix exactly once: F is streamed once per iteration, 8 p^2 bytes
    cover = np.zeros((p, p), dtype=np.int8)
    for ph in plan:
        cover[ph["row0"]:ph["row0"] + ph["nrows"], ph["col0"]:ph["col0"] + ph["ncols"]] += 1
    assert cover.min() == 1 and cover.max() == 1
    # a small SPD system with the structure of the path: rho I + low rank + smooth part
    Y = rng.normal(size=(30, p))
    x = np.sort(rng.uniform(0, 1, size=p))
    S = np.exp(-np.abs(x[:, None] - x[None, :]) * 5.0)
    A = 0.3 * S - (Y.T @ Y) / 40.0
    A[np.diag_indices(p)] += 3.0 + np.linalg.norm(Y, 2) ** 2 / 40.0
    F = build_factor(A, off)
    b = rng.normal(size=(p, 3))
    w, final_count = apply_plan(F, plan, b)
    want = np.linalg.solve(A, b)
    assert np.max(np.abs(w - want)) / np.max(np.abs(want)) <= 1e-11
    assert final_count.min() == 1 and final_count.max() == 1   # every row gets its epilogue exactly once


def test_default_partitions_are_consistent():
    for p in (2048, 3000, 4096, 10000, 12288, 12289, 40000):
        off = eng.block_offsets(p, 0)
        assert len(off) - 1 == eng.default_blocks(p)
        sizes = np.diff(off)
        assert sizes.max() - sizes[:-1].min() == 0 or len(sizes) == 1       # uniform except the last block
        assert sizes[-1] <= sizes[0]


# ---------------------------------------------------------------------------------------------
# NumPy model of the polar step's inverse square root (csrc/admm_kernels.cu: ns_inv_sqrt_cta): the same
# scaling, iteration and stopping rule, to pin the constants (24 steps, test at 3e-9 before a step).
# ---------------------------------------------------------------------------------------------
def ns_inv_sqrt_model(G, max_steps=24, tol=3e-9):
    K = G.shape[0]
    sc = np.max(np.sum(np.abs(G), axis=1))
    if not (sc > 0.0) or not np.isfinite(sc):
        return None, 0
    Y, Z = G / sc, np.eye(K)
    with np.errstate(all="ignore"):            # a diverging iteration overflows; it then simply never converges
        for it in range(max_steps):
            M = 3.0 * np.eye(K) - Z @ Y
            big = not (np.max(np.abs(M - 2.0 * np.eye(K))) < tol)
            Y, Z = 0.5 * (Y @ M), 0.5 * (M @ Z)
            if not big:
                return 0.5 * (Z + Z.T) / np.sqrt(sc), it + 1
    return None, max_steps


def _gram(K, cond, rng):
    Q, _ = np.linalg.qr(rng.normal(size=(K, K)))
    lam = np.geomspace(1.0, 1.0 / cond, K) * rng.uniform(0.5, 50.0)
    return (Q * lam) @ Q.T


@pytest.mark.parametrize("K", [1, 2, 5, 8, 16])
def test_newton_schulz_model_on_the_paths_gram_matrices(K):
    """cond(T'T) <= ~1.5 on this path: a handful of steps, eigen-solve accuracy."""
    rng = np.random.default_rng(K)
    for cond in (1.0, 1.05, 1.5, 4.0, 30.0):
        G = _gram(K, cond if K > 1 else 1.0, rng)
        P, steps = ns_inv_sqrt_model(G)
        assert P is not None and steps <= (8 if cond <= 1.5 else 16)
        lam, V = np.linalg.eigh(G)
        want = (V / np.sqrt(lam)) @ V.T
        assert np.max(np.abs(P - want)) / np.max(np.abs(want)) <= 1e-13 * max(cond, 1.0)
        assert np.max(np.abs(P @ G @ P - np.eye(K))) <= 1e-13 * max(cond, 1.0)


def test_newton_schulz_model_gives_up_on_bad_gram_matrices():
    """Ill-conditioned, singular, indefinite or non-finite input must end in 'not converged' (the kernel then
    runs the Jacobi eigen-solve, which classifies the failure) -- never in a wrong answer."""
    rng = np.random.default_rng(0)
    # the smallest eigenvalue of Z Y grows 2.25x per step, so cond = 1e6 still converges (~22 steps, accuracy
    # ~cond * eps like the eigen-solve); cond = 1e13 does not
    G6 = _gram(5, 1e6, rng)
    P6, steps6 = ns_inv_sqrt_model(G6)
    assert P6 is not None and 16 <= steps6 <= 24
    assert np.max(np.abs(P6 @ G6 @ P6 - np.eye(5))) <= 1e-8
    assert ns_inv_sqrt_model(_gram(5, 1e13, rng))[0] is None
    S = _gram(4, 2.0, rng)
    S[:, 0] = S[:, 1]; S[0, :] = S[1, :]                      # singular
    assert ns_inv_sqrt_model(S)[0] is None
    assert ns_inv_sqrt_model(-_gram(3, 2.0, rng))[0] is None    # negative definite
    N = _gram(3, 2.0, rng); N[1, 1] = np.nan
    assert ns_inv_sqrt_model(N)[0] is None
    assert ns_inv_sqrt_model(np.zeros((3, 3)))[0] is None
