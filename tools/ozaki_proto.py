"""CPU model of the int8-sliced (Ozaki-type) products of the Omega build: the slices are integer-valued float64 arrays, so
BLAS products of them are exact (sums < 2^53) and model the int8 x int8 -> int32 tensor-core GEMMs bit for bit.
  python tools/ozaki_proto.py p [S_resid] [S_corr] [x0_noise]"""
import sys, time
import numpy as np, scipy.linalg as sla
sys.path.insert(0, '.')
from oracle import spatpca_ref as ref

def expo(mx):
    e = np.zeros(mx.shape, dtype=np.int64)
    nz = mx > 0
    e[nz] = np.frexp(mx[nz])[1]          # mx = f * 2^e, f in [0.5, 1)  ->  |x| 2^-e < 1
    return e

def slices(M, e_cols, S):
    """digits d_s (|d| <= 64) with M[:, j] = 2^(e_j - 6) sum_s d_s 2^(-7 s) + remainder"""
    y = np.ldexp(M, (6 - e_cols)[None, :].astype(np.int32))
    out = []
    for s in range(S):
        d = np.rint(y)
        out.append(d)
        y = (y - d) * 128.0
    return out

def oz_product(A_sym, ea, B, eb, S, sa=None):
    """A_sym @ B with row scaling of A (= its column scaling, symmetric) and column scaling of B; diagonals s + t < S"""
    if sa is None: sa = slices(A_sym, ea, S)        # sa[s][:, i] = digits of row i
    sb = slices(B, eb, S)
    acc = np.zeros((A_sym.shape[0], B.shape[1]))
    for d in range(S - 1, -1, -1):
        C = np.zeros_like(acc)
        for s in range(d + 1):
            C += sa[s].T @ sb[d - s]
        assert np.abs(C).max() < 2**31
        acc += C * 2.0 ** (-7 * d)
    return np.ldexp(acc, (ea[:, None] + eb[None, :] - 12).astype(np.int32))

if __name__ == "__main__":
    p = int(sys.argv[1]); S_res = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    S_cor = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    noise = float(sys.argv[4]) if len(sys.argv) > 4 else 3e-4
    steps = int(sys.argv[5]) if len(sys.argv) > 5 else 2
    x = np.random.default_rng(2024).uniform(0, 1, (p, 2))
    A = ref.tps_bordered(x); q = A.shape[0]; A = A + 1e-8 * np.eye(q)
    lu, piv = sla.lu_factor(A)
    Xt = sla.lu_solve((lu, piv), np.eye(q)).astype(np.longdouble)
    Al = A.astype(np.longdouble); I = np.eye(q, dtype=np.longdouble)
    for it in range(3):
        R = I - Al @ Xt
        Xt = Xt + sla.lu_solve((lu, piv), R.astype(np.float64)).astype(np.longdouble)
    Lp = Xt[:p, :p]; M = Xt[:p, p:]
    truth = (Lp - np.longdouble(1e-8) * (Lp @ Lp) + np.longdouble(1e-8) * (M @ M.T)).astype(np.float64)
    rng = np.random.default_rng(1)
    # a symmetric start with |I - A X0| ~ noise: the exact inverse of a slightly perturbed A
    G = rng.standard_normal((q, q)); G = (G + G.T) * np.abs(A)
    dA = G
    X = sla.lu_solve(sla.lu_factor(A + 1e-14 * dA), np.eye(q)); X = np.triu(X) + np.triu(X, 1).T
    r0 = np.abs(np.eye(q) - A @ X).max()
    X = sla.lu_solve(sla.lu_factor(A + 1e-14 * dA * noise / r0), np.eye(q)); X = np.triu(X) + np.triu(X, 1).T
    ea = expo(np.abs(A).max(axis=0))
    print("sum_k|a||x| / (rowmax colmax) max:", float(((np.abs(A) / np.abs(A).max(axis=1)[:, None]) @ (np.abs(X) / np.abs(X).max(axis=0)[None, :])).max()))
    for it in range(steps):
        S = S_res
        R = np.eye(q) - oz_product(A, ea, X, expo(np.abs(X).max(axis=0)), S)
        Rex = (I - Al @ X.astype(np.longdouble)).astype(np.float64)
        print(f"step {it}: max|R| {np.abs(R).max():.3e}  residual error vs long double {np.abs(R - Rex).max():.3e}")
        ex = expo(np.abs(X).max(axis=0))
        T = oz_product(X, ex, R, expo(np.abs(R).max(axis=0)), S_cor)
        X = X + T
        X = np.triu(X) + np.triu(X, 1).T
        print(f"        X rel err {np.abs(X - Xt.astype(np.float64)).max() / np.abs(Xt).max():.3e}")
    Lq = X[:p, :p].copy(); Mq = X[:p, p:]
    el = expo(np.abs(Lq).max(axis=0))
    Om = Lq - 1e-8 * oz_product(Lq, el, Lq, el, S_cor) + 1e-8 * (Mq @ Mq.T)
    iu = np.triu_indices(p)
    print("Omega vs truth (upper, max-abs rel): %.3e" % (np.abs(Om[iu] - truth[iu]).max() / np.abs(truth).max()))
    orc = ref.thin_plate_spline_matrix(x)
    print("oracle vs truth: %.3e" % (np.abs(orc[iu] - truth[iu]).max() / np.abs(truth).max()))
