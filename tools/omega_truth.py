"""Extended-precision ground truth of thinPlateSplineMatrix for uniform 2-D sites (seed 2024): LU + iterative
refinement with 80-bit residuals, Omega by the identity formula.  python tools/omega_truth.py p [out.npy]"""
import sys, time
import numpy as np, scipy.linalg as sla
from concurrent.futures import ProcessPoolExecutor
sys.path.insert(0, '.')
from oracle import spatpca_ref as ref

def _mm(args):
    a, b = args
    return a @ b

def pmm(A, B, workers=12):
    """long double A @ B, rows of A split over processes"""
    chunks = np.array_split(np.arange(A.shape[0]), workers)
    with ProcessPoolExecutor(workers) as ex:
        parts = list(ex.map(_mm, [(A[c], B) for c in chunks]))
    return np.vstack(parts)

if __name__ == "__main__":
    p = int(sys.argv[1]); out = sys.argv[2] if len(sys.argv) > 2 else f"/tmp/omega_truth_p{p}.npy"
    x = np.random.default_rng(2024).uniform(0, 1, (p, 2))
    A = ref.tps_bordered(x); q = A.shape[0]; A = A + 1e-8 * np.eye(q)
    lu, piv = sla.lu_factor(A)
    Xt = sla.lu_solve((lu, piv), np.eye(q)).astype(np.longdouble)
    Al = A.astype(np.longdouble); I = np.eye(q, dtype=np.longdouble)
    for it in range(4):
        t = time.time(); R = I - pmm(Al, Xt)
        dX = sla.lu_solve((lu, piv), R.astype(np.float64)); Xt = Xt + dX.astype(np.longdouble)
        print("refine", it, float(np.abs(R).max()), float(np.abs(dX[:p, :p]).max() / np.abs(Xt[:p, :p]).max()), round(time.time() - t, 1), flush=True)
    Lp = Xt[:p, :p]; M = Xt[:p, p:]
    truth = (Lp - np.longdouble(1e-8) * pmm(Lp, Lp) + np.longdouble(1e-8) * (M @ M.T)).astype(np.float64)
    np.save(out, truth)
    print("saved", out)
