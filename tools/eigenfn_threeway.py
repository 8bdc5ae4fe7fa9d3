"""eigenFunction (src/RcppSpatPCA.cpp:94-147) three ways: the CUDA path, the oracle (LAPACK) and an 80-bit ground truth
(kernel and bordered solve in long double with iterative refinement).  Prints both errors relative to max |truth| and
cond * eps of the un-ridged bordered system."""
import json
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.linalg as sla


def truth_eigen_function(xn, xo, Phi):
    LD = np.longdouble
    xn, xo, Phi = xn.astype(LD), xo.astype(LD), Phi.astype(LD)
    p, d = xo.shape

    def kern(a, b):
        r = np.sqrt(((a[:, None, :] - b[None, :, :]) ** 2).sum(-1))
        with np.errstate(divide="ignore", invalid="ignore"):
            if d == 1:
                return np.where(r > 0, r * r * r / LD(12), LD(0))
            if d == 2:
                return np.where(r > 0, r * r * np.log(np.where(r > 0, r, LD(1))) / (LD(8) * LD(math.pi)), LD(0))
            return np.where(r > 0, -r / (LD(8) * LD(math.pi)), LD(0))

    q = p + d + 1
    L = np.zeros((q, q), dtype=LD)
    L[:p, :p] = kern(xo, xo)
    L[:p, p] = 1
    L[p, :p] = 1
    L[:p, p + 1:] = xo
    L[p + 1:, :p] = xo.T
    rhs = np.zeros((q, Phi.shape[1]), dtype=LD)
    rhs[:p] = Phi
    lu = sla.lu_factor(L.astype(np.float64))
    x = sla.lu_solve(lu, rhs.astype(np.float64)).astype(LD)
    for _ in range(8):
        x = x + sla.lu_solve(lu, (rhs - L @ x).astype(np.float64)).astype(LD)
    out = kern(xn, xo) @ x[:p] + x[p][None, :] + xn @ x[p + 1:]
    return out


if __name__ == "__main__":
    import spatpca_b200 as sp
    from spatpca_b200 import _lib
    from oracle import spatpca_ref as ref
    _lib.check(_lib.load().spb_set_device(0))
    for d in (1, 2, 3):
        rng = np.random.default_rng(40 + d)
        xo = rng.uniform(-2, 2, size=(80, d))
        xn = np.vstack([rng.uniform(-2, 2, size=(17, d)), xo[:3]])
        Phi = rng.normal(size=(80, 3))
        t = truth_eigen_function(xn, xo, Phi)
        scale = float(np.max(np.abs(t)))
        got = sp.eigenFunction(xn, xo, Phi)
        want = ref.eigen_function(xn, xo, Phi)
        cond = float(np.linalg.cond(ref.tps_bordered(xo)))
        print(json.dumps(dict(d=d, cond=cond, cond_eps=cond * 2.2e-16,
                              gpu_vs_truth=float(np.max(np.abs(got - t))) / scale,
                              oracle_vs_truth=float(np.max(np.abs(want - t))) / scale,
                              gpu_vs_oracle=float(np.max(np.abs(got - want))) / scale)))
