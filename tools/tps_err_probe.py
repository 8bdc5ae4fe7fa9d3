"""max-abs relative distance of thinPlateSplineMatrix (CUDA, refined) from the oracle (LAPACK dgetrf + dgetri) per size."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spatpca_b200 as sp
from spatpca_b200 import _lib
from oracle import spatpca_ref as ref
_lib.check(_lib.load().spb_set_device(0))
for d, p in [(1, 37), (2, 100), (3, 125), (2, 900), (1, 300), (3, 1000), (2, 2500)]:
    rng = np.random.default_rng(d * 100 + p)
    x = rng.uniform(-3, 3, size=(p, d))
    got = sp.thinPlateSplineMatrix(x)
    want = ref.thin_plate_spline_matrix(x)
    print(d, p, "rel", np.max(np.abs(got - want)) / np.max(np.abs(want)), "cond", np.linalg.cond(ref.tps_bordered(x) + 1e-8 * np.eye(p + d + 1)), flush=True)
