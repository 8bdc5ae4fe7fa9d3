"""How many INT8 digit planes does the Omega build need?  Builds thinPlateSplineMatrix at the C4 sites with the default
plane counts (residual 12 / correction 6 / product 7) and with smaller ones, and prints the largest deviation from the
default build relative to max |Omega| (the default is 1e-12 .. 4e-12 from an extended-precision ground truth,
profiles/r02_omega_threeway.md) and the build's device time.  SPB_TRACE=1 adds the per-step laps on stderr."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import spatpca_b200 as sp
from spatpca_b200 import _lib
from spatpca_b200.synth import make_case

_lib.check(_lib.load().spb_set_device(0))
name = sys.argv[1] if len(sys.argv) > 1 else "C4"
x = make_case(name)["x"]
variants = [(12, 6, 7), (11, 6, 7), (10, 6, 7), (12, 5, 7), (12, 4, 7), (12, 6, 6), (12, 6, 5), (11, 5, 6), (10, 5, 5)]
base = None
for S, Sc, Sp in variants:
    os.environ["SPB_OZAKI_SLICES"] = str(S)
    os.environ["SPB_OZAKI_CORR_SLICES"] = str(Sc)
    os.environ["SPB_OZAKI_PROD_SLICES"] = str(Sp)
    t0 = time.perf_counter()
    om = sp.thinPlateSplineMatrix(x)
    dt = time.perf_counter() - t0
    iu = np.triu_indices(om.shape[0])
    if base is None:
        base = om[iu].copy()
        scale = float(np.max(np.abs(base)))
    dev = float(np.max(np.abs(om[iu] - base))) / scale
    print(json.dumps(dict(case=name, residual_planes=S, correction_planes=Sc, product_planes=Sp,
                          max_dev_from_default=dev, call_s=round(dt, 3))), flush=True)
    del om
