"""thinPlateSplineMatrix() of the CUDA path under the environment's route settings: time and difference from the FP64
head/tail refinement (SPB_OMEGA_RESID=dgemm, a second process).  python tools/omega_route_probe.py save|cmp tag p..."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spatpca_b200 as sp
from spatpca_b200 import _lib
_lib.check(_lib.load().spb_set_device(0))
mode, tag = sys.argv[1], sys.argv[2]
for p in [int(a) for a in sys.argv[3:]]:
    x = np.random.default_rng(2024).uniform(0, 1, (p, 2))
    sp.thinPlateSplineMatrix(x[:200])
    t0 = time.time(); om = sp.thinPlateSplineMatrix(x); dt = time.time() - t0
    if mode == "save":
        np.save(f"/tmp/omega_{tag}_p{p}.npy", om)
        print(tag, p, "wall_ms", round(dt * 1e3, 1), flush=True)
    else:
        ref = np.load(f"gpurun_in/omega_truth_p{p}.npy") if mode == "truth" else np.load(f"/tmp/omega_{tag}_p{p}.npy")
        iu = np.triu_indices(p)
        print("cmp", p, "wall_ms", round(dt * 1e3, 1), "max-abs rel diff vs", tag, float(np.abs(om[iu] - ref[iu]).max() / np.abs(ref).max()), flush=True)
