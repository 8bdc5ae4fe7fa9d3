"""Dump thinPlateSplineMatrix() of the CUDA path for uniform 2-D sites (seed 2024), for the three-way accuracy table."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spatpca_b200 as sp
from spatpca_b200 import _lib
_lib.check(_lib.load().spb_set_device(0))
for p in [int(a) for a in sys.argv[1:]]:
    x = np.random.default_rng(2024).uniform(0, 1, (p, 2))
    om = sp.thinPlateSplineMatrix(x)
    np.save(f"gpurun_out/omega_gpu_p{p}.npy", om)
    print(p, om.shape, float(np.abs(om).max()))
