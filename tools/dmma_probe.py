"""Parity + timing probe of the TMA-fed DMMA matrix stream (U = A V)."""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spatpca_b200 import engine as eng, _lib
_lib.check(_lib.load().spb_set_device(0))
rng = np.random.default_rng(0)
for p, nc in ((64, 3), (200, 8), (1000, 5), (1003, 25), (2500, 32), (4100, 12), (777, 17)):
    A = rng.normal(size=(p, p)); A = A + A.T
    V = rng.normal(size=(p, nc))
    U = eng.dmma_matvec(A, V)
    W = A @ V
    print("parity", p, nc, float(np.abs(U - W).max() / np.abs(W).max()), flush=True)
for p, nc in ((10000, 25), (10000, 32), (10000, 5), (10000, 8), (10000, 16), (4096, 25), (4096, 5), (16000, 25)):
    ms = eng.bench_dmma_matvec(p, nc, 5)
    ncp = 8 if nc <= 8 else (16 if nc <= 16 else 32)
    print(json.dumps(dict(p=p, nc=nc, us_per_pass=ms * 1e3, gbs=8.0 * p * p / (ms * 1e-3) / 1e9,
                          tflops_padded=2.0 * p * p * ncp / (ms * 1e-3) / 1e12)), flush=True)
