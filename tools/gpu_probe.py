"""Scratch performance probe run on the GPU box (not part of the product or the tests)."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import spatpca_b200 as sp
from spatpca_b200 import engine as eng
from spatpca_b200.synth import make_case

out = {}
t = time.time()
out["dgemm_4096_tflops"] = eng.bench_dgemm(4096, 3)
out["dgemm_8192_tflops"] = eng.bench_dgemm(8192, 3)
for p in (1024, 4096, 10000):
    out[f"inverse_ms_p{p}"] = eng.bench_inverse(p, 2)
for p, K in ((900, 3), (4096, 5), (10000, 5), (10000, 8)):
    for mode in (2, 3):
        ms = eng.bench_admm(p, K, mode, 50, 3)
        byt = 8.0 * p * p + (72 if mode == 3 else 40) * p * K
        out[f"admm_p{p}_K{K}_m{mode}"] = {"ms_per_iter": ms, "GBps": byt / ms / 1e6}
print(json.dumps(out, indent=1), flush=True)

for name in sys.argv[1:]:
    c = make_case(name)
    for rep in range(2):
        t0 = time.time()
        r = sp.spatpcaCV(c["x"], c["Y"], c["M"], c["K"], c["tau1"], c["tau2"], c["gamma"], c["nk"], 100, 1e-4,
                         c["l2"], return_stats=True)
        dt = time.time() - t0
        print(name, "rep", rep, "wall_s", round(dt, 3), "sel", r["selected_tau1"], r["selected_tau2"],
              r["selected_gamma"], json.dumps(r["stats"]), flush=True)
