"""Forced-iteration mode of SURVEY.md section 8d: thr = 0, so every ADMM call runs exactly maxit iterations and
ADMM iterations/s has a deterministic denominator.  argv: config [maxit]"""
import json
import sys
import time

sys.path.insert(0, ".")
import spatpca_b200 as sp
from spatpca_b200.synth import make_case

name = sys.argv[1] if len(sys.argv) > 1 else "C4"
maxit = int(sys.argv[2]) if len(sys.argv) > 2 else 100
c = make_case(name)
args = (c["x"], c["Y"], c["M"], c["K"], c["tau1"], c["tau2"], c["gamma"], c["nk"])
sp.spatpcaCV(*args, 2, 0.0, c["l2"])                    # warm-up (module loading, allocator)
t0 = time.time()
r = sp.spatpcaCV(*args, maxit, 0.0, c["l2"], return_stats=True)
dt = time.time() - t0
st = r["stats"]
print(json.dumps({"config": name, "maxit": maxit, "thr": 0.0, "wall_s": dt, "admm_calls": st["n_admm_calls"],
                  "admm_iters": st["admm_iters"], "iters_per_s": st["admm_iters"] / dt,
                  "ms_admm_stream": st["ms_admm"], "ms_inverse_stream": st["ms_inverse"],
                  "selected": [r["selected_tau1"], r["selected_tau2"], r["selected_gamma"]]}))
