"""Run one spatpcaCV job on a named synthetic case and print stats (scratch tool)."""
import json
import sys
import time

sys.path.insert(0, ".")
import spatpca_b200 as sp
from spatpca_b200.synth import make_case

name = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
t0 = time.time()
c = make_case(name)
print("generated", name, c["x"].shape, c["Y"].shape, "in", round(time.time() - t0, 1), "s", flush=True)
for rep in range(reps):
    t0 = time.time()
    r = sp.spatpcaCV(c["x"], c["Y"], c["M"], c["K"], c["tau1"], c["tau2"], c["gamma"], c["nk"], 100, 1e-4, c["l2"],
                     return_stats=True)
    print(name, "rep", rep, "wall_s", round(time.time() - t0, 3), "selected", r["selected_tau1"], r["selected_tau2"],
          r["selected_gamma"], json.dumps(r["stats"]), flush=True)
