import sys, time
sys.path.insert(0, ".")
import spatpca_b200 as sp
from spatpca_b200.synth import make_case
c = make_case(sys.argv[1] if len(sys.argv) > 1 else "C3")
for rep in range(2):
    print("=== rep", rep, file=sys.stderr, flush=True)
    t0 = time.time()
    r = sp.spatpcaCV(c["x"], c["Y"], c["M"], c["K"], c["tau1"], c["tau2"], c["gamma"], c["nk"], 100, 1e-4, c["l2"], return_stats=True)
    print("wall", time.time() - t0, r["stats"], file=sys.stderr, flush=True)
