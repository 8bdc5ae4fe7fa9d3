"""K-selection loop timing: stateless spb_cv per K vs the session entry point (scratch tool)."""
import sys
import time

sys.path.insert(0, ".")
import spatpca_b200 as sp
from spatpca_b200 import engine as eng
from spatpca_b200.synth import make_case

c = make_case(sys.argv[1] if len(sys.argv) > 1 else "C3")
args = lambda K, **kw: sp.spatpcaCV(c["x"], c["Y"], c["M"], K, c["tau1"], c["tau2"], c["gamma"], c["nk"], 100, 1e-4,
                                    c["l2"], return_stats=True, **kw)
args(1)
for mode in (False, True):
    eng.session_clear()
    t0 = time.time()
    out = []
    for K in (1, 2, 3, 4, 5, 6):
        r = args(K, session=mode)
        out.append((K, round(r["selected_gamma"], 4), r["stats"]["n_inverses"]))
    print("session" if mode else "stateless", "K=1..6 total_s", round(time.time() - t0, 3), out, flush=True)
eng.session_clear()
