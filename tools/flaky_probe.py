"""Run-to-run determinism of spatpcaCV over the branch variants of test_cv_branches (debugging aid)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import spatpca_b200 as sp
from spatpca_b200 import _lib
from oracle import r_api
from synth import make_case
_lib.check(_lib.load().spb_set_device(0))
case = make_case("grid2d_small")
variants = [
    dict(tau1=[0.0], tau2=[0.0], gamma=[0.0, 0.5, 1.0]),
    dict(tau1=[10.0], tau2=[0.0], gamma=case["gamma"]),
    dict(tau1=[10.0], tau2=[100.0], gamma=[0.0, 1.0]),
    dict(tau1=[0.0], tau2=[0.0, 1.0, 10.0], gamma=case["gamma"]),
    dict(tau1=[0.5], tau2=[0.0, 1.0, 10.0], gamma=[2.0]),
    dict(tau1=case["tau1"], tau2=[3.0], gamma=case["gamma"]),
    dict(tau1=[0.0], tau2=[5.0], gamma=case["gamma"]),
]
ref = {}
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
    for vi, v in enumerate(variants):
        c = dict(case)
        c.update({k: np.asarray(val, dtype=float) for k, val in v.items()})
        c["l2"] = r_api.set_l2(c["tau2"])
        r = sp.spatpcaCV(c["x"], c["Y"], c["M"], c["K"], c["tau1"], c["tau2"], c["gamma"], c["nk"], 100, 1e-4, c["l2"])
        key = np.concatenate([np.ravel(r["cv_score_tau1"]), np.ravel(r["cv_score_tau2"]), np.ravel(r["cv_score_gamma"])])
        if vi not in ref:
            ref[vi] = key
        elif not np.array_equal(ref[vi], key):
            d = np.abs(ref[vi] - key)
            print("rep", rep, "variant", vi, "DIFFERS: max", d.max(), "at", int(d.argmax()), "of", len(key), flush=True)
print("done")
