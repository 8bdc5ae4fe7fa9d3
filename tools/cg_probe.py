"""Timing probe of the matrix-free Core2 kernel (B200): ms per launch, matrix passes, us per pass."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spatpca_b200 import engine as eng, _lib
_lib.check(_lib.load().spb_set_device(0))
out = []
for p, K, n_tr, iters in ((10000, 5, 160, 4), (4096, 5, 160, 4), (10000, 3, 160, 4), (10000, 8, 400, 2), (2048, 5, 160, 4)):
    ms, passes = eng.bench_core2_cg(p, K, n_tr, iters, 3)
    rec = dict(p=p, K=K, n_tr=n_tr, admm_iters=iters, ms_per_launch=ms, passes=passes, us_per_pass=ms * 1e3 / max(passes, 1),
               gbs=8.0 * p * p * passes / (ms * 1e-3) / 1e9)
    out.append(rec)
    print(json.dumps(rec), flush=True)
