"""GPU probe: block-factor vs explicit-inverse cost and the ADMM kernel's per-iteration time per block count."""
import json
import sys

from spatpca_b200 import _lib, engine as eng

_lib.check(_lib.load().spb_set_device(0))
out = {}
for p in [int(a) for a in sys.argv[1:]] or [4096, 10000]:
    row = {}
    for m in (1, 2, 3, 4, 6, 8):
        f_ms = eng.bench_factor(p, 2, m)
        a3 = eng.bench_admm_blocks(p, 5, 3, 10, 3, m)
        a2 = eng.bench_admm_blocks(p, 5, 2, 10, 3, m)
        row[m] = {"factor_ms": f_ms, "core3_us_per_iter": a3 * 1e3, "core2_us_per_iter": a2 * 1e3}
        print(p, m, row[m], flush=True)
    out[p] = row
print(json.dumps(out))
