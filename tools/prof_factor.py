"""ncu target: one block factor at size p with m blocks (after one warm-up)."""
import sys
sys.path.insert(0, ".")
from spatpca_b200 import _lib, engine as eng
_lib.check(_lib.load().spb_set_device(0))
p, m = int(sys.argv[1]), int(sys.argv[2])
print(eng.bench_factor(p, 1, m))
