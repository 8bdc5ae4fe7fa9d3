"""ncu target: a few passes of the TMA-fed DMMA stream at p = 10 000 (25 vectors and 5 vectors)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spatpca_b200 import engine as eng, _lib
_lib.check(_lib.load().spb_set_device(0))
nc = int(sys.argv[1]) if len(sys.argv) > 1 else 25
print(nc, eng.bench_dmma_matvec(10000, nc, 2))
