"""ncu target: forced Core3 iterations of the persistent ADMM kernel (argv: p K iterations [blocks])."""
import sys
sys.path.insert(0, ".")
from spatpca_b200 import _lib, engine as eng
_lib.check(_lib.load().spb_set_device(0))
p, K, it = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
blocks = int(sys.argv[4]) if len(sys.argv) > 4 else 1
print("us/iteration", 1e3 * eng.bench_admm_blocks(p, K, 3, it, 1, blocks), "blocks", blocks)
