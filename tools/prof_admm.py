import sys
sys.path.insert(0, ".")
from spatpca_b200 import engine as eng
p = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 5
it = int(sys.argv[3]) if len(sys.argv) > 3 else 10
print("admm ms/iter", eng.bench_admm(p, K, 3, it, 1))
