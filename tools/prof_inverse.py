import sys
sys.path.insert(0, ".")
from spatpca_b200 import engine as eng
p = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
print("inverse ms", eng.bench_inverse(p, 1))
