"""ncu target: two thinPlateSplineMatrix builds at the sites of a named case (the second one is the warm one)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spatpca_b200 as sp
from spatpca_b200 import _lib
from spatpca_b200.synth import make_case

_lib.check(_lib.load().spb_set_device(0))
x = make_case(sys.argv[1] if len(sys.argv) > 1 else "C4")["x"]
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 2):
    om = sp.thinPlateSplineMatrix(x)
print(om.shape, float(om[0, 0]))
