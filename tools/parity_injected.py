"""ADMM/CV parity with the SAME Omega on both sides (CPU-built Omega injected into the GPU engine)."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import spatpca_b200 as sp
from oracle import spatpca_ref as ref
from spatpca_b200.synth import make_case

name = sys.argv[1]
c = make_case(name)
om = ref.thin_plate_spline_matrix(c["x"])
M = c["M"]
e = sp.Engine(c["x"].shape[0], c["x"].shape[1], c["Y"].shape[0], M, c["K"], c["tau1"], c["tau2"], c["gamma"], 100, 1e-4,
              c["l2"])
e.set_omega(om)
e.upload(c["x"], c["Y"], c["nk"])
e.prepare()
folds = list(range(M))
r1 = e.stage1_folds(folds); cv1 = np.array([r1[k] for k in folds]); i1, s1 = sp.select_index(cv1)
e.stage1_full(i1)
r2 = e.stage2_folds(folds, i1); cv2 = np.array([r2[k] for k in folds]); i2, s2 = sp.select_index(cv2)
e.stage2_full(i1, i2)
r3 = e.stage3_folds(folds, i1, i2); cv3 = np.array([r3[k] for k in folds]); i3, s3 = sp.select_index(cv3)
A = e.get_eigenfn()
want = ref.spatpca_cv(c["x"], c["Y"], M, c["K"], c["tau1"], c["tau2"], c["gamma"], c["nk"], 100, 1e-4, c["l2"], omega=om)
B = want["estimated_eigenfn"]
sgn = np.sign(np.sum(A * B, axis=0))
rel = lambda a, b: float(np.max(np.abs(a - b) / np.abs(b).clip(1e-300)))
print(json.dumps({"config": name, "omega": "injected (CPU-built, identical on both sides)",
                  "eigenfn_max_rel_err": float(np.max(np.abs(A * sgn - B)) / np.max(np.abs(B))),
                  "cv_rel_err": [rel(s1, want["cv_score_tau1"]), rel(s2, want["cv_score_tau2"]), rel(s3, want["cv_score_gamma"])],
                  "selected_equal": [bool(c["tau1"][i1] == want["selected_tau1"]), bool(c["tau2"][i2] == want["selected_tau2"]),
                                     bool(c["gamma"][i3] == want["selected_gamma"])]}))
