"""Full-size parity run: GPU spatpcaCV vs the CPU oracle on a BASELINE config (run on the GPU box)."""
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
args = (c["x"], c["Y"], c["M"], c["K"], c["tau1"], c["tau2"], c["gamma"], c["nk"], 100, 1e-4, c["l2"])
sp.spatpcaCV(*args)                                    # warm-up (module loading)
t0 = time.time()
got = sp.spatpcaCV(*args, return_stats=True)
t_gpu = time.time() - t0
eng = sp.Engine(c["x"].shape[0], c["x"].shape[1], c["Y"].shape[0], c["M"], c["K"], c["tau1"], c["tau2"], c["gamma"],
                100, 1e-4, c["l2"])
eng.upload(c["x"], c["Y"], c["nk"])
eng.prepare()
om_gpu = eng.get_omega()
eng.close()
tr = ref.Trace()
t0 = time.time()
want = ref.spatpca_cv(*args, trace=tr)
t_cpu = time.time() - t0
A, B = got["estimated_eigenfn"], want["estimated_eigenfn"]
sgn = np.sign(np.sum(A * B, axis=0))
out = {
    "config": name, "p": int(c["x"].shape[0]), "gpu_s": t_gpu, "cpu_oracle_s": t_cpu, "speedup": t_cpu / t_gpu,
    "cpu_phase_s": tr.phase_s,
    "selected_gpu": [got["selected_tau1"], got["selected_tau2"], got["selected_gamma"]],
    "selected_cpu": [want["selected_tau1"], want["selected_tau2"], want["selected_gamma"]],
    "eigenfn_max_rel_err": float(np.max(np.abs(A * sgn - B)) / np.max(np.abs(B))),
    "cv_rel_err": {k: float(np.max(np.abs(got[k] - want[k]) / np.abs(want[k]).clip(1e-300))) for k in
                   ("cv_score_tau1", "cv_score_tau2", "cv_score_gamma")},
    "admm_iters_gpu": got["stats"]["admm_iters"], "admm_iters_cpu": tr.admm_iters,
    "admm_calls_gpu": got["stats"]["n_admm_calls"], "admm_calls_cpu": len(tr.calls),
}
om_cpu = ref.thin_plate_spline_matrix(c["x"])
iu = np.triu_indices(om_cpu.shape[0])
out["omega_upper_max_rel_err"] = float(np.max(np.abs(om_gpu[iu] - om_cpu[iu])) / np.max(np.abs(om_cpu)))
print(json.dumps(out), flush=True)
