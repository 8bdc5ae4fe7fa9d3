"""Parity of the CUDA path against ONE oracle run under several settings (environment variables differ per process).
  python tools/parity_sweep.py oracle C4        # runs the CPU oracle once, saves /tmp/oracle_C4.npz
  SPB_CG_TOL=1e-12 python tools/parity_sweep.py gpu C4"""
import json, sys, time, os
import numpy as np
sys.path.insert(0, ".")
from spatpca_b200.synth import make_case
mode, name = sys.argv[1], sys.argv[2]
c = make_case(name)
args = (c["x"], c["Y"], c["M"], c["K"], c["tau1"], c["tau2"], c["gamma"], c["nk"], 100, 1e-4, c["l2"])
path = f"/tmp/oracle_{name}.npz"
if mode == "oracle":
    from oracle import spatpca_ref as ref
    tr = ref.Trace(); t0 = time.time()
    want = ref.spatpca_cv(*args, trace=tr)
    np.savez(path, eig=want["estimated_eigenfn"], t1=want["cv_score_tau1"], t2=want["cv_score_tau2"], g=want["cv_score_gamma"],
             sel=np.array([want["selected_tau1"], want["selected_tau2"], want["selected_gamma"]]),
             calls=np.array(sorted((a, b, c_, d) for a, b, c_, d in [(str(k), f, i, it) for k, f, i, it in tr.calls]), dtype=object),
             iters=tr.admm_iters)
    print("oracle", name, "s", round(time.time() - t0, 1), "iters", tr.admm_iters, flush=True)
else:
    import spatpca_b200 as sp
    w = np.load(path, allow_pickle=True)
    sp.spatpcaCV(*args)
    t0 = time.time(); got = sp.spatpcaCV(*args, return_stats=True); dt = time.time() - t0
    A, B = got["estimated_eigenfn"], w["eig"]
    sgn = np.sign(np.sum(A * B, axis=0))
    rel = lambda a, b: float(np.max(np.abs(a - b) / np.abs(b).clip(1e-300)))
    out = dict(env={k: v for k, v in os.environ.items() if k.startswith("SPB_")}, gpu_s=round(dt, 4),
               eigenfn=float(np.max(np.abs(A * sgn - B)) / np.max(np.abs(B))),
               cv1=rel(got["cv_score_tau1"], w["t1"]), cv2=rel(got["cv_score_tau2"], w["t2"]), cvg=rel(got["cv_score_gamma"], w["g"]),
               sel_same=bool(np.array_equal(w["sel"], [got["selected_tau1"], got["selected_tau2"], got["selected_gamma"]])),
               iters=(int(got["stats"]["admm_iters"]), int(w["iters"])), passes=int(got["stats"]["cg_passes"]),
               ms_cg=round(got["stats"]["ms_cg"], 1))
    print(json.dumps(out), flush=True)
