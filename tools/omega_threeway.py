import numpy as np, scipy.linalg as sla, time, sys
sys.path.insert(0, '.')
from oracle import spatpca_ref as ref
p = int(sys.argv[1])
x = np.random.default_rng(2024).uniform(0, 1, (p, 2))
gpu = np.load(f'gpurun_out/omega_gpu_p{p}.npy')
t=time.time(); orc = ref.thin_plate_spline_matrix(x); print("oracle s", time.time()-t)
A = ref.tps_bordered(x); q = A.shape[0]
A = A + 1e-8*np.eye(q)
E = A[:p,:p] - 1e-8*np.eye(p)
lu, piv = sla.lu_factor(A)
Xt = sla.lu_solve((lu,piv), np.eye(q)).astype(np.longdouble)
Al = A.astype(np.longdouble); I = np.eye(q, dtype=np.longdouble)
for it in range(3):
    t=time.time(); R = I - Al @ Xt
    dX = sla.lu_solve((lu, piv), R.astype(np.float64)); Xt = Xt + dX.astype(np.longdouble)
    print("refine", it, float(np.abs(R).max()), float(np.abs(dX).max()/np.abs(Xt).max()), time.time()-t, flush=True)
Lp = Xt[:p,:p]; M = Xt[:p,p:]
truth = (Lp - np.longdouble(1e-8)*(Lp@Lp) + np.longdouble(1e-8)*(M@M.T)).astype(np.float64)
truth2 = (Lp.T @ (E.astype(np.longdouble) @ Lp)).astype(np.float64)
def rel(a,b): return np.abs(a-b).max()/np.abs(b).max()
print("truth formulas agree:", rel(truth, truth2))
iu = np.triu_indices(p)
print("p", p, "upper-triangle max-abs rel: gpu-vs-truth %.3e  oracle-vs-truth %.3e  gpu-vs-oracle %.3e" % (
    rel(gpu[iu], truth2[iu]), rel(orc[iu], truth2[iu]), rel(gpu[iu], orc[iu])))
print("fro rel: gpu %.3e oracle %.3e gpu-oracle %.3e" % (np.linalg.norm(gpu-truth2)/np.linalg.norm(truth2), np.linalg.norm(orc-truth2)/np.linalg.norm(truth2), np.linalg.norm(gpu-orc)/np.linalg.norm(orc)))
