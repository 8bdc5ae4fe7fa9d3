"""compute-sanitizer target: the kernels built on hand-rolled grid barriers / tickets / mbarriers on small inputs
(persistent ADMM kernel explicit + factored, matrix-free FMA kernel, TMA + DMMA stream, fold-batched kernel)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spatpca_b200 import engine as eng, _lib
_lib.check(_lib.load().spb_set_device(0))
rng = np.random.default_rng(3)
for K in (5, 20):
    p, n = 640, 40
    Y = rng.normal(size=(n, p))
    G = Y.T @ Y
    x = rng.uniform(0, 1, size=(p, 2))
    D = np.linalg.norm(x[:, None, :] - x[None, :, :], axis=2)
    Om = np.exp(-D * 8.0)                      # any SPD stand-in for Omega
    s = np.linalg.svd(Y, compute_uv=False)
    rho = 10 * s[0] ** 2
    Q = np.linalg.qr(rng.normal(size=(p, K)))[0]
    L2 = np.zeros((p, K))
    for blocks in (1, 3):
        r = eng.admm_system(Om, G, 0.1, rho, 2.0, blocks, 2, Q, L2, 6, 0.0)
        print("core2 K", K, "blocks", blocks, "iters", r[5], flush=True)
        r = eng.admm_system(Om, G, 0.1, rho, 1.0, blocks, 3, Q, L2, 6, 0.0, R=Q.copy(), Lambda1=L2.copy(), tau2=0.5)
        print("core3 K", K, "blocks", blocks, "iters", r[5], flush=True)
    if K <= 8:
        r = eng.core2_matrix_free(Om, Y, 0.1, rho, Q, Q, L2, 3, 0.0)
        print("matrix-free K", K, "iters", r[3], "passes", r[4], flush=True)
        U = eng.dmma_matvec(Om, rng.normal(size=(p, 25)))
        print("dmma stream", U.shape, flush=True)
        nk = rng.permutation(np.resize(np.arange(1, 4), n)).astype(float)
        rhos = [rho] * 3
        b = eng.core2_batched(Om, Y, nk, [1, 2, 3], rhos, [0.1, 0.5], [Q] * 3, [Q] * 3, [L2] * 3, 3, 0.0)
        print("batched iters", b["iters"].tolist(), "passes", b["passes"], flush=True)
