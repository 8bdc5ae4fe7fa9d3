"""Per-phase timing of the fold-batched matrix-free Core2 kernel (SPB_BCG_DBG=1): one tau1 step, 5 folds, p = 10 000."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SPB_BCG_DBG"] = "1"
from spatpca_b200 import engine as eng, _lib
_lib.check(_lib.load().spb_set_device(0))
p, n, K, M = int(sys.argv[1]) if len(sys.argv) > 1 else 10000, 200, 5, int(sys.argv[2]) if len(sys.argv) > 2 else 5
rng = np.random.default_rng(1)
# banded SPD stand-in for Omega_u (the kernel only streams it), data with a clear leading subspace
Om = np.zeros((p, p), order="F")
i = np.arange(p)
Om[i, i] = 0.2
for d in range(1, 6):
    Om[i[:-d], i[:-d] + d] = 0.05 / (1 + d)
    Om[i[:-d] + d, i[:-d]] = 0.05 / (1 + d)
load = np.linalg.qr(rng.normal(size=(p, K)))[0]
Y = (rng.normal(size=(n, K)) * np.linspace(3, 1, K)) @ load.T + 0.05 * rng.normal(size=(n, p))
nk = rng.permutation(np.resize(np.arange(1, M + 1), n)).astype(float)
starts, rhos = [], []
for k in range(M):
    Ytr = Y[nk != k + 1]
    s = np.linalg.svd(Ytr, compute_uv=False)
    rhos.append(10 * s[0] ** 2)
    Q = np.linalg.qr(Ytr.T @ rng.normal(size=(Ytr.shape[0], K)))[0]
    starts.append(Q)
got = eng.core2_batched(Om, Y, nk, list(range(1, M + 1)), rhos, [0.5], starts, starts, [np.zeros((p, K))] * M, 3, 0.0,
                        snapshots=False)
print("iters", got["iters"].tolist(), "passes", got["passes"])
