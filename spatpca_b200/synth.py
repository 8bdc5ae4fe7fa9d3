"""Seeded synthetic spatial fields for the parity tests and the benchmark.

    locations X;  K_true smooth bumps phi_j(x) = exp(-|x - c_j|^2 / w_j), orthonormalised by QR;
    Y = Z diag(lambda) Phi' + E,  Z, E ~ N(0, 1) i.i.d.,  lambda = (25, 16, 9, 4, 2, 1.5, ...) * sqrt(p) / 10;
    folds nk = permutation(rep(1:M, length.out = n)).

Pure NumPy host code that only *generates inputs*; it is shared by tests/ and bench.py so that the
GPU path, the oracle and the CPU baseline all see bit-identical data.
"""
from __future__ import annotations

import math

import numpy as np

_LAMBDA = np.array([25.0, 16.0, 9.0, 4.0, 2.0, 1.5, 1.2, 1.1])


def _log_grid(lo, hi, m):
    return np.concatenate([[0.0], np.exp(np.linspace(math.log(lo), math.log(hi), m))])


def field(x, n, k_true, seed):
    rng = np.random.default_rng(seed)
    p, d = x.shape
    lo, hi = x.min(axis=0), x.max(axis=0)
    centers = lo + (hi - lo) * rng.uniform(0.15, 0.85, size=(k_true, d))
    widths = (0.08 + 0.25 * rng.uniform(size=k_true)) * float(np.sum((hi - lo) ** 2))
    B = np.stack([np.exp(-np.sum((x - c) ** 2, axis=1) / w) for c, w in zip(centers, widths)], axis=1)
    Phi, _ = np.linalg.qr(B)
    lam = _LAMBDA[:k_true] * math.sqrt(p) / 10.0
    Z = rng.normal(size=(n, k_true))
    Y = (Z * lam[None, :]) @ Phi.T + rng.normal(size=(n, p))
    return np.asfortranarray(Y), Phi


def folds(n, M, seed):
    rng = np.random.default_rng(seed + 1)
    return rng.permutation(np.resize(np.arange(1, M + 1), n)).astype(float)


def default_gamma(Y, nk):
    """setGamma(NULL, Y[shuffle_split != 1, ]) of the reference (R/helper.R:207-217)."""
    d1 = np.linalg.svd(Y[nk != 1], compute_uv=False)[0]
    mx = d1 ** 2 / np.sum(nk != 1)
    return _log_grid(mx / 1e4, mx, 10)


def set_l2(tau2):
    tau2 = np.atleast_1d(np.asarray(tau2, dtype=float))
    if len(tau2) == 1 and tau2[0] > 0:
        return _log_grid(tau2[0] / 1e4, tau2[0], 10)
    return np.array([1.0])


def _grid(side, d, lo=-5.0, hi=5.0):
    s = np.linspace(lo, hi, side)
    mesh = np.meshgrid(*([s] * d), indexing="ij")
    return np.stack([m.ravel(order="F") for m in mesh], axis=1)


def _uniform(p, d, seed):
    return np.random.default_rng(seed + 2).uniform(size=(p, d))


# name -> (location builder, n, K, M, tau1 grid, tau2 grid, seed)
_DEFAULT_TAU1 = _log_grid(1e-6, 1.0, 10)
_CASES = {
    # small parity cases (oracle finishes in seconds)
    "grid2d_small": dict(x=lambda s: _grid(12, 2), n=60, K=3, M=5, tau1=_DEFAULT_TAU1,
                         tau2=np.array([0.0, 10.0, 100.0]), seed=11),
    "irregular2d_small": dict(x=lambda s: _uniform(300, 2, s), n=80, K=3, M=5, tau1=_log_grid(1e-6, 1.0, 5),
                              tau2=_log_grid(1.0, 400.0, 4), seed=12),
    "grid3d_small": dict(x=lambda s: _grid(5, 3), n=60, K=2, M=5, tau1=_log_grid(1e-4, 1.0, 4),
                         tau2=np.array([0.0, 20.0]), seed=13),
    # BASELINE.json configs
    "C2": dict(x=lambda s: _grid(30, 2), n=100, K=3, M=5, tau1=_DEFAULT_TAU1, tau2=np.array([0.0]), seed=2024),
    "C2_tau2": dict(x=lambda s: _grid(30, 2), n=100, K=3, M=5, tau1=_DEFAULT_TAU1,
                    tau2=_log_grid(10.0, 400.0, 10), seed=2024),
    "C3": dict(x=lambda s: _uniform(4096, 2, s), n=200, K=5, M=5, tau1=_log_grid(1e-6, 1.0, 9),
               tau2=_log_grid(1.0, 400.0, 9), seed=2024),
    "C4": dict(x=lambda s: _uniform(10000, 2, s), n=200, K=5, M=5, tau1=_DEFAULT_TAU1,
               tau2=_log_grid(1.0, 400.0, 10), seed=2024),
    "C5": dict(x=lambda s: _uniform(40000, 3, s), n=500, K=8, M=5, tau1=_log_grid(1e-6, 1.0, 7),
               tau2=_log_grid(1.0, 400.0, 7), seed=2024),
    # smallest irregular case whose Core2 / Core3 calls take the matrix-free (fold-batched) path (p >= 2048)
    "irregular2d_2304": dict(x=lambda s: _uniform(2304, 2, s), n=100, K=3, M=5, tau1=_log_grid(1e-5, 1.0, 4),
                             tau2=_log_grid(1.0, 400.0, 3), seed=77),
    # mid-size irregular case used by the slower parity test
    "irregular2d_1500": dict(x=lambda s: _uniform(1500, 2, s), n=200, K=5, M=5, tau1=_log_grid(1e-6, 1.0, 9),
                             tau2=_log_grid(1.0, 400.0, 9), seed=2024),
}


def case_names():
    return list(_CASES)


def make_case(name: str) -> dict:
    c = _CASES[name]
    seed = c["seed"]
    x = np.asfortranarray(c["x"](seed))
    Y, Phi_true = field(x, c["n"], max(c["K"], 3), seed)
    nk = folds(c["n"], c["M"], seed)
    tau2 = np.asarray(c["tau2"], dtype=float)
    return dict(name=name, x=x, Y=Y, M=c["M"], K=c["K"], tau1=np.asarray(c["tau1"], dtype=float), tau2=tau2,
                gamma=default_gamma(Y, nk), nk=nk, l2=set_l2(tau2), Phi_true=Phi_true)
