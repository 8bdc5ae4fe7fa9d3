"""Emulation of R's default RNG stream (test infrastructure only).

The reference's seeded goldens (tests/testthat/test-SpatPCA.R:2-13, 150-187)
come from `set.seed(1234)` followed by `rnorm()` draws and the `sample()` call
inside `spatpca()` (R/SpatPCA.R:184).  R itself is not installed here, so the
stream is re-created from R's published algorithms:

* `set.seed(s)`   -- Mersenne-Twister; the seed is scrambled with the LCG
                     s <- 69069*s + 1 (uint32): 50 warm-up steps, then 625 more;
                     the first of those is the (overwritten) position slot,
                     the next 624 are the MT19937 key, position = 624.
* `unif_rand()`   -- genrand_int32 * 2.3283064365386963e-10, clamped into (0,1).
* `rnorm()`       -- "Inversion": u = unif; u = int(2^27*u) + unif;
                     qnorm(u / 2^27).
* `sample(x)`     -- R >= 3.6 "Rejection" sampling: bits = ceil(log2(n));
                     draws floor(unif*65536) once per started 16 bits, masks to
                     `bits`, rejects while >= n; partial Fisher-Yates with
                     x[j] <- x[--n].

Checked against values every R user can reproduce (see tests/test_oracle_goldens.py):
`set.seed(1234); rnorm(3)` = -1.2070657, 0.2774292, 1.0844412 and
`set.seed(42); sample(1:10)` = 1 5 10 8 2 4 6 9 7 3.
"""
from __future__ import annotations

import math

import numpy as np
from scipy.special import ndtri

_I2_32M1 = 2.328306437080797e-10
_MT_SCALE = 2.3283064365386963e-10
_BIG = 134217728.0  # 2^27


class RRng:
    """R's Mersenne-Twister / Inversion / Rejection generator triple."""

    def __init__(self, seed: int):
        self.set_seed(seed)

    def set_seed(self, seed: int) -> None:
        s = np.uint32(seed & 0xFFFFFFFF)
        mul = np.uint32(69069)
        one = np.uint32(1)
        with np.errstate(over="ignore"):
            for _ in range(50):
                s = mul * s + one
            words = np.empty(625, dtype=np.uint32)
            for j in range(625):
                s = mul * s + one
                words[j] = s
        self._bg = np.random.MT19937()
        st = self._bg.state
        st["state"]["key"] = words[1:].copy()
        st["state"]["pos"] = 624
        self._bg.state = st

    # -- uniforms -------------------------------------------------------
    def unif_rand(self) -> float:
        v = float(self._bg.random_raw()) * _MT_SCALE
        if v <= 0.0:
            return 0.5 * _I2_32M1
        if 1.0 - v <= 0.0:
            return 1.0 - 0.5 * _I2_32M1
        return v

    def runif(self, n: int) -> np.ndarray:
        return np.array([self.unif_rand() for _ in range(n)])

    # -- normals --------------------------------------------------------
    def norm_rand(self) -> float:
        u = self.unif_rand()
        u = float(int(_BIG * u)) + self.unif_rand()
        return float(ndtri(u / _BIG))

    def rnorm(self, n: int, mean: float = 0.0, sd: float = 1.0) -> np.ndarray:
        return np.array([mean + sd * self.norm_rand() for _ in range(n)])

    # -- sample() -------------------------------------------------------
    def _rbits(self, bits: int) -> float:
        v = 0
        n = 0
        while n <= bits:
            v1 = int(math.floor(self.unif_rand() * 65536))
            v = 65536 * v + v1
            n += 16
        if bits < 64:
            v &= (1 << bits) - 1
        return float(v)

    def unif_index(self, dn: float) -> float:
        if dn <= 0:
            return 0.0
        bits = int(math.ceil(math.log2(dn)))
        while True:
            dv = self._rbits(bits)
            if dn > dv:
                return dv

    def sample(self, x, size: int | None = None) -> np.ndarray:
        """`sample(x)` / `sample(x, size)` without replacement (R >= 3.6)."""
        x = np.asarray(x)
        if x.ndim == 0:
            x = np.arange(1, int(x) + 1)
        n = len(x)
        k = n if size is None else int(size)
        idx = list(range(n))
        out = np.empty(k, dtype=np.int64)
        m = n
        for i in range(k):
            j = int(self.unif_index(float(m)))
            out[i] = idx[j]
            m -= 1
            idx[j] = idx[m]
        return x[out]


def r_seq_len(a: float, b: float, length: int) -> np.ndarray:
    """`seq(a, b, length = length)` (R's seq.default: from + (0:(n-1))*by)."""
    if length == 1:
        return np.array([a], dtype=float)
    by = (b - a) / (length - 1)
    out = a + np.arange(length, dtype=float) * by
    out[-1] = b
    return out
