"""Inputs of the reference's own testthat files, regenerated bit-for-bit through
the R-RNG emulator (tests/testthat/test-SpatPCA.R:2-13, 169-179 and
test-RcppExports.R:2-8).  Shared by the oracle-pinning tests and the GPU parity
tests, so both sides see exactly the reference's inputs."""
import numpy as np

from oracle.r_rng import RRng, r_seq_len


def expand_grid(*axes):
    """`expand.grid`: the first factor varies fastest."""
    mesh = np.meshgrid(*axes, indexing="ij")
    return np.stack([m.ravel(order="F") for m in mesh], axis=1)


def case_1d(p=10, seed=1234):
    """test-SpatPCA.R:2-11 (p=10) and test-helper.R:33-41 (p=4).  Returns
    (rng positioned after the data draws, x, Y)."""
    rng = RRng(seed)
    x = r_seq_len(-5, 5, p).reshape(-1, 1)
    phi = np.exp(-x ** 2)
    phi = phi / np.linalg.norm(phi)
    a = rng.rnorm(100, sd=3.0)
    noise = rng.rnorm(100 * p).reshape(p, 100).T        # matrix(rnorm(n*p), n, p): column-major
    return rng, x, np.outer(a, phi[:, 0]) + noise


def case_3d(seed=1234):
    """test-SpatPCA.R:169-179."""
    rng = RRng(seed)
    s = r_seq_len(-5, 5, 4)
    d = expand_grid(s, s, s)
    phi = np.exp(-d ** 2).sum(axis=1)
    phi = phi / np.linalg.norm(phi)
    a = rng.rnorm(100, sd=3.0)
    noise = rng.rnorm(100 * 64).reshape(64, 100).T
    return rng, d, np.outer(a, phi) + noise


def tps_locations():
    """test-RcppExports.R:2-8."""
    s = r_seq_len(-5, 5, 2)
    return expand_grid(s, s), expand_grid(s, s, s)
