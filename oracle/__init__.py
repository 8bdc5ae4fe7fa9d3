"""CPU oracle for the spatpcaCV hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, in NumPy/SciPy, the algorithm of the reference engine
(`/root/reference/src/RcppSpatPCA.cpp`) and of the thin R layer that feeds it
(`/root/reference/R/helper.R`, `/root/reference/R/SpatPCA.R`).  Every function
cites the reference file:line it follows.

It is the checker, never the product: only `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s CPU-baseline legs may import it.  Nothing under
`spatpca_b200/` imports it, and the product has no CPU fallback.

Where the arithmetic really lives.  The reference's factorizations are done by
RcppArmadillo (header-only, `LinkingTo` at DESCRIPTION:32, version unpinned)
dispatching to R's LAPACK/BLAS (src/Makevars:3, unpinned).  Neither is present
under /root/reference, and neither R nor Armadillo is installed in this image,
so the reference itself cannot be compiled here ("unbuildable": needs Rcpp,
RcppArmadillo and an R toolchain).  The oracle therefore calls the *same LAPACK
drivers* Armadillo would (dgetrf/dgetri, dpotrf/dpotri, dgesvd, dgesdd, dsyevd)
through SciPy's bundled OpenBLAS.

Pinning.  The oracle is pinned against every golden value the reference's own
testthat files hold for this path (tests/testthat/test-RcppExports.R,
test-SpatPCA.R, test-helper.R) in `tests/test_oracle_goldens.py`; the seeded
goldens are reproduced through an emulation of R's RNG stream (`r_rng.py`).
Above p = 64 nothing in the reference pins results, so there the oracle *is*
the reference stand-in.
"""
