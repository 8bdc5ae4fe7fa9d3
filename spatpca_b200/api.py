"""Host-side mirror of the reference's R layer (R/SpatPCA.R, R/helper.R) over the
GPU engine: same function names, argument meaning, defaults, clamping rules and
error strings, so that code (and tests) written against the R API read the same.

The R layer is thin scalar orchestration that stays on the host in the
reference too; everything that touches p or n
goes through libspatpca_b200.so.  The R package shell that binds the same ABI
from R is under r-package/.
"""
from __future__ import annotations

import math
import os
import warnings

import numpy as np

from . import engine as _eng


# ---- R/helper.R ------------------------------------------------------------------------------
def _seq_len(a: float, b: float, length: int) -> np.ndarray:
    if length == 1:
        return np.array([a], dtype=float)
    by = (b - a) / (length - 1)
    out = a + np.arange(length, dtype=float) * by
    out[-1] = b
    return out


def _as_matrix(x) -> np.ndarray:
    x = np.asarray(x, dtype=float)
    return x.reshape(len(x), -1) if x.ndim != 2 else x


def setCores(num_cores=None):
    """R/helper.R:10-32 -- a validator only; the worker pool is replaced by GPU batching."""
    if num_cores is None:
        return None
    if isinstance(num_cores, (list, tuple, np.ndarray)):
        if len(num_cores) != 1:
            raise ValueError("Please supply a single numeric value for num_cores.")
        num_cores = num_cores[0]
    if isinstance(num_cores, bool) or not isinstance(num_cores, (int, float, np.integer, np.floating)):
        raise ValueError(f"Please enter valid type - but got {type(num_cores).__name__}")
    default_number = os.cpu_count() or 1
    if num_cores > default_number:
        raise ValueError(f"The input number of cores is invalid - default is {default_number}")
    if num_cores < 1:
        raise ValueError(f"The number of cores is not greater than 1 - but got {num_cores}")
    return True


def scaleLocation(location):
    """R/helper.R:41-51: only 1-D locations are min-max scaled."""
    location = _as_matrix(location)
    if location.shape[1] == 1:
        lo, hi = location.min(), location.max()
        return (location - lo) / (hi - lo)
    return location


def checkInputData(Y, x, M):
    """R/helper.R:87-103."""
    x = _as_matrix(x)
    Y = np.asarray(Y)
    n, p = Y.shape
    if p < 3:
        raise ValueError("Number of locations must be larger than 2.")
    if x.shape[0] != p:
        raise ValueError("The number of rows of x should be equal to the number of columns of Y.")
    if x.shape[1] > 3:
        raise ValueError("Dimension of locations must be less than 4.")
    if M >= n:
        raise ValueError("Number of folds must be less than sample size.")


def checkNewLocationsForSpatpcaObject(spatpca_object, x_new):
    """R/helper.R:61-76."""
    if not isinstance(spatpca_object, SpatpcaObject):
        raise ValueError("Invalid object! Please enter a `spatpca` object")
    if x_new is None:
        raise ValueError("New locations cannot be NULL")
    x_new = _as_matrix(x_new)
    if x_new.shape[1] != spatpca_object.scaled_x.shape[1]:
        raise ValueError("Inconsistent dimension of locations - original dimension is "
                         f"{spatpca_object.scaled_x.shape[1]}")
    return None


def fetchUpperBoundNumberEigenfunctions(Y, M):
    """R/helper.R:113-117."""
    n, p = np.asarray(Y).shape
    return int(min(math.floor(n - n / M), p))


def setNumberEigenfunctions(K, Y, M):
    """R/helper.R:128-137."""
    upper_bound = fetchUpperBoundNumberEigenfunctions(Y, M)
    if K is not None and K > upper_bound:
        K = upper_bound
        warnings.warn(f"K must be smaller than min(floor(n - n/M), p). Set K as {K}")
    return K


def setTau1(tau1, M):
    """R/helper.R:147-159."""
    if tau1 is None:
        t = np.concatenate([[0.0], np.exp(_seq_len(math.log(1e-6), 0.0, 10))])
    else:
        t = np.atleast_1d(np.asarray(tau1, dtype=float))
    return np.array([t.max()]) if M < 2 else t


def setTau2(tau2, M):
    """R/helper.R:169-180."""
    t = np.array([0.0]) if tau2 is None else np.atleast_1d(np.asarray(tau2, dtype=float))
    return np.array([t.max()]) if M < 2 else t


def setL2(tau2):
    """R/helper.R:189-197."""
    tau2 = np.atleast_1d(np.asarray(tau2, dtype=float))
    if len(tau2) == 1 and tau2[0] > 0:
        t = float(tau2[0])
        return np.concatenate([[0.0], np.exp(_seq_len(math.log(t / 1e4), math.log(t), 10))])
    return np.array([1.0])


def setGamma(gamma, Y=None):
    """R/helper.R:207-217."""
    if gamma is None:
        d1 = np.linalg.svd(np.asarray(Y, dtype=float), compute_uv=False)[0]
        mx = d1 ** 2 / Y.shape[0]
        return np.concatenate([[0.0], np.exp(_seq_len(math.log(mx / 1e4), math.log(mx), 10))])
    return np.atleast_1d(np.asarray(gamma, dtype=float))


def detrend(Y, is_Y_detrended):
    """R/helper.R:226-232."""
    Y = np.asarray(Y, dtype=float)
    return Y - Y.mean(axis=0)[None, :] if is_Y_detrended else Y


# ---- the four .Call stubs (R/RcppExports.R) -----------------------------------------------------
thinPlateSplineMatrix = _eng.thinPlateSplineMatrix
eigenFunction = _eng.eigenFunction
spatpcaCV = _eng.spatpcaCV
spatialPrediction = _eng.spatialPrediction


# ---- R/SpatPCA.R ---------------------------------------------------------------------------------
class SpatpcaObject(dict):
    """The `spatpca` S3 object (R/SpatPCA.R:231-248): a named list with attribute access."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as e:
            raise AttributeError(name) from e


def spatpcaCVWithSelectedK(x, Y, M, tau1, tau2, gamma, shuffle_split, maxit, thr, l2):
    """R/SpatPCA.R:23-48: raise K until the selected gamma stops decreasing."""
    upper_bound = fetchUpperBoundNumberEigenfunctions(Y, M)
    # every call of the loop sees the same data, folds and grids: the session entry point keeps
    # Omega, the fold SVD starts and the K-independent p x p inverses between the calls
    try:
        prev_cv_selection = spatpcaCV(x, Y, M, 1, tau1, tau2, gamma, shuffle_split, maxit, thr, l2, session=True)
        cv_selection = prev_cv_selection
        ks = list(range(2, upper_bound + 1)) if upper_bound >= 2 else list(range(2, upper_bound - 1, -1))
        k = 2
        for k in ks:                      # `for (k in 2:upper_bound)`; 2:1 counts down in R
            cv_selection = spatpcaCV(x, Y, M, k, tau1, tau2, gamma, shuffle_split, maxit, thr, l2, session=True)
            difference = prev_cv_selection["selected_gamma"] - cv_selection["selected_gamma"]
            prev_cv_selection = cv_selection
            if difference <= 0 or abs(difference) <= 1e-8:
                break
    finally:
        _eng.session_clear()
    return {"cv_result": cv_selection, "selected_K": k - 1}


def _shuffle(M, n, rng):
    base = np.resize(np.arange(1, M + 1), n)               # rep(1:M, length.out = n)
    if rng is None:
        rng = np.random.default_rng()
    if hasattr(rng, "sample"):                              # an R-stream emulator
        return np.asarray(rng.sample(base))
    return rng.permutation(base)


def spatpca(x, Y, M=5, K=None, is_K_selected=None, tau1=None, tau2=None, gamma=None,
            is_Y_detrended=False, maxit=100, thr=1e-4, num_cores=None, rng=None, shuffle_split=None):
    """`spatpca()` -- R/SpatPCA.R:161-249.  `rng` supplies the fold shuffle of :184 (any object
    with `.sample(x)` or a NumPy Generator); `shuffle_split` overrides it."""
    if is_K_selected is None:
        is_K_selected = K is None
    Y = np.asarray(Y, dtype=float)
    checkInputData(Y, x, M)
    setCores(num_cores)
    x = _as_matrix(x)
    Y = detrend(Y, is_Y_detrended)
    K = setNumberEigenfunctions(K, Y, M)
    n = Y.shape[0]
    scaled_x = scaleLocation(x)
    if shuffle_split is None:
        shuffle_split = _shuffle(M, n, rng)
    shuffle_split = np.asarray(shuffle_split)
    tau1 = setTau1(tau1, M)
    tau2 = setTau2(tau2, M)
    l2 = setL2(tau2)
    gamma = setGamma(gamma, Y[shuffle_split != 1])
    if is_K_selected:
        sel = spatpcaCVWithSelectedK(scaled_x, Y, M, tau1, tau2, gamma, shuffle_split, maxit, thr, l2)
        cv_result = sel["cv_result"]
    else:
        cv_result = spatpcaCV(scaled_x, Y, M, K, tau1, tau2, gamma, shuffle_split, maxit, thr, l2)
    return SpatpcaObject(
        eigenfn=cv_result["estimated_eigenfn"],
        selected_K=K,                                       # sic, R/SpatPCA.R:234
        selected_tau1=cv_result["selected_tau1"],
        selected_tau2=cv_result["selected_tau2"],
        selected_gamma=cv_result["selected_gamma"],
        cv_score_tau1=cv_result["cv_score_tau1"],
        cv_score_tau2=cv_result["cv_score_tau2"],
        cv_score_gamma=cv_result["cv_score_gamma"],
        tau1=tau1, tau2=tau2, gamma=gamma,
        detrended_Y=Y, scaled_x=scaled_x,
    )


def predictEigenfunction(spatpca_object, x_new):
    """R/SpatPCA.R:272-280."""
    checkNewLocationsForSpatpcaObject(spatpca_object, x_new)
    scaled_x_new = scaleLocation(x_new)
    return eigenFunction(scaled_x_new, spatpca_object.scaled_x, spatpca_object.eigenfn)


def predict(spatpca_object, x_new, eigen_patterns_on_new_site=None):
    """R/SpatPCA.R:304-322."""
    checkNewLocationsForSpatpcaObject(spatpca_object, x_new)
    if eigen_patterns_on_new_site is None:
        eigen_patterns_on_new_site = predictEigenfunction(spatpca_object, x_new)
    sp = spatialPrediction(spatpca_object.eigenfn, spatpca_object.detrended_Y,
                           spatpca_object.selected_gamma, eigen_patterns_on_new_site)
    return sp["prediction"]
