"""Restatement of the reference's R layer (TEST INFRASTRUCTURE ONLY).

`/root/reference/R/helper.R` (validators, default tuning grids) and
`/root/reference/R/SpatPCA.R` (`spatpca`, `spatpcaCVWithSelectedK`,
`predictEigenfunction`, `predict`) restated over the NumPy engine of
`oracle/spatpca_ref.py`, so that the reference's testthat goldens -- which are
written against the R API -- can pin the oracle.

`rng` is an `oracle.r_rng.RRng`; `spatpca()` consumes exactly one `sample()`
from it (R/SpatPCA.R:184), like the reference does from R's global stream.
"""
from __future__ import annotations

import math
import warnings

import numpy as np

from . import spatpca_ref as ref
from .r_rng import r_seq_len


# ---- R/helper.R ---------------------------------------------------------
def scale_location(x):                                   # helper.R:41-51
    x = np.asarray(x, dtype=float)
    x = x.reshape(len(x), -1)
    if x.shape[1] == 1:
        lo, hi = x.min(), x.max()
        return (x - lo) / (hi - lo)
    return x


def check_input_data(Y, x, M):                            # helper.R:87-103
    x = np.asarray(x, dtype=float)
    x = x.reshape(len(x), -1)
    n, p = Y.shape
    if p < 3:
        raise ValueError("Number of locations must be larger than 2.")
    if x.shape[0] != p:
        raise ValueError("The number of rows of x should be equal to the number of columns of Y.")
    if x.shape[1] > 3:
        raise ValueError("Dimension of locations must be less than 4.")
    if M >= n:
        raise ValueError("Number of folds must be less than sample size.")


def fetch_upper_bound_number_eigenfunctions(Y, M):        # helper.R:113-117
    n, p = Y.shape
    return int(min(math.floor(n - n / M), p))


def set_number_eigenfunctions(K, Y, M):                   # helper.R:128-137
    ub = fetch_upper_bound_number_eigenfunctions(Y, M)
    if K is not None and K > ub:
        K = ub
        warnings.warn(f"K must be smaller than min(floor(n - n/M), p). Set K as {K}")
    return K


def set_tau1(tau1, M):                                    # helper.R:147-159
    if tau1 is None:
        t = np.concatenate([[0.0], np.exp(r_seq_len(math.log(1e-6), 0.0, 10))])
    else:
        t = np.atleast_1d(np.asarray(tau1, dtype=float))
    return np.array([t.max()]) if M < 2 else t


def set_tau2(tau2, M):                                    # helper.R:169-180
    t = np.array([0.0]) if tau2 is None else np.atleast_1d(np.asarray(tau2, dtype=float))
    return np.array([t.max()]) if M < 2 else t


def set_l2(tau2):                                         # helper.R:189-197
    tau2 = np.atleast_1d(np.asarray(tau2, dtype=float))
    if len(tau2) == 1 and tau2[0] > 0:
        t = float(tau2[0])
        return np.concatenate([[0.0], np.exp(r_seq_len(math.log(t / 1e4), math.log(t), 10))])
    return np.array([1.0])


def set_gamma(gamma, Y):                                  # helper.R:207-217
    if gamma is None:
        d1 = np.linalg.svd(Y, compute_uv=False)[0]
        mx = d1 ** 2 / Y.shape[0]
        return np.concatenate([[0.0], np.exp(r_seq_len(math.log(mx / 1e4), math.log(mx), 10))])
    return np.atleast_1d(np.asarray(gamma, dtype=float))


def detrend(Y, is_Y_detrended):                           # helper.R:226-232
    return Y - Y.mean(axis=0)[None, :] if is_Y_detrended else Y


# ---- R/SpatPCA.R --------------------------------------------------------
def spatpca_cv_with_selected_k(x, Y, M, tau1, tau2, gamma, shuffle_split, maxit, thr, l2,
                               trace=None):                # SpatPCA.R:23-48
    ub = fetch_upper_bound_number_eigenfunctions(Y, M)
    prev = ref.spatpca_cv(x, Y, M, 1, tau1, tau2, gamma, shuffle_split, maxit, thr, l2, trace)
    cur = prev
    k = 2
    # `for (k in 2:upper_bound)`: R's 2:1 counts DOWN (2, 1) when upper_bound == 1
    ks = list(range(2, ub + 1)) if ub >= 2 else list(range(2, ub - 1, -1))
    for k in ks:
        cur = ref.spatpca_cv(x, Y, M, k, tau1, tau2, gamma, shuffle_split, maxit, thr, l2, trace)
        diff = prev["selected_gamma"] - cur["selected_gamma"]
        prev = cur
        if diff <= 0 or abs(diff) <= 1e-8:
            break
    return {"cv_result": cur, "selected_K": k - 1}


def spatpca(x, Y, rng, M=5, K=None, is_K_selected=None, tau1=None, tau2=None, gamma=None,
            is_Y_detrended=False, maxit=100, thr=1e-4, trace=None):   # SpatPCA.R:161-249
    if is_K_selected is None:
        is_K_selected = K is None
    Y = np.asarray(Y, dtype=float)
    check_input_data(Y, x, M)
    x = np.asarray(x, dtype=float)
    x = x.reshape(len(x), -1)
    Y = detrend(Y, is_Y_detrended)
    K = set_number_eigenfunctions(K, Y, M)
    n, p = Y.shape
    scaled_x = scale_location(x)
    base = np.resize(np.arange(1, M + 1), n)              # rep(1:M, length.out = n)
    shuffle_split = rng.sample(base)
    tau1 = set_tau1(tau1, M)
    tau2 = set_tau2(tau2, M)
    l2 = set_l2(tau2)
    gamma = set_gamma(gamma, Y[shuffle_split != 1])
    if is_K_selected:
        sel = spatpca_cv_with_selected_k(scaled_x, Y, M, tau1, tau2, gamma, shuffle_split,
                                         maxit, thr, l2, trace)
        res = sel["cv_result"]
    else:
        res = ref.spatpca_cv(scaled_x, Y, M, K, tau1, tau2, gamma, shuffle_split, maxit, thr, l2,
                             trace)
    return {
        "eigenfn": res["estimated_eigenfn"],
        "selected_K": K,                                   # sic: SpatPCA.R:234
        "selected_tau1": res["selected_tau1"],
        "selected_tau2": res["selected_tau2"],
        "selected_gamma": res["selected_gamma"],
        "cv_score_tau1": res["cv_score_tau1"],
        "cv_score_tau2": res["cv_score_tau2"],
        "cv_score_gamma": res["cv_score_gamma"],
        "tau1": tau1, "tau2": tau2, "gamma": gamma,
        "detrended_Y": Y, "scaled_x": scaled_x,
        "shuffle_split": shuffle_split, "l2": l2,
    }


def predict_eigenfunction(obj, x_new):                     # SpatPCA.R:272-280
    x_new = np.asarray(x_new, dtype=float)
    x_new = x_new.reshape(len(x_new), -1)
    if x_new.shape[1] != obj["scaled_x"].shape[1]:
        raise ValueError("Inconsistent dimension of locations")
    return ref.eigen_function(scale_location(x_new), obj["scaled_x"], obj["eigenfn"])


def predict(obj, x_new, eigen_patterns_on_new_site=None):  # SpatPCA.R:304-322
    if eigen_patterns_on_new_site is None:
        eigen_patterns_on_new_site = predict_eigenfunction(obj, x_new)
    sp = ref.spatial_prediction(obj["eigenfn"], obj["detrended_Y"], obj["selected_gamma"],
                                eigen_patterns_on_new_site)
    return sp["prediction"]
