"""NumPy/SciPy restatement of the reference engine (TEST INFRASTRUCTURE ONLY).

Follows `/root/reference/src/RcppSpatPCA.cpp` function by function, calling the
LAPACK drivers Armadillo dispatches to (see `oracle/__init__.py`).  Matrices
are column-major (`order="F"`) like R/Armadillo.  All quirks of the reference
listed in SURVEY.md section 8(a) are kept on purpose.

`Trace` records every ADMM call (kind, fold, grid index, iterations) and the
time spent per phase, so that the CUDA path can be diffed call by call and the
CPU baseline can be reported per phase.
"""
from __future__ import annotations

import math
import time
from dataclasses import dataclass, field

import numpy as np
import scipy.linalg as sla
from scipy.linalg import lapack


# --------------------------------------------------------------------------
# bookkeeping
# --------------------------------------------------------------------------
@dataclass
class Trace:
    calls: list = field(default_factory=list)      # (kind, fold, grid_index, iters)
    phase_s: dict = field(default_factory=dict)    # phase -> seconds
    counts: dict = field(default_factory=dict)     # phase -> number of calls

    def add_call(self, kind, fold, idx, iters):
        self.calls.append((kind, int(fold), int(idx), int(iters)))

    def tick(self, phase, dt):
        self.phase_s[phase] = self.phase_s.get(phase, 0.0) + dt
        self.counts[phase] = self.counts.get(phase, 0) + 1

    @property
    def admm_iters(self):
        return sum(c[3] for c in self.calls)


class _Timer:
    def __init__(self, trace, phase):
        self.trace, self.phase = trace, phase

    def __enter__(self):
        self.t0 = time.perf_counter()

    def __exit__(self, *a):
        if self.trace is not None:
            self.trace.tick(self.phase, time.perf_counter() - self.t0)


def _F(a):
    return np.asfortranarray(a, dtype=np.float64)


# --------------------------------------------------------------------------
# LAPACK-backed primitives (the Armadillo calls of the reference)
# --------------------------------------------------------------------------
def arma_inv_general(A):
    """`inv(A)` for a general square matrix: dgetrf + dgetri (RcppSpatPCA.cpp:69)."""
    lu, piv, info = lapack.dgetrf(_F(A), overwrite_a=False)
    if info != 0:
        raise RuntimeError("inv(): matrix is singular")
    inv, info = lapack.dgetri(lu, piv, overwrite_lu=True)
    if info != 0:
        raise RuntimeError("inv(): matrix is singular")
    return inv


def arma_inv_sympd_symmatu(A):
    """`inv_sympd(symmatu(A))`: dpotrf + dpotri on the upper triangle, result
    mirrored (RcppSpatPCA.cpp:165, 203, 397, 621, 661, 673)."""
    c, info = lapack.dpotrf(_F(A), lower=0, clean=0, overwrite_a=False)
    if info != 0:
        raise RuntimeError("inv_sympd(): matrix is singular or not positive definite")
    inv, info = lapack.dpotri(c, lower=0, overwrite_c=True)
    if info != 0:
        raise RuntimeError("inv_sympd(): matrix is singular or not positive definite")
    iu = np.triu_indices(inv.shape[0], 1)
    inv.T[iu] = inv[iu]
    return inv


def arma_svd_right(X):
    """`svd_econ(U, s, V, X, "right")`: dgesvd, returns (s, V) with V = Vt.T."""
    _, s, vt = sla.svd(X, full_matrices=False, lapack_driver="gesvd", check_finite=False)
    return s, _F(vt.T)


def arma_polar(T):
    """`svd_econ(U,S,V,T); U.cols(0,K-1) * V.t()` (dgesdd) -- the Stiefel
    projection of RcppSpatPCA.cpp:169-171, 206-208, 251-253."""
    u, _, vt = sla.svd(T, full_matrices=False, lapack_driver="gesdd", check_finite=False)
    return u @ vt


def arma_accumulate(x):
    """Armadillo's `arrayops::accumulate`: two interleaved accumulators.  Used
    for `sum(cv, 0)` so that ties in `index_min` break as in the reference."""
    acc1 = 0.0
    acc2 = 0.0
    n = len(x)
    j = 1
    i = 0
    while j < n:
        acc1 += x[i]
        acc2 += x[i + 1]
        i += 2
        j += 2
    if (j - 1) < n:
        acc1 += x[i]
    return acc1 + acc2


def arma_colsum(cv):
    return np.array([arma_accumulate(cv[:, j]) for j in range(cv.shape[1])])


def fro(X):
    return math.sqrt(float(np.sum(np.square(X))))


# --------------------------------------------------------------------------
# thin-plate spline matrix  (RcppSpatPCA.cpp:12-76)
# --------------------------------------------------------------------------
def _pairwise_dist(A, B):
    """sqrt(sum_k (a_k - b_k)^2), accumulated dimension by dimension in the
    order of RcppSpatPCA.cpp:28, 32-34 (no m x q x d temporary)."""
    r2 = np.square(A[:, 0][:, None] - B[:, 0][None, :])
    for k in range(1, A.shape[1]):
        r2 += np.square(A[:, k][:, None] - B[:, k][None, :])
    return np.sqrt(r2, out=r2)


def tps_kernel_block(A, B):
    """Radial TPS kernel between rows of A (m x d) and rows of B (q x d)
    (RcppSpatPCA.cpp:23-36).  No r > 0 guard, as in the reference: coincident
    2-D points give 0 * log(0) = NaN."""
    d = A.shape[1]
    r = _pairwise_dist(A, B)
    if d == 1:
        return r ** 3 / 12.0
    if d == 2:
        with np.errstate(divide="ignore", invalid="ignore"):
            return r * r * np.log(r) / (8.0 * math.pi)
    if d == 3:
        return -r / (8.0 * math.pi)
    raise ValueError("d must be 1, 2 or 3")


def tps_bordered(location):
    """The symmetric bordered matrix [[E, T], [T', 0]], T = [1, X]
    (RcppSpatPCA.cpp:19-44, 62-67).  Only j > i is evaluated, so diag(E) = 0."""
    P = _F(location)
    p, d = P.shape
    q = p + d + 1
    L = np.zeros((q, q), order="F")
    E = tps_kernel_block(P, P)
    E[np.tril_indices(p)] = 0.0          # strictly upper part only (j > i)
    L[:p, :p] = E
    L[:p, p] = 1.0
    L[:p, p + 1:] = P
    iu = np.triu_indices(q, 1)
    L.T[iu] = L[iu]                       # symmatu
    return L


def thin_plate_spline_matrix(location, trace=None):
    """`thinPlateSplineMatrix` (RcppSpatPCA.cpp:58-76): Omega = Lp' (E Lp) with
    Lp the top-left p x p block of inv(L + 1e-8 I) (general LU inverse)."""
    with _Timer(trace, "omega"):
        P = _F(location)
        p = P.shape[0]
        L = tps_bordered(P)
        Linv = arma_inv_general(L + 1e-8 * np.eye(L.shape[0]))
        Lp = _F(Linv[:p, :p])
        E = _F(L[:p, :p])
        return _F(Lp.T @ (E @ Lp))


# --------------------------------------------------------------------------
# eigenFunction  (RcppSpatPCA.cpp:94-147)
# --------------------------------------------------------------------------
def eigen_function(new_location, original_location, Phi):
    Xn = _F(new_location)
    Xo = _F(original_location)
    Phi = _F(Phi)
    p, d = Xo.shape
    K = Phi.shape[1]
    L = tps_bordered(Xo)                  # L + L.t() of the strict upper part: :103
    rhs = np.zeros((p + d + 1, K), order="F")
    rhs[:p, :] = Phi
    para = sla.solve(L, rhs, check_finite=False)      # dgetrf/dgetrs: :108
    r = _pairwise_dist(Xn, Xo)
    with np.errstate(divide="ignore", invalid="ignore"):
        if d == 1:
            kern = np.where(r > 0, r * r * r / 12.0, 0.0)
        elif d == 2:
            kern = np.where(r > 0, r * r * np.log(np.where(r > 0, r, 1.0)) / (8.0 * math.pi), 0.0)
        else:
            kern = np.where(r > 0, -r / (8.0 * math.pi), 0.0)
    out = kern @ para[:p, :]
    out += para[p, :][None, :]
    for k in range(d):
        out += Xn[:, k][:, None] * para[p + 1 + k, :][None, :]
    return _F(out)


# --------------------------------------------------------------------------
# ADMM cores
# --------------------------------------------------------------------------
def core2(G, Phi, C, Lambda2, Omega, tau1, rho, maxit, tol, trace=None, tag=("core2", -1, -1)):
    """`spatpcaCore2` / `spatpcaCore2p` (RcppSpatPCA.cpp:150-222).
    Returns (Phi, C, Lambda2).  The p^3 inverse is taken on every call (:165)."""
    p, K = C.shape
    with _Timer(trace, "inv_sympd"):
        A = 2.0 * (tau1 * Omega - G)
        A[np.diag_indices(p)] += rho
        tempinv = arma_inv_sympd_symmatu(A)
        del A
    denom = math.sqrt(p * K / 1.0)
    Cold = C
    L2old = Lambda2
    iters = 0
    with _Timer(trace, "admm_iter"):
        for it in range(maxit):
            iters = it + 1
            Phi = tempinv @ (rho * Cold - L2old)
            C = arma_polar(Phi + L2old / rho)
            diff = Phi - C
            Lambda2 = L2old + rho * diff
            e0 = fro(diff) / denom
            e1 = fro(C - Cold) / denom
            if max(e0, e1) <= tol:
                break
            Cold = C
            L2old = Lambda2
    if trace is not None:
        trace.add_call(tag[0], tag[1], tag[2], iters)
    return Phi, C, Lambda2


def core3(tempinv, Phi, R, C, Lambda1, Lambda2, tau2, rho, maxit, tol, trace=None,
          tag=("core3", -1, -1)):
    """`spatpcaCore3` (RcppSpatPCA.cpp:224-270).  Returns the five updated
    arrays (Phi, R, C, Lambda1, Lambda2)."""
    p, K = Phi.shape
    Rold, Cold, L1old, L2old = R, C, Lambda1, Lambda2
    thr = tau2 / rho
    denom = math.sqrt(float(p) * float(K))
    iters = 0
    with _Timer(trace, "admm_iter"):
        for it in range(maxit):
            iters = it + 1
            Phi = 0.5 * (tempinv @ (rho * (Rold + Cold) - (L1old + L2old)))
            a = L1old / rho + Phi
            R = np.sign(a) * np.maximum(0.0, np.abs(a) - thr)
            C = arma_polar(Phi + L2old / rho)
            dR = Phi - R
            dC = Phi - C
            Lambda1 = L1old + rho * dR
            Lambda2 = L2old + rho * dC
            err = max(fro(dR) / denom, fro(R - Rold) / denom,
                      fro(dC) / denom, fro(C - Cold) / denom)
            if err <= tol:
                break
            Rold, Cold, L1old, L2old = R, C, Lambda1, Lambda2
    if trace is not None:
        trace.add_call(tag[0], tag[1], tag[2], iters)
    return Phi, R, C, Lambda1, Lambda2


def core3_tempinv(Omega, G, tau1, rho, trace=None):
    """`inv_sympd(symmatu(tau1*Omega - G + rho*I))` (RcppSpatPCA.cpp:397, 621, 661, 673)."""
    with _Timer(trace, "inv_sympd"):
        A = tau1 * Omega - G
        A[np.diag_indices(A.shape[0])] += rho
        return arma_inv_sympd_symmatu(A)


def cv_loss(Yva, Phi, trace=None):
    """`pow(norm(Y_validation * (Ip - Phi*Phi.t()), "fro"), 2)` (RcppSpatPCA.cpp:327, 331, 400).
    Evaluated as Yva - (Yva Phi) Phi' -- the same matrix without forming the
    p x p projector (differs from the reference only by rounding)."""
    with _Timer(trace, "cv_loss"):
        Rm = Yva - (Yva @ Phi) @ Phi.T
        return float(np.sum(Rm * Rm))


def pca_init(Ytr, K, sv_for_lambda=None):
    """SVD-right start of a fold or of the full data (RcppSpatPCA.cpp:321-326, 563-567).
    `sv_for_lambda` reproduces the :642 quirk (full-data singular values used
    for a fold's Lambda2)."""
    s, V = arma_svd_right(Ytr)
    rho = 10.0 * s[0] ** 2
    Phi = _F(V[:, :K])
    sv = s if sv_for_lambda is None else sv_for_lambda
    Lambda2 = _F(Phi * (rho - 1.0 / (rho - 2.0 * sv[:K] ** 2))[None, :])
    return s, rho, Phi, Lambda2


def water_filling(dvals, total_variance, gamma, p, K):
    """Noise level `error` of stage 3 / spatialPrediction
    (RcppSpatPCA.cpp:499-519 and :743-760); dvals are the descending
    eigenvalues of Phi' S Phi."""
    total = float(np.sum(dvals))       # accu() of the *unsorted* values; same set
    return _water_filling(dvals, total, total_variance, gamma, p, K)


def _water_filling(d, total, tv, g, p, K):
    tempL = K
    if d[0] > g:
        error = (tv - total + K * g) / (p - tempL)
        temp = d[tempL - 1]
        while temp - g < error:
            if tempL == 1:
                error = (tv - d[0] + g) / (p - 1)
                break
            total -= d[tempL - 1]
            tempL -= 1
            error = (tv - total + tempL * g) / (p - tempL)
            temp = d[tempL - 1]
        if d[0] - g < error:
            error = tv / p
    else:
        error = tv / p
    return error


def gamma_losses(Phi, Ytr, Yva, gammas, trace=None):
    """Stage-3 inner loop for one fold (RcppSpatPCA.cpp:489-523).  The p x p
    matrices are formed explicitly here (small-p oracle); `gamma_losses_lowrank`
    is the algebraically identical form used for large p."""
    p, K = Phi.shape
    n_tr = Ytr.shape[0]
    n_va = Yva.shape[0]
    with _Timer(trace, "gamma_loss"):
        S_tr = (Ytr.T @ Ytr) / n_tr
        S_va = (Yva.T @ Yva) / n_va
        tv = float(np.trace(S_tr))
        w, Q = sla.eigh(Phi.T @ S_tr @ Phi, driver="evd", lower=False, check_finite=False)
        total = arma_accumulate(w)
        order = np.argsort(-w, kind="stable")
        dvals = w[order]
        Qd = Q[:, order]
        out = np.empty(len(gammas))
        B = Phi @ Qd
        for gj, g in enumerate(gammas):
            error = _water_filling(dvals, total, tv, g, p, K)
            lam = np.maximum(dvals - (error + g), 0.0)
            Sig = (B * lam[None, :]) @ B.T
            D = S_va - Sig
            D[np.diag_indices(p)] -= error
            out[gj] = float(np.sum(D * D))
        return out


def gamma_losses_lowrank(Phi, Ytr, Yva, gammas, trace=None, block=2048):
    """Same numbers as `gamma_losses` without p x p temporaries: the residual
    S_va - Sigma_hat - error*I is evaluated tile by tile and its squares summed
    (no algebraic cancellation, only a different summation order)."""
    p, K = Phi.shape
    n_tr = Ytr.shape[0]
    n_va = Yva.shape[0]
    with _Timer(trace, "gamma_loss"):
        tv = float(np.sum(Ytr * Ytr)) / n_tr
        W = Ytr @ Phi
        w, Q = sla.eigh((W.T @ W) / n_tr, driver="evd", lower=False, check_finite=False)
        total = arma_accumulate(w)
        order = np.argsort(-w, kind="stable")
        dvals = w[order]
        B = Phi @ Q[:, order]
        errs = np.array([_water_filling(dvals, total, tv, g, p, K) for g in gammas])
        lams = np.maximum(dvals[None, :] - (errs + np.asarray(gammas))[:, None], 0.0)
        out = np.zeros(len(gammas))
        for i0 in range(0, p, block):
            i1 = min(p, i0 + block)
            for j0 in range(0, p, block):
                j1 = min(p, j0 + block)
                S = (Yva[:, i0:i1].T @ Yva[:, j0:j1]) / n_va
                for gj in range(len(gammas)):
                    D = S - (B[i0:i1] * lams[gj][None, :]) @ B[j0:j1].T
                    if i0 == j0:
                        D[np.diag_indices(i1 - i0)] -= errs[gj]
                    out[gj] += float(np.sum(D * D))
        return out


# --------------------------------------------------------------------------
# spatpcaCV  (RcppSpatPCA.cpp:543-701)
# --------------------------------------------------------------------------
def spatpca_cv(sxy, Y, M, K, tau1, tau2, gamma, nk, maxit, tol, l2, trace=None,
               omega=None, lowrank_gamma=None):
    """`spatpcaCV`.  Arguments as the `.Call` stub (R/RcppExports.R:51-53).
    `omega` optionally injects thinPlateSplineMatrix(sxy) (without the extra
    1e-8 I of :574) so that ADMM parity can be tested separately from the
    ill-conditioned bordered inverse.  Returns the reference's 7-entry list as
    a dict, plus `aux` with per-fold scores."""
    sxy = _F(np.asarray(sxy, dtype=float).reshape(len(sxy), -1))
    Y = _F(Y)
    n, p = Y.shape
    tau1 = np.atleast_1d(np.asarray(tau1, dtype=float))
    tau2 = np.atleast_1d(np.asarray(tau2, dtype=float))
    gamma = np.atleast_1d(np.asarray(gamma, dtype=float))
    nk = np.asarray(nk, dtype=float).ravel()
    l2 = np.atleast_1d(np.asarray(l2, dtype=float))
    if lowrank_gamma is None:
        lowrank_gamma = p > 1500
    gl = gamma_losses_lowrank if lowrank_gamma else gamma_losses

    tr_rows = [np.flatnonzero(nk != (k + 1)) for k in range(M)]
    va_rows = [np.flatnonzero(nk == (k + 1)) for k in range(M)]

    with _Timer(trace, "prologue"):
        G = _F(Y.T @ Y)                                           # :561
        sv_full, rho_hat, Phi_hat, L2_hat = pca_init(Y, K)        # :563-567
        C_hat = Phi_hat.copy(order="F")

    G_k = [None] * M
    Phi_cv = [None] * M
    L2_cv = [None] * M
    tinv_cv = [None] * M
    rho_cv = np.zeros(M)
    Omega = None

    any_tau = (tau1.max() != 0) or (tau2.max() != 0)
    if any_tau:                                                   # :573-575
        if omega is None:
            Omega = thin_plate_spline_matrix(sxy, trace)
        else:
            Omega = _F(omega).copy(order="F")
        omega_raw_diag = Omega[np.diag_indices(p)].copy()
        Omega[np.diag_indices(p)] += 1e-8
    elif len(gamma) > 1:                                          # :576-590
        Omega = np.eye(p, order="F")
        with _Timer(trace, "prologue"):
            for k in range(M):
                Ytr = Y[tr_rows[k]]
                _, V = arma_svd_right(Ytr)
                Phi_cv[k] = _F(V[:, :K])
                G_k[k] = _F(Ytr.T @ Ytr)

    cv1 = np.zeros((M, len(tau1)))
    cv3 = np.zeros((M, len(gamma)))

    # ---------------- stage 1: tau1 path ----------------  :592-650
    if len(tau1) > 1:
        for k in range(M):                                        # :312-334
            Ytr, Yva = Y[tr_rows[k]], Y[va_rows[k]]
            with _Timer(trace, "prologue"):
                _, rho_cv[k], Phi0, L20 = pca_init(Ytr, K)
                G_k[k] = _F(Ytr.T @ Ytr)
            Phi_cv[k], L2_cv[k] = Phi0, L20
            Phi, C, L2 = Phi0, Phi0, L20
            cv1[k, 0] = cv_loss(Yva, Phi0, trace)
            for i in range(1, len(tau1)):
                Phi, C, L2 = core2(G_k[k], Phi, C, L2, Omega, tau1[i], rho_cv[k],
                                   maxit, tol, trace, ("s1.core2", k, i))
                cv1[k, i] = cv_loss(Yva, Phi, trace)
        cv1_sum = arma_colsum(cv1)
        index1 = int(np.argmin(cv1_sum))
        sel_tau1 = float(tau1[index1])
        if index1 > 0:                                            # :598-599
            Phi_hat, C_hat, L2_hat = core2(G, None, C_hat, L2_hat, Omega, sel_tau1, rho_hat,
                                           maxit, tol, trace, ("full.core2p", -1, index1))
        score1 = cv1_sum / M
    else:
        sel_tau1 = float(tau1.max())
        if sel_tau1 != 0 and tau2.max() == 0:                     # :604-623
            Phi_hat, C_hat, L2_hat = core2(G, None, C_hat, L2_hat, Omega, sel_tau1, rho_hat,
                                           maxit, tol, trace, ("full.core2p", -1, 0))
            for k in range(M):
                Ytr = Y[tr_rows[k]]
                with _Timer(trace, "prologue"):
                    G_k[k] = _F(Ytr.T @ Ytr)
                    _, rho_cv[k], Phi0, L20 = pca_init(Ytr, K)
                Pg, Cg, Lg = core2(G_k[k], Phi0, Phi0, L20, Omega, sel_tau1, rho_cv[k],
                                   maxit, tol, trace, ("s1x.core2", k, 0))
                Phi_cv[k], L2_cv[k] = Pg, Lg
                tinv_cv[k] = core3_tempinv(Omega, G_k[k], sel_tau1, rho_cv[k], trace)
        elif sel_tau1 == 0 and tau2.max() == 0:                   # :624-629
            pass                                                  # Phi_hat already the PCA solution
        else:                                                     # :630-648
            for k in range(M):
                Ytr = Y[tr_rows[k]]
                with _Timer(trace, "prologue"):
                    G_k[k] = _F(Ytr.T @ Ytr)
                    # quirk :642 -- Lambda2 uses the FULL-data singular values
                    _, rho_cv[k], Phi0, L20 = pca_init(Ytr, K, sv_for_lambda=sv_full)
                Pg, Cg, Lg = Phi0, Phi0, L20
                if sel_tau1 != 0:
                    Pg, Cg, Lg = core2(G_k[k], Pg, Cg, Lg, Omega, sel_tau1, rho_cv[k],
                                       maxit, tol, trace, ("s1x.core2", k, 0))
                Phi_cv[k], L2_cv[k] = Pg, Lg
        score1 = np.zeros(1)

    # ---------------- stage 2: tau2 path ----------------  :652-680
    index2 = 0
    cv2 = None
    if len(tau2) > 1:
        cv2 = np.zeros((M, len(tau2)))
        for k in range(M):                                        # :384-403
            Yva = Y[va_rows[k]]
            Phi = C = Phi_cv[k]
            L1 = np.zeros_like(Phi)
            L2 = L2_cv[k]
            if sel_tau1 != 0:
                Phi, C, L2 = core2(G_k[k], Phi, C, L2, Omega, sel_tau1, rho_cv[k],
                                   maxit, tol, trace, ("s2.core2", k, 0))
            R = Phi
            Phi_cv[k], L2_cv[k] = Phi, L2
            tinv_cv[k] = core3_tempinv(Omega, G_k[k], sel_tau1, rho_cv[k], trace)
            for i in range(len(tau2)):
                Phi, R, C, L1, L2 = core3(tinv_cv[k], Phi, R, C, L1, L2, tau2[i], rho_cv[k],
                                          maxit, tol, trace, ("s2.core3", k, i))
                cv2[k, i] = cv_loss(Yva, Phi, trace)
        cv2_sum = arma_colsum(cv2)
        index2 = int(np.argmin(cv2_sum))
        sel_tau2 = float(tau2[index2])
        tinv = core3_tempinv(Omega, G, sel_tau1, rho_hat, trace)  # :661
        R_hat = Phi_hat
        L1_hat = np.zeros_like(Phi_hat)
        for i in range(index2 + 1):                               # :665-666
            Phi_hat, R_hat, C_hat, L1_hat, L2_hat = core3(
                tinv, Phi_hat, R_hat, C_hat, L1_hat, L2_hat, tau2[i], rho_hat,
                maxit, tol, trace, ("full.core3", -1, i))
        del tinv
        score2 = cv2_sum / M
    else:
        sel_tau2 = float(tau2.max())
        if sel_tau2 > 0:                                          # :672-678
            tinv = core3_tempinv(Omega, G, sel_tau1, rho_hat, trace)
            R_hat = Phi_hat
            L1_hat = np.zeros_like(Phi_hat)
            for i in range(len(l2)):
                Phi_hat, R_hat, C_hat, L1_hat, L2_hat = core3(
                    tinv, Phi_hat, R_hat, C_hat, L1_hat, L2_hat, l2[i], rho_hat,
                    maxit, tol, trace, ("full.core3", -1, i))
            del tinv
        score2 = np.zeros(1)

    # ---------------- stage 3: gamma ----------------  :682-692, 459-525
    for k in range(M):
        Ytr, Yva = Y[tr_rows[k]], Y[va_rows[k]]
        if Phi_cv[k] is None:
            # reference reads an uninitialised cube here (:569, 576-590 not taken);
            # the oracle defines it as the fold's PCA solution.
            _, V = arma_svd_right(Ytr)
            Phi_cv[k] = _F(V[:, :K])
        Phi = C = Phi_cv[k]
        L2 = L2_cv[k]
        if tau2.max() != 0 or sel_tau1 != 0:                      # :473
            if len(tau2) == 1:
                Phi, C, L2 = core2(G_k[k], Phi, C, L2, Omega, sel_tau1, rho_cv[k],
                                   maxit, tol, trace, ("s3.core2", k, 0))
            else:
                R = Phi
                L1 = np.zeros_like(Phi)
                for i in range(index2 + 1):
                    Phi, R, C, L1, L2 = core3(tinv_cv[k], Phi, R, C, L1, L2, tau2[i], rho_cv[k],
                                              maxit, tol, trace, ("s3.core3", k, i))
        cv3[k, :] = gl(Phi, Ytr, Yva, gamma, trace)
    cv3_sum = arma_colsum(cv3)
    if len(gamma) > 1:
        index3 = int(np.argmin(cv3_sum))
        sel_gamma = float(gamma[index3])
    else:
        sel_gamma = float(gamma.max())
    score3 = cv3_sum / M

    return {
        "cv_score_tau1": score1,
        "cv_score_tau2": score2,
        "cv_score_gamma": score3,
        "estimated_eigenfn": _F(Phi_hat),
        "selected_tau1": sel_tau1,
        "selected_tau2": sel_tau2,
        "selected_gamma": sel_gamma,
        "aux": {"cv1": cv1, "cv2": cv2, "cv3": cv3, "rho_cv": rho_cv, "rho": rho_hat,
                "omega": _omega_without_ridge(Omega, omega_raw_diag if any_tau else None)},
    }


def _omega_without_ridge(Omega, raw_diag):
    """thinPlateSplineMatrix(sxy) as this run used it (the +1e-8 I of :574 undone), for injected-Omega parity runs."""
    if Omega is None or raw_diag is None:
        return None
    out = Omega.copy(order="F")
    out[np.diag_indices(out.shape[0])] = raw_diag
    return out


# --------------------------------------------------------------------------
# spatialPrediction  (RcppSpatPCA.cpp:715-770)
# --------------------------------------------------------------------------
def spatial_prediction(phi, Y, gamma, predicted_phi):
    phi = _F(phi)
    Y = _F(Y)
    predicted_phi = _F(predicted_phi)
    n = Y.shape[0]
    p, K = phi.shape
    cov = (Y.T @ Y) / n
    w, Q = sla.eigh(phi.T @ cov @ phi, driver="evd", lower=False, check_finite=False)
    tv = float(np.trace(cov))
    total = arma_accumulate(w)
    order = np.argsort(-w, kind="stable")
    dvals = w[order]
    Qd = Q[:, order]
    error = _water_filling(dvals, total, tv, gamma, p, K)
    eigenvalue = np.maximum(dvals - (error + gamma), 0.0)
    eigenvalue2 = eigenvalue + error
    est_cov = (phi @ Qd) @ np.diag(eigenvalue / eigenvalue2) @ (predicted_phi @ Qd).T
    return {
        "prediction": _F(Y @ est_cov),
        "estimated_covariance": _F(est_cov),
        "eigenvalue": eigenvalue,
        "error": error,
    }
