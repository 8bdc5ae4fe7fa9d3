"""The four `.Call` entry points of the reference (R/RcppExports.R:14-68) over the
C ABI, with the reference's names and argument meaning, plus the engine handle
used to place folds on different GPUs.  Pure marshalling: all computing happens
in libspatpca_b200.so on the GPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import CvStats, check, fmat, fvec, ptr


def thinPlateSplineMatrix(location) -> np.ndarray:
    """`thinPlateSplineMatrix(location)` -- src/RcppSpatPCA.cpp:58-76."""
    lib = _lib.load()
    x = fmat(location)
    p, d = x.shape
    out = np.empty((p, p), order="F")
    check(lib.spb_tps_matrix(ptr(x), p, d, ptr(out)))
    return out


def eigenFunction(new_location, original_location, Phi) -> np.ndarray:
    """`eigenFunction(new_location, original_location, Phi)` -- src/RcppSpatPCA.cpp:94-147."""
    lib = _lib.load()
    xo = fmat(original_location)
    xn = fmat(new_location, ncol=xo.shape[1])
    phi = fmat(Phi)
    if xn.shape[1] != xo.shape[1]:
        raise ValueError("Inconsistent dimension of locations")
    if phi.shape[0] != xo.shape[0]:
        raise ValueError("Phi must have one row per original location")
    out = np.empty((xn.shape[0], phi.shape[1]), order="F")
    check(lib.spb_eigen_function(ptr(xn), xn.shape[0], ptr(xo), xo.shape[0], xo.shape[1], ptr(phi),
                                 phi.shape[1], ptr(out)))
    return out


def spatialPrediction(phir, Yr, gamma, predicted_eignefunction) -> dict:
    """`spatialPrediction(phir, Yr, gamma, predicted_eignefunction)` -- src/RcppSpatPCA.cpp:715-770."""
    lib = _lib.load()
    phi = fmat(phir)
    Y = fmat(Yr)
    pp = fmat(predicted_eignefunction, ncol=phi.shape[1])
    p, K = phi.shape
    n = Y.shape[0]
    if Y.shape[1] != p or pp.shape[1] != K:
        raise ValueError("inconsistent shapes")
    p2 = pp.shape[0]
    pred = np.empty((n, p2), order="F")
    cov = np.empty((p, p2), order="F")
    ev = np.empty(K)
    err = np.empty(1)
    check(lib.spb_spatial_prediction(ptr(phi), p, K, ptr(Y), n, float(gamma), ptr(pp), p2, ptr(pred),
                                     ptr(cov), ptr(ev), ptr(err)))
    return {"prediction": pred, "estimated_covariance": cov, "eigenvalue": ev, "error": float(err[0])}


def spatpcaCV(sxyr, Yr, M, K, tau1r, tau2r, gammar, nkr, maxit, tol, l2r, return_stats=False,
              session=False) -> dict:
    """`spatpcaCV(...)` -- src/RcppSpatPCA.cpp:543-701; returns the reference's 7-entry list.
    `session=True` goes through `spb_cv_session`: consecutive calls on the same data and grids
    (the K-selection loop) reuse Omega, the SVD starts and the K-independent inverses."""
    lib = _lib.load()
    Y = fmat(Yr)
    n, p = Y.shape
    x = fmat(sxyr)
    if x.shape[0] != p:
        raise ValueError("The number of rows of x should be equal to the number of columns of Y.")
    t1, t2, g, nk, l2 = fvec(tau1r), fvec(tau2r), fvec(gammar), fvec(nkr), fvec(l2r)
    if len(nk) != n:
        raise ValueError("nk must have one fold id per row of Y")
    s1 = np.zeros(max(len(t1), 1))
    s2 = np.zeros(max(len(t2), 1))
    s3 = np.zeros(len(g))
    eig = np.empty((p, int(K)), order="F")
    sel = np.zeros(3)
    stats = CvStats()
    entry = lib.spb_cv_session if session else lib.spb_cv
    check(entry(ptr(x), p, x.shape[1], ptr(Y), n, int(M), int(K), ptr(t1), len(t1), ptr(t2), len(t2),
                ptr(g), len(g), ptr(nk), int(maxit), float(tol), ptr(l2), len(l2),
                ptr(s1), ptr(s2), ptr(s3), ptr(eig), ptr(sel), C.byref(stats)))
    out = {
        "cv_score_tau1": s1 if len(t1) > 1 else np.zeros(1),
        "cv_score_tau2": s2 if len(t2) > 1 else np.zeros(1),
        "cv_score_gamma": s3,
        "estimated_eigenfn": eig,
        "selected_tau1": float(sel[0]),
        "selected_tau2": float(sel[1]),
        "selected_gamma": float(sel[2]),
    }
    if return_stats:
        out["stats"] = stats.as_dict()
    return out


def session_clear():
    check(_lib.load().spb_session_clear())


def select_index(cv: np.ndarray):
    """`index_min(sum(cv, 0))`, `sum(cv, 0) / M` with Armadillo's accumulation order."""
    lib = _lib.load()
    cv = fmat(cv)
    M, ng = cv.shape
    score = np.empty(ng)
    idx = check(lib.spb_select_index(ptr(cv), M, ng, ptr(score)))
    return idx, score


class Engine:
    """Handle API (include/spatpca_b200.h): one engine per GPU / process; folds are the
    independent units (src/RcppSpatPCA.cpp:315, 385, 460)."""

    def __init__(self, p, d, n, M, K, tau1, tau2, gamma, maxit, tol, l2):
        self._lib = _lib.load()
        self.p, self.d, self.n, self.M, self.K = int(p), int(d), int(n), int(M), int(K)
        self.tau1, self.tau2, self.gamma, self.l2 = fvec(tau1), fvec(tau2), fvec(gamma), fvec(l2)
        self._h = C.c_void_p()
        check(self._lib.spb_engine_create(C.byref(self._h), self.p, self.d, self.n, self.M, self.K,
                                          ptr(self.tau1), len(self.tau1), ptr(self.tau2), len(self.tau2),
                                          ptr(self.gamma), len(self.gamma), int(maxit), float(tol),
                                          ptr(self.l2), len(self.l2)))

    def close(self):
        if self._h:
            self._lib.spb_engine_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, sxy, Y, nk):
        x, Yf, nkf = fmat(sxy), fmat(Y), fvec(nk)
        assert Yf.shape == (self.n, self.p) and x.shape == (self.p, self.d) and len(nkf) == self.n
        check(self._lib.spb_engine_upload(self._h, ptr(x), ptr(Yf), ptr(nkf)))

    def set_omega(self, omega):
        om = fmat(omega)
        assert om.shape == (self.p, self.p)
        check(self._lib.spb_engine_set_omega(self._h, ptr(om)))

    def prepare(self):
        check(self._lib.spb_engine_prepare(self._h))

    def get_omega(self):
        out = np.empty((self.p, self.p), order="F")
        check(self._lib.spb_engine_get_omega(self._h, ptr(out)))
        return out

    def stage1_fold(self, k):
        row = np.zeros(max(len(self.tau1), 1))
        check(self._lib.spb_engine_stage1_fold(self._h, int(k), ptr(row)))
        return row

    def stage1_full(self, index1):
        check(self._lib.spb_engine_stage1_full(self._h, int(index1)))

    def stage2_fold(self, k, index1):
        row = np.zeros(max(len(self.tau2), 1))
        check(self._lib.spb_engine_stage2_fold(self._h, int(k), int(index1), ptr(row)))
        return row

    def stage2_full(self, index1, index2):
        check(self._lib.spb_engine_stage2_full(self._h, int(index1), int(index2)))

    def stage3_fold(self, k, index1, index2):
        row = np.zeros(len(self.gamma))
        check(self._lib.spb_engine_stage3_fold(self._h, int(k), int(index1), int(index2), ptr(row)))
        return row

    # ---- the same stages for a list of folds (run concurrently on the GPU's lanes) ----
    def _folds(self, folds):
        return np.ascontiguousarray(np.asarray(list(folds), dtype=np.int32))

    def stage1_folds(self, folds):
        f = self._folds(folds)
        rows = np.zeros((len(f), max(len(self.tau1), 1)))
        check(self._lib.spb_engine_stage1_folds(self._h, f.ctypes.data_as(C.POINTER(C.c_int)), len(f), ptr(rows)))
        return {int(k): rows[i] for i, k in enumerate(f)}

    def stage2_folds(self, folds, index1):
        f = self._folds(folds)
        rows = np.zeros((len(f), max(len(self.tau2), 1)))
        check(self._lib.spb_engine_stage2_folds(self._h, f.ctypes.data_as(C.POINTER(C.c_int)), len(f), int(index1),
                                                ptr(rows)))
        return {int(k): rows[i] for i, k in enumerate(f)}

    def stage3_folds(self, folds, index1, index2):
        f = self._folds(folds)
        rows = np.zeros((len(f), len(self.gamma)))
        check(self._lib.spb_engine_stage3_folds(self._h, f.ctypes.data_as(C.POINTER(C.c_int)), len(f), int(index1),
                                                int(index2), ptr(rows)))
        return {int(k): rows[i] for i, k in enumerate(f)}

    # ---- (fold, tau1 index) inverses as jobs that any GPU can build (include/spatpca_b200.h) ----
    def can_share_inverses(self) -> bool:
        return check(self._lib.spb_engine_can_share_inverses(self._h)) == 1

    def prepare_fold(self, k):
        check(self._lib.spb_engine_prepare_fold(self._h, int(k)))

    def prepare_starts(self, folds):
        """SVD starts of the full data and of `folds` ahead of prepare() (they then hide under the Omega build)."""
        f = self._folds(folds)
        check(self._lib.spb_engine_prepare_starts(self._h, f.ctypes.data_as(C.POINTER(C.c_int)), len(f)))

    def inverse_buffer(self, k, i):
        """(device pointer, bytes) of the buffer that holds / will hold inverse (fold k, tau1[i])."""
        p = C.c_void_p()
        nb = C.c_longlong()
        check(self._lib.spb_engine_inverse_buffer(self._h, int(k), int(i), C.byref(p), C.byref(nb)))
        return int(p.value or 0), int(nb.value)

    def build_inverse(self, k, i):
        check(self._lib.spb_engine_build_inverse(self._h, int(k), int(i)))

    # ---- Omega in column blocks (shared between the ranks of a multi-GPU job) ----
    def omega_begin(self):
        needed = C.c_int(0)
        check(self._lib.spb_engine_omega_begin(self._h, C.byref(needed)))
        return bool(needed.value)

    def omega_solve_cols(self, c0, c1):
        check(self._lib.spb_engine_omega_solve_cols(self._h, int(c0), int(c1)))

    def omega_refine_cols(self, c0, c1) -> bool:
        active = C.c_int(0)
        check(self._lib.spb_engine_omega_refine_cols(self._h, int(c0), int(c1), C.byref(active)))
        return bool(active.value)

    def omega_product_cols(self, c0, c1):
        check(self._lib.spb_engine_omega_product_cols(self._h, int(c0), int(c1)))

    def omega_buffer(self, which, c0, c1):
        p = C.c_void_p()
        nb = C.c_longlong()
        check(self._lib.spb_engine_omega_buffer(self._h, int(which), int(c0), int(c1), C.byref(p), C.byref(nb)))
        return int(p.value or 0), int(nb.value)

    def omega_sync(self):
        check(self._lib.spb_engine_omega_sync(self._h))

    def omega_finish(self):
        check(self._lib.spb_engine_omega_finish(self._h))

    def mark_inverse(self, k, i):
        check(self._lib.spb_engine_mark_inverse(self._h, int(k), int(i)))

    def sync(self):
        check(self._lib.spb_engine_sync(self._h))

    def reset_k(self, K):
        check(self._lib.spb_engine_reset_k(self._h, int(K)))
        self.K = int(K)

    def get_eigenfn(self):
        out = np.empty((self.p, self.K), order="F")
        check(self._lib.spb_engine_get_eigenfn(self._h, ptr(out)))
        return out

    def stats(self) -> dict:
        st = CvStats()
        check(self._lib.spb_engine_get_stats(self._h, C.byref(st)))
        return st.as_dict()

    def call_log(self):
        n = check(self._lib.spb_engine_call_log(self._h, None, 0))
        buf = np.zeros((max(n, 1), 4), dtype=np.int32)
        check(self._lib.spb_engine_call_log(self._h, buf.ctypes.data_as(C.POINTER(C.c_int)), n))
        return [tuple(int(v) for v in buf[i]) for i in range(n)]

    def release_fold(self, k):
        check(self._lib.spb_engine_release_fold(self._h, int(k)))


# ---- kernel-level entry points (parity tests, benchmarks) ----------------------------------
def inv_sympd(omega, G, tau1, rho, scale):
    lib = _lib.load()
    om, g = fmat(omega), fmat(G)
    p = om.shape[0]
    out = np.empty((p, p), order="F")
    check(lib.spb_inv_sympd(ptr(om), ptr(g), p, float(tau1), float(rho), float(scale), ptr(out)))
    return out


def admm_core2(tempinv, C_, Lambda2, rho, maxit, tol):
    lib = _lib.load()
    A = fmat(tempinv)
    Cm, L2 = fmat(C_).copy(order="F"), fmat(Lambda2).copy(order="F")
    p, K = Cm.shape
    Phi = np.zeros((p, K), order="F")
    it = C.c_int(0)
    check(lib.spb_admm_core2(ptr(A), p, K, float(rho), int(maxit), float(tol), ptr(Phi), ptr(Cm), ptr(L2),
                             C.byref(it)))
    return Phi, Cm, L2, it.value


def admm_core3(tempinv, R, C_, Lambda1, Lambda2, tau2, rho, maxit, tol):
    lib = _lib.load()
    A = fmat(tempinv)
    Rm, Cm = fmat(R).copy(order="F"), fmat(C_).copy(order="F")
    L1, L2 = fmat(Lambda1).copy(order="F"), fmat(Lambda2).copy(order="F")
    p, K = Cm.shape
    Phi = np.zeros((p, K), order="F")
    it = C.c_int(0)
    check(lib.spb_admm_core3(ptr(A), p, K, float(tau2), float(rho), int(maxit), float(tol), ptr(Phi), ptr(Rm),
                             ptr(Cm), ptr(L1), ptr(L2), C.byref(it)))
    return Phi, Rm, Cm, L1, L2, it.value


def admm_system(omega, G, tau1, rho, scale, blocks, mode, C_, Lambda2, maxit, tol, R=None, Lambda1=None, tau2=0.0):
    """spatpcaCore2 (mode 2) / spatpcaCore3 (mode 3) with the system built and inverted -- or block-factored
    (`blocks` > 1; 0 = the library's default for p) -- on the device (spb_admm_system)."""
    lib = _lib.load()
    om, g = fmat(omega), fmat(G)
    Cm, L2 = fmat(C_).copy(order="F"), fmat(Lambda2).copy(order="F")
    p, K = Cm.shape
    Rm = fmat(R).copy(order="F") if R is not None else np.zeros((p, K), order="F")
    L1 = fmat(Lambda1).copy(order="F") if Lambda1 is not None else np.zeros((p, K), order="F")
    Phi = np.zeros((p, K), order="F")
    it = C.c_int(0)
    check(lib.spb_admm_system(ptr(om), ptr(g), p, float(tau1), float(rho), float(scale), int(blocks), int(mode), K,
                              float(tau2), int(maxit), float(tol), ptr(Phi), ptr(Rm), ptr(Cm), ptr(L1), ptr(L2),
                              C.byref(it)))
    return Phi, Rm, Cm, L1, L2, it.value


def core2_matrix_free(omega, Ytr, tau1, rho, Phi0, C_, Lambda2, maxit, tol, cg_tol=0.0):
    """spatpcaCore2 without inv_sympd (spb_core2_matrix_free): returns (Phi, C, Lambda2, iterations, matrix passes)."""
    lib = _lib.load()
    om, yt = fmat(omega), fmat(Ytr)
    Ph, Cm, L2 = fmat(Phi0).copy(order="F"), fmat(C_).copy(order="F"), fmat(Lambda2).copy(order="F")
    p, K = Cm.shape
    it, ps = C.c_int(0), C.c_int(0)
    check(lib.spb_core2_matrix_free(ptr(om), ptr(yt), yt.shape[0], p, K, float(tau1), float(rho), int(maxit),
                                    float(tol), float(cg_tol), ptr(Ph), ptr(Cm), ptr(L2), C.byref(it), C.byref(ps)))
    return Ph, Cm, L2, it.value, ps.value


def bench_core2_cg(p, K, n_tr, iters, repeats=3):
    lib = _lib.load()
    out = np.zeros(1)
    ps = C.c_int(0)
    check(lib.spb_bench_core2_cg(int(p), int(K), int(n_tr), int(iters), int(repeats), ptr(out), C.byref(ps)))
    return float(out[0]), ps.value


def bench_core2_batched(p, n, K, F, iters, repeats=3):
    lib = _lib.load()
    out = np.zeros(1)
    ps = C.c_int(0)
    check(lib.spb_bench_core2_batched(int(p), int(n), int(K), int(F), int(iters), int(repeats), ptr(out), C.byref(ps)))
    return float(out[0]), ps.value


def dmma_matvec(A, V):
    """U = A V on the TMA-fed DMMA stream (A symmetric p x p, V p x nc, nc <= 32)."""
    lib = _lib.load()
    Am, Vm = fmat(A), fmat(V)
    p, nc = Vm.shape
    U = np.empty((p, nc), order="F")
    check(lib.spb_dmma_matvec(ptr(Am), p, ptr(Vm), nc, ptr(U)))
    return U


def bench_dmma_matvec(p, nc, repeats=5):
    lib = _lib.load()
    out = np.zeros(1)
    check(lib.spb_bench_dmma_matvec(int(p), int(nc), int(repeats), ptr(out)))
    return float(out[0])


def core2_batched_supported(p, n, K, F) -> bool:
    return check(_lib.load().spb_core2_batched_supported(int(p), int(n), int(K), int(F))) == 1


def core2_batched(omega, Y, nk, fold_ids, rho, tau1, Phi0, C0, L20, maxit, tol, cg_tol=0.0, snapshots=True):
    """F matrix-free spatpcaCore2 problems in lock-step along a tau1 path (spb_core2_batched).
    Phi0 / C0 / L20: lists of F (p x K) arrays.  Returns dict(Phi, C, L2: lists; iters: nsteps x F; snap; passes)."""
    lib = _lib.load()
    om, Ym, nkf = fmat(omega), fmat(Y), fvec(nk)
    n, p = Ym.shape
    F = len(Phi0)
    K = np.asarray(Phi0[0]).shape[1]
    pack = lambda xs: np.ascontiguousarray(np.concatenate([np.asfortranarray(x).ravel(order="F") for x in xs]))
    Ph, Cm, L2 = pack(Phi0), pack(C0), pack(L20)
    t1 = fvec(tau1)
    ns = len(t1)
    fid = np.ascontiguousarray(np.asarray(fold_ids, dtype=np.int32))
    rh = fvec(rho)
    it = np.zeros(ns * F, dtype=np.int32)
    snap = np.zeros(ns * F * p * K) if snapshots else None
    ps = C.c_int(0)
    check(lib.spb_core2_batched(ptr(om), ptr(Ym), ptr(nkf), p, n, K, F, fid.ctypes.data_as(C.POINTER(C.c_int)), ptr(rh),
                                ptr(t1), ns, int(maxit), float(tol), float(cg_tol), ptr(Ph), ptr(Cm), ptr(L2),
                                it.ctypes.data_as(C.POINTER(C.c_int)), ptr(snap) if snapshots else None, C.byref(ps)))
    unpack = lambda v: [v[f * p * K:(f + 1) * p * K].reshape((p, K), order="F") for f in range(F)]
    out = dict(Phi=unpack(Ph), C=unpack(Cm), L2=unpack(L2), iters=it.reshape(ns, F), passes=ps.value)
    if snapshots:
        out["snap"] = snap.reshape(ns, F, K, p).transpose(0, 1, 3, 2)
    return out


def core3_batched(omega, Y, nk, fold_ids, rho, tau1, tau2, Phi0, R0, C0, L10, L20, maxit, tol, cg_tol=0.0, snapshots=True):
    """F matrix-free spatpcaCore3 problems in lock-step along a tau2 path at one tau1 (spb_core3_batched).
    Returns dict(Phi, R, C, L1, L2: lists of F (p x K) arrays; iters: nsteps x F; snap; passes)."""
    lib = _lib.load()
    om, Ym, nkf = fmat(omega), fmat(Y), fvec(nk)
    n, p = Ym.shape
    F = len(Phi0)
    K = np.asarray(Phi0[0]).shape[1]
    pack = lambda xs: np.ascontiguousarray(np.concatenate([np.asfortranarray(x).ravel(order="F") for x in xs]))
    Ph, Rm, Cm, L1, L2 = pack(Phi0), pack(R0), pack(C0), pack(L10), pack(L20)
    t2 = fvec(tau2)
    ns = len(t2)
    fid = np.ascontiguousarray(np.asarray(fold_ids, dtype=np.int32))
    rh = fvec(rho)
    it = np.zeros(ns * F, dtype=np.int32)
    snap = np.zeros(ns * F * p * K) if snapshots else None
    ps = C.c_int(0)
    check(lib.spb_core3_batched(ptr(om), ptr(Ym), ptr(nkf), p, n, K, F, fid.ctypes.data_as(C.POINTER(C.c_int)), ptr(rh),
                                float(tau1), ptr(t2), ns, int(maxit), float(tol), float(cg_tol), ptr(Ph), ptr(Rm), ptr(Cm),
                                ptr(L1), ptr(L2), it.ctypes.data_as(C.POINTER(C.c_int)), ptr(snap) if snapshots else None,
                                C.byref(ps)))
    unpack = lambda v: [v[f * p * K:(f + 1) * p * K].reshape((p, K), order="F") for f in range(F)]
    out = dict(Phi=unpack(Ph), R=unpack(Rm), C=unpack(Cm), L1=unpack(L1), L2=unpack(L2), iters=it.reshape(ns, F),
               passes=ps.value)
    if snapshots:
        out["snap"] = snap.reshape(ns, F, K, p).transpose(0, 1, 3, 2)
    return out


def default_blocks(p):
    return check(_lib.load().spb_default_blocks(int(p)))


def default_path_blocks(p):
    return check(_lib.load().spb_default_path_blocks(int(p)))


def block_offsets(p, blocks=0):
    """Row offsets off[0..m] of the block partition the library uses for a system of order p (host-only)."""
    off = np.zeros(64, dtype=np.int32)
    m = check(_lib.load().spb_block_offsets(int(p), int(blocks), off.ctypes.data_as(C.POINTER(C.c_int)), 64))
    return [int(v) for v in off[:m + 1]]


def admm_phase_plan(p, K, blocks=0):
    """The ADMM kernel's stream phases for one iteration: dicts with kind / rows / cols / final rows (host-only)."""
    buf = np.zeros(8 * 64, dtype=np.int32)
    n = check(_lib.load().spb_admm_phase_plan(int(p), int(K), int(blocks), buf.ctypes.data_as(C.POINTER(C.c_int)), 64))
    keys = ("kind", "row0", "nrows", "col0", "ncols", "fin0", "fin1", "segments")
    return [dict(zip(keys, (int(v) for v in buf[8 * f:8 * f + 8]))) for f in range(n)]


def svd_right(Y, K):
    """Singular values and the K leading right singular vectors of Y (m x p) from the device SVD start
    (src/RcppSpatPCA.cpp:321, :563; the quantity setGamma needs, R/helper.R:207-217)."""
    lib = _lib.load()
    Ym = fmat(Y)
    m, p = Ym.shape
    sv = np.zeros(min(m, p))
    V = np.zeros((p, int(K)), order="F")
    check(lib.spb_svd_right(ptr(Ym), m, p, int(K), ptr(sv), ptr(V)))
    return sv, V


def cv_loss(Yva, Phi):
    lib = _lib.load()
    Y, P = fmat(Yva), fmat(Phi)
    out = np.zeros(1)
    check(lib.spb_cv_loss(ptr(Y), Y.shape[0], Y.shape[1], ptr(P), P.shape[1], ptr(out)))
    return float(out[0])


def gamma_loss(Phi, Ytr, Yva, gamma):
    lib = _lib.load()
    P, tr, va, g = fmat(Phi), fmat(Ytr), fmat(Yva), fvec(gamma)
    out = np.zeros(len(g))
    check(lib.spb_gamma_loss(ptr(P), P.shape[0], P.shape[1], ptr(tr), tr.shape[0], ptr(va), va.shape[0],
                             ptr(g), len(g), ptr(out)))
    return out


def bench_admm(p, K, mode, iters, repeats=3):
    lib = _lib.load()
    out = np.zeros(1)
    check(lib.spb_bench_admm(int(p), int(K), int(mode), int(iters), int(repeats), ptr(out)))
    return float(out[0])


def bench_admm_blocks(p, K, mode, iters, repeats=3, blocks=0):
    lib = _lib.load()
    out = np.zeros(1)
    check(lib.spb_bench_admm_blocks(int(p), int(K), int(mode), int(iters), int(repeats), int(blocks), ptr(out)))
    return float(out[0])


def bench_factor(p, repeats=2, blocks=0):
    lib = _lib.load()
    out = np.zeros(1)
    check(lib.spb_bench_factor(int(p), int(repeats), int(blocks), ptr(out)))
    return float(out[0])


def bench_dgemm(m, repeats=3):
    lib = _lib.load()
    out = np.zeros(1)
    check(lib.spb_bench_dgemm(int(m), int(repeats), ptr(out)))
    return float(out[0])


def bench_inverse(p, repeats=2):
    lib = _lib.load()
    out = np.zeros(1)
    check(lib.spb_bench_inverse(int(p), int(repeats), ptr(out)))
    return float(out[0])
