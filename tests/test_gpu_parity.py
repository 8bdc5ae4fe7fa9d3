"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same
seeded inputs.  Tolerances are the ones BASELINE.json's north_star states:
  eigenfunctions up to column sign, max-abs error <= 1e-8 relative;
  CV losses <= 1e-9 relative; identical selections; identical per-call iteration counts.
"""
import numpy as np
import pytest

from oracle import r_api
from oracle import spatpca_ref as ref
from oracle.r_rng import RRng, r_seq_len

from ref_cases import case_1d, case_3d, tps_locations
from synth import make_case

pytestmark = pytest.mark.gpu

EIG_RTOL = 1e-8
CV_RTOL = 1e-9


@pytest.fixture(scope="module")
def sp():
    import spatpca_b200 as sp
    from spatpca_b200 import _lib
    _lib.check(_lib.load().spb_set_device(0))
    return sp


def eig_err(A, B):
    sgn = np.sign(np.sum(A * B, axis=0))
    sgn[sgn == 0] = 1
    return np.max(np.abs(A * sgn - B)) / np.max(np.abs(B))


def assert_cv_parity(got, want, eig_rtol=EIG_RTOL, cv_rtol=CV_RTOL):
    for key in ("selected_tau1", "selected_tau2", "selected_gamma"):
        assert got[key] == want[key], (key, got[key], want[key])
    for key in ("cv_score_tau1", "cv_score_tau2", "cv_score_gamma"):
        np.testing.assert_allclose(got[key], want[key], rtol=cv_rtol, atol=0, err_msg=key)
    assert eig_err(got["estimated_eigenfn"], want["estimated_eigenfn"]) <= eig_rtol


# ---------------------------------------------------------------------------------------------
# reference goldens through the GPU path (tests/testthat/*.R)
# ---------------------------------------------------------------------------------------------
def test_tps_matrix_goldens(sp):                       # test-RcppExports.R:10-15
    l2d, l3d = tps_locations()
    assert abs(np.linalg.norm(sp.thinPlateSplineMatrix(l2d)) - 0.362588) <= 1e-6
    assert abs(np.linalg.norm(sp.thinPlateSplineMatrix(l3d)) - 8.191034) <= 1e-6


def test_eigen_function_golden(sp):                    # test-RcppExports.R:18-23
    l2d, _ = tps_locations()
    v = sp.eigenFunction(np.array([[0.1, 0.2]]), l2d, np.array([[1.0], [0], [0], [0]]))
    assert abs(v[0, 0] - 0.2352884) <= 1e-6


@pytest.fixture(scope="module")
def seq_1d(sp):
    """test-SpatPCA.R:13-88 on one shared R-RNG stream, through the GPU engine."""
    rng, x, Y = case_1d()
    out = {"x": x, "Y": Y}
    out["cv_1D"] = sp.spatpca(x, Y, rng=rng, num_cores=1)
    out["multi_tau2"] = sp.spatpca(x, Y, K=1, tau2=[0, 1], rng=rng)
    out["multi_gamma"] = sp.spatpca(x, Y, K=1, gamma=[0, 1], rng=rng)
    out["fixed_both"] = sp.spatpca(x, Y, K=1, tau1=10, tau2=100, rng=rng)
    out["zero_multi_gamma"] = sp.spatpca(x, Y, K=1, tau1=0, tau2=0, gamma=[0, 0.5, 1], rng=rng)
    out["fixed_tau1"] = sp.spatpca(x, Y, K=1, tau1=10, rng=rng)
    out["zero_zero_K5"] = sp.spatpca(x, Y, K=5, tau1=0, tau2=0, rng=rng)
    return out


def test_selected_tuning_parameters(seq_1d):           # test-SpatPCA.R:97-122
    s = seq_1d
    tol = 1e-6
    assert abs(s["cv_1D"].selected_tau1 - 0.00046416) <= tol
    assert abs(s["cv_1D"].selected_tau2) <= tol
    assert abs(s["cv_1D"].selected_gamma - 0.44503397) <= 1.0001 * tol
    assert s["cv_1D"].selected_K is None
    assert s["cv_1D"].eigenfn.shape == (10, 2)
    assert s["multi_tau2"].selected_K == 1
    assert abs(s["multi_tau2"].selected_tau1 - 0.002154435) <= tol
    assert s["multi_tau2"].selected_tau2 == 1
    assert s["multi_gamma"].selected_gamma == 0
    assert s["fixed_both"].selected_tau1 == 10 and s["fixed_both"].selected_tau2 == 100
    assert s["zero_multi_gamma"].selected_gamma == 0
    assert s["fixed_tau1"].selected_tau1 == 10
    assert np.sum(s["fixed_tau1"].cv_score_tau1) == 0


def test_spatial_prediction_eigenvalues(sp, seq_1d):   # test-SpatPCA.R:52-88, 123-127
    f, z = seq_1d["fixed_tau1"], seq_1d["zero_zero_K5"]
    assert sp.spatialPrediction(f.eigenfn, f.detrended_Y, 1000, f.eigenfn)["eigenvalue"].sum() == 0
    assert sp.spatialPrediction(z.eigenfn, z.detrended_Y, 10, z.eigenfn)["eigenvalue"].sum() == 0
    assert abs(sp.spatialPrediction(z.eigenfn, z.detrended_Y, 0.5, z.eigenfn)["eigenvalue"].sum()
               - 9.397599) <= 1e-6
    assert sp.spatialPrediction(f.eigenfn, f.detrended_Y, 5, f.eigenfn)["eigenvalue"].sum() - 0.1066821 <= 1e-6


@pytest.mark.parametrize("p,K,n,p_new,gamma", [(60, 3, 40, 25, 0.05), (200, 5, 30, 200, 0.5), (130, 2, 50, 7, 0.0)])
def test_spatial_prediction_outputs_vs_oracle(sp, p, K, n, p_new, gamma):
    """`prediction` and `estimated_covariance` (:764-769), with p_new != p and a gamma that keeps eigenvalues alive."""
    rng = np.random.default_rng(1000 + p)
    load = np.linalg.qr(rng.normal(size=(p, K)))[0]
    Y = (rng.normal(size=(n, K)) * np.linspace(5, 2, K)) @ load.T + 0.3 * rng.normal(size=(n, p))
    phi = load + 0.01 * rng.normal(size=(p, K))
    phi_new = rng.normal(size=(p_new, K))
    got = sp.spatialPrediction(phi, Y, gamma, phi_new)
    want = ref.spatial_prediction(phi, Y, gamma, phi_new)
    assert np.count_nonzero(want["eigenvalue"]) > 0
    np.testing.assert_allclose(got["eigenvalue"], np.ravel(want["eigenvalue"]), rtol=1e-10, atol=1e-12)
    assert abs(got["error"] - float(want["error"])) <= 1e-10 * abs(float(want["error"]))
    for key in ("estimated_covariance", "prediction"):
        assert got[key].shape == want[key].shape
        assert np.max(np.abs(got[key] - want[key])) / np.max(np.abs(want[key])) <= 1e-9, key


def test_predict_shapes(sp, seq_1d):                   # test-SpatPCA.R:138-146
    xn = r_seq_len(6, 7, 4).reshape(-1, 1)
    assert sp.predict(seq_1d["cv_1D"], xn).shape[1] == 4
    assert sp.predictEigenfunction(seq_1d["cv_1D"], xn).shape[0] == 4


def test_k_selection_aux(sp):                          # test-SpatPCA.R:150-165
    _, x, Y = case_1d()
    rng = RRng(1234)
    M = 3
    split = rng.sample(np.resize(np.arange(1, M + 1), 100))
    tau2 = sp.setTau2(None, M)
    r = sp.spatpcaCVWithSelectedK(x, Y, M, sp.setTau1(None, M), tau2, [1.0], split, 10, 1e-4, sp.setL2(tau2))
    assert r["selected_K"] == 1
    assert r["cv_result"]["selected_gamma"] == 1
    assert r["cv_result"]["selected_tau1"] == 1
    assert r["cv_result"]["selected_tau2"] == 0


def test_3d_case(sp):                                  # test-SpatPCA.R:169-187
    rng, d, Y = case_3d()
    cv = sp.spatpca(d, Y, rng=rng)
    assert cv.eigenfn.shape == (64, 2)
    pred = sp.eigenFunction(np.zeros((1, 3)), d, cv.eigenfn)
    # the reference pins the sign LAPACK's SVD happened to pick; parity is up to column sign
    want = np.array([0.232199, 0.007501031])
    assert np.sum(np.abs(np.abs(pred[0]) - np.abs(want))) <= 1e-6


# ---------------------------------------------------------------------------------------------
# kernel-level parity against the oracle
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("d,p", [(1, 37), (2, 100), (3, 125), (2, 900), (3, 1000), (2, 2500)])
def test_tps_matrix_vs_oracle(sp, d, p):
    rng = np.random.default_rng(d * 100 + p)
    x = rng.uniform(-3, 3, size=(p, d))
    got = sp.thinPlateSplineMatrix(x)
    want = ref.thin_plate_spline_matrix(x)
    # The CUDA build is refined to ~1e-12 of the exact matrix (csrc/omega_refine.cu, profiles/r02_omega_threeway.md), so
    # what is measured here is the ORACLE's own LAPACK rounding error: 0.01-0.02 cond(A) eps on these inputs
    # (7.7e-9 / 3.5e-12 / 1.2e-14 / 9.7e-10 / 1.5e-13 / 1.7e-9).  The bound leaves a factor ~5-10.
    cond = np.linalg.cond(ref.tps_bordered(x) + 1e-8 * np.eye(p + d + 1))
    assert np.max(np.abs(got - want)) / np.max(np.abs(want)) <= max(1e-13, 0.1 * cond * np.finfo(float).eps)
    assert np.array_equal(got, got.T)          # the refined product is exactly symmetric


@pytest.mark.parametrize("m,p,K", [(40, 300, 3), (160, 2304, 5), (200, 200, 8), (30, 20, 4)])
def test_svd_start_vs_lapack(sp, m, p, K):
    """The device SVD start (QR of the long side + one-sided Jacobi + back-transformation; src/RcppSpatPCA.cpp:321, :563)
    against LAPACK: all singular values to 1e-12 of the largest, the K leading right singular vectors to 1e-10 up to
    sign, and setGamma's maximum d[1]^2 / nrow (R/helper.R:207-217) to 1e-12."""
    from spatpca_b200 import engine as eng
    rng = np.random.default_rng(m * 1000 + p)
    r = min(m, p)
    load = np.linalg.qr(rng.normal(size=(p, r)))[0]
    Y = (rng.normal(size=(m, r)) * np.linspace(30.0, 1.0, r)) @ load.T + 0.1 * rng.normal(size=(m, p))
    sv, V = eng.svd_right(Y, K)
    _, want_s, want_vt = np.linalg.svd(Y, full_matrices=False)
    assert np.max(np.abs(sv - want_s)) <= 1e-12 * want_s[0]
    assert eig_err(V, want_vt[:K].T) <= 1e-10
    assert abs(sv[0] ** 2 / m - want_s[0] ** 2 / m) <= 1e-12 * want_s[0] ** 2 / m
    assert np.max(np.abs(V.T @ V - np.eye(K))) <= 1e-12


@pytest.mark.parametrize("d", [1, 2, 3])
def test_eigen_function_vs_oracle(sp, d):
    """eigenFunction (src/RcppSpatPCA.cpp:94-147) three ways: the CUDA path, the oracle and an 80-bit ground truth
    (tools/eigenfn_threeway.py).  The un-ridged bordered system has cond 1e11 / 1e5 / 4e3 for 80 random sites in
    1 / 2 / 3 dimensions, so two FP64 LU solves differ by a fraction of cond * eps; measured on B200
    (profiles/r02d_eigenfn_threeway.txt): GPU vs oracle 0.002 / 0.08 / 0.15 cond * eps, and the GPU is as close to the
    ground truth as LAPACK is (5.7e-8 vs 6.5e-8, 1.1e-12 vs 9.9e-13, 2.3e-14 vs 1.4e-13)."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    from eigenfn_threeway import truth_eigen_function
    rng = np.random.default_rng(40 + d)
    xo = rng.uniform(-2, 2, size=(80, d))
    xn = np.vstack([rng.uniform(-2, 2, size=(17, d)), xo[:3]])     # includes coincident sites (r = 0 guard)
    Phi = rng.normal(size=(80, 3))
    got = sp.eigenFunction(xn, xo, Phi)
    want = ref.eigen_function(xn, xo, Phi)
    cond = np.linalg.cond(ref.tps_bordered(xo))
    scale = np.max(np.abs(want))
    assert np.max(np.abs(got - want)) / scale <= max(1e-13, 0.5 * cond * np.finfo(float).eps)
    truth = truth_eigen_function(xn, xo, Phi).astype(np.float64)
    err_gpu = np.max(np.abs(got - truth)) / scale
    err_oracle = np.max(np.abs(want - truth)) / scale
    assert err_gpu <= max(1e-13, 3.0 * err_oracle), (err_gpu, err_oracle)


def _spd_problem(p, K, seed, n=40):
    rng = np.random.default_rng(seed)
    Yt = rng.normal(size=(n, p))
    G = Yt.T @ Yt
    x = rng.uniform(0, 1, size=(p, 2))
    Om = ref.thin_plate_spline_matrix(x)
    Om[np.diag_indices(p)] += 1e-8
    s, rho, Phi0, L20 = ref.pca_init(Yt, K)
    return Yt, G, Om, rho, Phi0, L20


@pytest.mark.parametrize("p,K", [(50, 1), (300, 3), (1000, 5), (513, 8), (700, 12), (640, 20)])
def test_inverse_and_core2_vs_oracle(sp, p, K):
    from spatpca_b200 import engine as eng
    Yt, G, Om, rho, Phi0, L20 = _spd_problem(p, K, 3 * p + K)
    tau1 = 0.01
    A = 2.0 * (tau1 * Om - G)
    A[np.diag_indices(p)] += rho
    want_inv = ref.arma_inv_sympd_symmatu(A)
    got_inv = eng.inv_sympd(Om, G, tau1, rho, 2.0)
    assert np.max(np.abs(got_inv - want_inv)) / np.max(np.abs(want_inv)) <= 1e-10
    tr = ref.Trace()
    wPhi, wC, wL2 = ref.core2(G, None, Phi0, L20, Om, tau1, rho, 100, 1e-4, tr)
    gPhi, gC, gL2, it = eng.admm_core2(want_inv, Phi0, L20, rho, 100, 1e-4)
    assert it == tr.calls[0][3]
    for g, w in ((gPhi, wPhi), (gC, wC), (gL2, wL2)):
        assert np.max(np.abs(g - w)) / np.max(np.abs(w)) <= 1e-10


@pytest.mark.parametrize("p,K,tau2", [(60, 2, 0.0), (300, 3, 5.0), (1000, 5, 50.0), (513, 8, 1.0)])
def test_core3_vs_oracle(sp, p, K, tau2):
    from spatpca_b200 import engine as eng
    Yt, G, Om, rho, Phi0, L20 = _spd_problem(p, K, 7 * p + K)
    tinv = ref.core3_tempinv(Om, G, 0.01, rho)
    R0 = Phi0.copy()
    L10 = np.zeros_like(Phi0)
    tr = ref.Trace()
    want = ref.core3(tinv, Phi0, R0, Phi0, L10, L20, tau2, rho, 100, 1e-4, tr)
    got = eng.admm_core3(tinv, R0, Phi0, L10, L20, tau2, rho, 100, 1e-4)
    assert got[5] == tr.calls[0][3]
    for g, w in zip(got[:5], want):
        scale = max(np.max(np.abs(w)), 1e-300)
        assert np.max(np.abs(g - w)) / scale <= 1e-10


@pytest.mark.parametrize("scales", [(1.0, 0.2, 0.05), (1.0, 0.03, 0.001)])
def test_polar_step_on_ill_conditioned_iterate(sp, scales):
    """The polar step (svd_econ's U V', :169-171) when T'T is far from the identity: with Ainv = I / rho and
    Lambda2 = 0 the first T equals the start C, whose columns are scaled to cond(T'T) = 400 resp. 1e6.  The
    Newton-Schulz iteration needs ~12 resp. ~22 of its 24 steps there (beyond that the kernel falls back to the
    Jacobi eigen-solve); either way the iterates have to match the oracle's SVD-based ones."""
    from spatpca_b200 import engine as eng
    rng = np.random.default_rng(5)
    p, K, rho = 700, 3, 7.0
    Q, _ = np.linalg.qr(rng.normal(size=(p, K)))
    mix = np.array([[1.0, 0.3, 0.1], [0.0, 1.0, 0.4], [0.0, 0.0, 1.0]])
    C0 = (Q * np.array(scales)) @ mix
    L20 = np.zeros((p, K))
    Z = np.zeros((p, p))
    tr = ref.Trace()
    wPhi, wC, wL2 = ref.core2(Z, None, C0, L20, Z, 0.0, rho, 3, 0.0, tr)
    gPhi, gC, gL2, it = eng.admm_core2(np.eye(p) / rho, C0, L20, rho, 3, 0.0)
    assert it == 3 == tr.calls[0][3]
    scale = np.max(np.abs(wPhi))            # Lambda2 ends at rounding-noise level: compare on Phi's scale
    for g, w in ((gPhi, wPhi), (gC, wC), (gL2, wL2)):
        assert np.max(np.abs(g - w)) / max(np.max(np.abs(w)), scale) <= 1e-8


def test_core3_forced_iterations(sp):
    """thr = 0: exactly maxit iterations, still matching the oracle iterate by iterate."""
    from spatpca_b200 import engine as eng
    Yt, G, Om, rho, Phi0, L20 = _spd_problem(400, 4, 99)
    tinv = ref.core3_tempinv(Om, G, 0.1, rho)
    tr = ref.Trace()
    want = ref.core3(tinv, Phi0, Phi0.copy(), Phi0, np.zeros_like(Phi0), L20, 2.0, rho, 25, 0.0, tr)
    got = eng.admm_core3(tinv, Phi0.copy(), Phi0, np.zeros_like(Phi0), L20, 2.0, rho, 25, 0.0)
    assert got[5] == 25 == tr.calls[0][3]
    assert np.max(np.abs(got[0] - want[0])) / np.max(np.abs(want[0])) <= 1e-10


# block U'DU factor + multi-phase stream (csrc/spd_inverse.cu, csrc/admm_kernels.cu) against the oracle's
# explicit inv_sympd: same iterates to 1e-10, same iteration counts, for every block count incl. ragged
# last blocks (p not a multiple of 512), K > 8 (256-row tiles) and blocks = 1 (explicit inverse).
@pytest.mark.parametrize("p,K,blocks", [(1000, 5, 2), (1300, 3, 3), (1537, 5, 4), (2048, 2, 4), (1100, 12, 3),
                                        (4100, 5, 8), (700, 4, 1)])
def test_factored_core2_vs_oracle(sp, p, K, blocks):
    from spatpca_b200 import engine as eng
    Yt, G, Om, rho, Phi0, L20 = _spd_problem(p, K, 11 * p + K)
    tau1 = 0.01
    tr = ref.Trace()
    wPhi, wC, wL2 = ref.core2(G, None, Phi0, L20, Om, tau1, rho, 100, 1e-4, tr)
    gPhi, _, gC, _, gL2, it = eng.admm_system(Om, G, tau1, rho, 2.0, blocks, 2, Phi0, L20, 100, 1e-4)
    assert it == tr.calls[0][3]
    for g, w in ((gPhi, wPhi), (gC, wC), (gL2, wL2)):
        assert np.max(np.abs(g - w)) / np.max(np.abs(w)) <= 1e-10


@pytest.mark.parametrize("p,K,blocks,tau2,maxit,tol", [(1300, 3, 3, 5.0, 100, 1e-4), (1537, 5, 4, 50.0, 100, 1e-4),
                                                       (2000, 4, 2, 2.0, 25, 0.0), (1100, 9, 3, 1.0, 100, 1e-4)])
def test_factored_core3_vs_oracle(sp, p, K, blocks, tau2, maxit, tol):
    from spatpca_b200 import engine as eng
    Yt, G, Om, rho, Phi0, L20 = _spd_problem(p, K, 13 * p + K)
    tinv = ref.core3_tempinv(Om, G, 0.01, rho)
    R0, L10 = Phi0.copy(), np.zeros_like(Phi0)
    tr = ref.Trace()
    want = ref.core3(tinv, Phi0, R0, Phi0, L10, L20, tau2, rho, maxit, tol, tr)
    got = eng.admm_system(Om, G, 0.01, rho, 1.0, blocks, 3, Phi0, L20, maxit, tol, R=R0, Lambda1=L10, tau2=tau2)
    assert got[5] == tr.calls[0][3]
    for g, w in zip(got[:5], want):
        scale = max(np.max(np.abs(w)), 1e-300)
        assert np.max(np.abs(g - w)) / scale <= 1e-10


def test_factored_matches_explicit_and_is_reproducible(sp):
    """blocks = 4 vs the explicit inverse on the same system; two runs of the factored path are bitwise equal."""
    from spatpca_b200 import engine as eng
    Yt, G, Om, rho, Phi0, L20 = _spd_problem(2100, 5, 4242)
    a = eng.admm_system(Om, G, 0.1, rho, 2.0, 4, 2, Phi0, L20, 30, 0.0)
    b = eng.admm_system(Om, G, 0.1, rho, 2.0, 4, 2, Phi0, L20, 30, 0.0)
    e = eng.admm_system(Om, G, 0.1, rho, 2.0, 1, 2, Phi0, L20, 30, 0.0)
    assert a[5] == b[5] == e[5] == 30
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2]) and np.array_equal(a[4], b[4])
    assert np.max(np.abs(a[0] - e[0])) / np.max(np.abs(e[0])) <= 1e-11


# Matrix-free Core2 (conjugate gradients inside the persistent kernel, no inv_sympd) against the oracle's explicit
# inverse: identical iteration counts, iterates to 1e-10, for ragged sizes, every K instance and warm / cold starts.
@pytest.mark.parametrize("p,K,tau1,n", [(64, 1, 0.01, 40), (300, 3, 0.01, 40), (1000, 5, 0.01, 40), (513, 8, 0.1, 70),
                                        (1537, 5, 1.0, 300), (2100, 2, 0.001, 33), (4100, 5, 0.1, 160)])
def test_core2_matrix_free_vs_oracle(sp, p, K, tau1, n):
    from spatpca_b200 import engine as eng
    Yt, G, Om, rho, Phi0, L20 = _spd_problem(p, K, 17 * p + K, n=n)
    tr = ref.Trace()
    wPhi, wC, wL2 = ref.core2(G, None, Phi0, L20, Om, tau1, rho, 100, 1e-4, tr)
    gPhi, gC, gL2, it, passes = eng.core2_matrix_free(Om, Yt, tau1, rho, Phi0, Phi0, L20, 100, 1e-4)
    assert it == tr.calls[0][3]
    assert passes >= 2 * it
    for g, w in ((gPhi, wPhi), (gC, wC), (gL2, wL2)):
        assert np.max(np.abs(g - w)) / np.max(np.abs(w)) <= 1e-10
    # a second, warm-started call (the next tau1 of a path, :329-332) from the oracle's own state
    tr2 = ref.Trace()
    w2 = ref.core2(G, wPhi, wC, wL2, Om, 3.0 * tau1, rho, 100, 1e-4, tr2)
    g2 = eng.core2_matrix_free(Om, Yt, 3.0 * tau1, rho, wPhi, wC, wL2, 100, 1e-4)
    assert g2[3] == tr2.calls[0][3]
    for g, w in zip(g2[:3], w2):
        assert np.max(np.abs(g - w)) / np.max(np.abs(w)) <= 1e-10


def test_core2_matrix_free_forced_and_reproducible(sp):
    """thr = 0: exactly maxit iterations; two runs are bitwise equal (fixed-shape reductions, no atomics on data)."""
    from spatpca_b200 import engine as eng
    Yt, G, Om, rho, Phi0, L20 = _spd_problem(1300, 4, 777, n=50)
    tr = ref.Trace()
    want = ref.core2(G, None, Phi0, L20, Om, 0.05, rho, 12, 0.0, tr)
    a = eng.core2_matrix_free(Om, Yt, 0.05, rho, Phi0, Phi0, L20, 12, 0.0)
    b = eng.core2_matrix_free(Om, Yt, 0.05, rho, Phi0, Phi0, L20, 12, 0.0)
    assert a[3] == b[3] == 12 == tr.calls[0][3]
    for x, y in zip(a[:3], b[:3]):
        assert np.array_equal(x, y)
    assert np.max(np.abs(a[0] - want[0])) / np.max(np.abs(want[0])) <= 1e-10


def test_core2_matrix_free_rejects_indefinite_system(sp):
    """rho far below the Gram matrix's top eigenvalue: inv_sympd fails in the reference, r.Ar <= 0 here."""
    from spatpca_b200 import engine as eng, SpatpcaError
    Yt, G, Om, rho, Phi0, L20 = _spd_problem(300, 2, 5)
    with pytest.raises(SpatpcaError):
        eng.core2_matrix_free(Om, Yt, 0.01, rho * 1e-3, Phi0, Phi0, L20, 5, 1e-4)


@pytest.mark.parametrize("p,K,M,n", [(300, 2, 3, 40), (1000, 5, 5, 60), (2100, 3, 5, 45), (4100, 5, 5, 200), (700, 8, 4, 50)])
def test_core2_batched_path_vs_oracle(sp, p, K, M, n):
    """The folds of a stage in lock-step on the TMA-fed DMMA stream: a tau1 path of 4 values, warm starts carried,
    against the oracle's spatpcaCore2 per fold (explicit inv_sympd): identical iteration counts per (step, fold),
    every snapshot of Phi to 1e-10."""
    from spatpca_b200 import engine as eng
    if not eng.core2_batched_supported(p, n, K, M):
        pytest.skip("shape not supported by the batched kernel on this device")
    rng = np.random.default_rng(31 * p + K)
    x = rng.uniform(0, 1, size=(p, 2))
    Om = ref.thin_plate_spline_matrix(x)
    Om[np.diag_indices(p)] += 1e-8
    load = np.linalg.qr(rng.normal(size=(p, K)))[0]
    Y = (rng.normal(size=(n, K)) * np.linspace(30, 10, K)) @ load.T + rng.normal(size=(n, p))
    nk = rng.permutation(np.resize(np.arange(1, M + 1), n)).astype(float)
    tau1 = np.array([1e-4, 1e-2, 0.1, 1.0])
    starts, rhos = [], []
    want_snap, want_it = [], []
    for k in range(M):
        Ytr = Y[nk != k + 1]
        _, rho, Phi0, L20 = ref.pca_init(Ytr, K)
        starts.append((Phi0, L20))
        rhos.append(rho)
        G = np.asfortranarray(Ytr.T @ Ytr)
        Phi, Cc, L2 = Phi0, Phi0, L20
        snaps, its = [], []
        for t in tau1:
            tr = ref.Trace()
            Phi, Cc, L2 = ref.core2(G, Phi, Cc, L2, Om, t, rho, 100, 1e-4, tr)
            snaps.append(Phi)
            its.append(tr.calls[0][3])
        want_snap.append(snaps)
        want_it.append(its)
    got = eng.core2_batched(Om, Y, nk, list(range(1, M + 1)), rhos, tau1, [s[0] for s in starts], [s[0] for s in starts],
                            [s[1] for s in starts], 100, 1e-4)
    assert got["iters"].tolist() == np.array(want_it).T.tolist()
    for k in range(M):
        for i in range(len(tau1)):
            w = want_snap[k][i]
            assert np.max(np.abs(got["snap"][i, k] - w)) / np.max(np.abs(w)) <= 1e-10, (k, i)
    assert got["passes"] > 0


def test_core2p_rides_with_the_folds(sp):
    """spatpcaCore2p (src/RcppSpatPCA.cpp:187-222: the full-data fit at the selected tau1, :598-599) in the SAME launch
    as the folds' spatpcaCore2 calls -- fold id 0 = all rows -- as stage 2 of the engine runs it: against the oracle's
    core2 on the full Gram matrix and on each fold's, identical iteration counts, Phi to 1e-10."""
    from spatpca_b200 import engine as eng
    p, K, M, n = 2100, 3, 3, 60
    if not eng.core2_batched_supported(p, n, K, M + 1):
        pytest.skip("shape not supported by the batched kernel on this device")
    rng = np.random.default_rng(99)
    x = rng.uniform(0, 1, size=(p, 2))
    Om = ref.thin_plate_spline_matrix(x)
    Om[np.diag_indices(p)] += 1e-8
    load = np.linalg.qr(rng.normal(size=(p, K)))[0]
    Y = (rng.normal(size=(n, K)) * np.linspace(30, 10, K)) @ load.T + rng.normal(size=(n, p))
    nk = rng.permutation(np.resize(np.arange(1, M + 1), n)).astype(float)
    tau1 = np.array([0.05])
    ids = [0] + list(range(1, M + 1))
    starts, rhos, want, want_it = [], [], [], []
    for fid in ids:
        Ytr = Y if fid == 0 else Y[nk != fid]
        _, rho, Phi0, L20 = ref.pca_init(Ytr, K)
        starts.append((Phi0, L20))
        rhos.append(rho)
        tr = ref.Trace()
        Phi, _, _ = ref.core2(np.asfortranarray(Ytr.T @ Ytr), Phi0, Phi0, L20, Om, tau1[0], rho, 100, 1e-4, tr)
        want.append(Phi)
        want_it.append(tr.calls[0][3])
    got = eng.core2_batched(Om, Y, nk, ids, rhos, tau1, [s[0] for s in starts], [s[0] for s in starts],
                            [s[1] for s in starts], 100, 1e-4)
    assert got["iters"].tolist() == [want_it]
    for f in range(len(ids)):
        assert np.max(np.abs(got["snap"][0, f] - want[f])) / np.max(np.abs(want[f])) <= 1e-10, f


@pytest.mark.gpu
@pytest.mark.parametrize("p,K,M,n", [(640, 3, 4, 40), (1500, 5, 5, 60), (2300, 2, 1, 30)])
def test_core3_batched_path_vs_oracle(sp, p, K, M, n):
    """spatpcaCore3 along a tau2 path for all folds in one launch of the matrix-free kernel (CG solves with
    2 tau1 Omega - 2 Ytr'Ytr + 2 rho I instead of the tempinv product of :247) against the oracle's core3 with its
    explicit inv_sympd: identical iteration counts per (step, fold), every snapshot of Phi and the final
    (R, C, Lambda1, Lambda2) to 1e-10."""
    from spatpca_b200 import engine as eng
    if not eng.core2_batched_supported(p, n, K, M):
        pytest.skip("shape not supported by the batched kernel on this device")
    rng = np.random.default_rng(17 * p + K)
    x = rng.uniform(0, 1, size=(p, 2))
    Om = ref.thin_plate_spline_matrix(x)
    Om[np.diag_indices(p)] += 1e-8
    load = np.linalg.qr(rng.normal(size=(p, K)))[0]
    Y = (rng.normal(size=(n, K)) * np.linspace(30, 10, K)) @ load.T + rng.normal(size=(n, p))
    nk = rng.permutation(np.resize(np.arange(1, M + 1), n)).astype(float)
    tau1 = 0.05
    tau2 = np.array([0.0, 1e-3, 1e-2, 0.1, 1.0])
    fold_ids = list(range(1, M + 1)) if M > 1 else [0]
    starts, rhos, want = [], [], []
    for k in range(M):
        Ytr = Y[nk != fold_ids[k]] if fold_ids[k] else Y
        _, rho, Phi0, L20 = ref.pca_init(Ytr, K)
        G = np.asfortranarray(Ytr.T @ Ytr)
        Phi, Cc, L2 = ref.core2(G, Phi0, Phi0, L20, Om, tau1, rho, 100, 1e-4)          # :392-393
        starts.append((Phi, Phi.copy(), Cc, np.zeros_like(Phi), L2))                  # :394 R = Phi, :390 Lambda1 = 0
        rhos.append(rho)
        tinv = ref.core3_tempinv(Om, G, tau1, rho)
        R, L1 = Phi.copy(), np.zeros_like(Phi)
        snaps, its = [], []
        for t in tau2:
            tr = ref.Trace()
            Phi, R, Cc, L1, L2 = ref.core3(tinv, Phi, R, Cc, L1, L2, t, rho, 100, 1e-4, tr)
            snaps.append(Phi)
            its.append(tr.calls[0][3])
        want.append(dict(snaps=snaps, its=its, R=R, C=Cc, L1=L1, L2=L2))
    got = eng.core3_batched(Om, Y, nk, fold_ids, rhos, tau1, tau2, *[[s[i] for s in starts] for i in range(5)], 100, 1e-4)
    assert got["iters"].tolist() == np.array([w["its"] for w in want]).T.tolist()
    rel = lambda a, b: np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)
    for k in range(M):
        for i in range(len(tau2)):
            assert rel(got["snap"][i, k], want[k]["snaps"][i]) <= 1e-10, (k, i)
        for name in ("R", "C", "L1", "L2"):
            scale = max(np.max(np.abs(want[k][name])), 1e-12)
            assert np.max(np.abs(got[name][k] - want[k][name])) / scale <= 1e-9, (k, name)
    assert got["passes"] > 0


@pytest.mark.parametrize("p,nc", [(64, 3), (1003, 25), (2500, 32), (4100, 12), (777, 17), (10000, 25)])
def test_dmma_stream_vs_numpy(sp, p, nc):
    """U = A V on the TMA + DMMA stream (csrc/tma_dmma.cuh) against NumPy: swizzled tiles, fragment permutations,
    ragged edges (p not a multiple of 8 / 64), every vector-tile count."""
    from spatpca_b200 import engine as eng
    rng = np.random.default_rng(p + nc)
    A = rng.normal(size=(p, p))
    A = A + A.T
    V = rng.normal(size=(p, nc))
    U = eng.dmma_matvec(A, V)
    W = A @ V
    assert np.max(np.abs(U - W)) / np.max(np.abs(W)) <= 1e-13


@pytest.mark.parametrize("name", ["irregular2d_small", "grid3d_small"])
def test_cv_matrix_free_forced_vs_oracle(sp, name, monkeypatch):
    """SPB_CORE2=cg: every spatpcaCore2 / Core2p call of the job runs matrix-free (the automatic choice only does so
    from p = 2048 on); same gates as the factored path, and SPB_CORE2=factor gives the same selections."""
    case = make_case(name)
    monkeypatch.setenv("SPB_CORE2", "cg")
    got, want, tr = run_both(sp, case)
    assert_cv_parity(got, want)
    assert got["stats"]["n_cg_calls"] > 0 and got["stats"]["cg_passes"] > 0
    assert got["stats"]["admm_iters"] == tr.admm_iters
    monkeypatch.setenv("SPB_CORE2", "factor")
    fac = _run_cv(sp, case)
    assert fac["stats"]["n_cg_calls"] == 0
    assert_cv_parity(fac, want)


@pytest.mark.parametrize("m,p,K", [(1, 10, 1), (20, 50, 2), (40, 1000, 5), (100, 333, 8), (7, 4097, 3)])
def test_cv_loss_vs_oracle(sp, m, p, K):
    from spatpca_b200 import engine as eng
    rng = np.random.default_rng(m + p + K)
    Yva = rng.normal(size=(m, p))
    Phi = np.linalg.qr(rng.normal(size=(p, K)))[0] * (1 + 0.01 * rng.normal(size=(1, K)))
    want = float(np.linalg.norm(Yva @ (np.eye(p) - Phi @ Phi.T)) ** 2)    # the reference's literal form (:327)
    got = eng.cv_loss(Yva, Phi)
    assert abs(got - want) / want <= 1e-12


@pytest.mark.parametrize("p,K,ng", [(10, 1, 3), (130, 3, 11), (600, 5, 11), (257, 4, 20)])
def test_gamma_loss_vs_oracle(sp, p, K, ng):
    from spatpca_b200 import engine as eng
    rng = np.random.default_rng(p + K)
    n_tr, n_va = 64, 16
    load = np.linalg.qr(rng.normal(size=(p, K)))[0]
    Z = rng.normal(size=(n_tr + n_va, K)) * np.linspace(6, 2, K)
    Yall = Z @ load.T + rng.normal(size=(n_tr + n_va, p))
    Ytr, Yva = Yall[:n_tr], Yall[n_tr:]
    Phi = load + 0.01 * rng.normal(size=(p, K))
    gam = np.concatenate([[0.0], np.exp(np.linspace(np.log(1e-3), np.log(40.0), ng - 1))])
    want = ref.gamma_losses(Phi, Ytr, Yva, gam)
    got = eng.gamma_loss(Phi, Ytr, Yva, gam)
    np.testing.assert_allclose(got, want, rtol=1e-11)


# ---------------------------------------------------------------------------------------------
# end-to-end spatpcaCV parity on synthetic fields (reduced sizes)
# ---------------------------------------------------------------------------------------------
def run_both(sp, case, **kw):
    args = (case["x"], case["Y"], case["M"], case["K"], case["tau1"], case["tau2"], case["gamma"],
            case["nk"], kw.get("maxit", 100), kw.get("tol", 1e-4), case["l2"])
    got = sp.spatpcaCV(*args, return_stats=True)
    tr = ref.Trace()
    want = ref.spatpca_cv(*args, trace=tr)
    return got, want, tr


def test_cv_1d_demo_config(sp):
    """BASELINE config 1: 1-D demo, p=50, n=100, default grids (R/SpatPCA.R:94-98)."""
    rng = RRng(1234)
    x = r_seq_len(-5, 5, 50).reshape(-1, 1)
    phi = np.exp(-x ** 2)
    phi /= np.linalg.norm(phi)
    Y = np.outer(rng.rnorm(100, sd=3.0), phi[:, 0]) + rng.rnorm(100 * 50).reshape(50, 100).T
    nk = rng.sample(np.resize(np.arange(1, 6), 100)).astype(float)
    sx = r_api.scale_location(x)
    case = dict(x=sx, Y=Y, M=5, K=2, tau1=r_api.set_tau1(None, 5), tau2=np.array([0.0]),
                gamma=r_api.set_gamma(None, Y[nk != 1]), nk=nk, l2=np.array([1.0]))
    got, want, _ = run_both(sp, case)
    assert_cv_parity(got, want)


@pytest.mark.parametrize("name", ["grid2d_small", "irregular2d_small", "grid3d_small"])
def test_cv_synthetic_vs_oracle(sp, name):
    case = make_case(name)
    got, want, tr = run_both(sp, case)
    assert_cv_parity(got, want)
    assert got["stats"]["n_admm_calls"] == len(tr.calls)
    assert got["stats"]["admm_iters"] == tr.admm_iters


def test_cv_call_log_matches_oracle(sp):
    """Per-call ADMM iteration counts must be identical (trap 1)."""
    case = make_case("irregular2d_small")
    e = sp.Engine(case["x"].shape[0], case["x"].shape[1], case["Y"].shape[0], case["M"], case["K"],
                  case["tau1"], case["tau2"], case["gamma"], 100, 1e-4, case["l2"])
    e.upload(case["x"], case["Y"], case["nk"])
    e.prepare()
    M = case["M"]
    cv1 = np.array([e.stage1_fold(k) for k in range(M)])
    i1, _ = sp.select_index(cv1)
    e.stage1_full(i1)
    cv2 = np.array([e.stage2_fold(k, i1) for k in range(M)])
    i2, _ = sp.select_index(cv2)
    e.stage2_full(i1, i2)
    cv3 = np.array([e.stage3_fold(k, i1, i2) for k in range(M)])
    log = e.call_log()
    tr = ref.Trace()
    want = ref.spatpca_cv(case["x"], case["Y"], M, case["K"], case["tau1"], case["tau2"], case["gamma"],
                          case["nk"], 100, 1e-4, case["l2"], trace=tr)
    # the engine runs fold work stage by stage in the same order as the reference's workers
    got_seq = [(3 if k == 3 else 2, f, i, it) for (k, f, i, it) in log]
    want_seq = [(3 if "core3" in c[0] else 2, c[1], c[2], c[3]) for c in tr.calls]
    assert sorted(got_seq) == sorted(want_seq)
    np.testing.assert_allclose(cv1, want["aux"]["cv1"], rtol=CV_RTOL)
    np.testing.assert_allclose(cv2, want["aux"]["cv2"], rtol=CV_RTOL)
    np.testing.assert_allclose(cv3, want["aux"]["cv3"], rtol=CV_RTOL)
    e.close()


def test_cv_with_injected_omega(sp):
    """ADMM parity separated from the ill-conditioned bordered inverse: both sides get the same Omega."""
    case = make_case("irregular2d_small")
    om = ref.thin_plate_spline_matrix(case["x"])
    e = sp.Engine(case["x"].shape[0], case["x"].shape[1], case["Y"].shape[0], case["M"], case["K"],
                  case["tau1"], case["tau2"], case["gamma"], 100, 1e-4, case["l2"])
    e.set_omega(om)
    e.upload(case["x"], case["Y"], case["nk"])
    e.prepare()
    M = case["M"]
    cv1 = np.array([e.stage1_fold(k) for k in range(M)])
    i1, s1 = sp.select_index(cv1)
    e.stage1_full(i1)
    cv2 = np.array([e.stage2_fold(k, i1) for k in range(M)])
    i2, s2 = sp.select_index(cv2)
    e.stage2_full(i1, i2)
    phi = e.get_eigenfn()
    want = ref.spatpca_cv(case["x"], case["Y"], M, case["K"], case["tau1"], case["tau2"], case["gamma"],
                          case["nk"], 100, 1e-4, case["l2"], omega=om)
    np.testing.assert_allclose(s1, want["cv_score_tau1"], rtol=1e-11)
    np.testing.assert_allclose(s2, want["cv_score_tau2"], rtol=1e-11)
    assert eig_err(phi, want["estimated_eigenfn"]) <= 1e-10
    e.close()


def test_cv_branches(sp):
    """Every branch of spatpcaCV :592-692 on one small input (mirrors test-SpatPCA.R:14-51)."""
    case = make_case("grid2d_small")
    base = dict(case)
    variants = [
        dict(tau1=[0.0], tau2=[0.0], gamma=[0.0, 0.5, 1.0]),
        dict(tau1=[10.0], tau2=[0.0], gamma=case["gamma"]),
        dict(tau1=[10.0], tau2=[100.0], gamma=[0.0, 1.0]),
        dict(tau1=[0.0], tau2=[0.0, 1.0, 10.0], gamma=case["gamma"]),
        dict(tau1=[0.5], tau2=[0.0, 1.0, 10.0], gamma=[2.0]),
        dict(tau1=case["tau1"], tau2=[3.0], gamma=case["gamma"]),
        dict(tau1=[0.0], tau2=[5.0], gamma=case["gamma"]),
    ]
    for v in variants:
        c = dict(base)
        c.update({k: np.asarray(val, dtype=float) for k, val in v.items()})
        c["l2"] = r_api.set_l2(c["tau2"])
        got, want, _ = run_both(sp, c)
        assert_cv_parity(got, want)


def test_error_paths(sp):
    from spatpca_b200 import SpatpcaError
    with pytest.raises(SpatpcaError):
        sp.thinPlateSplineMatrix(np.zeros((5, 4)))              # d > 3
    x = np.random.default_rng(0).uniform(size=(20, 2))
    Y = np.random.default_rng(1).normal(size=(10, 20))
    with pytest.raises(SpatpcaError):                            # K larger than the training block
        sp.spatpcaCV(x, Y, 5, 9, [0, 1.0], [0.0], [0.0], np.resize(np.arange(1, 6), 10), 10, 1e-4, [1.0])
    # duplicate 2-D sites: 0 * log(0) = NaN in the reference too (:29); the LU then fails or yields NaN
    # ... at a size that takes the LU start (p = 21) and at one that takes the SPD / Cholesky start with the INT8
    # refinement (p = 2301: the digit-plane kernels poison their output on a non-finite operand, the fallback LU
    # reports a singular matrix -- "inv(): matrix is singular", as Armadillo would)
    for xs in (x, np.random.default_rng(0).uniform(size=(2300, 2))):
        xd = np.vstack([xs, xs[:1]])
        try:
            om = sp.thinPlateSplineMatrix(xd)
            assert not np.all(np.isfinite(om))
        except SpatpcaError:
            pass


# ---------------------------------------------------------------------------------------------
# determinism, sharding, larger sizes
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["irregular2d_small", "irregular2d_2304"])
def test_omega_column_blocks_match_single_build(sp, name):
    """The multi-GPU Omega build (column blocks of Lp and of Lp'(E Lp)) against the one-piece build: the same
    start, the same right-hand sides, only the GEMM / solve calls are split by columns.  p = 300 takes the LU start,
    p = 2304 the replicated SPD / Cholesky start with the INT8 refinement and product; the SVD starts asked for right
    after omega_begin() (prepare_starts) run under the enqueued first stage."""
    from spatpca_b200 import distributed as D
    case = make_case(name)
    p = case["x"].shape[0]
    whole = sp.thinPlateSplineMatrix(case["x"])
    e = sp.Engine(p, case["x"].shape[1], case["Y"].shape[0], case["M"], case["K"], case["tau1"], case["tau2"],
                  case["gamma"], 100, 1e-4, case["l2"])
    try:
        e.upload(case["x"], case["Y"], case["nk"])
        assert e.omega_begin()
        e.prepare_starts([0, 2])
        cols = D.omega_columns(p, 3)
        for c0, c1 in cols:
            e.omega_solve_cols(c0, c1)
        for c0, c1 in cols:
            e.omega_refine_cols(c0, c1)              # refined build (default): residual + correction per block
        for c0, c1 in cols:
            e.omega_product_cols(c0, c1)
        e.omega_finish()
        assert not e.omega_begin()                     # already there
        e.prepare()
        got = e.get_omega()
    finally:
        e.close()
    iu = np.triu_indices(p)
    assert np.max(np.abs(got[iu] - whole[iu])) / np.max(np.abs(whole)) <= 1e-11
    # the engine keeps symmatu(Omega): the lower triangle is the mirror of the upper one
    il = np.tril_indices(p, -1)
    assert np.array_equal(got[il], got.T[il])


def _run_cv(sp, case, **kw):
    return sp.spatpcaCV(case["x"], case["Y"], case["M"], case["K"], case["tau1"], case["tau2"], case["gamma"],
                        case["nk"], kw.get("maxit", 100), kw.get("tol", 1e-4), case["l2"], return_stats=True)


def _same_bits(a, b):
    for key in ("cv_score_tau1", "cv_score_tau2", "cv_score_gamma", "estimated_eigenfn"):
        assert np.array_equal(a[key], b[key]), key
    for key in ("selected_tau1", "selected_tau2", "selected_gamma"):
        assert a[key] == b[key]


def test_bitwise_reproducible_across_lanes_and_cache(sp, monkeypatch):
    """Results must not depend on how many folds run concurrently, on whether stage-1 inverses are
    reused, or on the run: every reduction has a fixed shape (no floating-point atomics)."""
    case = make_case("irregular2d_small")
    base = _run_cv(sp, case)
    again = _run_cv(sp, case)
    _same_bits(base, again)
    monkeypatch.setenv("SPB_LANES", "1")
    one_lane = _run_cv(sp, case)
    _same_bits(base, one_lane)
    monkeypatch.setenv("SPB_LANES", "2")
    monkeypatch.setenv("SPB_CACHE_INVERSES", "0")
    no_cache = _run_cv(sp, case)
    _same_bits(base, no_cache)
    assert no_cache["stats"]["n_inverses"] > base["stats"]["n_inverses"]


def test_inverse_matches_cusolver_path(sp, monkeypatch):
    """The recursive GEMM-rate inverse against cuSOLVER potrf + potri (the reference's algorithm class)
    on the same device: same matrix to ~cond * eps."""
    import subprocess, sys, os, json
    code = (
        "import sys, json, numpy as np; sys.path.insert(0, '.'); sys.path.insert(0, 'tests');"
        "from synth import make_case; import spatpca_b200 as sp;"
        "c = make_case('grid2d_small');"
        "r = sp.spatpcaCV(c['x'], c['Y'], c['M'], c['K'], c['tau1'], c['tau2'], c['gamma'], c['nk'], 100, 1e-4, c['l2']);"
        "print(json.dumps({'s1': list(r['cv_score_tau1']), 's2': list(r['cv_score_tau2']), 's3': list(r['cv_score_gamma']),"
        "'sel': [r['selected_tau1'], r['selected_tau2'], r['selected_gamma']], 'phi': r['estimated_eigenfn'].ravel().tolist()}))")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for mode in ("recursive", "cusolver"):
        env = dict(os.environ, SPB_INVERSE=mode)
        res = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, check=True)
        outs.append(json.loads(res.stdout.strip().splitlines()[-1]))
    a, b = outs
    assert a["sel"] == b["sel"]
    for k in ("s1", "s2", "s3"):
        np.testing.assert_allclose(a[k], b[k], rtol=1e-11)
    pa, pb = np.array(a["phi"]), np.array(b["phi"])
    assert np.max(np.abs(pa - pb)) / np.max(np.abs(pb)) <= 1e-10


def test_threaded_starts_match_serial(sp):
    """The fold SVD starts on one host thread each (default) against the serial path (SPB_PROLOGUE_THREADS=0): the same
    kernels on the same inputs, so the whole job must agree bit for bit.  The switch is read once per process, hence
    the subprocesses."""
    import subprocess, sys, os, json
    code = (
        "import sys, json, numpy as np; sys.path.insert(0, '.'); sys.path.insert(0, 'tests');"
        "from synth import make_case; import spatpca_b200 as sp;"
        "c = make_case('irregular2d_small');"
        "r = sp.spatpcaCV(c['x'], c['Y'], c['M'], c['K'], c['tau1'], c['tau2'], c['gamma'], c['nk'], 100, 1e-4, c['l2']);"
        "print(json.dumps({'s1': list(r['cv_score_tau1']), 's2': list(r['cv_score_tau2']), 's3': list(r['cv_score_gamma']),"
        "'sel': [r['selected_tau1'], r['selected_tau2'], r['selected_gamma']], 'phi': r['estimated_eigenfn'].ravel().tolist()}))")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for threads in ("1", "0"):
        env = dict(os.environ, SPB_PROLOGUE_THREADS=threads)
        res = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, check=True)
        outs.append(json.loads(res.stdout.strip().splitlines()[-1]))
    assert outs[0] == outs[1]


def test_sharded_two_engines_match_single(sp):
    """Fold sharding as bench.py does under torchrun, emulated with two engines in one process: the
    per-fold rows, the selections and the final fit must equal the single-engine run bit for bit."""
    from spatpca_b200 import distributed as D
    case = make_case("grid3d_small")
    single = _run_cv(sp, case)
    M, world = case["M"], 2
    engines = []
    for r in range(world):
        e = sp.Engine(case["x"].shape[0], case["x"].shape[1], case["Y"].shape[0], M, case["K"], case["tau1"],
                      case["tau2"], case["gamma"], 100, 1e-4, case["l2"])
        e.upload(case["x"], case["Y"], case["nk"])
        e.prepare()
        engines.append(e)
    mine = [[k for k in range(M) if D.fold_owner(k, world) == r] for r in range(world)]
    full = D.full_fit_owner(M, world)

    def gather(per_rank_rows, width):
        cv = np.zeros((M, width))
        for rows in per_rank_rows:
            for k, row in rows.items():
                cv[k] = row
        return cv

    cv1 = gather([engines[r].stage1_folds(mine[r]) for r in range(world)], len(case["tau1"]))
    i1, s1 = sp.select_index(cv1)
    engines[full].stage1_full(i1)
    cv2 = gather([engines[r].stage2_folds(mine[r], i1) for r in range(world)], len(case["tau2"]))
    i2, s2 = sp.select_index(cv2)
    engines[full].stage2_full(i1, i2)
    cv3 = gather([engines[r].stage3_folds(mine[r], i1, i2) for r in range(world)], len(case["gamma"]))
    i3, s3 = sp.select_index(cv3)
    phi = engines[full].get_eigenfn()
    assert np.array_equal(s1, single["cv_score_tau1"])
    assert np.array_equal(s2, single["cv_score_tau2"])
    assert np.array_equal(s3, single["cv_score_gamma"])
    assert np.array_equal(phi, single["estimated_eigenfn"])
    assert case["tau1"][i1] == single["selected_tau1"] and case["gamma"][i3] == single["selected_gamma"]
    for e in engines:
        e.close()


@pytest.mark.parametrize("world", [2, 3, 8])
def test_sharded_matrix_free_worlds(sp, world):
    """Fold sharding on a case whose ADMM calls take the fold-batched matrix-free kernel (p >= 2048), emulated with one
    engine per rank that owns work: world = 2 / 3 batch different numbers of folds per launch, world = 8 has single-fold
    ranks, idle ranks and a full-data fit on a rank WITHOUT folds (its Core2p / tau2 path has no fold batch to ride
    with).  Batches of different width sum in different orders, so the comparison with the single-engine job is to
    the parity tolerances, with identical selections and identical per-rank iteration totals."""
    from spatpca_b200 import distributed as D
    case = make_case("irregular2d_2304")
    single = _run_cv(sp, case)
    assert single["stats"]["n_cg_calls"] > 0            # the case must exercise the matrix-free path
    M = case["M"]
    mine = [[k for k in range(M) if D.fold_owner(k, world) == r] for r in range(world)]
    full = D.full_fit_owner(M, world)
    engines = {}
    for r in range(world):
        if not mine[r] and r != full:
            continue                                     # a rank without work only takes part in the collectives
        e = sp.Engine(case["x"].shape[0], case["x"].shape[1], case["Y"].shape[0], M, case["K"], case["tau1"],
                      case["tau2"], case["gamma"], 100, 1e-4, case["l2"])
        e.upload(case["x"], case["Y"], case["nk"])
        e.prepare()
        engines[r] = e

    def gather(stage, width, *idx):
        cv = np.zeros((M, width))
        for r, e in engines.items():
            if mine[r]:
                for k, row in getattr(e, stage)(mine[r], *idx).items():
                    cv[k] = row
        return cv

    try:
        i1, s1 = sp.select_index(gather("stage1_folds", len(case["tau1"])))
        engines[full].stage1_full(i1)
        i2, s2 = sp.select_index(gather("stage2_folds", len(case["tau2"]), i1))
        engines[full].stage2_full(i1, i2)
        i3, s3 = sp.select_index(gather("stage3_folds", len(case["gamma"]), i1, i2))
        phi = engines[full].get_eigenfn()
        iters = sum(int(e.stats()["admm_iters"]) for e in engines.values())
    finally:
        for e in engines.values():
            e.close()
    assert case["tau1"][i1] == single["selected_tau1"]
    assert case["tau2"][i2] == single["selected_tau2"]
    assert case["gamma"][i3] == single["selected_gamma"]
    np.testing.assert_allclose(s1, single["cv_score_tau1"], rtol=CV_RTOL)
    np.testing.assert_allclose(s2, single["cv_score_tau2"], rtol=CV_RTOL)
    np.testing.assert_allclose(s3, single["cv_score_gamma"], rtol=CV_RTOL)
    assert eig_err(phi, single["estimated_eigenfn"]) <= EIG_RTOL
    assert iters == int(single["stats"]["admm_iters"])


def test_cv_forced_iterations(sp):
    """thr = 0: every ADMM call runs exactly maxit iterations (the deterministic-denominator mode of
    SURVEY.md section 8d); iterates still match the oracle."""
    case = make_case("grid3d_small")
    args = (case["x"], case["Y"], case["M"], case["K"], case["tau1"], case["tau2"], case["gamma"], case["nk"], 7, 0.0,
            case["l2"])
    got = sp.spatpcaCV(*args, return_stats=True)
    tr = ref.Trace()
    want = ref.spatpca_cv(*args, trace=tr)
    assert got["stats"]["admm_iters"] == 7 * got["stats"]["n_admm_calls"] == tr.admm_iters
    assert_cv_parity(got, want)


def test_cv_irregular_1500(sp):
    """Mid-size irregular 2-D case (p = 1500, n = 200, K = 5, 10 x 10 x 11 grids): the largest size the
    oracle finishes in well under a minute."""
    case = make_case("irregular2d_1500")
    got, want, tr = run_both(sp, case)
    assert_cv_parity(got, want)
    assert got["stats"]["admm_iters"] == tr.admm_iters


def test_session_reuse_is_bitwise_identical(sp):
    """spb_cv_session keeps Omega, SVD starts and inverses between calls that differ only in K; every
    result must equal the stateless spb_cv bit for bit, in any order of K, and fewer inverses are built."""
    from spatpca_b200 import engine as eng
    case = make_case("irregular2d_small")
    args = lambda K: (case["x"], case["Y"], case["M"], K, case["tau1"], case["tau2"], case["gamma"], case["nk"],
                      100, 1e-4, case["l2"])
    fresh = {K: sp.spatpcaCV(*args(K), return_stats=True) for K in (1, 2, 4)}
    eng.session_clear()
    built = []
    for K in (1, 4, 2, 4):
        r = sp.spatpcaCV(*args(K), return_stats=True, session=True)
        _same_bits(fresh[K], r)
        built.append(r["stats"]["n_inverses"])
    assert built[0] == fresh[1]["stats"]["n_inverses"]
    assert built[1] < built[0] and built[3] <= built[1]
    # a different data set must not hit the session
    other = make_case("grid2d_small")
    r = sp.spatpcaCV(other["x"], other["Y"], other["M"], other["K"], other["tau1"], other["tau2"], other["gamma"],
                     other["nk"], 100, 1e-4, other["l2"], return_stats=True, session=True)
    want = sp.spatpcaCV(other["x"], other["Y"], other["M"], other["K"], other["tau1"], other["tau2"], other["gamma"],
                        other["nk"], 100, 1e-4, other["l2"], return_stats=True)
    _same_bits(want, r)
    eng.session_clear()


def test_engine_reset_k(sp):
    case = make_case("grid3d_small")
    e = sp.Engine(case["x"].shape[0], case["x"].shape[1], case["Y"].shape[0], case["M"], 1, case["tau1"],
                  case["tau2"], case["gamma"], 100, 1e-4, case["l2"])
    e.upload(case["x"], case["Y"], case["nk"])
    folds = list(range(case["M"]))
    for K in (1, 3, 2):
        e.reset_k(K)
        e.prepare()
        r1 = e.stage1_folds(folds)
        i1, s1 = sp.select_index(np.array([r1[k] for k in folds]))
        e.stage1_full(i1)
        r2 = e.stage2_folds(folds, i1)
        i2, s2 = sp.select_index(np.array([r2[k] for k in folds]))
        e.stage2_full(i1, i2)
        r3 = e.stage3_folds(folds, i1, i2)
        i3, s3 = sp.select_index(np.array([r3[k] for k in folds]))
        phi = e.get_eigenfn()
        want = sp.spatpcaCV(case["x"], case["Y"], case["M"], K, case["tau1"], case["tau2"], case["gamma"], case["nk"],
                            100, 1e-4, case["l2"])
        assert phi.shape == (case["x"].shape[0], K)
        assert np.array_equal(phi, want["estimated_eigenfn"])
        assert np.array_equal(s1, want["cv_score_tau1"]) and np.array_equal(s3, want["cv_score_gamma"])
    e.close()


@pytest.mark.parametrize("p,K", [(125, 1), (77, 3), (129, 5), (255, 7)])
def test_core3_odd_sizes(sp, p, K):
    """Odd p with odd p*K: every workspace sub-buffer must stay 16-byte aligned."""
    from spatpca_b200 import engine as eng
    Yt, G, Om, rho, Phi0, L20 = _spd_problem(p, K, 11 * p + K, n=max(40, K + 5))
    tinv = ref.core3_tempinv(Om, G, 0.05, rho)
    tr = ref.Trace()
    want = ref.core3(tinv, Phi0, Phi0.copy(), Phi0, np.zeros_like(Phi0), L20, 1.0, rho, 100, 1e-4, tr)
    got = eng.admm_core3(tinv, Phi0.copy(), Phi0, np.zeros_like(Phi0), L20, 1.0, rho, 100, 1e-4)
    assert got[5] == tr.calls[0][3]
    for g, w in zip(got[:5], want):
        assert np.max(np.abs(g - w)) / max(np.max(np.abs(w)), 1e-300) <= 1e-10


def test_stage3_strided_fold_list_with_two_lanes(sp, monkeypatch):
    """A rank's fold list is strided (k % world == rank): with fewer lanes than folds two of them share a lane and
    its chain buffers.  Stage 3 must run them one after the other (it grouped folds by list position)."""
    case = make_case("grid3d_small")
    M = case["M"]
    mk = lambda: sp.Engine(case["x"].shape[0], case["x"].shape[1], case["Y"].shape[0], M, case["K"], case["tau1"],
                           case["tau2"], case["gamma"], 100, 1e-4, case["l2"])

    def run(folds_groups):
        e = mk()
        e.upload(case["x"], case["Y"], case["nk"])
        e.prepare()
        r1 = {}
        for g in folds_groups:
            r1.update(e.stage1_folds(g))
        i1, _ = sp.select_index(np.array([r1[k] for k in range(M)]))
        r2 = {}
        for g in folds_groups:
            r2.update(e.stage2_folds(g, i1))
        i2, _ = sp.select_index(np.array([r2[k] for k in range(M)]))
        r3 = {}
        for g in folds_groups:
            r3.update(e.stage3_folds(g, i1, i2))
        e.close()
        return np.array([r3[k] for k in range(M)])

    base = run([list(range(M))])
    monkeypatch.setenv("SPB_LANES", "2")
    strided = run([[0, 2, 4], [1, 3]])
    assert np.array_equal(base, strided)


# ---------------------------------------------------------------------------------------------
# BASELINE.json configs at full size (C2, C3, C4) against the oracle, in the driver-run suite
# ---------------------------------------------------------------------------------------------
def _full_config(sp, name):
    case = make_case(name)
    got, want, tr = run_both(sp, case)
    return case, got, want, tr


@pytest.mark.parametrize("name", ["C2", "C2_tau2"])
def test_cv_config2_full(sp, name):
    """BASELINE config 2: 30 x 30 grid, p = 900, n = 100, K = 3, default tau1 grid (and the vignette's tau2 grid)."""
    case, got, want, tr = _full_config(sp, name)
    assert_cv_parity(got, want)
    assert got["stats"]["n_admm_calls"] == len(tr.calls)
    assert got["stats"]["admm_iters"] == tr.admm_iters


@pytest.fixture(scope="module")
def c3_runs(sp):
    return _full_config(sp, "C3")


def test_cv_config3_full(sp, c3_runs):
    """BASELINE config 3: p = 4096 irregular 2-D, n = 200, K = 5, 10 x 10 x 11 grids (oracle: ~40 s of CPU)."""
    case, got, want, tr = c3_runs
    assert_cv_parity(got, want)
    assert got["stats"]["n_admm_calls"] == len(tr.calls)
    assert got["stats"]["admm_iters"] == tr.admm_iters


def _injected(sp, case, om):
    e = sp.Engine(case["x"].shape[0], case["x"].shape[1], case["Y"].shape[0], case["M"], case["K"],
                  case["tau1"], case["tau2"], case["gamma"], 100, 1e-4, case["l2"])
    try:
        e.set_omega(om)
        e.upload(case["x"], case["Y"], case["nk"])
        e.prepare()
        M = case["M"]
        folds = list(range(M))
        r1 = e.stage1_folds(folds)
        i1, s1 = sp.select_index(np.array([r1[k] for k in folds]))
        e.stage1_full(i1)
        r2 = e.stage2_folds(folds, i1)
        i2, s2 = sp.select_index(np.array([r2[k] for k in folds]))
        e.stage2_full(i1, i2)
        r3 = e.stage3_folds(folds, i1, i2)
        i3, s3 = sp.select_index(np.array([r3[k] for k in folds]))
        return dict(s1=s1, s2=s2, s3=s3, phi=e.get_eigenfn(), idx=(i1, i2, i3), log=e.call_log())
    finally:
        e.close()


def test_cv_config3_injected_omega(sp, c3_runs):
    """The same job with the ORACLE's Omega on both sides: everything after thinPlateSplineMatrix to 1e-10."""
    case, _, want, tr = c3_runs
    r = _injected(sp, case, want["aux"]["omega"])
    np.testing.assert_allclose(r["s1"], want["cv_score_tau1"], rtol=1e-11)
    np.testing.assert_allclose(r["s2"], want["cv_score_tau2"], rtol=1e-11)
    np.testing.assert_allclose(r["s3"], want["cv_score_gamma"], rtol=1e-11)
    assert eig_err(r["phi"], want["estimated_eigenfn"]) <= 1e-10
    assert sum(c[3] for c in r["log"]) == tr.admm_iters


@pytest.fixture(scope="module")
def c4_runs(sp):
    return _full_config(sp, "C4")


def test_cv_config4_full(sp, c4_runs):
    """BASELINE config 4, the headline: p = 10 000, n = 200, K = 5, 11 x 11 x 11 grids (oracle: ~3 min of CPU)."""
    case, got, want, tr = c4_runs
    assert_cv_parity(got, want)
    assert got["stats"]["n_admm_calls"] == len(tr.calls)
    assert got["stats"]["admm_iters"] == tr.admm_iters


def test_cv_config4_injected_omega(sp, c4_runs):
    case, _, want, tr = c4_runs
    r = _injected(sp, case, want["aux"]["omega"])
    np.testing.assert_allclose(r["s1"], want["cv_score_tau1"], rtol=1e-11)
    np.testing.assert_allclose(r["s2"], want["cv_score_tau2"], rtol=1e-11)
    np.testing.assert_allclose(r["s3"], want["cv_score_gamma"], rtol=1e-11)
    assert eig_err(r["phi"], want["estimated_eigenfn"]) <= 1e-10
    assert sum(c[3] for c in r["log"]) == tr.admm_iters
