"""Pins the CPU oracle against every golden the reference's testthat files hold
for the hot path.  CPU only."""
import numpy as np
import pytest

from oracle import r_api
from oracle import spatpca_ref as ref
from oracle.r_rng import RRng, r_seq_len

import json
import os

from ref_cases import case_1d, case_3d, tps_locations

with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_goldens.json")) as _f:
    GOLD = json.load(_f)          # the reference's own known-answer values, each with its testthat file:line
TOL = GOLD["tolerance"]           # the reference's own tolerance (test-SpatPCA.R:3)


def gold(name):
    return GOLD[name]["value"]


def test_r_rng_known_values():
    r = RRng(1234)
    np.testing.assert_allclose(r.rnorm(3), GOLD["r_rng"]["set_seed_1234_rnorm3"], atol=5e-8)
    r = RRng(42)
    assert list(r.sample(np.arange(1, 11))) == GOLD["r_rng"]["set_seed_42_sample_1_10"]


def test_tps_matrix_goldens():                       # test-RcppExports.R:10-15
    l2d, l3d = tps_locations()
    assert abs(np.linalg.norm(ref.thin_plate_spline_matrix(l2d)) - gold("tps_matrix_frobenius_norm_2d")) <= TOL
    assert abs(np.linalg.norm(ref.thin_plate_spline_matrix(l3d)) - gold("tps_matrix_frobenius_norm_3d")) <= TOL


def test_eigen_function_golden():                    # test-RcppExports.R:18-23
    l2d, _ = tps_locations()
    v = ref.eigen_function(np.array([[0.1, 0.2]]), l2d, np.array([[1.0], [0], [0], [0]]))
    assert abs(v[0, 0] - gold("eigen_function_value")) <= TOL


@pytest.fixture(scope="module")
def seq_1d():
    """The whole call sequence of test-SpatPCA.R:13-88 on one shared RNG stream."""
    rng, x, Y = case_1d()
    out = {"x": x, "Y": Y}
    out["cv_1D"] = r_api.spatpca(x, Y, rng)
    out["multi_tau2"] = r_api.spatpca(x, Y, rng, K=1, tau2=[0, 1])
    out["multi_gamma"] = r_api.spatpca(x, Y, rng, K=1, gamma=[0, 1])
    out["fixed_both"] = r_api.spatpca(x, Y, rng, K=1, tau1=10, tau2=100)
    out["zero_multi_gamma"] = r_api.spatpca(x, Y, rng, K=1, tau1=0, tau2=0, gamma=[0, 0.5, 1])
    out["fixed_tau1"] = r_api.spatpca(x, Y, rng, K=1, tau1=10)
    out["zero_zero_K5"] = r_api.spatpca(x, Y, rng, K=5, tau1=0, tau2=0)
    return out


def test_selected_tuning_parameters(seq_1d):         # test-SpatPCA.R:97-122
    s = seq_1d
    assert abs(s["cv_1D"]["selected_tau1"] - gold("cv_1D_selected_tau1")) <= TOL
    assert abs(s["cv_1D"]["selected_tau2"]) <= TOL
    assert abs(s["cv_1D"]["selected_gamma"] - gold("cv_1D_selected_gamma")) <= 1.0001 * TOL
    assert s["cv_1D"]["selected_K"] is None
    assert s["cv_1D"]["eigenfn"].shape == (10, 2)     # quirk: fit for selected_K + 1
    assert s["multi_tau2"]["selected_K"] == 1
    assert abs(s["multi_tau2"]["selected_tau1"] - gold("multi_tau2_selected_tau1")) <= TOL
    assert s["multi_tau2"]["selected_tau2"] == gold("multi_tau2_selected_tau2")
    assert s["multi_gamma"]["selected_gamma"] == 0
    assert s["fixed_both"]["selected_tau1"] == 10
    assert s["fixed_both"]["selected_tau2"] == 100
    assert s["zero_multi_gamma"]["selected_gamma"] == 0
    assert s["fixed_tau1"]["selected_tau1"] == 10
    assert np.sum(s["fixed_tau1"]["cv_score_tau1"]) == 0


def test_spatial_prediction_eigenvalues(seq_1d):     # test-SpatPCA.R:52-88, 123-127
    f = seq_1d["fixed_tau1"]
    z = seq_1d["zero_zero_K5"]
    sp = ref.spatial_prediction
    assert sp(f["eigenfn"], f["detrended_Y"], 1000, f["eigenfn"])["eigenvalue"].sum() == 0
    assert sp(z["eigenfn"], z["detrended_Y"], 10, z["eigenfn"])["eigenvalue"].sum() == 0
    assert abs(sp(z["eigenfn"], z["detrended_Y"], 0.5, z["eigenfn"])["eigenvalue"].sum()
               - gold("zero_zero_K5_eigenvalue_sum_gamma_0.5")) <= TOL
    # one-sided in the reference (expect_lte without abs)
    assert sp(f["eigenfn"], f["detrended_Y"], 5, f["eigenfn"])["eigenvalue"].sum() - gold("fixed_tau1_eigenvalue_sum_gamma_5_upper") <= TOL


def test_predict_shapes(seq_1d):                     # test-SpatPCA.R:138-146
    xn = r_seq_len(6, 7, 4).reshape(-1, 1)
    assert r_api.predict(seq_1d["cv_1D"], xn).shape[1] == 4
    assert r_api.predict_eigenfunction(seq_1d["cv_1D"], xn).shape[0] == 4


def test_k_selection_aux():                          # test-SpatPCA.R:150-165
    _, x, Y = case_1d()
    rng = RRng(1234)
    M = 3
    split = rng.sample(np.resize(np.arange(1, M + 1), 100))
    tau2 = r_api.set_tau2(None, M)
    r = r_api.spatpca_cv_with_selected_k(x, Y, M, r_api.set_tau1(None, M), tau2, [1.0], split,
                                         10, 1e-4, r_api.set_l2(tau2))
    k = GOLD["k_selection"]
    assert r["selected_K"] == k["selected_K"]
    assert r["cv_result"]["selected_gamma"] == k["selected_gamma"]
    assert r["cv_result"]["selected_tau1"] == k["selected_tau1"]
    assert r["cv_result"]["selected_tau2"] == k["selected_tau2"]


def test_3d_case():                                  # test-SpatPCA.R:169-187
    rng, d, Y = case_3d()
    cv = r_api.spatpca(d, Y, rng)
    assert cv["eigenfn"].shape == (64, 2)
    pred = ref.eigen_function(np.zeros((1, 3)), d, cv["eigenfn"])
    assert np.sum(np.abs(pred[0] - np.array(gold("predict_3d_at_origin")))) <= TOL


def test_helper_grids():                             # test-helper.R:20-25, 85-120
    _, x4, Y4 = case_1d(p=4)
    x = r_seq_len(-5, 5, 10).reshape(-1, 1)
    sx = r_api.scale_location(x)
    assert abs(sx.sum() - 5) < 1e-12 and sx.min() == 0 and sx.max() == 1
    assert r_api.fetch_upper_bound_number_eigenfunctions(Y4, 5) == 4
    t1 = r_api.set_tau1(None, 5)
    assert t1.min() == 0 and abs(t1.max() - 1) < 1e-15 and abs(np.median(t1) - 0.0004641589) < 1e-9
    assert r_api.set_tau1(None, 1)[0] == t1.max()
    assert list(r_api.set_tau1([1, 2], 5)) == [1, 2] and list(r_api.set_tau1([1, 2], 1)) == [2]
    assert list(r_api.set_tau2(None, 5)) == [0] and list(r_api.set_tau2([1, 2], 1)) == [2]
    assert list(r_api.set_l2([1, 2])) == [1] and list(r_api.set_l2(-1)) == [1]
    l2 = r_api.set_l2(1)
    assert abs(l2.max() - 1) < 1e-15 and l2.min() == 0 and abs(np.median(l2) - 0.005994843) <= TOL
    g = r_api.set_gamma(None, Y4)
    assert abs(g.max() - 11.14708) <= 1e-5 and g.min() == 0 and abs(np.median(g) - 0.06682497) <= TOL
    assert abs(r_api.detrend(Y4, True).sum()) <= TOL
