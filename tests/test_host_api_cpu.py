"""The host-side mirror of the reference's R layer (spatpca_b200/api.py): validators, default tuning
grids and clamping rules, against the reference's own testthat expectations (tests/testthat/test-helper.R).
CPU only; nothing here launches a kernel."""
import os
import warnings

import numpy as np
import pytest

import spatpca_b200 as sp
from oracle.r_rng import r_seq_len

from ref_cases import case_1d


def test_set_cores():                                   # test-helper.R:5-16
    with pytest.raises(ValueError, match="Please enter valid type - but got"):
        sp.setCores("test")
    with pytest.raises(ValueError, match="The number of cores is not greater than 1 - but got 0"):
        sp.setCores(0)
    n = os.cpu_count() or 1
    with pytest.raises(ValueError, match=f"The input number of cores is invalid - default is {n}"):
        sp.setCores(n + 1)
    assert sp.setCores(n) is True
    assert sp.setCores() is None


def test_scale_locations():                             # test-helper.R:20-25
    x1 = r_seq_len(-5, 5, 10).reshape(-1, 1)
    s = sp.scaleLocation(x1)
    assert abs(s.sum() - 5) < 1e-12 and s.min() == 0 and s.max() == 1
    x2 = np.array([[1.0, 2.0]])
    assert np.array_equal(sp.scaleLocation(x2), x2)


def test_check_input():                                 # test-helper.R:62-79
    _, x4, Y4 = case_1d(p=4)
    with pytest.raises(ValueError, match="The number of rows of x should be equal"):
        sp.checkInputData(Y4, np.ones((1, 1)), 5)
    with pytest.raises(ValueError, match="Number of locations must be larger than 2"):
        sp.checkInputData(np.ones((1, 1)), np.arange(10.0).reshape(1, 10), 5)
    with pytest.raises(ValueError, match="Dimension of locations must be less than 4"):
        sp.checkInputData(Y4, np.arange(40.0).reshape(4, 10), 5)
    with pytest.raises(ValueError, match="Number of folds must be less than sample size"):
        sp.checkInputData(Y4, x4, 1000)


def test_detrend_and_k_bounds():                        # test-helper.R:82-93
    _, _, Y4 = case_1d(p=4)
    assert np.array_equal(sp.detrend(Y4, False), Y4)
    assert abs(sp.detrend(Y4, True).sum()) <= 1e-6
    assert sp.fetchUpperBoundNumberEigenfunctions(Y4, 5) == 4
    assert sp.setNumberEigenfunctions(None, Y4, 5) is None
    with pytest.warns(UserWarning):
        assert sp.setNumberEigenfunctions(300, Y4, 5) == 4
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        assert sp.setNumberEigenfunctions(3, Y4, 5) == 3


def test_default_grids():                               # test-helper.R:95-120
    _, _, Y4 = case_1d(p=4)
    t1 = sp.setTau1(None, 5)
    assert t1.min() == 0 and abs(t1.max() - 1) < 1e-15 and abs(np.median(t1) - 0.0004641589) < 1e-9
    assert sp.setTau1(None, 1)[0] == t1.max()
    assert list(sp.setTau1([1, 2], 5)) == [1, 2] and list(sp.setTau1([1, 2], 1)) == [2]
    assert list(sp.setTau2(None, 5)) == [0] and list(sp.setTau2(None, 1)) == [0]
    assert list(sp.setTau2([1, 2], 5)) == [1, 2] and list(sp.setTau2([1, 2], 1)) == [2]
    assert list(sp.setL2([1, 2])) == [1] and list(sp.setL2(-1)) == [1]
    l2 = sp.setL2(1)
    assert abs(l2.max() - 1) < 1e-15 and l2.min() == 0 and abs(np.median(l2) - 0.005994843) <= 1e-6
    assert list(sp.setGamma([1, 2])) == [1, 2]
    g = sp.setGamma(None, Y4)
    assert abs(g.max() - 11.14708) <= 1e-5 and g.min() == 0 and abs(np.median(g) - 0.06682497) <= 1e-6


def test_new_location_checks():                         # test-helper.R:44-59
    obj = sp.SpatpcaObject(scaled_x=np.zeros((4, 1)), eigenfn=np.zeros((4, 1)))
    with pytest.raises(ValueError, match="Invalid object"):
        sp.checkNewLocationsForSpatpcaObject(None, None)
    with pytest.raises(ValueError, match="New locations cannot be NULL"):
        sp.checkNewLocationsForSpatpcaObject(obj, None)
    with pytest.raises(ValueError, match="Inconsistent dimension of locations - original dimension is 1"):
        sp.checkNewLocationsForSpatpcaObject(obj, np.array([[1.0, 2.0]]))
    assert sp.checkNewLocationsForSpatpcaObject(obj, r_seq_len(6, 7, 4).reshape(-1, 1)) is None


def test_synthetic_generator_is_seeded():
    from spatpca_b200.synth import make_case
    a, b = make_case("grid2d_small"), make_case("grid2d_small")
    assert np.array_equal(a["Y"], b["Y"]) and np.array_equal(a["nk"], b["nk"])
    assert a["Y"].flags.f_contiguous and sorted(set(a["nk"])) == [1, 2, 3, 4, 5]
