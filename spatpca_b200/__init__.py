"""spatpca_b200 -- B200-native cross-validated regularized-PCA engine (drop-in for the hot path of
the R package SpatPCA).  The compute lives in `libspatpca_b200.so` (CUDA, sm_100a) behind the C ABI
of `include/spatpca_b200.h`; this package is the Python-side binding and the mirror of the R API."""
from .api import (  # noqa: F401
    spatpca, spatpcaCV, spatpcaCVWithSelectedK, thinPlateSplineMatrix, eigenFunction,
    spatialPrediction, predict, predictEigenfunction, setTau1, setTau2, setL2, setGamma,
    setNumberEigenfunctions, fetchUpperBoundNumberEigenfunctions, setCores, scaleLocation,
    detrend, checkInputData, checkNewLocationsForSpatpcaObject, SpatpcaObject,
)
from .engine import Engine, select_index  # noqa: F401
from ._lib import SpatpcaError, load as load_library  # noqa: F401

__version__ = "0.1.0"
