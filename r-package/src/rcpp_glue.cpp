// Rcpp glue: re-exports the reference's four `.Call` symbols on top of the spatpca_b200 C ABI.
//
// Replaces src/RcppSpatPCA.cpp + src/RcppExports.cpp of SpatPCA 1.3.8 for the hot path: the symbol
// names, arities and routine registration are those of src/RcppExports.cpp:16-85, the argument
// meaning and the returned lists are those of src/RcppSpatPCA.cpp:58, 94, 543-551/694-700, 715-718/766-769.
// All numerical work happens in libspatpca_b200.so; this file only marshals SEXPs and maps negative
// statuses onto R conditions (the job BEGIN_RCPP / END_RCPP did for Armadillo exceptions).
//
// NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no R, Rcpp or R headers (SURVEY.md
// section 8b).  The same ABI is exercised from Python (spatpca_b200/_lib.py) by the test-suite.
#include <Rcpp.h>

#include "spatpca_b200.h"

using namespace Rcpp;

namespace {
inline void check(int rc) {
  if (rc < 0) Rcpp::stop("%s", spb_last_error());
}
}  // namespace

// thinPlateSplineMatrix(location)                                  reference: src/RcppSpatPCA.cpp:58
RcppExport SEXP _SpatPCA_thinPlateSplineMatrix(SEXP locationSEXP) {
BEGIN_RCPP
  NumericMatrix location(locationSEXP);
  const int p = location.nrow(), d = location.ncol();
  NumericMatrix omega(p, p);
  check(spb_tps_matrix(location.begin(), p, d, omega.begin()));
  return omega;
END_RCPP
}

// eigenFunction(new_location, original_location, Phi)              reference: src/RcppSpatPCA.cpp:94
RcppExport SEXP _SpatPCA_eigenFunction(SEXP new_locationSEXP, SEXP original_locationSEXP, SEXP PhiSEXP) {
BEGIN_RCPP
  NumericMatrix xn(new_locationSEXP), xo(original_locationSEXP), Phi(PhiSEXP);
  if (xn.ncol() != xo.ncol()) Rcpp::stop("Inconsistent dimension of locations");
  NumericMatrix out(xn.nrow(), Phi.ncol());
  check(spb_eigen_function(xn.begin(), xn.nrow(), xo.begin(), xo.nrow(), xo.ncol(), Phi.begin(), Phi.ncol(),
                           out.begin()));
  return out;
END_RCPP
}

// spatpcaCV(sxyr, Yr, M, K, tau1r, tau2r, gammar, nkr, maxit, tol, l2r)   reference: src/RcppSpatPCA.cpp:543
RcppExport SEXP _SpatPCA_spatpcaCV(SEXP sxyrSEXP, SEXP YrSEXP, SEXP MSEXP, SEXP KSEXP, SEXP tau1rSEXP,
                                   SEXP tau2rSEXP, SEXP gammarSEXP, SEXP nkrSEXP, SEXP maxitSEXP, SEXP tolSEXP,
                                   SEXP l2rSEXP) {
BEGIN_RCPP
  NumericMatrix sxy(sxyrSEXP), Y(YrSEXP);
  NumericVector tau1(tau1rSEXP), tau2(tau2rSEXP), gamma(gammarSEXP), nk(nkrSEXP), l2(l2rSEXP);
  const int M = as<int>(MSEXP), K = as<int>(KSEXP), maxit = as<int>(maxitSEXP);
  const double tol = as<double>(tolSEXP);
  const int n = Y.nrow(), p = Y.ncol(), d = sxy.ncol();
  const int n1 = tau1.size(), n2 = tau2.size(), n3 = gamma.size();
  // shapes of the reference's return value (:600/:649, :668/:679, :692, :697)
  NumericMatrix cv1(1, n1 > 1 ? n1 : 1), cv2(1, n2 > 1 ? n2 : 1), cv3(1, n3), eigenfn(p, K);
  double sel[3] = {0, 0, 0};
  // spb_cv_session, not the stateless spb_cv: R's K-selection loop (R/SpatPCA.R:35-39) calls this entry for
  // K = 1, 2, ... on the SAME data and grids, and the session keeps Omega, the fold SVD starts and the K-independent
  // factorisations between such calls.  The session is keyed by the CONTENT of every argument except K (data,
  // locations, fold ids, the three grids and their lengths, maxit, tol), so a call with anything else rebuilds;
  // results are bitwise those of spb_cv (tests/test_gpu_parity.py::test_session_reuse_is_bitwise_identical).
  check(spb_cv_session(sxy.begin(), p, d, Y.begin(), n, M, K, tau1.begin(), n1, tau2.begin(), n2, gamma.begin(), n3,
                       nk.begin(), maxit, tol, l2.begin(), (int)l2.size(), cv1.begin(), cv2.begin(), cv3.begin(),
                       eigenfn.begin(), sel, NULL));
  return List::create(Named("cv_score_tau1") = cv1, Named("cv_score_tau2") = cv2, Named("cv_score_gamma") = cv3,
                      Named("estimated_eigenfn") = eigenfn, Named("selected_tau1") = sel[0],
                      Named("selected_tau2") = sel[1], Named("selected_gamma") = sel[2]);
END_RCPP
}

// spatialPrediction(phir, Yr, gamma, predicted_eignefunction)      reference: src/RcppSpatPCA.cpp:715
RcppExport SEXP _SpatPCA_spatialPrediction(SEXP phirSEXP, SEXP YrSEXP, SEXP gammaSEXP,
                                           SEXP predicted_eignefunctionSEXP) {
BEGIN_RCPP
  NumericMatrix phi(phirSEXP), Y(YrSEXP), pphi(predicted_eignefunctionSEXP);
  const double gamma = as<double>(gammaSEXP);
  const int p = phi.nrow(), K = phi.ncol(), n = Y.nrow(), p2 = pphi.nrow();
  NumericMatrix prediction(n, p2), cov(p, p2);
  NumericVector eigenvalue(K);
  double error = 0.0;
  check(spb_spatial_prediction(phi.begin(), p, K, Y.begin(), n, gamma, pphi.begin(), p2, prediction.begin(),
                               cov.begin(), eigenvalue.begin(), &error));
  return List::create(Named("prediction") = prediction, Named("estimated_covariance") = cov,
                      Named("eigenvalue") = eigenvalue, Named("error") = error);
END_RCPP
}

// same registration table as src/RcppExports.cpp:74-85
static const R_CallMethodDef CallEntries[] = {
    {"_SpatPCA_thinPlateSplineMatrix", (DL_FUNC)&_SpatPCA_thinPlateSplineMatrix, 1},
    {"_SpatPCA_eigenFunction", (DL_FUNC)&_SpatPCA_eigenFunction, 3},
    {"_SpatPCA_spatpcaCV", (DL_FUNC)&_SpatPCA_spatpcaCV, 11},
    {"_SpatPCA_spatialPrediction", (DL_FUNC)&_SpatPCA_spatialPrediction, 4},
    {NULL, NULL, 0}};

RcppExport void R_init_SpatPCA(DllInfo* dll) {
  R_registerRoutines(dll, NULL, CallEntries, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}

// library.dynam.unload(): give the session's device memory (Omega, cached factors) back before the DLL goes away
RcppExport void R_unload_SpatPCA(DllInfo* dll) {
  (void)dll;
  spb_session_clear();
  spb_trim_memory();
}
