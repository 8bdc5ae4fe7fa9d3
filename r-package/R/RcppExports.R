# `.Call` stubs, identical in name and arity to R/RcppExports.R:14-68 of SpatPCA 1.3.8.
# R/SpatPCA.R, R/helper.R, R/zzz.R and NAMESPACE of the reference are used unchanged on top of these.

thinPlateSplineMatrix <- function(location) {
    .Call(`_SpatPCA_thinPlateSplineMatrix`, location)
}

eigenFunction <- function(new_location, original_location, Phi) {
    .Call(`_SpatPCA_eigenFunction`, new_location, original_location, Phi)
}

spatpcaCV <- function(sxyr, Yr, M, K, tau1r, tau2r, gammar, nkr, maxit, tol, l2r) {
    .Call(`_SpatPCA_spatpcaCV`, sxyr, Yr, M, K, tau1r, tau2r, gammar, nkr, maxit, tol, l2r)
}

spatialPrediction <- function(phir, Yr, gamma, predicted_eignefunction) {
    .Call(`_SpatPCA_spatialPrediction`, phir, Yr, gamma, predicted_eignefunction)
}
