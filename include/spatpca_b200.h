/*
 * spatpca_b200 -- C ABI of the B200-native cross-validated regularized-PCA engine.
 *
 * This is the drop-in boundary for the hot path of the R package SpatPCA
 * (reference: src/RcppSpatPCA.cpp, bound to R through src/RcppExports.cpp:16-85
 * and R/RcppExports.R:14-68).  Every entry point below names the reference
 * interface it replaces.  The Rcpp glue that re-exports the reference's four
 * `.Call` symbols on top of this ABI is in r-package/src/rcpp_glue.cpp; the
 * ctypes binding used by the tests is spatpca_b200/_lib.py.
 *
 * Conventions
 *   - every pointer is a HOST pointer to caller-owned memory unless its name
 *     ends in `_dev`;
 *   - every matrix is column-major FP64 with leading dimension = its row count,
 *     exactly as R / Armadillo hand them over (src/RcppSpatPCA.cpp:552-558);
 *   - every function returns SPB_OK (0) or a negative status; no C++ exception
 *     crosses the ABI.  `spb_last_error()` returns a thread-local message for
 *     the last failing call on the calling thread (the text the glue passes to
 *     Rcpp::stop, mirroring END_RCPP at src/RcppExports.cpp:23,36,57,71);
 *   - there is no CPU fallback: without a CUDA device every compute entry
 *     point fails with SPB_ERR_NO_DEVICE.
 */
#ifndef SPATPCA_B200_H
#define SPATPCA_B200_H

#if defined(SPB_BUILDING) && defined(__GNUC__)
#define SPB_API __attribute__((visibility("default")))
#else
#define SPB_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define SPB_ABI_VERSION 2

/* status codes */
#define SPB_OK                 0
#define SPB_ERR_INVALID_ARG   -1   /* bad size / NULL pointer / unsupported d or K            */
#define SPB_ERR_NO_DEVICE     -2   /* no usable CUDA device                                   */
#define SPB_ERR_CUDA          -3   /* CUDA / cuBLAS / cuSOLVER runtime failure                */
#define SPB_ERR_NOT_POSDEF    -4   /* inv_sympd(): matrix is singular or not positive definite */
#define SPB_ERR_SINGULAR      -5   /* inv()/solve(): matrix is singular                        */
#define SPB_ERR_POLAR         -6   /* Stiefel projection met a rank-deficient p x K iterate    */
#define SPB_ERR_SVD           -7   /* SVD / symmetric eigensolver failed to converge           */
#define SPB_ERR_INTERNAL      -8

#define SPB_MAX_K             32   /* largest number of eigenfunctions the kernels are built for */

/* ------------------------------------------------------------------------ */
/* library / device                                                          */
/* ------------------------------------------------------------------------ */
SPB_API int          spb_abi_version(void);
SPB_API const char*  spb_last_error(void);
SPB_API int          spb_device_count(void);           /* < 0 on error */
SPB_API int          spb_set_device(int ordinal);      /* device used by subsequent calls of this thread */
/* The library keeps released device buffers for reuse (cudaMalloc/cudaFree synchronise the device);
 * this returns them to the driver. */
SPB_API int          spb_trim_memory(void);

/* ------------------------------------------------------------------------ */
/* thinPlateSplineMatrix(location)                                           */
/*   reference: src/RcppSpatPCA.cpp:58-76, stub R/RcppExports.R:14-16,       */
/*              symbol _SpatPCA_thinPlateSplineMatrix (src/RcppExports.cpp:16)*/
/*   location : p x d, d in {1,2,3};  omega_out : p x p (NOT symmetrized,    */
/*   like the reference's return value)                                      */
/* ------------------------------------------------------------------------ */
SPB_API int spb_tps_matrix(const double* location, int p, int d, double* omega_out);

/* ------------------------------------------------------------------------ */
/* eigenFunction(new_location, original_location, Phi)                       */
/*   reference: src/RcppSpatPCA.cpp:94-147, stub R/RcppExports.R:33-35,      */
/*              symbol _SpatPCA_eigenFunction (src/RcppExports.cpp:27)       */
/*   out : p_new x K                                                         */
/* ------------------------------------------------------------------------ */
SPB_API int spb_eigen_function(const double* new_location, int p_new,
                       const double* original_location, int p, int d,
                       const double* Phi, int K, double* out);

/* ------------------------------------------------------------------------ */
/* spatialPrediction(phir, Yr, gamma, predicted_eignefunction)               */
/*   reference: src/RcppSpatPCA.cpp:715-770, stub R/RcppExports.R:66-68,     */
/*              symbol _SpatPCA_spatialPrediction (src/RcppExports.cpp:61)   */
/*   prediction : n x p_new; estimated_covariance : p x p_new;               */
/*   eigenvalue : K; error : 1.  Any output pointer may be NULL.             */
/* ------------------------------------------------------------------------ */
SPB_API int spb_spatial_prediction(const double* phi, int p, int K,
                           const double* Y, int n, double gamma,
                           const double* predicted_phi, int p_new,
                           double* prediction, double* estimated_covariance,
                           double* eigenvalue, double* error);

/* ------------------------------------------------------------------------ */
/* spatpcaCV(sxyr, Yr, M, K, tau1r, tau2r, gammar, nkr, maxit, tol, l2r)     */
/*   reference: src/RcppSpatPCA.cpp:543-701, stub R/RcppExports.R:51-53,     */
/*              symbol _SpatPCA_spatpcaCV (src/RcppExports.cpp:40)           */
/* ------------------------------------------------------------------------ */
typedef struct spb_cv_stats {
  int       n_admm_calls;     /* spatpcaCore2/2p/3 calls made                          */
  long long admm_iters;       /* ADMM iterations summed over those calls               */
  int       n_inverses;       /* p x p SPD inverses (inv_sympd call sites)             */
  double    ms_total;         /* wall time of the call (host clock)                    */
  double    ms_h2d;           /* host->device copies                                   */
  double    ms_d2h;           /* device->host copies of results                        */
  double    ms_omega;         /* thin-plate-spline matrix (device time)                */
  double    ms_prologue;      /* fold gathers, SVD starts, Gram matrices               */
  double    ms_inverse;       /* assembly + Cholesky inverse + mirror                  */
  double    ms_admm;          /* persistent ADMM kernel                                */
  double    ms_cv_loss;       /* hold-out reconstruction losses                        */
  double    ms_gamma;         /* stage-3 eigen step + covariance losses                */
  long long bytes_h2d;
  long long bytes_d2h;
  int       gpu_launches;     /* kernels of this library launched (library calls not counted) */
  int       n_cg_calls;       /* spatpcaCore2/2p calls solved matrix-free (conjugate gradients, no inv_sympd) */
  long long cg_passes;        /* p x p matrix passes streamed by those calls                                   */
  double    ms_cg;            /* matrix-free Core2 kernel (device time, not included in ms_admm)               */
  double    admm_bytes;       /* algorithmic bytes of all ADMM / matrix-free launches (8p^2 + 72pK per
                                 Core3 iteration, 8p^2 + 40pK per Core2 iteration, 8p^2 + 16 n_tr p per CG pass) */
  double    p3_flops;         /* flop of the p^3 work enqueued: Omega build + SPD factorisations                */
  int       omega_route;      /* start of the Omega inverse: 0 none / plain LU, 1 SPD route, 2 pivoted LU,
                                 3 SPD route broke down and the build was repeated with the LU                  */
  double    omega_x0_resid;   /* max |I - A X0| of that start, before the Newton-Schulz step                     */
} spb_cv_stats;

/*
 * sxy   : p x d locations (already scaled by the caller, R/SpatPCA.R:183)
 * Y     : n x p data
 * nk    : n fold ids in 1..M, as doubles (R/SpatPCA.R:184)
 * outputs (all caller-allocated):
 *   cv_score_tau1  : n_tau1  values if n_tau1 > 1, else 1 value (0)   (:600, :649)
 *   cv_score_tau2  : n_tau2  values if n_tau2 > 1, else 1 value (0)   (:668, :679)
 *   cv_score_gamma : n_gamma values                                    (:692)
 *   estimated_eigenfn : p x K                                          (:697)
 *   selected       : 3 values {selected_tau1, selected_tau2, selected_gamma}
 *   stats          : may be NULL
 */
SPB_API int spb_cv(const double* sxy, int p, int d,
           const double* Y, int n, int M, int K,
           const double* tau1, int n_tau1,
           const double* tau2, int n_tau2,
           const double* gamma, int n_gamma,
           const double* nk, int maxit, double tol,
           const double* l2, int n_l2,
           double* cv_score_tau1, double* cv_score_tau2, double* cv_score_gamma,
           double* estimated_eigenfn, double* selected, spb_cv_stats* stats);

/* spb_cv with a one-entry session: identical arguments and results, but the engine of the previous
 * call on this thread is kept, and when the next call brings the same data, folds and grids (content
 * hash) only K may differ -- Omega, the fold SVD starts and every p x p inverse are K-independent and
 * are reused.  This is what the K-selection loop needs (R/SpatPCA.R:23-48 calls spatpcaCV for
 * K = 1, 2, ... on the same inputs).  The kept engine holds device memory until the next different
 * call, spb_session_clear() or spb_trim_memory().                                              */
SPB_API int spb_cv_session(const double* sxy, int p, int d,
           const double* Y, int n, int M, int K,
           const double* tau1, int n_tau1,
           const double* tau2, int n_tau2,
           const double* gamma, int n_gamma,
           const double* nk, int maxit, double tol,
           const double* l2, int n_l2,
           double* cv_score_tau1, double* cv_score_tau2, double* cv_score_gamma,
           double* estimated_eigenfn, double* selected, spb_cv_stats* stats);
SPB_API int spb_session_clear(void);

/* ------------------------------------------------------------------------ */
/* Engine handle: the same computation as spb_cv, cut at the reference's     */
/* natural seams (per fold inside each stage, src/RcppSpatPCA.cpp:315, 385,  */
/* 460 -- the old parallelFor axis) so that folds can be placed on different */
/* GPUs / processes.  The caller gathers the per-fold score rows, takes the  */
/* argmin (spb_select_index) and hands the index to the next stage.          */
/* spb_cv() is exactly: create, upload, prepare, then every fold and the     */
/* full-data fit of each stage in order on one device.                       */
/* ------------------------------------------------------------------------ */
typedef struct spb_engine spb_engine;

SPB_API int spb_engine_create(spb_engine** out, int p, int d, int n, int M, int K,
                      const double* tau1, int n_tau1,
                      const double* tau2, int n_tau2,
                      const double* gamma, int n_gamma,
                      int maxit, double tol,
                      const double* l2, int n_l2);
SPB_API int spb_engine_destroy(spb_engine* e);

/* host -> device copy of the inputs (the only H2D traffic of a run) */
SPB_API int spb_engine_upload(spb_engine* e, const double* sxy, const double* Y, const double* nk);
/* optional: inject thinPlateSplineMatrix(sxy) (p x p, host) instead of building it */
SPB_API int spb_engine_set_omega(spb_engine* e, const double* omega);
/* Omega (if any tau is non-zero) and the full-data SVD start (:561-575).   */
SPB_API int spb_engine_prepare(spb_engine* e);
/* copy the engine's Omega (without the extra 1e-8 I of :574) to the host   */
SPB_API int spb_engine_get_omega(spb_engine* e, double* omega_out);

/* stage 1 (:592-650).  fold k in 0..M-1; cv_row receives n_tau1 values when
 * n_tau1 > 1 (row k of `cv`, :327-331), otherwise it is not written.       */
SPB_API int spb_engine_stage1_fold(spb_engine* e, int k, double* cv_row);
SPB_API int spb_engine_stage1_full(spb_engine* e, int index1);
/* stage 2 (:652-680).  cv_row receives n_tau2 values when n_tau2 > 1.      */
SPB_API int spb_engine_stage2_fold(spb_engine* e, int k, int index1, double* cv_row);
SPB_API int spb_engine_stage2_full(spb_engine* e, int index1, int index2);
/* stage 3 (:682-692, worker :459-525).  cv_row receives n_gamma values.    */
SPB_API int spb_engine_stage3_fold(spb_engine* e, int k, int index1, int index2, double* cv_row);

/* The same stages for a LIST of folds at once: the folds run concurrently on independent streams
 * ("lanes"), which is how one GPU replaces the reference's fold worker pool.  cv_rows is
 * nfolds x n_grid, row-major (row f = scores of folds[f]).  The *_full calls only enqueue the
 * full-data fit on its own lane; it overlaps with the next stage and is waited for by
 * spb_engine_get_eigenfn / spb_engine_get_stats.                                              */
SPB_API int spb_engine_stage1_folds(spb_engine* e, const int* folds, int nfolds, double* cv_rows);
SPB_API int spb_engine_stage2_folds(spb_engine* e, const int* folds, int nfolds, int index1, double* cv_rows);
SPB_API int spb_engine_stage3_folds(spb_engine* e, const int* folds, int nfolds, int index1, int index2,
                                    double* cv_rows);

/* (fold k, tau1 index i) inverses as separately schedulable jobs (i in 1..n_tau1-1).  The matrix
 * inv_sympd(2 (tau1[i] Omega - G_k) + rho_k I) (src/RcppSpatPCA.cpp:165) does not depend on the ADMM
 * iterates, so with more GPUs than folds any GPU can build it and copy it to the fold's home GPU over
 * NVLink (the returned pointer is a DEVICE pointer to ld x p doubles, ld = roundup(p, 16)); the home
 * engine then marks it present and stage 1 only runs the ADMM chain.  Needs the stage-1 inverse cache
 * (spb_engine_can_share_inverses() == 1, i.e. enough device memory).                               */
SPB_API int spb_engine_can_share_inverses(spb_engine* e);
SPB_API int spb_engine_prepare_fold(spb_engine* e, int k);
SPB_API int spb_engine_inverse_buffer(spb_engine* e, int k, int i, void** dev_ptr, long long* bytes);
SPB_API int spb_engine_build_inverse(spb_engine* e, int k, int i);
SPB_API int spb_engine_mark_inverse(spb_engine* e, int k, int i);
SPB_API int spb_engine_sync(spb_engine* e);
/* thinPlateSplineMatrix (:58-76) built in column blocks so that the GPUs of one job share its p^3 work:
 * every rank calls omega_begin (LU of the bordered system; *needed = 0 when this job needs no Omega or it
 * is already there -- then skip the rest), solves its columns [c0, c1) of Lp, receives the other ranks'
 * column blocks into omega_buffer(which = 0), forms its columns of Omega = Lp'(E Lp), receives the other
 * ranks' into omega_buffer(which = 1), and calls omega_finish (mirror + the ridge of :574).  The buffers
 * are device memory; call omega_sync before handing them to a collective.  spb_engine_prepare() afterwards
 * only adds the full-data SVD start.                                          */
SPB_API int spb_engine_omega_begin(spb_engine* e, int* needed);
/* The SVD starts (svd_econ of src/RcppSpatPCA.cpp:563 for the full data, :321 for each listed fold) ahead of
 * spb_engine_prepare() / stage1_folds(): called right after omega_begin(), they run under the replicated first
 * stage of the column-split Omega build instead of after it.  Optional: prepare() and stage1_folds() do whatever
 * has not been done yet.                                                                               */
SPB_API int spb_engine_prepare_starts(spb_engine* e, const int* folds, int nfolds);
SPB_API int spb_engine_omega_solve_cols(spb_engine* e, int c0, int c1);
SPB_API int spb_engine_omega_product_cols(spb_engine* e, int c0, int c1);
/* Refined build (the default, csrc/omega_refine.cu): between the exchange of buffer 0 (X0 = A^-1 columns) and
 * omega_product_cols, every rank forms the exactly-split residual and the Newton-Schulz correction of its own
 * columns (*active = 1) and the corrected columns (buffer 2) are exchanged.  With SPB_OMEGA_REFINE=0 this is
 * a no-op (*active = 0) and buffer 2 is empty.                                                              */
SPB_API int spb_engine_omega_refine_cols(spb_engine* e, int c0, int c1, int* active);
SPB_API int spb_engine_omega_buffer(spb_engine* e, int which, int c0, int c1, void** dev_ptr, long long* bytes);
SPB_API int spb_engine_omega_sync(spb_engine* e);
SPB_API int spb_engine_omega_finish(spb_engine* e);
/* Re-arm the engine for another K on the same data and grids (then prepare + the stages again).   */
SPB_API int spb_engine_reset_k(spb_engine* e, int K);
/* estimated_eigenfn (p x K) of the full-data fit, device -> host */
SPB_API int spb_engine_get_eigenfn(spb_engine* e, double* eigenfn_out);
SPB_API int spb_engine_get_stats(spb_engine* e, spb_cv_stats* stats);
/* per-call ADMM log: 4 ints per call {kind, fold, grid index, iterations};
 * kind: 2 = Core2, 20 = Core2p, 3 = Core3; fold -1 = full data.
 * Returns the number of calls logged (writes at most `cap` of them).       */
SPB_API int spb_engine_call_log(spb_engine* e, int* log4, int cap);
/* release the per-fold p x p matrices early (optional)                     */
SPB_API int spb_engine_release_fold(spb_engine* e, int k);

/* `index_min(sum(cv, 0))` and `sum(cv, 0) / M` over an M x n_grid column-major
 * score matrix, with Armadillo's accumulation order (:595-600, 657-668,
 * 685-692).  score_out (n_grid values) may be NULL.  Host-only helper.     */
SPB_API int spb_select_index(const double* cv, int M, int n_grid, double* score_out);

/* ------------------------------------------------------------------------ */
/* Kernel-level entry points (parity tests, benchmarks).                     */
/* ------------------------------------------------------------------------ */
/* inv_sympd(symmatu(scale * (tau1 * Omega - G) + rho * I))   (:165, :397)   */
SPB_API int spb_inv_sympd(const double* omega, const double* G, int p,
                  double tau1, double rho, double scale, double* out);
/* spatpcaCore2 (:150-185) with the inverse supplied; Phi is output only.    */
SPB_API int spb_admm_core2(const double* tempinv, int p, int K, double rho, int maxit, double tol,
                   double* Phi, double* C, double* Lambda2, int* iters);
/* spatpcaCore3 (:224-270)                                                   */
SPB_API int spb_admm_core3(const double* tempinv, int p, int K, double tau2, double rho,
                   int maxit, double tol,
                   double* Phi, double* R, double* C, double* Lambda1, double* Lambda2,
                   int* iters);
/* spatpcaCore2 / spatpcaCore3 as the reference calls them: the system
 * scale * (tau1 * Omega - G) + rho * I is assembled and inverted on the device
 * (:164-165, :397) and the ADMM loop runs on the result.  `blocks` selects how
 * the inverse is held: 1 = explicit inverse, m > 1 = block U'DU factor over m
 * row blocks applied as 2m - 1 matrix-stream phases (csrc/spd_inverse.cu,
 * csrc/admm_kernels.cu), 0 = the library's default for this p.
 * mode 2: R / Lambda1 are ignored (may be NULL); Phi is output only.         */
SPB_API int spb_admm_system(const double* omega, const double* G, int p,
                    double tau1, double rho, double scale, int blocks,
                    int mode, int K, double tau2, int maxit, double tol,
                    double* Phi, double* R, double* C, double* Lambda1, double* Lambda2,
                    int* iters);
/* spatpcaCore2 (:150-185) WITHOUT inv_sympd: the Phi-update solves
 * (2 (tau1 Omega_u - Ytr'Ytr) + rho I) Phi = rho C - Lambda2 by conjugate gradients inside the persistent
 * kernel (csrc/admm_kernels.cu, cg_core2_kernel); Omega_u = symmatu(omega) is streamed once per CG step, the
 * Gram matrix is applied through Ytr (n_tr x p).  Phi is the warm start on entry.  cg_tol: relative residual
 * per column (<= 0: the library default 1e-14).  passes (may be NULL): p x p matrix passes streamed.
 * K <= 8.                                                                                              */
SPB_API int spb_core2_matrix_free(const double* omega, const double* Ytr, int n_tr, int p, int K,
                          double tau1, double rho, int maxit, double tol, double cg_tol,
                          double* Phi, double* C, double* Lambda2, int* iters, int* passes);
/* U (p x nc) = A (p x p, symmetric) * V (p x nc), 1 <= nc <= 32, on the TMA-fed FP64 DMMA stream that the
 * fold-batched matrix-free Core2 is built on (csrc/tma_dmma.cuh): kernel-level parity entry.                 */
SPB_API int spb_dmma_matvec(const double* A, int p, const double* V, int nc, double* U);
/* the same stream on a synthetic matrix in device memory: device time per pass (ms, mean of `repeats`)      */
SPB_API int spb_bench_dmma_matvec(int p, int nc, int repeats, double* ms_per_pass);
/* F matrix-free spatpcaCore2 problems in lock-step on the DMMA stream (csrc/batched_cg.cu): problem f leaves out the
 * rows of Y with nk == fold_ids[f] (0: none), uses rho[f], and is solved for tau1[0..nsteps) one after the other
 * with its state (Phi, C, Lambda2: F blocks of p x K) carried -- the tau1 path of one stage for all folds in ONE
 * launch.  iters: nsteps x F (row = step); snap (may be NULL): Phi after every step, nsteps x F x (p x K).
 * spb_core2_batched_supported: 1 if this (p, n, K, F) fits the kernel on the current device, else 0.            */
SPB_API int spb_core2_batched(const double* omega, const double* Y, const double* nk, int p, int n, int K, int F,
                      const int* fold_ids, const double* rho, const double* tau1, int nsteps, int maxit,
                      double tol, double cg_tol, double* Phi, double* C, double* Lambda2, int* iters,
                      double* snap, int* passes);
SPB_API int spb_core2_batched_supported(int p, int n, int K, int F);
/* The same kernel in spatpcaCore3 form (src/RcppSpatPCA.cpp:224-270): F problems at one tau1, solved for
 * tau2[0..nsteps) one after the other with the state (Phi, R, C, Lambda1, Lambda2) carried -- the tau2 path of stage 2
 * (:398-401) for all folds in ONE launch, matrix-free: Phi = (2 tau1 Omega - 2 Ytr'Ytr + 2 rho I)^-1 (rho (R + C) -
 * (Lambda1 + Lambda2)) by conjugate gradients instead of the reference's tempinv product (:247).                     */
SPB_API int spb_core3_batched(const double* omega, const double* Y, const double* nk, int p, int n, int K, int F,
                      const int* fold_ids, const double* rho, double tau1, const double* tau2, int nsteps,
                      int maxit, double tol, double cg_tol, double* Phi, double* R, double* C, double* Lambda1,
                      double* Lambda2, int* iters, double* snap, int* passes);
/* block count the library uses for systems of order p (1 = explicit inverse) */
SPB_API int spb_default_blocks(int p);
/* the same for the systems reused along a whole tau2 path (tempinv, :397)    */
SPB_API int spb_default_path_blocks(int p);
/* Host-only introspection of the factored solve (no device needed; used by the CPU tests):
 * block offsets off[0..m] of the partition (returns m), and the per-iteration stream phases of the ADMM
 * kernel, 8 ints each {kind (0 single fused phase, 1 forward, 2 backward), row0, nrows, col0, ncols,
 * first / one-past-last row that becomes final in the phase, segments} (returns the phase count).        */
SPB_API int spb_block_offsets(int p, int blocks, int* off, int cap);
SPB_API int spb_admm_phase_plan(int p, int K, int blocks, int* plan8, int cap);
/* svd_econ(U, s, V, Y, "right") as the engine runs it for the ADMM starts (src/RcppSpatPCA.cpp:321, :563) and as
 * setGamma() needs it (R/helper.R:207-217: d[1]^2 / nrow of the fold-1 training block): Householder QR of the long
 * side, one-sided Jacobi SVD of the triangular factor, back-transformation, all on the device.
 *   Y : m x p;  sv_out : min(m, p) singular values, descending;  V_out : p x K leading right singular vectors
 *   (may be NULL).  Signs of the vectors are arbitrary, as in LAPACK.                                        */
SPB_API int spb_svd_right(const double* Y, int m, int p, int K, double* sv_out, double* V_out);
/* pow(norm(Yva * (I - Phi Phi'), "fro"), 2)   (:327, :331, :400)            */
SPB_API int spb_cv_loss(const double* Yva, int n_va, int p, const double* Phi, int K, double* loss);
/* stage-3 losses of one fold for all gammas (:489-523)                      */
SPB_API int spb_gamma_loss(const double* Phi, int p, int K,
                   const double* Ytr, int n_tr, const double* Yva, int n_va,
                   const double* gamma, int n_gamma, double* out);

/* ------------------------------------------------------------------------ */
/* Device-resident benchmark hooks (bench.py).                               */
/* ------------------------------------------------------------------------ */
/* The matrix-free Core2 kernel on a synthetic system in device memory (Omega_u = a banded SPD matrix with
 * lambda_max ~ rho / 2, Ytr random-sign, n_tr rows): `iters` forced ADMM iterations per launch; returns the
 * device time per launch (ms, mean of `repeats` after a warm-up) and the p x p matrix passes per launch.      */
SPB_API int spb_bench_core2_cg(int p, int K, int n_tr, int iters, int repeats, double* ms_per_launch,
                       int* passes_per_launch);
/* The fold-batched kernel (csrc/batched_cg.cu) on the same synthetic system: F problems of K columns in lock-step,
 * `iters` forced ADMM iterations; device time per launch and the matrix passes (each serves all F problems).   */
SPB_API int spb_bench_core2_batched(int p, int n, int K, int F, int iters, int repeats, double* ms_per_launch,
                            int* passes_per_launch);
/* Runs `iters` forced iterations on a synthetic p x p system that lives in
 * device memory (identity: the bytes streamed are what matters; the iterate
 * starts from non-orthogonal columns so that every polar step does real
 * work) and returns the device time per iteration averaged over `repeats`
 * launches after one warm-up launch (CUDA events on the launching stream) --
 * the roofline probe of the persistent ADMM kernel.
 * mode: 2 = Core2 update, 3 = Core3 update.                                  */
SPB_API int spb_bench_admm(int p, int K, int mode, int iters, int repeats, double* ms_per_iter);
/* the same with the matrix treated as a block factor of `blocks` blocks
 * (0 = the library's default for p, 1 = explicit inverse = spb_bench_admm)  */
SPB_API int spb_bench_admm_blocks(int p, int K, int mode, int iters, int repeats, int blocks,
                          double* ms_per_iter);
/* FP64 GEMM peak of this device measured with cuBLAS DGEMM m^3 (TFLOP/s)    */
SPB_API int spb_bench_dgemm(int m, int repeats, double* tflops);
/* average device time of one inv_sympd at size p (assembly + potrf + potri + mirror) */
SPB_API int spb_bench_inverse(int p, int repeats, double* ms);
/* the same for the block factor (0 = default for p, 1 = explicit inverse)   */
SPB_API int spb_bench_factor(int p, int repeats, int blocks, double* ms);

#ifdef __cplusplus
}
#endif
#endif /* SPATPCA_B200_H */
