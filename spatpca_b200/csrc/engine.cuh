// Host-side engine: owns device memory, the stream and the cuBLAS/cuSOLVER handles, and
// sequences the stages of spatpcaCV (reference src/RcppSpatPCA.cpp:543-701).
#pragma once

#include <chrono>
#include <cstdint>
#include <memory>
#include <vector>

#include "kernels.cuh"

namespace spb {

enum Phase { PH_OMEGA = 0, PH_PROLOGUE, PH_INVERSE, PH_ADMM, PH_CVLOSS, PH_GAMMA, PH_CG, PH_COUNT };

// Where work is enqueued: a stream with its own cuBLAS handle and inverse workspace.
struct Exec {
  cudaStream_t stream;
  cublasHandle_t blas;
  int* flag;                       // sticky device-side error flag (shared by every lane of the device)
  DevBuf<double>* inv_work;        // workspace of the recursive SPD inverse
};

// Pooled per device: creating streams / cuBLAS handles costs milliseconds, engines come and go.
struct LaneRes {
  cudaStream_t stream = nullptr;
  cublasHandle_t blas = nullptr;
  DevBuf<double> inv_work;
  ~LaneRes();
};

// Stream + library handles + workspaces of the thin-plate-spline build, separate from the main stream so
// that the Omega build (260 ms at p = 10 000) overlaps with the SVD prologues of the folds.
struct OmegaRes {
  cudaStream_t stream = nullptr;
  cublasHandle_t blas = nullptr;
  cusolverDnHandle_t solver = nullptr;
  DevBuf<double> work;
  DevBuf<int> ipiv, info;
  ~OmegaRes();
};
// temporaries of one build; must outlive the enqueued work
struct OmegaTemps {
  DevBuf<double> L, Bm, T1;
  DevBuf<double> C1, C2, vec, hh, resid;   // refined build only (omega_refine.cu)
  DevBuf<int8_t> oz_left, oz_right;        // INT8 digit planes of the refinement products (ozaki_int8.cu)
  DevBuf<int> oz_c, oz_ea, oz_eb, oz_bad;
  DevBuf<int> spdflag;            // failure flag of the SPD route to X0
  bool force_lu = false;          // take the LU route to X0 (the fallback after a failed SPD route)
  bool used_spd = false;
  int steps_done = 0;             // Newton-Schulz steps enqueued (residual records in `resid`)
  int col_step = 0;               // column-split build: steps done so far
};

// Per-thread, per-device CUDA context (stream + library handles + solver workspaces).
struct DeviceCtx {
  int device = -1;
  cudaStream_t stream = nullptr;
  cublasHandle_t blas = nullptr;
  cusolverDnHandle_t solver = nullptr;
  DevBuf<double> solver_work;     // cuSOLVER workspace (grown on demand)
  DevBuf<int> info;               // devInfo
  DevBuf<int> flag;               // sticky error flag
  DevBuf<int> ipiv;
  DevBuf<double> inv_work;        // workspace of the recursive SPD inverse

  std::unique_ptr<OmegaRes> omega_res;
  OmegaRes& omega();              // created on first use
  std::vector<std::unique_ptr<LaneRes>> lane_pool;
  LaneRes& lane_res(int i);       // i-th pooled lane of this device (created on first use)
  Exec exec() { return Exec{stream, blas, flag.get(), &inv_work}; }

  static DeviceCtx& get();        // context of the calling thread's current device
  static std::unique_ptr<DeviceCtx> create(int dev, bool with_blas);
  std::vector<std::unique_ptr<DeviceCtx>> worker_pool;
  DeviceCtx& worker(int i);       // i-th pooled context for a host worker thread (created on first use)
  void sync_and_check();          // stream sync + translate the sticky flag into an Error
  double* work(size_t doubles) { solver_work.ensure(doubles); return solver_work.get(); }
  ~DeviceCtx();
};

// Event-based phase timing on the launching stream (no synchronisation until collect()).
class PhaseClock {
 public:
  void begin(Phase ph, cudaStream_t s);
  void end(cudaStream_t s);
  void collect(double ms[PH_COUNT]);     // stream must be idle
  void reset();
  ~PhaseClock();
 private:
  cudaEvent_t grab();
  std::vector<cudaEvent_t> pool_;
  size_t used_ = 0;
  struct Span { Phase ph; cudaEvent_t a, b; };
  std::vector<Span> spans_;
  int open_ = -1;
  double acc_[PH_COUNT] = {0, 0, 0, 0, 0, 0, 0};
};

struct ChainState {              // ADMM iterate of one chain (p x K each)
  DevBuf<double> Phi, R, C, L1, L2;
  void alloc(size_t n) { Phi.alloc(n); R.alloc(n); C.alloc(n); L1.alloc(n); L2.alloc(n); }
};

struct FoldData {
  bool prepared = false;
  int n_tr = 0, n_va = 0;
  std::vector<int> rows_tr, rows_va;
  DevBuf<double> Ytr, Yva;        // n_tr x p, n_va x p
  DevBuf<double> Ytt;             // p x n_tr: transpose of Ytr (the matrix-free Core2 kernel reads both layouts)
  DevBuf<double> V;               // leading right singular vectors of Ytr (p x kcap), the PCA start for any K
  int kcap = 0;
  DevBuf<double> Phi0, L20;       // Phi_cv / Lambd2_cv slices of the reference (:569)
  DevBuf<double> tinv;            // tempinv_cv slice (p x p), allocated in stage 2
  int tinv_index = -1;            // tau1 index `tinv` was built for (reused across K)
  bool tinv_valid = false;        // `tinv` holds that tempinv (false while the tau2 paths run matrix-free)
  std::vector<DevBuf<double>> inv_cache;   // stage-1 inverses by tau1 index (kept when memory allows)
  std::vector<char> inv_valid;
  std::vector<double> sv;         // singular values of Ytr
  double rho = 0.0;
  double tv = 0.0;                // trace(Ytr'Ytr) / n_tr
};

// device temporaries of one stage-3 evaluation (kept per lane so that the work can stay asynchronous)
struct GammaWork {
  DevBuf<double> W, scr, H, Q, B, lam, err, gscr;
  DevBuf<double> Yvt, Gv, Wv, Cv, BtB, cst, red;      // low-rank form of the loss: Yva', Yva Yva' / m, Yva B, ...
  void ensure(int p, int K, int n_tr, int n_va, int ng, bool tiles);
};

struct CallRec { int kind, fold, idx, slot, n_tr; };
// One launch of the fold-batched matrix-free Core2 kernel: its per-(step, problem) iteration counts live in device
// memory until the next flush.
struct BatchRec {
  int kind;                        // call-log kind of its calls (2 or 20)
  std::vector<int> folds;          // problem f -> fold (-1: full data)
  std::vector<int> idx;            // step -> grid index reported in the call log
  std::vector<int> kinds;          // per problem, when they differ (a full-data Core2p riding with the folds' Core2)
  std::vector<int> idx_override;   // per problem: grid index of its (single) step, or -1
  size_t it_off, st_off;           // offsets into bcg_iters_ / bcg_status_
  int n;                           // rows of Y seen by the kernel (for the byte accounting)
};

// One independent chain of work (a fold's tau path, or the full-data fit): its own stream, cuBLAS
// handle and working set, so that the latency-bound parts of one chain (leaf inverses, small
// GEMMs, the ADMM serial tail) overlap with the GEMM-bound parts of the others.  This is what
// replaces the reference's (already removed) RcppParallel fold pool (src/RcppSpatPCA.cpp:312, 384, 459).
struct Lane {
  LaneRes* res = nullptr;
  DevBuf<double> work, gram;       // assembled system -> inverse; Gram matrix of `gram_owner`
  int gram_owner = -2;
  ChainState chain;
  DevBuf<char> admm_ws;            // sized for the matrix-free Core2 kernel too (cg_workspace_bytes)
  DevBuf<double> lossW, loss_scratch;
  DevBuf<int> stat_slots;
  DevBuf<double> err_slots;
  int next_slot = 0;
  std::vector<CallRec> pending;
  PhaseClock clock;
  GammaWork gwork;                 // stage-3 temporaries
  cudaEvent_t evt = nullptr;       // ordering between this lane and a batched launch on another stream
  bool dirty = false;              // work enqueued since the last flush
  bool omega_synced = false;       // this lane's stream already waits for the Omega build
  Exec exec(int* flag) { return Exec{res->stream, res->blas, flag, &res->inv_work}; }
};

class Engine {
 public:
  Engine(int p, int d, int n, int M, int K, const double* tau1, int n1, const double* tau2, int n2,
         const double* gamma, int n3, int maxit, double tol, const double* l2, int nl2);
  ~Engine();

  void upload(const double* sxy, const double* Y, const double* nk);
  void set_omega(const double* omega_host);
  void prepare();
  void get_omega(double* out);

  // per-stage work of a set of folds; the folds run concurrently on their lanes.
  // rows: nfolds x n_grid, row-major (row f = scores of folds[f])
  void stage1_folds(const int* folds, int nfolds, double* rows);
  void stage1_full(int index1);
  void stage2_folds(const int* folds, int nfolds, int index1, double* rows);
  void stage2_full(int index1, int index2);
  void stage3_folds(const int* folds, int nfolds, int index1, int index2, double* rows);
  void stage1_fold(int k, double* cv_row) { stage1_folds(&k, 1, cv_row); }
  void stage2_fold(int k, int index1, double* cv_row) { stage2_folds(&k, 1, index1, cv_row); }
  void stage3_fold(int k, int index1, int index2, double* cv_row) { stage3_folds(&k, 1, index1, index2, cv_row); }
  // (fold, tau1 index) inverses as separately schedulable jobs: they depend on Omega, G_k, rho_k and
  // tau1[i] only -- not on the ADMM iterates -- so any GPU can build them and ship them to the fold's
  // home GPU.  Require the stage-1 inverse cache (memory permitting).
  bool can_share_inverses();      // false when stage 1 runs matrix-free (there is no inverse to share)
  void prepare_fold(int k) { fold_prepare(k); }
  void prepare_starts(const int* folds, int nfolds);   // full-data + fold SVD starts ahead of prepare()
  void* inverse_buffer(int k, int i, long long* bytes);   // device buffer of inverse (k, tau1[i]), allocated on demand
  void build_inverse(int k, int i);                        // enqueue on fold k's lane
  void mark_inverse(int k, int i);                         // the buffer was filled from outside (peer copy)
  void sync() { flush(); }
  // Omega built in column blocks, so that the ranks of a multi-GPU job share its 6.7 p^3 flop (SURVEY.md
  // section 8e): every rank factors the bordered system (omega_begin), solves for its own columns of
  // Lp (omega_solve_cols), receives the others' (omega_buffer(0, ..) is the NCCL target), forms its
  // columns of Omega = Lp'(E Lp) (omega_product_cols), receives the others' (omega_buffer(1, ..)) and
  // finishes with the mirror + ridge (omega_finish).  Returns false from omega_begin when no Omega is needed.
  bool omega_begin();
  void omega_solve_cols(int c0, int c1);
  void omega_product_cols(int c0, int c1);
  // refined build only (returns false and does nothing otherwise): residual + Newton-Schulz correction of this rank's
  // columns; runs between the exchange of buffer 0 (X0) and the exchange of buffer 2 (X1)
  bool omega_refine_cols(int c0, int c1);
  void* omega_buffer(int which, int c0, int c1, long long* bytes);   // which: 0 = X0 / Lp columns, 1 = Omega columns, 2 = refined X1 columns
  void omega_stream_sync();
  void omega_finish();
  // Re-run with another number of eigenfunctions on the same data and grids (the K-selection loop of
  // R/SpatPCA.R:23-48).  Omega, the fold gathers, the SVD starts and every cached inverse are
  // K-independent and are kept.
  void reset_k(int K);
  void get_eigenfn(double* out);
  void get_stats(spb_cv_stats* st);
  int call_log(int* log4, int cap);
  void release_fold(int k);

  int p() const { return p_; }
  int K() const { return K_; }
  int M() const { return M_; }
  int n_tau1() const { return (int)tau1_.size(); }
  int n_tau2() const { return (int)tau2_.size(); }
  int n_gamma() const { return (int)gamma_.size(); }
  double tau1_at(int i) const { return tau1_[i]; }
  double tau2_at(int i) const { return tau2_[i]; }
  double gamma_at(int i) const { return gamma_[i]; }
  double selected_tau1(int index1) const;
  std::chrono::steady_clock::time_point t_created;

  // building blocks shared with the kernel-level C entry points
  // Enqueues thinPlateSplineMatrix on R.stream (asynchronous; `tmp` must stay alive until that stream has
  // drained).  upper_only: only the tiles on / above the diagonal of the final product are formed.
  static void build_omega_raw(OmegaRes& R, int* flag, const double* x_dev, int p, int d, double* omega_dev,
                              size_t ld, bool upper_only, OmegaTemps& tmp);
  // the same matrix with this side's rounding error removed (omega_refine.cu): LU + one Newton-Schulz step on an
  // exactly formed residual + the cancellation-free product; the default (SPB_OMEGA_REFINE=0: the plain pipeline)
  static void build_omega_refined(OmegaRes& R, int* flag, const double* x_dev, int p, int d, double* omega_dev,
                                  size_t ld, bool upper_only, OmegaTemps& tmp);
  static bool omega_refine_enabled();
  // blk.m == 1: explicit inverse; blk.m > 1: block U'DU factor for the multi-phase ADMM stream
  static void spd_inverse(const Exec& ex, const double* omega_u, size_t ldo, const double* G,
                          size_t ldg, int p, double tau1, double rho, double scale, double* out,
                          size_t ldout, const Blocking& blk = Blocking());
  static void gram_upper(const Exec& ex, const double* Ymat, int m, int p, double* G, size_t ldg);
  // SVD-right start: singular values (host) and the first K right singular vectors (device)
  static void svd_right(DeviceCtx& ctx, const double* Ymat, int m, int p, int K,
                        std::vector<double>& sv, double* Phi_dev);
  static void cv_loss(cudaStream_t s, const double* Yva, int m, int p, const double* Phi, int K,
                      double* W, double* scratch, double* out_dev);
  // all-gamma covariance losses of one fold, asynchronous on `s`: eigen step and water filling on the device
  static void gamma_losses_async(cudaStream_t s, GammaWork& w, const double* Phi, int p, int K, const double* Ytr,
                                 int n_tr, double tv, const double* Yva, int n_va, const double* gamma_dev, int ng,
                                 double* out_dev);
  // the same, synchronous: `out_host` receives n_gamma values
  static void gamma_losses(DeviceCtx& ctx, cudaStream_t s, const double* Phi, int p, int K, const double* Ytr,
                           int n_tr, double tv, const double* Yva, int n_va, const std::vector<double>& gamma,
                           double* out_host);

 private:
  void fold_prepare(int k);
  void fold_prepare_on(DeviceCtx& c, int k);
  void full_svd_start();
  void omega_lu_factor();
  void omega_x0_resolve();
  bool omega_x0_check_pending_ = false;   // column build: the SPD route's flag has not been read yet
  void prepare_folds(const int* folds, int nfolds);
  double prologue_wall_ms_ = 0.0;   // host wall time of the threaded fold starts (added to ms_prologue)
  Lane& lane_of(int fold);                       // fold < 0: the full-data lane
  void ensure_gram(Lane& L, int k);              // k = -1: full data
  void init_lambda(cudaStream_t s, const double* Phi, const std::vector<double>& sv, double rho, double* L2);
  // buf: matrix buffer to build the inverse in (nullptr: the lane's work buffer);
  // have: buf already holds inv_sympd(2(tau1 Omega - G) + rho I) for this (fold, tau1)
  void run_core2(Lane& L, int kind, int fold, int idx, double tau1, double rho, ChainState& st,
                 double* buf = nullptr, bool have = false, int expected_iters = 8);
  // Matrix-free Core2 (conjugate gradients, admm_kernels.cu) or the factored path?  Decided per call from a
  // bound on cond(c_omega Omega_u + rho I - c_gram Ytr'Ytr) and the iterations the call is expected to run.
  bool use_cg(double c_omega, double rho, double gram_top, int expected_iters, double pass_scale = 1.0);
  bool core2_matrix_free(int fold, double tau1, double rho, int expected_iters, double pass_scale = 1.0);   // the decision run_core2 takes
  double omega_lmax();                           // power-iteration estimate of lambda_max(Omega_u) (host; waits for Omega)
  void enqueue_lmax(cudaStream_t s, cublasHandle_t blas);
  double* cached_inverse(FoldData& f, int index1);   // stage-1 inverse of the selected tau1, or nullptr
  void drop_cached_inverses(FoldData& f, int keep);
  void run_core3(Lane& L, int kind, int fold, int idx, const double* tinv, double tau2, double rho, ChainState& st);
  // fold-batched matrix-free Core2 (batched_cg.cu) for `folds` (each on its own lane) along `taus`: returns false
  // (nothing enqueued) when the shape / policy does not allow it.  snap: Phi after every step, or nullptr.
  // mode 3: spatpcaCore3 along the tau2 values `taus` at `fixed_tau1` (R and Lambda1 of the chains carried as well).
  bool run_core2_batched(const std::vector<int>& folds, const std::vector<double>& taus, const std::vector<int>& idx,
                         int kind, int expected_iters, double** snap_out, int mode = 2, double fixed_tau1 = 0.0,
                         const std::vector<int>* kinds = nullptr, const std::vector<int>* idx_override = nullptr);
  // Full-data work that can ride with the folds' next batch (the padded columns of the DMMA tiles are free): the
  // Core2p call of stage1_full() joins stage 2's cold Core2 batch, the tau2 path of stage2_full() joins stage 3's
  // refit batch.  run_pending_full() runs whatever is still pending on its own.
  int pend_core2p_ = -1;           // tau1 index of a deferred Core2p (:598-599), or -1
  bool pend_path_ = false;         // deferred tau2 path of the full-data fit (:662-666)
  int pend_path_i1_ = 0, pend_path_upto_ = 0;
  int my_stage1_folds_ = 0;        // folds this engine ran in stage 1 (0: a dedicated full-fit rank, nothing to ride with)
  void run_pending_full();
  void run_full_path(int index1, const std::vector<double>& path, int upto);
  // the checks of run_core2_batched without side effects: would this batch run matrix-free?
  bool batch_eligible(const std::vector<int>& folds, int nsteps, double tau1_max, int expected_iters, int mode);
  bool core3_matrix_free(int fold, double tau1, double rho, int expected_iters, double pass_scale = 1.0);
  void ensure_fold_tempinv(int k, int index1);   // the fold's tempinv (:397) on its lane, unless already there
  void ensure_batch_buffers(int F, int nsteps);
  void launch_loss(Lane& L, const FoldData& f, const double* Phi, double* out_dev);
  std::vector<std::vector<int>> lane_rounds(const int* folds, int nfolds) const;
  void enqueue_stage1(int k, int step);
  void enqueue_stage2(int k, int index1, int step);
  void enqueue_stage3_refit(int k, int index1, int index2, bool chain = true);
  double* full_tempinv(int index1);              // tempinv of the full-data fit, built on first request
  void flush_lane(Lane& L);                      // sync the lane, move its call stats to the host log
  void flush();                                  // every lane + the sticky error flag
  void copy_pk(cudaStream_t s, double* dst, const double* src);
  void zero_pk(cudaStream_t s, double* dst);
  double* fold_scores(int k) { return loss_out_.get() + (size_t)k * max_grid_; }

  int p_, d_, n_, M_, K_, maxit_;
  double tol_;
  std::vector<double> tau1_, tau2_, gamma_, l2_;
  bool any_tau_ = false;
  size_t ld_ = 0;                                // leading dimension of every p x p matrix
  Blocking blocks_;                              // partition of the Core2 systems (each used by ONE ADMM call, 1-5 iterations)
  Blocking blocks_path_;                         // partition of the tempinv systems (each reused along a whole tau2 path)
  int max_grid_ = 1;
  DeviceCtx& ctx_;

  DevBuf<double> dY_, dX_;
  DevBuf<double> dYt_;                           // p x n: transpose of Y (full-data matrix-free Core2p)
  DevBuf<int> dNk_;                              // fold ids of the rows of Y (device)
  DevBuf<char> bcg_ws_;                          // workspace of the batched kernel
  DevBuf<double> bcg_snap_;                      // snapshots of Phi: steps x problems x (p x K)
  DevBuf<int> bcg_iters_, bcg_status_;
  size_t bcg_it_used_ = 0, bcg_st_used_ = 0;
  std::vector<BatchRec> batch_pending_;
  cudaStream_t batch_stream_ = nullptr;          // stream of the batches enqueued since the last flush
  bool snap_in_use_ = false;
  int bcg_mode_ = 0;                             // SPB_BATCH: 0 = auto, 2 = never batch the folds
  int core3_mode_ = 0;                           // SPB_CORE3: 0 = auto (matrix-free tau2 paths when the policy allows), 2 = always tempinv
  DevBuf<double> lmax_vec_, lmax_out_;           // power iteration: 2 p-vectors, norms
  double lmax_ = -1.0;                           // < 0: not read back yet
  bool lmax_enqueued_ = false;
  int cg_mode_ = 0;                              // SPB_CORE2: 0 = auto, 1 = always matrix-free (K <= 8), 2 = never
  double cg_tol_ = 1e-12;                        // SPB_CG_TOL: relative residual of the CG solves (CV scores move by ~1e-13 at C4, the parity bar is 1e-9)
  std::vector<double> nk_;
  DevBuf<double> omega_u_;                       // symmatu(Omega) + 1e-8 I
  DevBuf<double> omega_diag_raw_;                // diagonal before the ridge (for get_omega)
  bool omega_ready_ = false, omega_injected_ = false;
  bool omega_pending_ = false;                   // the build is in flight on the Omega stream
  cudaEvent_t omega_evt_ = nullptr;
  std::unique_ptr<OmegaTemps> omega_tmp_;
  bool omega_cols_open_ = false;                 // between omega_begin() and omega_finish()
  bool omega_x0_is_spd_ = false;                 // column build: X0 came from the SPD route (replicated, no first exchange)
  bool omega_begin_refined();
  void omega_product_cols_refined(int c0, int c1);
  void wait_omega(Lane& L);                      // first use of Omega on a lane: wait for the build
  void finish_omega();                           // host wait + release of the temporaries
  std::vector<FoldData> folds_;
  // full-data fit (:563-567)
  std::vector<double> sv_full_;
  double rho_full_ = 0.0;
  ChainState full_;
  bool full_ready_ = false;                      // full_ holds the PCA start for the current K
  bool full_svd_done_ = false;
  DevBuf<double> Vfull_;                         // leading right singular vectors of Y (p x kcap_full_)
  int kcap_full_ = 0;
  DevBuf<double> full_inv2_, full_tinv_;         // full-data inverses kept across K (when cache_inverses_)
  int full_inv2_index_ = -1, full_tinv_index_ = -1;
  void alloc_k_state();                          // everything whose size depends on K

  std::vector<std::unique_ptr<Lane>> lanes_;     // n_fold_lanes_ fold lanes + 1 full-data lane
  int n_fold_lanes_ = 1;
  bool cache_inverses_ = false;                  // keep the (fold, tau1) inverses of stage 1 for stages 2/3
  int slot_cap_ = 0;
  DevBuf<double> loss_out_;                      // M x max_grid scores, (M+1)-th row = scratch
  DevBuf<double> gamma_dev_, gamma_out_;         // the gamma grid on the device; M x n_gamma stage-3 losses
  DevBuf<double> prep_scratch_;
  std::vector<int> log_;                         // 4 ints per call

  PhaseClock clock_;
  spb_cv_stats stats_;
  long long launches_at_start_ = 0;
};

// recursive GEMM-rate SPD inverse (spd_inverse.cu): A (upper valid) -> A^-1 (full storage)
size_t spd_inverse_workspace_doubles(int n);
void spd_inverse_recursive(const Exec& ex, double* A, int n, size_t ld, double* workspace);
// upper(C) += alpha * A' * B  (A, B: k x n; C: n x n symmetric result) as a recursion of cuBLAS GEMMs
void gemm_upper_tn(cublasHandle_t blas, int n, int k, double alpha, const double* A, size_t lda, const double* B,
                   size_t ldb, double* C, size_t ldc);
// block U'DU factor for the multi-phase ADMM stream (blk.m == 1: the explicit inverse)
void spd_factor_blocked(const Exec& ex, double* A, int n, size_t ld, double* workspace, const Blocking& blk);

double omega_refined_flops(int p);
bool omega_x0_spd_enabled(int p, int d);   // X0 of the refined Omega build through the SPD part of the system (no LU)
void omega_x0_spd(OmegaRes& R, const double* x_dev, int p, int d, double* Pbuf, double* X0, double* inv_ws, OmegaTemps& tmp);
bool omega_build_ok(OmegaTemps& tmp, double* resid_out = nullptr);
int omega_refine_steps(bool spd_start);
bool omega_int8_enabled();


// ---- ozaki_int8.cu: extended-precision products on the INT8 tensor cores ----
struct OzakiGeom {
  size_t kq = 0;      // contraction length of one digit plane (q rounded up to 128)
  int mp = 0, np = 0; // padded row / column counts of the product (multiples of 16)
};
OzakiGeom ozaki_geom(int q, int nc);
int ozaki_env_slices(const char* name, int dflt, int lo, int hi);
size_t ozaki_workspace_bytes(int q, int nc, int S, int Sc);
void ozaki_transpose(cudaStream_t s, const double* src, size_t lds, int n, double* dst, size_t ldd);
void ozaki_slice(cudaStream_t s, const double* M, size_t ld, int rows, int cols, int colsp, int S, bool reversed,
                 int8_t* out, size_t kq, int* expo, int* bad);
void ozaki_product(cublasHandle_t blas, cudaStream_t s, const OzakiGeom& g, const int8_t* left, int Sa,
                   const int8_t* right, int Sb, int S, int m, int n, int* C, double* acc, size_t lda, const int* ea,
                   const int* eb, int mode, int col0, double coef, const double* base, size_t ldb, double* out,
                   size_t ldo, const int* bad);
// the same for a symmetric result of which only the upper triangle is used: column strips, rows above the strip's end only
void ozaki_product_upper(cublasHandle_t blas, cudaStream_t s, const OzakiGeom& g, const int8_t* left, int Sa,
                         const int8_t* right, int Sb, int S, int m, int n, int* C, double* acc, size_t lda, const int* ea,
                         const int* eb, int mode, int col0, double coef, const double* base, size_t ldb, double* out,
                         size_t ldo, const int* bad, int strip);

// small host helpers (host_math.cpp-like, defined in engine.cu)
double arma_accumulate(const double* x, int n);

}  // namespace spb
