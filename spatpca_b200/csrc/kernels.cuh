// Launchers of the hand-written sm_100a kernels (definitions in *_kernels.cu).
#pragma once

#include "common.cuh"

namespace spb {

#ifdef __CUDACC__
// Reciprocal / reciprocal square root for LATENCY-bound scalar chains (leaf pivots, Jacobi
// rotations).  IEEE division and sqrt are ~50-instruction dependent sequences; one warp working
// alone pays their full latency (ncu: a 128 x 128 leaf spent 340 k cycles, >90 % of them in the
// per-pivot divisions).  Hardware approximation (>= 20 bits) + two Newton steps: error ~1 ulp.
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  return r;
}
__device__ __forceinline__ double fast_rsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double h = 0.5 * x;
  y = y * fma(-h * y, y, 1.5);
  y = y * fma(-h * y, y, 1.5);
  y = y * fma(-h * y, y, 1.5);
  return y;
}
#endif

// ---- thin-plate-spline / matrix assembly (tps_kernels.cu) -------------------
// Bordered TPS system [[E + ridge I, T], [T', ridge I]], T = [1 X], q = p + d + 1
// (reference: struct thinPlateSpline, src/RcppSpatPCA.cpp:12-45 + symmatu :67 + ridge :69).
// border = false writes only the p x p kernel block E (diag = ridge).
void launch_tps_fill(const double* x_dev, int p, int d, double* L, size_t ld, double ridge,
                     bool border, cudaStream_t s);
// In-place A <- symmatu(A) + diag_add * I for the leading n x n block (tile-pair transpose).
void launch_symmatu_inplace(double* A, int n, size_t ld, double diag_add, cudaStream_t s);
// Upper triangle of  scale * (tau1 * Omega_u - G) + rho * I  (reference :164-165, :397),
// same operation order as the reference, no FMA contraction.  G may alias out.
void launch_assemble_upper(const double* omega_u, size_t ldo, const double* G, size_t ldg,
                           int p, double tau1, double scale, double rho,
                           double* out, size_t ldout, cudaStream_t s);
// out (m x ncol) <- rows `rows[0..m)` of Y (n x ncol)
void launch_gather_rows(const double* Y, int n, int ncol, const int* rows, int m, double* out,
                        cudaStream_t s);
// out (cols x rows, ld = cols) <- transpose of in (rows x cols, ld = rows)
void launch_transpose(const double* in, int rows, int cols, double* out, cudaStream_t s);
// dst (n x n, ldd) <- upper triangle of src (leading n x n, lds), zeros below the diagonal
void launch_copy_upper(const double* src, size_t lds, int n, double* dst, size_t ldd, cudaStream_t s);
// identity columns: B (q x ncol, ld) <- [I_ncol; 0]
void launch_set_identity(double* B, int q, int ncol, size_t ld, cudaStream_t s);
// out[0] <- sum of squares of x[0..n)   (deterministic two-level tree)
void launch_sumsq(const double* x, size_t n, double* scratch, double* out, cudaStream_t s);
// Lambda2 = Phi * diag(coef_k), coef_k = rho - 1 / (rho - 2 sv_k^2) computed by the host in the
// reference's operation order (reference :326, :567)
void launch_lambda_init_host(const double* Phi, int p, int K, const double* coef_host, double* Lambda2,
                             cudaStream_t s);
// dst[i] = alpha * src[i]
void launch_scale_copy(const double* src, double alpha, double* dst, size_t n, cudaStream_t s);
// sticky device-side status: if (*info != 0) *flag = code (first failure wins)
void launch_check_info(const int* info, int* flag, int code, cudaStream_t s);
// power iteration on a symmetric matrix: fixed start vector; x <- y / |y| with |y| stored to norm_out[0]
void launch_power_init(double* x, int p, cudaStream_t s);
void launch_power_normalize(const double* y, int p, double* x, double* norm_out, cudaStream_t s);
// B = Phi * Q  (p x K times K x K, Q on device)
void launch_small_rmul(const double* Phi, int p, int K, const double* Q_dev, double* out, cudaStream_t s);

// ---- persistent ADMM kernel (admm_kernels.cu) --------------------------------
struct AdmmWorkspace;   // partial-sum buffers, sized by (p, K)

// Row / column partition of a p x p system into m diagonal blocks (offsets are multiples of 512 = the
// ADMM kernel's row tile, off[m] = p).  m = 1 means "explicit inverse".  m > 1 selects the block
// U'DU factor of spd_inverse.cu, which costs ~p^3/3 + p^3/m flop instead of p^3 and is applied by the
// ADMM kernel as 2m - 1 dependent matrix-stream phases over the same 8 p^2 bytes.
constexpr int kMaxBlocks = 12;     // 2 * 12 - 1 = 23 stream phases at most
struct Blocking {
  int m = 1;
  int off[kMaxBlocks + 1] = {};
};
// Default partition for a system of order p (SPB_FACTOR_BLOCKS overrides the block count; 1 = explicit
// inverse everywhere).  Deterministic in p, so every rank of a multi-GPU job agrees on it.
Blocking spd_blocking(int p);
// Partition for systems that are reused by many ADMM calls (the tempinv of a tau2 path): fewer blocks,
// because there the per-iteration phases cost more than the factorisation saves (SPB_PATH_BLOCKS overrides).
Blocking spd_blocking_path(int p);
Blocking spd_blocking_m(int p, int m_request);

struct AdmmCall {
  const double* Ainv;   // p x p symmetric, full storage: the inverse, or (blocks.m > 1) its block factor F
  Blocking blocks;      // partition F was built with (default: m = 1, explicit inverse)
  size_t ld;
  int p, K;
  int mode;             // 2 = spatpcaCore2 / 2p update, 3 = spatpcaCore3 update
  double rho, tau2;
  int maxit;
  double tol;
  double *Phi, *R, *C, *L1, *L2;   // p x K, ld = p.  R/L1 unused in mode 2
  int* stat_slot;       // device: {iterations, status, sweeps}
  double* err_slot;     // device: last max error
  unsigned long long* dbg = nullptr;   // optional phase time stamps (performance debugging)
};

size_t admm_workspace_bytes(int p, int K);
// `ws` must hold admm_workspace_bytes(p, K) bytes.  Asynchronous on `s`.
void launch_admm(const AdmmCall& c, void* ws, cudaStream_t s);
// deterministic segment count used by phase 1 (exposed for DESIGN/bench reporting)
int admm_num_segments(int p, int K);
// The stream phases the kernel runs per iteration for this (p, K, partition): 8 ints per phase
// {kind (0 = single fused phase, 1 = FWD, 2 = BWD), row0, nrows, col0, ncols, fin0, fin1, segments}; returns the
// number of phases.  Host-only (tests check that the rectangles tile the matrix and solve the system).
int admm_phase_plan(int p, int K, const Blocking& blk, int* out, int cap);

// ---- matrix-free spatpcaCore2 (admm_kernels.cu): conjugate gradients on the ADMM system ------------
// A = c_omega * Omega_u + rho I - c_gram * Ytr'Ytr is applied, never factored: every CG step streams Omega_u
// once and uses the L2-resident data block for the rank-n_tr part.  The whole call (CG solves, Stiefel
// projection, dual update, stopping rule) is one cooperative launch, like launch_admm.
constexpr int kCgMaxK = 8;
struct CgCall {
  const double* omega_u;   // p x p symmetric, full storage (symmatu(Omega) + 1e-8 I)
  size_t ld;
  int p, K;
  const double* Ytr;       // n_tr x p (ld n_tr)
  const double* Yt;        // p x n_tr (ld p)
  int n_tr;
  double c_omega, c_gram;  // spatpcaCore2: 2 tau1 and 2 (:164)
  double rho;
  int maxit;
  double tol;
  double cg_tol;           // relative residual per column (default 1e-14)
  int cg_cap;              // CG steps per solve before the call is failed
  double *Phi, *C, *L2;    // p x K; Phi is the warm start of the first solve
  int* stat_slot;          // device: {iterations, status, polar steps, matrix passes}
  double* err_slot;
};
size_t cg_workspace_bytes(int p, int K, int n_tr);
void launch_core2_cg(const CgCall& c, void* ws, cudaStream_t s);

// ---- TMA-fed FP64 DMMA matrix stream (batched_cg.cu, tma_dmma.cuh) ---------------------------------
// U (p x nc, ld = p) = A (p x p symmetric, full storage) * V (p x nc), 1 <= nc <= 32; device pointers
void launch_dmma_matvec(const double* A, size_t ld, int p, const double* V, size_t ldv, int nc, double* U,
                        cudaStream_t s);

// ---- fold-batched matrix-free spatpcaCore2 (batched_cg.cu): F problems in lock-step on the DMMA stream --------
struct BatchedCore2Call {
  const double* omega_u;     // p x p symmetric, full storage
  size_t ld;
  int p, n, K, F;            // F problems of K columns each, F * K <= 32, K <= 8
  const double* Y;           // n x p (ld n): ALL rows of the data
  const double* Yt;          // p x n (ld p)
  const int* nk;             // n fold ids (device)
  int fold_id[8];            // problem f leaves out the rows with nk == fold_id[f] (0: none, i.e. the full-data fit)
  double rho[8];
  double* Phi[8];            // per problem, p x K (ld p): warm start in, last iterate out
  double* C[8];
  double* L2[8];
  int mode = 2;              // 2: spatpcaCore2 / 2p, one step per tau1 value; 3: spatpcaCore3, one step per tau2 value
  double* R[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};    // mode 3
  double* L1[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // mode 3
  const double* tau2 = nullptr;   // host, nsteps values (mode 3; tau1[] then repeats the fixed tau1)
  int nsteps;                // grid values solved one after the other, state carried (the tau1 path, :329-332)
  const double* tau1;        // host, nsteps values
  int maxit;
  double tol, cg_tol;
  int cg_cap;
  double* snap;              // device, nsteps x F x (p x K): Phi after every step (may be nullptr)
  int* iters_out;            // device, nsteps x F
  int* status_out;           // device, {status, matrix passes}
  unsigned long long* dbg = nullptr;   // optional: globaltimer stamps of CTA 0, 9 per CG step (400 slots)
};
bool batched_core2_supported(int p, int n, int K, int F);
size_t batched_core2_workspace_bytes(int p, int n, int K, int F);
void launch_batched_core2(const BatchedCore2Call& c, void* ws, cudaStream_t s);

// ---- losses (loss_kernels.cu) --------------------------------------------------
// W (m x K) <- Ymat (m x p) * Phi (p x K); scratch must hold cvloss_scratch_doubles(m, p, K)
size_t yphi_scratch_doubles(int m, int p, int K);
void launch_yphi(const double* Ymat, int m, int p, const double* Phi, int K, double* W,
                 double* scratch, cudaStream_t s);
// out[0] <- || Yva - W Phi' ||_F^2   with W = Yva Phi precomputed
void launch_residual_sumsq(const double* Yva, int m, int p, const double* Phi, int K,
                           const double* W, double* scratch, double* out, cudaStream_t s);
// H (K x K) <- W' W * alpha
void launch_small_gram(const double* W, int m, int K, double alpha, double* H, cudaStream_t s);
// stage-3 eigen step on the device: eig_sym(H) sorted descending, accu(), water-filling `err` and the kept eigenvalues
// per gamma (reference :492-520); Qd: K x K eigenvectors (descending), lam: ng x K, err: ng, dvals_out: K or nullptr
void launch_gamma_eigen(const double* H, int K, double tv, int p, const double* gamma_dev, int ng, double* Qd,
                        double* lam, double* err, double* dvals_out, cudaStream_t s);
// stage-3 covariance losses for up to SPB_GAMMA_BATCH gammas per launch:
// out[g] = || Yva'Yva / m - B diag(lam_g) B' - err_g I ||_F^2
constexpr int kGammaBatch = 16;
size_t gamma_scratch_doubles(int p);
void launch_gamma_combine(const double* cst, const double* Cv, const double* BtB, const double* lam, const double* err,
                          int ng, int K, int p, int m, double* out, cudaStream_t s);
void launch_gamma_loss(const double* Yva, int m, int p, const double* B, int K,
                       const double* lam_dev /* ng x K, row g contiguous */,
                       const double* err_dev /* ng */, int ng, double* scratch, double* out,
                       cudaStream_t s);

// ---- eigenFunction evaluation (tps_kernels.cu) ----------------------------------
// out (p_new x K) = kernel(new, orig) * para[0:p] + affine part (reference :113-145)
void launch_tps_eval(const double* xnew, int pnew, const double* xorig, int p, int d,
                     const double* para, size_t ldpara, int K, double* out, cudaStream_t s);

}  // namespace spb
