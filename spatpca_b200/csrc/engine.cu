// Engine implementation.  Control flow follows the *semantics* of the reference's
// spatpcaCV (src/RcppSpatPCA.cpp:543-701) including its quirks, but is
// organised per (stage, fold) so that folds can live on different devices; the data path is
// device-resident from upload() to get_eigenfn().
#include "engine.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <cstdio>
#include <map>
#include <mutex>
#include <numeric>
#include <thread>

namespace spb {

thread_local long long g_launch_count = 0;

// SPB_TRACE=1: host-side timeline of the prologue (stream-synchronised), for performance debugging
static bool trace_on() {
  static const bool on = [] { const char* e = std::getenv("SPB_TRACE"); return e && e[0] == '1'; }();
  return on;
}
static void trace(DeviceCtx& ctx, const char* label) {
  if (!trace_on()) return;
  static auto last = std::chrono::steady_clock::now();
  cudaStreamSynchronize(ctx.stream);
  auto now = std::chrono::steady_clock::now();
  fprintf(stderr, "[spb trace] %-28s +%8.3f ms\n", label, std::chrono::duration<double, std::milli>(now - last).count());
  last = now;
}

// ---------------------------------------------------------------------------------------
// caching allocator
// ---------------------------------------------------------------------------------------
namespace {
struct MemCache {
  std::mutex mu;
  std::map<std::pair<int, size_t>, std::vector<void*>> free_lists;   // (device, bytes) -> buffers
};
MemCache& mem_cache() {
  static MemCache* c = new MemCache();   // leaked on purpose: outlives every static destructor
  return *c;
}
size_t bucket(size_t bytes) { return (bytes + 511) / 512 * 512; }
}  // namespace

void cached_trim() {
  MemCache& c = mem_cache();
  std::lock_guard<std::mutex> lock(c.mu);
  int cur = 0;
  cudaGetDevice(&cur);
  for (auto& kv : c.free_lists) {
    if (kv.second.empty()) continue;
    cudaSetDevice(kv.first.first);
    for (void* p : kv.second) cudaFree(p);
    kv.second.clear();
  }
  cudaSetDevice(cur);
}

void* cached_alloc(size_t bytes) {
  const size_t b = bucket(bytes);
  int dev = 0;
  SPB_CUDA(cudaGetDevice(&dev));
  MemCache& c = mem_cache();
  {
    std::lock_guard<std::mutex> lock(c.mu);
    auto it = c.free_lists.find({dev, b});
    if (it != c.free_lists.end() && !it->second.empty()) {
      void* p = it->second.back();
      it->second.pop_back();
      return p;
    }
  }
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, b);
  if (e != cudaSuccess) {
    cudaGetLastError();
    cached_trim();
    e = cudaMalloc(&p, b);
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    throw Error(SPB_ERR_CUDA, std::string("cudaMalloc of ") + std::to_string(b) + " bytes failed: " +
                                  cudaGetErrorString(e));
  }
  return p;
}

void cached_free(void* ptr, size_t bytes) {
  if (!ptr) return;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return; }
  MemCache& c = mem_cache();
  std::lock_guard<std::mutex> lock(c.mu);
  c.free_lists[{dev, bucket(bytes)}].push_back(ptr);
}

// ---------------------------------------------------------------------------------------
// DeviceCtx
// ---------------------------------------------------------------------------------------
// SPB_PRIO=swap: the Omega build on the highest-priority stream and the SVD starts on the lowest (experiment: with the
// fold starts running concurrently on host threads they have slack, the Omega build is the critical path)
static bool prio_swapped() {
  static const bool v = [] { const char* e = std::getenv("SPB_PRIO"); return e && std::string(e) == "swap"; }();
  return v;
}

DeviceCtx& DeviceCtx::get() {
  static thread_local std::map<int, std::unique_ptr<DeviceCtx>> ctxs;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0) {
    cudaGetLastError();
    throw Error(SPB_ERR_NO_DEVICE,
                "no CUDA device available (spatpca_b200 has no CPU fallback)");
  }
  int dev = 0;
  SPB_CUDA(cudaGetDevice(&dev));
  auto it = ctxs.find(dev);
  if (it != ctxs.end()) return *it->second;
  std::unique_ptr<DeviceCtx> c = create(dev, true);
  DeviceCtx& ref = *c;
  ctxs[dev] = std::move(c);
  return ref;
}

std::unique_ptr<DeviceCtx> DeviceCtx::create(int dev, bool with_blas) {
  std::unique_ptr<DeviceCtx> c(new DeviceCtx());
  c->device = dev;
  // highest priority: the host-synchronous SVD starts (many small kernels) run on this stream while the Omega build
  // keeps every SM busy with GEMM waves on a lowest-priority stream; without the priorities each small kernel waits
  // for a whole GEMM grid to be dispatched (measured at p = 10 000: 38 ms per fold start instead of 9 ms)
  int prio_least = 0, prio_greatest = 0;
  SPB_CUDA(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
  SPB_CUDA(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio_swapped() ? prio_least : prio_greatest));
  if (with_blas) {
    SPB_CUBLAS(cublasCreate(&c->blas));
    SPB_CUBLAS(cublasSetStream(c->blas, c->stream));
    SPB_CUBLAS(cublasSetPointerMode(c->blas, CUBLAS_POINTER_MODE_HOST));
  }
  SPB_CUSOLVER(cusolverDnCreate(&c->solver));
  SPB_CUSOLVER(cusolverDnSetStream(c->solver, c->stream));
  c->info.alloc(4);
  c->flag.alloc(1);
  SPB_CUDA(cudaMemsetAsync(c->flag.get(), 0, sizeof(int), c->stream));
  return c;
}

// i-th pooled context for a host worker thread (own high-priority stream, cuSOLVER handle, workspace, error flag);
// created on the owning thread, kept for the life of this context
DeviceCtx& DeviceCtx::worker(int i) {
  while ((int)worker_pool.size() <= i) worker_pool.push_back(create(device, false));
  return *worker_pool[i];
}

OmegaRes::~OmegaRes() {
  if (solver) cusolverDnDestroy(solver);
  if (blas) cublasDestroy(blas);
  if (stream) cudaStreamDestroy(stream);
}

OmegaRes& DeviceCtx::omega() {
  if (!omega_res) {
    std::unique_ptr<OmegaRes> r(new OmegaRes());
    int prio_least = 0, prio_greatest = 0;
    SPB_CUDA(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
    SPB_CUDA(cudaStreamCreateWithPriority(&r->stream, cudaStreamNonBlocking, prio_swapped() ? prio_greatest : prio_least));
    SPB_CUBLAS(cublasCreate(&r->blas));
    SPB_CUBLAS(cublasSetStream(r->blas, r->stream));
    SPB_CUBLAS(cublasSetPointerMode(r->blas, CUBLAS_POINTER_MODE_HOST));
    SPB_CUSOLVER(cusolverDnCreate(&r->solver));
    SPB_CUSOLVER(cusolverDnSetStream(r->solver, r->stream));
    r->info.alloc(4);
    omega_res = std::move(r);
  }
  return *omega_res;
}

LaneRes::~LaneRes() {
  if (blas) cublasDestroy(blas);
  if (stream) cudaStreamDestroy(stream);
}

LaneRes& DeviceCtx::lane_res(int i) {
  while ((int)lane_pool.size() <= i) {
    std::unique_ptr<LaneRes> r(new LaneRes());
    // (distinct stream priorities per lane were tried and change nothing: profiles/r01_notes.md section 6)
    SPB_CUDA(cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking));
    SPB_CUBLAS(cublasCreate(&r->blas));
    SPB_CUBLAS(cublasSetStream(r->blas, r->stream));
    SPB_CUBLAS(cublasSetPointerMode(r->blas, CUBLAS_POINTER_MODE_HOST));
    lane_pool.push_back(std::move(r));
  }
  return *lane_pool[i];
}

DeviceCtx::~DeviceCtx() {
  lane_pool.clear();
  omega_res.reset();
  if (solver) cusolverDnDestroy(solver);
  if (blas) cublasDestroy(blas);
  if (stream) cudaStreamDestroy(stream);
}

void DeviceCtx::sync_and_check() {
  SPB_CUDA(cudaStreamSynchronize(stream));
  int f = 0;
  SPB_CUDA(cudaMemcpy(&f, flag.get(), sizeof(int), cudaMemcpyDeviceToHost));
  if (f != 0) {
    SPB_CUDA(cudaMemset(flag.get(), 0, sizeof(int)));
    switch (f) {
      case SPB_ERR_NOT_POSDEF:
        throw Error(f, "inv_sympd(): matrix is singular or not positive definite");
      case SPB_ERR_SINGULAR:
        throw Error(f, "inv(): matrix is singular");
      case SPB_ERR_SVD:
        throw Error(f, "svd(): decomposition failed");
      default:
        throw Error(f, "device-side failure");
    }
  }
}

// ---------------------------------------------------------------------------------------
// PhaseClock
// ---------------------------------------------------------------------------------------
cudaEvent_t PhaseClock::grab() {
  if (used_ == pool_.size()) {
    cudaEvent_t e;
    SPB_CUDA(cudaEventCreate(&e));
    pool_.push_back(e);
  }
  return pool_[used_++];
}
void PhaseClock::begin(Phase ph, cudaStream_t s) {
  if (open_ >= 0) return;   // nested spans are attributed to the outer phase
  Span sp; sp.ph = ph; sp.a = grab(); sp.b = grab();
  SPB_CUDA(cudaEventRecord(sp.a, s));
  spans_.push_back(sp);
  open_ = (int)spans_.size() - 1;
}
void PhaseClock::end(cudaStream_t s) {
  if (open_ < 0) return;
  SPB_CUDA(cudaEventRecord(spans_[open_].b, s));
  open_ = -1;
}
void PhaseClock::collect(double ms[PH_COUNT]) {
  for (auto& sp : spans_) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, sp.a, sp.b) == cudaSuccess) acc_[sp.ph] += t;
  }
  spans_.clear();
  used_ = 0;
  for (int i = 0; i < PH_COUNT; ++i) ms[i] = acc_[i];
}
void PhaseClock::reset() {
  spans_.clear(); used_ = 0; open_ = -1;
  for (int i = 0; i < PH_COUNT; ++i) acc_[i] = 0;
}
PhaseClock::~PhaseClock() { for (auto e : pool_) cudaEventDestroy(e); }

struct Scoped {
  PhaseClock& c; cudaStream_t s;
  Scoped(PhaseClock& c_, Phase ph, cudaStream_t s_) : c(c_), s(s_) { c.begin(ph, s); }
  ~Scoped() { try { c.end(s); } catch (...) {} }
};

// ---------------------------------------------------------------------------------------
// small host math
// ---------------------------------------------------------------------------------------
double arma_accumulate(const double* x, int n) {
  // Armadillo arrayops::accumulate: two interleaved accumulators (used by sum(cv, 0), accu()).
  double a1 = 0.0, a2 = 0.0;
  int i = 0, j = 1;
  for (; j < n; j += 2, i += 2) { a1 += x[i]; a2 += x[i + 1]; }
  if ((j - 1) < n) a1 += x[i];
  return a1 + a2;
}

// LU 2/3 + solve for [I; 0] 2 + E Lp 2 + upper half of Lp'(E Lp) 1   (x p^3)
static double omega_flops(int p) { return (2.0 / 3.0 + 2.0 + 2.0 + 1.0) * (double)p * p * p; }
static double factor_flops(int p, const Blocking& b) {
  const double p3 = (double)p * p * p;
  if (b.m <= 1) return p3;
  return p3 / 3.0 + p3 / b.m + p3 / ((double)b.m * b.m);
}

// ---------------------------------------------------------------------------------------
// static building blocks
// ---------------------------------------------------------------------------------------
bool Engine::omega_refine_enabled() {
  static const bool on = [] { const char* e = std::getenv("SPB_OMEGA_REFINE"); return !(e && e[0] == '0'); }();
  return on;
}

void Engine::build_omega_raw(OmegaRes& R, int* flag, const double* x_dev, int p, int d, double* omega_dev,
                             size_t ld, bool upper_only, OmegaTemps& tmp) {
  NvtxRange nvtx_range("spb::thinPlateSplineMatrix");
  if (omega_refine_enabled()) {
    build_omega_refined(R, flag, x_dev, p, d, omega_dev, ld, upper_only, tmp);
    return;
  }
  // thinPlateSplineMatrix (:58-76): Lp = top-left block of inv(L + 1e-8 I) by LU with partial
  // pivoting (the reference's dgetrf/dgetri class; here getrf + getrs on [I_p; 0]); Omega = Lp'(E Lp).
  const int q = p + d + 1;
  const size_t ldq = padded_ld(q);
  tmp.L.ensure(ldq * (size_t)q);
  tmp.Bm.ensure(ldq * (size_t)p);
  tmp.T1.ensure(ld * (size_t)p);
  double *L = tmp.L.get(), *Bm = tmp.Bm.get(), *T1 = tmp.T1.get();
  cudaStream_t s = R.stream;
  auto lap = [&](const char* label) {               // SPB_TRACE=1: synchronous per-step timing
    if (!trace_on()) return;
    static auto last = std::chrono::steady_clock::now();
    cudaStreamSynchronize(s);
    auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[spb omega] %-12s +%8.3f ms\n", label, std::chrono::duration<double, std::milli>(now - last).count());
    last = now;
  };
  lap("enter");
  launch_tps_fill(x_dev, p, d, L, ldq, 1e-8, true, s);
  launch_set_identity(Bm, q, p, ldq, s);
  int lwork = 0;
  SPB_CUSOLVER(cusolverDnDgetrf_bufferSize(R.solver, q, q, L, (int)ldq, &lwork));
  R.ipiv.ensure(q);
  R.work.ensure((size_t)lwork);
  SPB_CUSOLVER(cusolverDnDgetrf(R.solver, q, q, L, (int)ldq, R.work.get(), R.ipiv.get(), R.info.get()));
  launch_check_info(R.info.get(), flag, SPB_ERR_SINGULAR, s);
  lap("fill+getrf");
  SPB_CUSOLVER(cusolverDnDgetrs(R.solver, CUBLAS_OP_N, q, p, L, (int)ldq, R.ipiv.get(), Bm, (int)ldq, R.info.get()));
  lap("getrs");
  // E (p x p, zero diagonal) regenerated into L's storage
  launch_tps_fill(x_dev, p, d, L, ldq, 0.0, false, s);
  const double one = 1.0, zero = 0.0;
  SPB_CUBLAS(cublasDgemm(R.blas, CUBLAS_OP_N, CUBLAS_OP_N, p, p, p, &one, L, (int)ldq, Bm, (int)ldq, &zero, T1,
                         (int)ld));
  lap("E*Lp");
  if (upper_only) {
    // every downstream use of Omega is under symmatu (:165, 203, 397, 621, 661, 673)
    SPB_CUDA(cudaMemsetAsync(omega_dev, 0, sizeof(double) * ld * p, s));
    gemm_upper_tn(R.blas, p, p, 1.0, Bm, ldq, T1, ld, omega_dev, ld);
  } else {
    SPB_CUBLAS(cublasDgemm(R.blas, CUBLAS_OP_T, CUBLAS_OP_N, p, p, p, &one, Bm, (int)ldq, T1, (int)ld, &zero,
                           omega_dev, (int)ld));
  }
  lap("Lp'*(E*Lp)");
}

void Engine::spd_inverse(const Exec& ex, const double* omega_u, size_t ldo, const double* G, size_t ldg,
                         int p, double tau1, double rho, double scale, double* out, size_t ldout,
                         const Blocking& blk) {
  // inv_sympd(symmatu(scale*(tau1*Omega - G) + rho*I)) (:164-165, :397): upper-triangle assembly, then
  // the inverse -- or its block factor -- in full storage for the matrix-stream kernel.
  cudaStream_t s = ex.stream;
  launch_assemble_upper(omega_u, ldo, G, ldg, p, tau1, scale, rho, out, ldout, s);
  static const bool use_cusolver = [] {
    const char* e = std::getenv("SPB_INVERSE");
    return e != nullptr && std::string(e) == "cusolver";
  }();
  if (!use_cusolver) {
    // GEMM-rate recursive Schur-complement inverse (spd_inverse.cu); result is in full storage
    ex.inv_work->ensure(spd_inverse_workspace_doubles(p));
    if (blk.m > 1) spd_factor_blocked(ex, out, p, ldout, ex.inv_work->get(), blk);
    else spd_inverse_recursive(ex, out, p, ldout, ex.inv_work->get());
    return;
  }
  // cross-check path (SPB_INVERSE=cusolver): the reference's own algorithm class, dpotrf + dpotri.
  // cuSOLVER is bound to the main stream, so this path serialises the lanes.
  DeviceCtx& ctx = DeviceCtx::get();
  SPB_CUDA(cudaStreamSynchronize(s));
  int lw1 = 0, lw2 = 0;
  SPB_CUSOLVER(cusolverDnDpotrf_bufferSize(ctx.solver, CUBLAS_FILL_MODE_UPPER, p, out, (int)ldout, &lw1));
  SPB_CUSOLVER(cusolverDnDpotri_bufferSize(ctx.solver, CUBLAS_FILL_MODE_UPPER, p, out, (int)ldout, &lw2));
  int lw = std::max(lw1, lw2);
  double* work = ctx.work((size_t)lw);
  SPB_CUSOLVER(cusolverDnDpotrf(ctx.solver, CUBLAS_FILL_MODE_UPPER, p, out, (int)ldout, work, lw, ctx.info.get()));
  launch_check_info(ctx.info.get(), ctx.flag.get(), SPB_ERR_NOT_POSDEF, ctx.stream);
  SPB_CUSOLVER(cusolverDnDpotri(ctx.solver, CUBLAS_FILL_MODE_UPPER, p, out, (int)ldout, work, lw, ctx.info.get()));
  launch_check_info(ctx.info.get(), ctx.flag.get(), SPB_ERR_NOT_POSDEF, ctx.stream);
  launch_symmatu_inplace(out, p, ldout, 0.0, ctx.stream);
  SPB_CUDA(cudaStreamSynchronize(ctx.stream));
}

void Engine::gram_upper(const Exec& ex, const double* Ymat, int m, int p, double* G, size_t ldg) {
  const double one = 1.0, zero = 0.0;
  SPB_CUBLAS(cublasDsyrk(ex.blas, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_T, p, m, &one, Ymat, m, &zero, G, (int)ldg));
}

void Engine::svd_right(DeviceCtx& ctx, const double* Ymat, int m, int p, int K, std::vector<double>& sv,
                       double* Phi_dev) {
  // svd_econ(U, s, V, Y, "right") (:321, :563): singular values and right singular vectors of the
  // m x p block.  The long side is removed by a Householder QR first (what LAPACK's dgesvd does
  // for very rectangular inputs), the r x r triangular factor (r = min(m, p)) goes through
  // cuSOLVER's one-sided Jacobi SVD, and the K leading vectors are mapped back with Q.
  // (cusolverDnDgesvd on the full block took > 100 ms per fold at p = 900: profiles/r01_notes.)
  cudaStream_t s = ctx.stream;
  const int r = std::min(m, p);
  SPB_REQUIRE(K <= r, "K must not exceed min(rows of the training block, p)");
  const bool wide = (p >= m);                  // wide: factorise Y' (p x m); else Y (m x p)
  const int rows = wide ? p : m;
  trace(ctx, "svd: enter");
  DevBuf<double> A((size_t)rows * r), tau((size_t)r), Rm((size_t)r * r), S((size_t)r), U((size_t)r * r),
      V((size_t)r * r);
  if (wide) launch_transpose(Ymat, m, p, A.get(), s);
  else SPB_CUDA(cudaMemcpyAsync(A.get(), Ymat, sizeof(double) * (size_t)m * p, cudaMemcpyDeviceToDevice, s));
  int lq = 0, lo = 0, lj = 0;
  SPB_CUSOLVER(cusolverDnDgeqrf_bufferSize(ctx.solver, rows, r, A.get(), rows, &lq));
  double* work = ctx.work((size_t)lq);
  SPB_CUSOLVER(cusolverDnDgeqrf(ctx.solver, rows, r, A.get(), rows, tau.get(), work, lq, ctx.info.get()));
  launch_check_info(ctx.info.get(), ctx.flag.get(), SPB_ERR_SVD, s);
  trace(ctx, "svd: alloc+transpose+geqrf");
  launch_copy_upper(A.get(), (size_t)rows, r, Rm.get(), (size_t)r, s);           // R, zero below the diagonal
  gesvdjInfo_t gi = nullptr;
  SPB_CUSOLVER(cusolverDnCreateGesvdjInfo(&gi));
  try {
    SPB_CUSOLVER(cusolverDnXgesvdjSetMaxSweeps(gi, 100));
    SPB_CUSOLVER(cusolverDnXgesvdjSetSortEig(gi, 1));
    SPB_CUSOLVER(cusolverDnDgesvdj_bufferSize(ctx.solver, CUSOLVER_EIG_MODE_VECTOR, 1, r, r, Rm.get(), r, S.get(),
                                              U.get(), r, V.get(), r, &lj, gi));
    work = ctx.work((size_t)lj);
    SPB_CUSOLVER(cusolverDnDgesvdj(ctx.solver, CUSOLVER_EIG_MODE_VECTOR, 1, r, r, Rm.get(), r, S.get(), U.get(), r,
                                   V.get(), r, work, lj, ctx.info.get(), gi));
    launch_check_info(ctx.info.get(), ctx.flag.get(), SPB_ERR_SVD, s);
    trace(ctx, "svd: gesvdj");
    sv.assign(r, 0.0);
    if (wide) {
      // Y' = Q R = Q U_R S V_R'  =>  right singular vectors of Y are Q U_R
      DevBuf<double> Z((size_t)p * K);
      SPB_CUDA(cudaMemsetAsync(Z.get(), 0, sizeof(double) * (size_t)p * K, s));
      SPB_CUDA(cudaMemcpy2DAsync(Z.get(), (size_t)p * sizeof(double), U.get(), (size_t)r * sizeof(double),
                                 (size_t)r * sizeof(double), K, cudaMemcpyDeviceToDevice, s));
      SPB_CUSOLVER(cusolverDnDormqr_bufferSize(ctx.solver, CUBLAS_SIDE_LEFT, CUBLAS_OP_N, p, K, r, A.get(), rows,
                                               tau.get(), Z.get(), p, &lo));
      work = ctx.work((size_t)lo);
      SPB_CUSOLVER(cusolverDnDormqr(ctx.solver, CUBLAS_SIDE_LEFT, CUBLAS_OP_N, p, K, r, A.get(), rows, tau.get(),
                                    Z.get(), p, work, lo, ctx.info.get()));
      launch_check_info(ctx.info.get(), ctx.flag.get(), SPB_ERR_SVD, s);
      SPB_CUDA(cudaMemcpyAsync(Phi_dev, Z.get(), sizeof(double) * (size_t)p * K, cudaMemcpyDeviceToDevice, s));
      SPB_CUDA(cudaMemcpyAsync(sv.data(), S.get(), sizeof(double) * r, cudaMemcpyDeviceToHost, s));
      ctx.sync_and_check();
      trace(ctx, "svd: ormqr+copy");
    } else {
      // Y = Q R = Q U_R S V_R'  =>  right singular vectors of Y are V_R (r = p)
      SPB_CUDA(cudaMemcpyAsync(Phi_dev, V.get(), sizeof(double) * (size_t)p * K, cudaMemcpyDeviceToDevice, s));
      SPB_CUDA(cudaMemcpyAsync(sv.data(), S.get(), sizeof(double) * r, cudaMemcpyDeviceToHost, s));
      ctx.sync_and_check();
    }
  } catch (...) {
    cusolverDnDestroyGesvdjInfo(gi);
    throw;
  }
  cusolverDnDestroyGesvdjInfo(gi);
}

void Engine::cv_loss(cudaStream_t s, const double* Yva, int m, int p, const double* Phi, int K, double* W,
                     double* scratch, double* out_dev) {
  launch_yphi(Yva, m, p, Phi, K, W, scratch, s);
  launch_residual_sumsq(Yva, m, p, Phi, K, W, scratch, out_dev, s);
}

void GammaWork::ensure(int p, int K, int n_tr, int n_va, int ng, bool tiles) {
  const int rows = std::max(n_tr, n_va);
  W.ensure((size_t)n_tr * K);
  scr.ensure(yphi_scratch_doubles(rows, p, std::max(K, std::min(n_va, SPB_MAX_K))));
  H.ensure((size_t)K * K);
  Q.ensure((size_t)K * K);
  B.ensure((size_t)p * K);
  lam.ensure((size_t)ng * K);
  err.ensure((size_t)ng);
  if (tiles) {
    gscr.ensure(gamma_scratch_doubles(p));
  } else {
    Yvt.ensure((size_t)p * n_va);
    Gv.ensure((size_t)n_va * n_va);
    Wv.ensure((size_t)n_va * K);
    Cv.ensure((size_t)K * K);
    BtB.ensure((size_t)K * K);
    cst.ensure(2);
    red.ensure(1024);
  }
}

void Engine::gamma_losses_async(cudaStream_t s, GammaWork& w, const double* Phi, int p, int K, const double* Ytr, int n_tr,
                                double tv, const double* Yva, int n_va, const double* gamma_dev, int ng, double* out_dev) {
  // stage-3 body for one fold (:489-523), entirely on stream `s`: H = Phi' S_tr Phi through W = Ytr Phi, its
  // eigen-decomposition + water filling in a one-warp kernel (no host round trip), B = Phi Q, then the covariance
  // losses of all gammas without ever forming a p x p matrix.
  // SPB_GAMMA_LOSS=tiles evaluates the p x p residual tile by tile (round 1 / 2a: 5.2 ms per fold at p = 10 000,
  // FMA-bound); the default expands the norm into O(K^2) statistics (launch_gamma_combine: ~0.1 ms).
  static const bool tiles = [] { const char* e = std::getenv("SPB_GAMMA_LOSS"); return e && std::string(e) == "tiles"; }();
  const bool use_tiles = tiles || n_va == 0;
  w.ensure(p, K, n_tr, n_va, ng, use_tiles);
  launch_yphi(Ytr, n_tr, p, Phi, K, w.W.get(), w.scr.get(), s);
  launch_small_gram(w.W.get(), n_tr, K, 1.0 / (double)n_tr, w.H.get(), s);
  launch_gamma_eigen(w.H.get(), K, tv, p, gamma_dev, ng, w.Q.get(), w.lam.get(), w.err.get(), nullptr, s);
  launch_small_rmul(Phi, p, K, w.Q.get(), w.B.get(), s);
  if (use_tiles) {
    for (int g0 = 0; g0 < ng; g0 += kGammaBatch) {
      const int nb = std::min(kGammaBatch, ng - g0);
      launch_gamma_loss(Yva, n_va, p, w.B.get(), K, w.lam.get() + (size_t)g0 * K, w.err.get() + g0, nb, w.gscr.get(),
                        out_dev + g0, s);
    }
    return;
  }
  // Gv = Yva Yva' / m in column chunks of <= SPB_MAX_K (the W = Y Phi kernel with Phi = a block of Yva')
  launch_transpose(Yva, n_va, p, w.Yvt.get(), s);                               // p x n_va
  for (int c0 = 0; c0 < n_va; c0 += SPB_MAX_K) {
    const int nc = std::min(SPB_MAX_K, n_va - c0);
    launch_yphi(Yva, n_va, p, w.Yvt.get() + (size_t)c0 * p, nc, w.Gv.get() + (size_t)c0 * n_va, w.scr.get(), s);
  }
  launch_sumsq(w.Gv.get(), (size_t)n_va * n_va, w.red.get(), w.cst.get(), s);                 // m^2 |S_va|_F^2
  launch_sumsq(Yva, (size_t)n_va * p, w.red.get() + 512, w.cst.get() + 1, s);                 // m tr(S_va)
  launch_yphi(Yva, n_va, p, w.B.get(), K, w.Wv.get(), w.scr.get(), s);
  launch_small_gram(w.Wv.get(), n_va, K, 1.0 / (double)n_va, w.Cv.get(), s);                  // B' S_va B
  launch_small_gram(w.B.get(), p, K, 1.0, w.BtB.get(), s);
  launch_gamma_combine(w.cst.get(), w.Cv.get(), w.BtB.get(), w.lam.get(), w.err.get(), ng, K, p, n_va, out_dev, s);
}

void Engine::gamma_losses(DeviceCtx& ctx, cudaStream_t s, const double* Phi, int p, int K, const double* Ytr,
                          int n_tr, double tv, const double* Yva, int n_va, const std::vector<double>& gamma,
                          double* out_host) {
  const int ng = (int)gamma.size();
  GammaWork w;
  DevBuf<double> gdev((size_t)ng), out_dev((size_t)ng);
  SPB_CUDA(cudaMemcpyAsync(gdev.get(), gamma.data(), sizeof(double) * ng, cudaMemcpyHostToDevice, s));
  gamma_losses_async(s, w, Phi, p, K, Ytr, n_tr, tv, Yva, n_va, gdev.get(), ng, out_dev.get());
  SPB_CUDA(cudaMemcpyAsync(out_host, out_dev.get(), sizeof(double) * ng, cudaMemcpyDeviceToHost, s));
  SPB_CUDA(cudaStreamSynchronize(s));
  ctx.sync_and_check();
}

// ---------------------------------------------------------------------------------------
// Engine
// ---------------------------------------------------------------------------------------
Engine::Engine(int p, int d, int n, int M, int K, const double* tau1, int n1, const double* tau2, int n2,
               const double* gamma, int n3, int maxit, double tol, const double* l2, int nl2)
    : p_(p), d_(d), n_(n), M_(M), K_(K), maxit_(maxit), tol_(tol), ctx_(DeviceCtx::get()) {
  SPB_REQUIRE(p >= 1 && n >= 1, "p and n must be positive");
  SPB_REQUIRE(d >= 1 && d <= 3, "Dimension of locations must be less than 4.");
  SPB_REQUIRE(M >= 1, "M must be positive");
  SPB_REQUIRE(K >= 1 && K <= SPB_MAX_K, "K must be in 1..SPB_MAX_K");
  SPB_REQUIRE(K <= p && K <= n, "K must not exceed min(n, p)");
  SPB_REQUIRE(n1 >= 1 && n2 >= 1 && n3 >= 1, "tuning grids must not be empty");
  SPB_REQUIRE(tau1 && tau2 && gamma, "NULL tuning grid");
  SPB_REQUIRE(maxit >= 0, "maxit must be non-negative");
  tau1_.assign(tau1, tau1 + n1);
  tau2_.assign(tau2, tau2 + n2);
  gamma_.assign(gamma, gamma + n3);
  if (l2 && nl2 > 0) l2_.assign(l2, l2 + nl2);
  any_tau_ = (*std::max_element(tau1_.begin(), tau1_.end()) != 0.0) ||
             (*std::max_element(tau2_.begin(), tau2_.end()) != 0.0);
  ld_ = padded_ld((size_t)p);
  blocks_ = spd_blocking(p);
  blocks_path_ = spd_blocking_path(p);
  folds_.resize(M);
  max_grid_ = std::max({n1, n2, n3, 1});
  loss_out_.alloc((size_t)(M + 1) * max_grid_);
  prep_scratch_.alloc(512);
  slot_cap_ = (n1 + 2 * n2 + nl2 + 8) * (M + 2);

  // lanes: one per fold if memory allows (each needs ~2.5 p x p matrices), plus the full-data lane
  const double per_lane = 2.5 * (double)ld_ * p * sizeof(double) + 64e6;
  size_t free_b = 0, total_b = 0;
  SPB_CUDA(cudaMemGetInfo(&free_b, &total_b));
  const double fixed = (double)(M + 2) * ld_ * p * sizeof(double);      // Omega + per-fold tempinv + slack
  int nl = (int)std::floor(((double)free_b * 0.8 - fixed) / per_lane) - 1;
  if (const char* e = std::getenv("SPB_LANES")) nl = std::atoi(e);
  n_fold_lanes_ = std::max(1, std::min(M, nl));
  // The stage-1 inverse at the selected tau1 is needed again by stage 2 (:393) and, when |tau2| = 1,
  // by stage 3 (:475): keep the (fold, tau1) inverses when they fit comfortably.
  {
    const double cache_bytes = (double)(n1 > 1 ? n1 - 1 : 0) * M * ld_ * p * sizeof(double);
    const double left = (double)free_b * 0.8 - fixed - (n_fold_lanes_ + 1) * per_lane;
    cache_inverses_ = n1 > 1 && cache_bytes <= 0.7 * left;
    if (const char* e = std::getenv("SPB_CACHE_INVERSES")) cache_inverses_ = (std::atoi(e) != 0) && n1 > 1;
  }
  for (int i = 0; i <= n_fold_lanes_; ++i) {
    std::unique_ptr<Lane> L(new Lane());
    L->res = &ctx_.lane_res(i);
    L->stat_slots.alloc((size_t)slot_cap_ * 4);
    L->err_slots.alloc((size_t)slot_cap_);
    lanes_.push_back(std::move(L));
  }
  if (const char* e = std::getenv("SPB_CORE2")) {
    const std::string m(e);
    cg_mode_ = (m == "cg") ? 1 : ((m == "factor") ? 2 : 0);
  }
  if (const char* e = std::getenv("SPB_BATCH")) bcg_mode_ = (std::string(e) == "0") ? 2 : 0;
  if (const char* e = std::getenv("SPB_CORE3")) core3_mode_ = (std::string(e) == "factor") ? 2 : 0;
  if (const char* e = std::getenv("SPB_CG_TOL")) { const double t = std::atof(e); if (t > 0.0 && t < 1e-6) cg_tol_ = t; }
  alloc_k_state();
  std::memset(&stats_, 0, sizeof(stats_));
  launches_at_start_ = g_launch_count;
  t_created = std::chrono::steady_clock::now();
}

void Engine::alloc_k_state() {
  const size_t pk = (size_t)p_ * K_;
  full_.alloc(pk);
  for (auto& L : lanes_) {
    L->chain.alloc(pk);
    L->admm_ws.alloc(std::max({admm_workspace_bytes(p_, K_), cg_workspace_bytes(p_, K_, n_),
                               batched_core2_workspace_bytes(p_, n_, K_, 1)}));
    L->lossW.alloc((size_t)n_ * K_);
    L->loss_scratch.alloc(yphi_scratch_doubles(n_, p_, K_) + 512);
  }
  for (auto& f : folds_) {
    if (f.prepared) { f.Phi0.alloc(pk); f.L20.alloc(pk); }
  }
}

void Engine::reset_k(int K) {
  SPB_REQUIRE(K >= 1 && K <= SPB_MAX_K, "K must be in 1..SPB_MAX_K");
  SPB_REQUIRE(K <= p_ && K <= n_, "K must not exceed min(n, p)");
  for (auto& f : folds_) SPB_REQUIRE(K <= std::min(f.n_tr, p_), "K must not exceed min(training rows, p)");
  flush();
  K_ = K;
  pend_core2p_ = -1; pend_path_ = false;
  if (!cache_inverses_) for (auto& f : folds_) f.tinv_valid = false;
  alloc_k_state();
  full_ready_ = false;                          // prepare() re-derives the full-data start for the new K
  log_.clear();
  const long long h2d = stats_.bytes_h2d;
  std::memset(&stats_, 0, sizeof(stats_));
  stats_.bytes_h2d = h2d;
  clock_.reset();
  prologue_wall_ms_ = 0.0;
  for (auto& L : lanes_) L->clock.reset();
  launches_at_start_ = g_launch_count;
  t_created = std::chrono::steady_clock::now();
}

Engine::~Engine() {
  // lanes may still have work in flight if a call failed half-way: drain before buffers are recycled
  if (omega_pending_) cudaStreamSynchronize(ctx_.omega().stream);
  for (auto& L : lanes_) cudaStreamSynchronize(L->res->stream);
  cudaStreamSynchronize(ctx_.stream);
  if (omega_evt_) cudaEventDestroy(omega_evt_);
  for (auto& L : lanes_) if (L->evt) cudaEventDestroy(L->evt);
}

Lane& Engine::lane_of(int fold) {
  if (fold < 0) return *lanes_[n_fold_lanes_];
  return *lanes_[fold % n_fold_lanes_];
}

double Engine::selected_tau1(int index1) const {
  if (tau1_.size() > 1) return tau1_[index1];
  return *std::max_element(tau1_.begin(), tau1_.end());
}

void Engine::copy_pk(cudaStream_t s, double* dst, const double* src) {
  SPB_CUDA(cudaMemcpyAsync(dst, src, sizeof(double) * (size_t)p_ * K_, cudaMemcpyDeviceToDevice, s));
}
void Engine::zero_pk(cudaStream_t s, double* dst) {
  SPB_CUDA(cudaMemsetAsync(dst, 0, sizeof(double) * (size_t)p_ * K_, s));
}

void Engine::upload(const double* sxy, const double* Y, const double* nk) {
  SPB_REQUIRE(sxy && Y && nk, "NULL input");
  auto t0 = std::chrono::steady_clock::now();
  dY_.alloc((size_t)n_ * p_);
  dX_.alloc((size_t)p_ * d_);
  SPB_CUDA(cudaMemcpyAsync(dY_.get(), Y, sizeof(double) * (size_t)n_ * p_, cudaMemcpyHostToDevice, ctx_.stream));
  SPB_CUDA(cudaMemcpyAsync(dX_.get(), sxy, sizeof(double) * (size_t)p_ * d_, cudaMemcpyHostToDevice, ctx_.stream));
  nk_.assign(nk, nk + n_);
  for (int k = 0; k < M_; ++k) {
    FoldData& f = folds_[k];
    f = FoldData();
    for (int r = 0; r < n_; ++r) {
      if (nk_[r] == (double)(k + 1)) f.rows_va.push_back(r); else f.rows_tr.push_back(r);   // :318-319
    }
    f.n_tr = (int)f.rows_tr.size();
    f.n_va = (int)f.rows_va.size();
    SPB_REQUIRE(f.n_tr >= 1, "a fold has no training rows");
    SPB_REQUIRE(K_ <= std::min(f.n_tr, p_), "K must not exceed min(training rows, p)");
  }
  dYt_.alloc((size_t)n_ * p_);
  launch_transpose(dY_.get(), n_, p_, dYt_.get(), ctx_.stream);             // p x n, for the matrix-free Core2p
  {
    std::vector<int> nki(n_);
    for (int r = 0; r < n_; ++r) nki[r] = (int)nk[r];
    dNk_.alloc(n_);
    SPB_CUDA(cudaMemcpyAsync(dNk_.get(), nki.data(), sizeof(int) * n_, cudaMemcpyHostToDevice, ctx_.stream));
    SPB_CUDA(cudaStreamSynchronize(ctx_.stream));
  }
  SPB_CUDA(cudaStreamSynchronize(ctx_.stream));
  stats_.bytes_h2d += (long long)sizeof(double) * ((size_t)n_ * p_ + (size_t)p_ * d_ + n_);
  stats_.ms_h2d += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  omega_ready_ = omega_injected_;
  if (!omega_injected_) { lmax_ = -1.0; lmax_enqueued_ = false; }
  full_ready_ = false;
  full_svd_done_ = false;
  full_inv2_index_ = full_tinv_index_ = -1;
  pend_core2p_ = -1; pend_path_ = false; my_stage1_folds_ = 0;
  for (auto& L : lanes_) L->gram_owner = -2;
}

void Engine::set_omega(const double* omega_host) {
  SPB_REQUIRE(omega_host, "NULL omega");
  omega_u_.alloc(ld_ * (size_t)p_);
  omega_diag_raw_.alloc(p_);
  SPB_CUDA(cudaMemsetAsync(omega_u_.get(), 0, sizeof(double) * ld_ * p_, ctx_.stream));
  SPB_CUDA(cudaMemcpy2DAsync(omega_u_.get(), ld_ * sizeof(double), omega_host, (size_t)p_ * sizeof(double),
                             (size_t)p_ * sizeof(double), p_, cudaMemcpyHostToDevice, ctx_.stream));
  SPB_CUDA(cudaMemcpy2DAsync(omega_diag_raw_.get(), sizeof(double), omega_u_.get(), (ld_ + 1) * sizeof(double),
                             sizeof(double), p_, cudaMemcpyDeviceToDevice, ctx_.stream));
  launch_symmatu_inplace(omega_u_.get(), p_, ld_, 1e-8, ctx_.stream);     // symmatu + the ridge of :574
  lmax_ = -1.0;
  enqueue_lmax(ctx_.stream, ctx_.blas);
  SPB_CUDA(cudaStreamSynchronize(ctx_.stream));
  omega_ready_ = true;
  omega_injected_ = true;
}

// lambda_max(Omega_u) by 12 steps of power iteration from a fixed start (DGEMV + one-CTA normalisation, ~2 ms at
// p = 10 000, on the stream that built Omega).  Only the choice between the matrix-free and the factored Core2
// depends on it (use_cg multiplies the estimate by 1.5), never a result.
void Engine::enqueue_lmax(cudaStream_t s, cublasHandle_t blas) {
  constexpr int kSteps = 12;
  lmax_vec_.ensure(2 * (size_t)p_);
  lmax_out_.ensure(kSteps);
  double* x = lmax_vec_.get();
  double* y = x + p_;
  launch_power_init(x, p_, s);
  const double one = 1.0, zero = 0.0;
  for (int it = 0; it < kSteps; ++it) {
    SPB_CUBLAS(cublasDgemv(blas, CUBLAS_OP_N, p_, p_, &one, omega_u_.get(), (int)ld_, x, 1, &zero, y, 1));
    launch_power_normalize(y, p_, x, lmax_out_.get() + it, s);
  }
  lmax_enqueued_ = true;
}

double Engine::omega_lmax() {
  if (lmax_ >= 0.0) return lmax_;
  if (!any_tau_) { lmax_ = 1.0; return lmax_; }                    // Omega = I when all taus are 0 (:578)
  if (!lmax_enqueued_) return 1e300;                                // no estimate: the factored path
  finish_omega();
  SPB_CUDA(cudaStreamSynchronize(ctx_.stream));
  double h[12];
  SPB_CUDA(cudaMemcpy(h, lmax_out_.get(), sizeof(h), cudaMemcpyDeviceToHost));
  double m = 0.0;
  for (int i = 0; i < 12; ++i) if (h[i] == h[i]) m = std::max(m, h[i]);   // |Omega x| with |x| = 1 (the first entry has |x| != 1)
  lmax_ = std::max(m, h[11]);
  return lmax_;
}

bool Engine::use_cg(double c_omega, double rho, double gram_top, int expected_iters, double pass_scale) {
  if (K_ > kCgMaxK || cg_mode_ == 2) return false;
  if (cg_mode_ == 1) return true;
  if (p_ < 2048) return false;                   // small systems: latency-bound, the explicit inverse wins
  const double lo = rho - gram_top;              // lambda_min(A) >= rho - c_gram sigma_1^2 (Omega_u is PSD up to rounding)
  if (!(lo > 0.0)) return false;
  const double kappa = (rho + c_omega * 1.5 * omega_lmax()) / lo;
  if (!(kappa < 1e6)) return false;
  const double sq = std::sqrt(kappa);
  const double rate = (sq - 1.0) / (sq + 1.0);
  const double cold = rate <= 1e-12 ? 1.0 : std::ceil(std::log(0.05 * cg_tol_) / std::log(rate));
  // Cost model in units of one p x p matrix pass, calibrated on B200 at p = 10 000 (profiles/r02_notes.md): a block
  // factorisation costs ~65 + 0.0085 p passes and each ADMM iteration on it ~2.4.  A CG solve from the PCA start takes
  // ~0.7 of the cold-start bound; every later solve of the chain is warm-started by the previous iterate and takes
  // about a quarter of it (measured: 19 / 10 / 7 passes per ADMM iteration for the cold Core2 call, a tau2 path and a
  // tau1 path), plus one pass for the residual now and then (it is recycled between solves on the same matrix).
  const double factor_cost = 65.0 + 0.0085 * p_ + 2.4 * expected_iters;
  const double warm = 0.25 * cold + 1.0;
  const double cg_cost = 0.7 * cold + (expected_iters - 1) * warm + 0.25 * expected_iters + 1.0;
  // pass_scale: cost of a pass per problem relative to a single-problem pass (the fold-batched DMMA stream serves F
  // problems with one pass that takes ~2.1 single passes)
  return cg_cost * pass_scale < factor_cost;
}

bool Engine::core2_matrix_free(int fold, double tau1, double rho, int expected_iters, double pass_scale) {
  if (maxit_ <= 0) return false;
  const double sv1 = fold < 0 ? sv_full_[0] : folds_[fold].sv[0];
  return use_cg(2.0 * tau1, rho, 2.0 * sv1 * sv1, tol_ > 0.0 ? expected_iters : maxit_, pass_scale);
}

bool Engine::core3_matrix_free(int fold, double tau1, double rho, int expected_iters, double pass_scale) {
  // spatpcaCore3's system (:247 with tempinv of :397) is 2 tau1 Omega - 2 G + 2 rho I: twice Core2's shift
  if (maxit_ <= 0) return false;
  const double sv1 = fold < 0 ? sv_full_[0] : folds_[fold].sv[0];
  return use_cg(2.0 * tau1, 2.0 * rho, 2.0 * sv1 * sv1, tol_ > 0.0 ? expected_iters : maxit_, pass_scale);
}

bool Engine::can_share_inverses() {
  if (!cache_inverses_) return false;
  if (!full_svd_done_) return true;
  const double tmax = *std::max_element(tau1_.begin(), tau1_.end());
  return !use_cg(2.0 * tmax, rho_full_, 2.0 * sv_full_[0] * sv_full_[0], tol_ > 0.0 ? 3 : maxit_);
}

void Engine::init_lambda(cudaStream_t s, const double* Phi, const std::vector<double>& sv, double rho, double* L2) {
  // Phi * diagmat(rho - 1 / (rho - 2 * pow(sv, 2)))  (:326, :567, :617, :642)
  double coef[SPB_MAX_K];
  for (int k = 0; k < K_; ++k) {
    volatile double sq = sv[k] * sv[k];
    volatile double den = rho - 2.0 * sq;
    volatile double inv = 1.0 / den;
    coef[k] = rho - inv;
  }
  launch_lambda_init_host(Phi, p_, K_, coef, L2, s);
}

void Engine::prepare() {
  NvtxRange nvtx_range("spb::prepare (SVD start, Omega build enqueue)");
  SPB_REQUIRE(dY_.get() != nullptr, "upload() must be called before prepare()");
  trace(ctx_, "prepare: enter");
  cudaStream_t s = ctx_.stream;
  // Omega first: its build (GEMM-bound, own stream) runs under the host-synchronous SVD starts below
  const bool need_omega = any_tau_ || tau1_.size() > 1 || tau2_.size() > 1;
  if (need_omega && !omega_ready_) {
    omega_u_.alloc(ld_ * (size_t)p_);
    omega_diag_raw_.alloc(p_);
    if (any_tau_) {
      // enqueued on the Omega stream and not waited for here: the fold SVD prologues run meanwhile, every
      // lane waits for `omega_evt_` before its first use of Omega (wait_omega)
      OmegaRes& R = ctx_.omega();
      SPB_CUDA(cudaStreamSynchronize(s));                                     // dX_ was uploaded on the main stream
      omega_tmp_.reset(new OmegaTemps());
      {
        Scoped so(clock_, PH_OMEGA, R.stream);
        build_omega_raw(R, ctx_.flag.get(), dX_.get(), p_, d_, omega_u_.get(), ld_, true, *omega_tmp_);   // :574
        stats_.p3_flops += omega_refine_enabled() ? omega_refined_flops(p_) : omega_flops(p_);
        SPB_CUDA(cudaMemcpy2DAsync(omega_diag_raw_.get(), sizeof(double), omega_u_.get(),
                                   (ld_ + 1) * sizeof(double), sizeof(double), p_, cudaMemcpyDeviceToDevice,
                                   R.stream));
        launch_symmatu_inplace(omega_u_.get(), p_, ld_, 1e-8, R.stream);
        enqueue_lmax(R.stream, R.blas);
      }
      if (!omega_evt_) SPB_CUDA(cudaEventCreateWithFlags(&omega_evt_, cudaEventDisableTiming));
      SPB_CUDA(cudaEventRecord(omega_evt_, R.stream));
      omega_pending_ = true;
      for (auto& L : lanes_) L->omega_synced = false;
      // the temporaries (3 p x p matrices) stay alive until the first flush; when memory is tight
      // (C5: 38 GB) finish the build here instead so that the lanes can have it
      size_t free_b = 0, total_b = 0;
      SPB_CUDA(cudaMemGetInfo(&free_b, &total_b));
      // 3 (plain) / 5 (refined) p x p FP64 temporaries, plus 2 x 12 INT8 planes and the int32 product
      const double tmp_mats = !omega_refine_enabled() ? 3.0 : (omega_int8_enabled() ? 8.5 : 5.0);
      if (tmp_mats * (double)ld_ * p_ * sizeof(double) > 0.15 * (double)free_b) finish_omega();
    } else {
      // all taus are zero but a grid has several entries: Omega = I (:578); tau * Omega vanishes.
      launch_set_identity(omega_u_.get(), p_, p_, ld_, s);
    }
    omega_ready_ = true;
  }
  full_svd_start();
  if (!full_ready_) {
    copy_pk(s, full_.Phi.get(), Vfull_.get());                                // :565 (first K columns)
    copy_pk(s, full_.C.get(), full_.Phi.get());                               // :566
    init_lambda(s, full_.Phi.get(), sv_full_, rho_full_, full_.L2.get());     // :567
    full_ready_ = true;
  }
  ctx_.sync_and_check();            // the lanes read Omega and the full-data start from other streams
  trace(ctx_, "prepare: done");
}

void Engine::full_svd_start() {
  if (full_svd_done_) return;
  Scoped sc(clock_, PH_PROLOGUE, ctx_.stream);
  kcap_full_ = std::min(std::min(n_, p_), SPB_MAX_K);
  Vfull_.alloc((size_t)p_ * kcap_full_);
  svd_right(ctx_, dY_.get(), n_, p_, kcap_full_, sv_full_, Vfull_.get());     // :563
  rho_full_ = 10.0 * (sv_full_[0] * sv_full_[0]);                             // :564
  full_svd_done_ = true;
}

// The SVD starts of the full data and of the given folds ahead of prepare() / stage1_folds().  The column-split Omega
// build of a multi-GPU job is driven from the host (synchronise + exchange between its phases), so nothing can hide
// under it unless the caller asks for it: distributed.prepare_sharded() calls this right after omega_begin(), and
// the starts (~25 ms + ~20 ms, latency-bound) then run under the replicated X0 stage instead of after the build.
void Engine::prepare_starts(const int* folds, int nfolds) {
  SPB_REQUIRE(dY_.get() != nullptr, "upload() must be called before prepare_starts()");
  SPB_REQUIRE(nfolds == 0 || folds != nullptr, "invalid fold list");
  for (int f = 0; f < nfolds; ++f) SPB_REQUIRE(folds[f] >= 0 && folds[f] < M_, "fold index out of range");
  SPB_CUDA(cudaStreamSynchronize(ctx_.stream));
  full_svd_start();
  prepare_folds(folds, nfolds);
}

void Engine::wait_omega(Lane& L) {
  // The first consumer of Omega waits for the build on the HOST (it is on the critical path anyway) so that the
  // result can be checked -- and rebuilt by the LU route if the SPD route broke down -- before anything reads it.
  finish_omega();
}

void Engine::finish_omega() {
  if (!omega_pending_) return;
  OmegaRes& R = ctx_.omega();
  SPB_CUDA(cudaStreamSynchronize(R.stream));
  bool ok = true;
  if (omega_tmp_ && omega_refine_enabled()) {
    ok = omega_build_ok(*omega_tmp_, &stats_.omega_x0_resid);
    stats_.omega_route = omega_tmp_->used_spd ? (ok ? 1 : 3) : 2;
    if (trace_on()) fprintf(stderr, "[spb omega] route %d, max|I - A X0| = %.3e\n", stats_.omega_route, stats_.omega_x0_resid);
  }
  if (!ok) {
    // the SPD route to X0 broke down in FP64 (cond(B) ~ 1e11, e.g. one-dimensional sites): rebuild with the pivoted LU
    omega_tmp_->force_lu = true;
    build_omega_raw(R, ctx_.flag.get(), dX_.get(), p_, d_, omega_u_.get(), ld_, true, *omega_tmp_);
    SPB_CUDA(cudaMemcpy2DAsync(omega_diag_raw_.get(), sizeof(double), omega_u_.get(), (ld_ + 1) * sizeof(double),
                               sizeof(double), p_, cudaMemcpyDeviceToDevice, R.stream));
    launch_symmatu_inplace(omega_u_.get(), p_, ld_, 1e-8, R.stream);
    lmax_ = -1.0;
    enqueue_lmax(R.stream, R.blas);
    SPB_CUDA(cudaStreamSynchronize(R.stream));
  }
  omega_tmp_.reset();
  omega_pending_ = false;
}

// ---- column-sharded Omega build (multi-GPU) ---------------------------------------------------
bool Engine::omega_begin() {
  SPB_REQUIRE(dX_.get() != nullptr, "upload() must be called before omega_begin()");
  const bool need_omega = any_tau_ || tau1_.size() > 1 || tau2_.size() > 1;
  if (!need_omega || !any_tau_ || omega_ready_) return false;
  if (omega_refine_enabled()) return omega_begin_refined();
  OmegaRes& R = ctx_.omega();
  cudaStream_t s = R.stream;
  SPB_CUDA(cudaStreamSynchronize(ctx_.stream));
  const int q = p_ + d_ + 1;
  const size_t ldq = padded_ld(q);
  omega_u_.alloc(ld_ * (size_t)p_);
  omega_diag_raw_.alloc(p_);
  omega_tmp_.reset(new OmegaTemps());
  OmegaTemps& T = *omega_tmp_;
  T.L.ensure(ldq * (size_t)q);
  T.Bm.ensure(ldq * (size_t)p_);
  T.T1.ensure(ld_ * (size_t)p_);
  clock_.begin(PH_OMEGA, s);
  launch_tps_fill(dX_.get(), p_, d_, T.L.get(), ldq, 1e-8, true, s);
  launch_set_identity(T.Bm.get(), q, p_, ldq, s);
  int lwork = 0;
  SPB_CUSOLVER(cusolverDnDgetrf_bufferSize(R.solver, q, q, T.L.get(), (int)ldq, &lwork));
  R.ipiv.ensure(q);
  R.work.ensure((size_t)lwork);
  SPB_CUSOLVER(cusolverDnDgetrf(R.solver, q, q, T.L.get(), (int)ldq, R.work.get(), R.ipiv.get(), R.info.get()));
  launch_check_info(R.info.get(), ctx_.flag.get(), SPB_ERR_SINGULAR, s);
  clock_.end(s);
  omega_cols_open_ = true;
  return true;
}

void Engine::omega_solve_cols(int c0, int c1) {
  SPB_REQUIRE(omega_cols_open_, "omega_begin() must be called first");
  SPB_REQUIRE(c0 >= 0 && c0 <= c1 && c1 <= p_, "column range out of bounds");
  omega_x0_resolve();
  OmegaRes& R = ctx_.omega();
  const int q = p_ + d_ + 1;
  const size_t ldq = padded_ld(q);
  if (omega_refine_enabled() && omega_x0_is_spd_) return;   // X0 was built whole by omega_begin()
  if (omega_refine_enabled() && c1 == p_) c1 = q;       // refined build: the last block's owner also solves the border columns
  if (c0 == c1) return;
  clock_.begin(PH_OMEGA, R.stream);
  SPB_CUSOLVER(cusolverDnDgetrs(R.solver, CUBLAS_OP_N, q, c1 - c0, omega_tmp_->L.get(), (int)ldq, R.ipiv.get(),
                                omega_tmp_->Bm.get() + (size_t)c0 * ldq, (int)ldq, R.info.get()));
  clock_.end(R.stream);
}

void Engine::omega_product_cols(int c0, int c1) {
  SPB_REQUIRE(omega_cols_open_, "omega_begin() must be called first");
  SPB_REQUIRE(c0 >= 0 && c0 <= c1 && c1 <= p_, "column range out of bounds");
  omega_x0_resolve();
  if (omega_refine_enabled()) { omega_product_cols_refined(c0, c1); return; }
  OmegaRes& R = ctx_.omega();
  cudaStream_t s = R.stream;
  const int q = p_ + d_ + 1;
  const size_t ldq = padded_ld(q);
  OmegaTemps& T = *omega_tmp_;
  clock_.begin(PH_OMEGA, s);
  // E (p x p, zero diagonal) regenerated into L's storage: the LU factors are not needed any more
  launch_tps_fill(dX_.get(), p_, d_, T.L.get(), ldq, 0.0, false, s);
  if (c1 > c0) {
    const double one = 1.0, zero = 0.0;
    const int nc = c1 - c0;
    SPB_CUBLAS(cublasDgemm(R.blas, CUBLAS_OP_N, CUBLAS_OP_N, p_, nc, p_, &one, T.L.get(), (int)ldq,
                           T.Bm.get() + (size_t)c0 * ldq, (int)ldq, &zero, T.T1.get() + (size_t)c0 * ld_, (int)ld_));
    // rows 0 .. c1-1 of these columns cover the upper triangle, which is all that is used downstream
    SPB_CUBLAS(cublasDgemm(R.blas, CUBLAS_OP_T, CUBLAS_OP_N, c1, nc, p_, &one, T.Bm.get(), (int)ldq,
                           T.T1.get() + (size_t)c0 * ld_, (int)ld_, &zero, omega_u_.get() + (size_t)c0 * ld_, (int)ld_));
  }
  clock_.end(s);
}

void* Engine::omega_buffer(int which, int c0, int c1, long long* bytes) {
  SPB_REQUIRE(omega_cols_open_, "omega_begin() must be called first");
  SPB_REQUIRE(c0 >= 0 && c0 <= c1 && c1 <= p_, "column range out of bounds");
  SPB_REQUIRE(which >= 0 && which <= 2, "which must be 0 (X0 / Lp), 1 (Omega) or 2 (refined X1)");
  omega_x0_resolve();
  const size_t ldq = padded_ld(p_ + d_ + 1);
  if (which == 0 || which == 2) {
    const bool refined = omega_refine_enabled();
    if (which == 2 && !refined) { if (bytes) *bytes = 0; return nullptr; }
    if (which == 0 && refined && omega_x0_is_spd_) { if (bytes) *bytes = 0; return nullptr; }   // replicated
    const int c1e = (refined && c1 == p_) ? p_ + d_ + 1 : c1;
    if (bytes) *bytes = (long long)((size_t)(c1e - c0) * ldq * sizeof(double));
    return (which == 0 ? omega_tmp_->Bm.get() : omega_tmp_->C2.get()) + (size_t)c0 * ldq;
  }
  if (bytes) *bytes = (long long)((size_t)(c1 - c0) * ld_ * sizeof(double));
  return omega_u_.get() + (size_t)c0 * ld_;
}

void Engine::omega_stream_sync() {
  omega_x0_resolve();
  SPB_CUDA(cudaStreamSynchronize(ctx_.omega().stream));
}

void Engine::omega_finish() {
  SPB_REQUIRE(omega_cols_open_, "omega_begin() must be called first");
  omega_x0_resolve();
  OmegaRes& R = ctx_.omega();
  cudaStream_t s = R.stream;
  clock_.begin(PH_OMEGA, s);
  SPB_CUDA(cudaMemcpy2DAsync(omega_diag_raw_.get(), sizeof(double), omega_u_.get(), (ld_ + 1) * sizeof(double),
                             sizeof(double), p_, cudaMemcpyDeviceToDevice, s));
  launch_symmatu_inplace(omega_u_.get(), p_, ld_, 1e-8, s);                 // symmatu + the ridge of :574
  enqueue_lmax(s, R.blas);
  clock_.end(s);
  SPB_CUDA(cudaStreamSynchronize(s));
  omega_tmp_.reset();
  omega_cols_open_ = false;
  omega_ready_ = true;
}

void Engine::get_omega(double* out) {
  SPB_REQUIRE(omega_ready_ && any_tau_, "no thin-plate-spline matrix has been built");
  finish_omega();
  SPB_CUDA(cudaStreamSynchronize(ctx_.stream));
  SPB_CUDA(cudaMemcpy2D(out, (size_t)p_ * sizeof(double), omega_u_.get(), ld_ * sizeof(double),
                        (size_t)p_ * sizeof(double), p_, cudaMemcpyDeviceToHost));
  std::vector<double> dg(p_);
  SPB_CUDA(cudaMemcpy(dg.data(), omega_diag_raw_.get(), sizeof(double) * p_, cudaMemcpyDeviceToHost));
  for (int i = 0; i < p_; ++i) out[(size_t)i * p_ + i] = dg[i];
}

void Engine::fold_prepare(int k) {
  NvtxRange nvtx_range("spb::fold_prepare (gather, SVD start)");
  FoldData& f = folds_[k];
  if (f.prepared) return;
  trace(ctx_, "fold_prepare: enter");
  Scoped sc(clock_, PH_PROLOGUE, ctx_.stream);
  fold_prepare_on(ctx_, k);
  trace(ctx_, "fold_prepare: done");
}

// The start of one fold (:318-326) on the stream / cuSOLVER handle of `c`: safe to run for DIFFERENT folds from
// different host threads (each with its own context): it touches only folds_[k], read-only engine state and the
// (mutex-protected) caching allocator.
void Engine::fold_prepare_on(DeviceCtx& c, int k) {
  FoldData& f = folds_[k];
  cudaStream_t s = c.stream;
  DevBuf<int> rows((size_t)n_);
  DevBuf<double> scratch(512 + 8);
  f.Ytr.alloc((size_t)f.n_tr * p_);
  SPB_CUDA(cudaMemcpyAsync(rows.get(), f.rows_tr.data(), sizeof(int) * f.n_tr, cudaMemcpyHostToDevice, s));
  launch_gather_rows(dY_.get(), n_, p_, rows.get(), f.n_tr, f.Ytr.get(), s);
  f.Ytt.alloc((size_t)f.n_tr * p_);
  launch_transpose(f.Ytr.get(), f.n_tr, p_, f.Ytt.get(), s);                // p x n_tr
  SPB_CUDA(cudaStreamSynchronize(s));
  if (f.n_va > 0) {
    f.Yva.alloc((size_t)f.n_va * p_);
    SPB_CUDA(cudaMemcpyAsync(rows.get(), f.rows_va.data(), sizeof(int) * f.n_va, cudaMemcpyHostToDevice, s));
    launch_gather_rows(dY_.get(), n_, p_, rows.get(), f.n_va, f.Yva.get(), s);
  }
  f.Phi0.alloc((size_t)p_ * K_);
  f.L20.alloc((size_t)p_ * K_);
  f.kcap = std::min(std::min(f.n_tr, p_), SPB_MAX_K);
  f.V.alloc((size_t)p_ * f.kcap);
  svd_right(c, f.Ytr.get(), f.n_tr, p_, f.kcap, f.sv, f.V.get());             // :321
  f.rho = 10.0 * (f.sv[0] * f.sv[0]);                                         // :322
  // trace(Ytr'Ytr) / n_tr (:489, :491)
  double* out = scratch.get() + 512;
  launch_sumsq(f.Ytr.get(), (size_t)f.n_tr * p_, scratch.get(), out, s);
  double ss = 0.0;
  SPB_CUDA(cudaMemcpyAsync(&ss, out, sizeof(double), cudaMemcpyDeviceToHost, s));
  SPB_CUDA(cudaStreamSynchronize(s));
  f.tv = ss / (double)f.n_tr;
  f.prepared = true;
}

// The starts of several folds at once.  One start is ~270 dependent small launches (transpose, Householder QR of the
// p x n_tr block, one-sided Jacobi SVD of its triangular factor with a host read per sweep, back-transformation) and
// is latency-bound: ~9 ms alone, ~23 ms while the Omega build fills the SMs with GEMM waves -- 5 folds one after the
// other took as long as the whole Omega build (141 of 152 ms at p = 10 000) and slowed it by their SM share.  cuSOLVER's
// gesvdj synchronises with the host inside the call, so the overlap needs host threads: one per fold, each with a
// pooled worker context.  Same kernels on the same inputs as the serial path: bitwise identical starts.
// SPB_PROLOGUE_THREADS=0 keeps the serial path.
void Engine::prepare_folds(const int* folds, int nfolds) {
  std::vector<int> todo;
  for (int i = 0; i < nfolds; ++i)
    if (!folds_[folds[i]].prepared && std::find(todo.begin(), todo.end(), folds[i]) == todo.end()) todo.push_back(folds[i]);
  static const bool threaded = [] { const char* e = std::getenv("SPB_PROLOGUE_THREADS"); return !(e && e[0] == '0'); }();
  if (todo.size() < 2 || !threaded) {
    for (int k : todo) fold_prepare(k);
    return;
  }
  NvtxRange nvtx_range("spb::prepare_folds (SVD starts of all folds, one host thread each)");
  trace(ctx_, "prepare_folds: enter");
  const int nt = (int)todo.size();
  std::vector<DeviceCtx*> res;
  for (int i = 0; i < nt; ++i) res.push_back(&ctx_.worker(i));                // created on this thread
  SPB_CUDA(cudaStreamSynchronize(ctx_.stream));                               // Y was uploaded on the main stream
  std::vector<std::exception_ptr> errs((size_t)nt);
  std::vector<long long> launches((size_t)nt, 0);
  const int dev = ctx_.device;
  const auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> workers;
  for (int i = 0; i < nt; ++i) {
    workers.emplace_back([this, i, dev, &todo, &res, &errs, &launches] {
      try {
        SPB_CUDA(cudaSetDevice(dev));
        fold_prepare_on(*res[i], todo[i]);
      } catch (...) {
        errs[i] = std::current_exception();
      }
      launches[i] = g_launch_count;                                           // thread-local counter of this worker
    });
  }
  for (auto& w : workers) w.join();
  for (int i = 0; i < nt; ++i) g_launch_count += launches[i];
  prologue_wall_ms_ += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  for (int i = 0; i < nt; ++i)
    if (errs[i]) std::rethrow_exception(errs[i]);
  trace(ctx_, "prepare_folds: done");
}

void Engine::ensure_gram(Lane& L, int k) {
  if (L.gram_owner == k) return;
  Scoped sc(L.clock, PH_PROLOGUE, L.res->stream);
  L.gram.ensure(ld_ * (size_t)p_);
  Exec ex = L.exec(ctx_.flag.get());
  if (k < 0) gram_upper(ex, dY_.get(), n_, p_, L.gram.get(), ld_);            // :561
  else gram_upper(ex, folds_[k].Ytr.get(), folds_[k].n_tr, p_, L.gram.get(), ld_);   // :328
  L.gram_owner = k;
}

void Engine::run_core2(Lane& L, int kind, int fold, int idx, double tau1, double rho, ChainState& st,
                       double* buf, bool have, int expected_iters) {
  // spatpcaCore2 / spatpcaCore2p (:150-222).  The reference rebuilds the inverse on every call (:165);
  // here a call may reuse the identical matrix built earlier for the same (fold, tau1) -- or, when the system is
  // well conditioned and serves few iterations, never build it: the matrix-free kernel solves with it by CG.
  cudaStream_t s = L.res->stream;
  if (!have && core2_matrix_free(fold, tau1, rho, expected_iters)) {
    wait_omega(L);
    if (L.next_slot >= slot_cap_) flush_lane(L);
    const int n_rows = fold < 0 ? n_ : folds_[fold].n_tr;
    if (bcg_mode_ != 2 && batched_core2_supported(p_, n_rows, K_, 1)) {
      // one problem on the TMA-fed DMMA stream (126 us per pass at p = 10 000 against 200 us of the FMA stream)
      BatchedCore2Call b;
      b.omega_u = omega_u_.get(); b.ld = ld_; b.p = p_; b.n = n_rows; b.K = K_; b.F = 1;
      b.Y = fold < 0 ? dY_.get() : folds_[fold].Ytr.get();
      b.Yt = fold < 0 ? dYt_.get() : folds_[fold].Ytt.get();
      b.nk = dNk_.get();
      for (int f = 0; f < 8; ++f) { b.fold_id[f] = 0; b.rho[f] = 1.0; b.Phi[f] = b.C[f] = b.L2[f] = nullptr; }
      b.rho[0] = rho; b.Phi[0] = st.Phi.get(); b.C[0] = st.C.get(); b.L2[0] = st.L2.get();
      b.nsteps = 1; b.tau1 = &tau1; b.maxit = maxit_; b.tol = tol_; b.cg_tol = cg_tol_; b.cg_cap = 2000;
      b.snap = nullptr;
      int* slot = L.stat_slots.get() + (size_t)L.next_slot * 4;
      b.iters_out = slot; b.status_out = slot + 1;            // {iterations, status, matrix passes, -}
      {
        Scoped sc(L.clock, PH_CG, s);
        SPB_CUDA(cudaMemsetAsync(slot, 0, 4 * sizeof(int), s));
        launch_batched_core2(b, L.admm_ws.get(), s);
      }
      L.pending.push_back({kind + 200, fold, idx, L.next_slot, n_rows});   // + 200: passes in slot [2]
      L.next_slot++;
      L.dirty = true;
      return;
    }
    CgCall c;
    c.omega_u = omega_u_.get(); c.ld = ld_; c.p = p_; c.K = K_;
    if (fold < 0) { c.Ytr = dY_.get(); c.Yt = dYt_.get(); c.n_tr = n_; }
    else { c.Ytr = folds_[fold].Ytr.get(); c.Yt = folds_[fold].Ytt.get(); c.n_tr = folds_[fold].n_tr; }
    c.c_omega = 2.0 * tau1; c.c_gram = 2.0; c.rho = rho; c.maxit = maxit_; c.tol = tol_;
    c.cg_tol = cg_tol_; c.cg_cap = 2000;
    c.Phi = st.Phi.get(); c.C = st.C.get(); c.L2 = st.L2.get();
    c.stat_slot = L.stat_slots.get() + (size_t)L.next_slot * 4;
    c.err_slot = L.err_slots.get() + L.next_slot;
    {
      Scoped sc(L.clock, PH_CG, s);
      SPB_CUDA(cudaMemsetAsync(c.stat_slot, 0, 4 * sizeof(int), s));
      launch_core2_cg(c, L.admm_ws.get(), s);
    }
    L.pending.push_back({kind + 100, fold, idx, L.next_slot, c.n_tr});     // + 100: solved matrix-free (flush_lane strips it)
    L.next_slot++;
    L.dirty = true;
    return;
  }
  if (buf == nullptr) {
    L.work.ensure(ld_ * (size_t)p_);
    buf = L.work.get();
    have = false;
  }
  if (!have) {
    ensure_gram(L, fold);
    wait_omega(L);
    Scoped sc(L.clock, PH_INVERSE, s);
    spd_inverse(L.exec(ctx_.flag.get()), omega_u_.get(), ld_, L.gram.get(), ld_, p_, tau1, rho, 2.0, buf, ld_, blocks_);
    stats_.n_inverses++;
    stats_.p3_flops += factor_flops(p_, blocks_);
  }
  if (L.next_slot >= slot_cap_) flush_lane(L);
  AdmmCall c;
  c.Ainv = buf; c.blocks = blocks_; c.ld = ld_; c.p = p_; c.K = K_; c.mode = 2; c.rho = rho; c.tau2 = 0.0;
  c.maxit = maxit_; c.tol = tol_;
  c.Phi = st.Phi.get(); c.R = nullptr; c.C = st.C.get(); c.L1 = nullptr; c.L2 = st.L2.get();
  c.stat_slot = L.stat_slots.get() + (size_t)L.next_slot * 4;
  c.err_slot = L.err_slots.get() + L.next_slot;
  {
    Scoped sc(L.clock, PH_ADMM, s);
    SPB_CUDA(cudaMemsetAsync(c.stat_slot, 0, 4 * sizeof(int), s));
    if (maxit_ > 0) launch_admm(c, L.admm_ws.get(), s);
  }
  L.pending.push_back({kind, fold, idx, L.next_slot, 0});
  L.next_slot++;
  L.dirty = true;
}

void Engine::ensure_batch_buffers(int F, int nsteps) {
  bcg_ws_.ensure(batched_core2_workspace_bytes(p_, n_, K_, F));
  bcg_snap_.ensure((size_t)nsteps * F * p_ * K_);
  const size_t cap_it = (size_t)(tau1_.size() + 8) * (M_ + 1) * 4, cap_st = 64;
  if (bcg_iters_.count < cap_it) { bcg_iters_.alloc(cap_it); bcg_it_used_ = 0; }
  if (bcg_status_.count < cap_st) { bcg_status_.alloc(cap_st); bcg_st_used_ = 0; }
}

bool Engine::batch_eligible(const std::vector<int>& folds, int ns, double tau1_max, int expected_iters, int mode) {
  const int F = (int)folds.size();
  if (bcg_mode_ == 2 || cg_mode_ == 2 || F < 1 || F > 8 || ns < 1 || ns > 64 || maxit_ <= 0 || K_ > 8 || F * K_ > 32)
    return false;
  if (mode == 3 && core3_mode_ == 2) return false;
  if (cg_mode_ != 1 && p_ < 2048) return false;
  if (!batched_core2_supported(p_, n_, K_, F)) return false;
  // distinct lanes (problems sharing a lane share its chain buffers) and the matrix-free policy for every problem
  std::vector<int> lanes_seen;
  for (int k : folds) {
    const int li = k < 0 ? n_fold_lanes_ : k % n_fold_lanes_;
    if (std::find(lanes_seen.begin(), lanes_seen.end(), li) != lanes_seen.end()) return false;
    lanes_seen.push_back(li);
    const double rho = k < 0 ? rho_full_ : folds_[k].rho;
    const double scale = F > 1 ? 2.1 / F : 1.0;
    const bool ok = mode == 3 ? core3_matrix_free(k, tau1_max, rho, expected_iters, scale)
                              : core2_matrix_free(k, tau1_max, rho, expected_iters, scale);
    if (!ok) return false;
  }
  return true;
}

bool Engine::run_core2_batched(const std::vector<int>& folds, const std::vector<double>& taus, const std::vector<int>& idx,
                               int kind, int expected_iters, double** snap_out, int mode, double fixed_tau1,
                               const std::vector<int>* kinds, const std::vector<int>* idx_override) {
  const int F = (int)folds.size(), ns = (int)taus.size();
  if (ns < 1) return false;
  const double tmax = mode == 3 ? fixed_tau1 : *std::max_element(taus.begin(), taus.end());
  if (!batch_eligible(folds, ns, tmax, expected_iters, mode)) return false;
  // one workspace and one snapshot buffer: a batch that would overlap with an earlier one on another stream, or
  // overwrite snapshots that loss kernels may still read, waits for everything enqueued so far (rare: only when
  // folds share lanes, i.e. when device memory limits the lane count)
  if (!batch_pending_.empty() &&
      (bcg_it_used_ + (size_t)ns * F > bcg_iters_.count || bcg_st_used_ + 2 > bcg_status_.count ||
       batch_stream_ != lane_of(folds[0]).res->stream || (snap_out != nullptr && snap_in_use_)))
    flush();
  ensure_batch_buffers(F, ns);
  batch_stream_ = lane_of(folds[0]).res->stream;
  if (snap_out != nullptr) snap_in_use_ = true;
  Lane& L0 = lane_of(folds[0]);
  cudaStream_t s = L0.res->stream;
  wait_omega(L0);
  // the other lanes' pending work (their starts) must be visible to the batch stream
  for (int k : folds) {
    Lane& L = lane_of(k);
    if (&L == &L0) continue;
    if (!L.evt) SPB_CUDA(cudaEventCreateWithFlags(&L.evt, cudaEventDisableTiming));
    SPB_CUDA(cudaEventRecord(L.evt, L.res->stream));
    SPB_CUDA(cudaStreamWaitEvent(s, L.evt, 0));
  }
  BatchedCore2Call c;
  c.omega_u = omega_u_.get(); c.ld = ld_; c.p = p_; c.n = n_; c.K = K_; c.F = F;
  c.Y = dY_.get(); c.Yt = dYt_.get(); c.nk = dNk_.get();
  for (int f = 0; f < 8; ++f) {
    c.fold_id[f] = 0; c.rho[f] = 1.0; c.Phi[f] = c.C[f] = c.L2[f] = nullptr;
  }
  for (int f = 0; f < F; ++f) {
    const int k = folds[f];
    ChainState& st = k < 0 ? full_ : lane_of(k).chain;
    c.fold_id[f] = k < 0 ? 0 : k + 1;                       // nk holds 1-based fold ids (:318-319)
    c.rho[f] = k < 0 ? rho_full_ : folds_[k].rho;
    c.Phi[f] = st.Phi.get(); c.C[f] = st.C.get(); c.L2[f] = st.L2.get();
    if (mode == 3) { c.R[f] = st.R.get(); c.L1[f] = st.L1.get(); }
  }
  const std::vector<double> t1rep(ns, fixed_tau1);
  c.mode = mode;
  c.nsteps = ns; c.tau1 = mode == 3 ? t1rep.data() : taus.data(); c.tau2 = mode == 3 ? taus.data() : nullptr;
  c.maxit = maxit_; c.tol = tol_; c.cg_tol = cg_tol_; c.cg_cap = 2000;
  c.snap = snap_out ? bcg_snap_.get() : nullptr;
  c.iters_out = bcg_iters_.get() + bcg_it_used_;
  c.status_out = bcg_status_.get() + bcg_st_used_;
  {
    Scoped sc(L0.clock, PH_CG, s);
    SPB_CUDA(cudaMemsetAsync(c.iters_out, 0, sizeof(int) * (size_t)ns * F, s));
    SPB_CUDA(cudaMemsetAsync(c.status_out, 0, 2 * sizeof(int), s));
    launch_batched_core2(c, bcg_ws_.get(), s);
  }
  BatchRec rec;
  rec.kind = kind; rec.folds = folds; rec.idx = idx; rec.it_off = bcg_it_used_; rec.st_off = bcg_st_used_; rec.n = n_;
  if (kinds) rec.kinds = *kinds;
  if (idx_override) rec.idx_override = *idx_override;
  batch_pending_.push_back(rec);
  bcg_it_used_ += (size_t)ns * F;
  bcg_st_used_ += 2;
  // the other lanes continue after the batch
  if (!L0.evt) SPB_CUDA(cudaEventCreateWithFlags(&L0.evt, cudaEventDisableTiming));
  SPB_CUDA(cudaEventRecord(L0.evt, s));
  for (int k : folds) {
    Lane& L = lane_of(k);
    L.dirty = true;
    if (&L == &L0) continue;
    SPB_CUDA(cudaStreamWaitEvent(L.res->stream, L0.evt, 0));
  }
  if (snap_out) *snap_out = bcg_snap_.get();
  return true;
}

double* Engine::cached_inverse(FoldData& f, int index1) {
  if (!cache_inverses_ || index1 < 1 || index1 >= (int)f.inv_cache.size()) return nullptr;
  if (index1 >= (int)f.inv_valid.size() || !f.inv_valid[index1]) return nullptr;
  return f.inv_cache[index1].get();
}

void Engine::drop_cached_inverses(FoldData& f, int keep) {
  // Kept for the lifetime of the engine: the memory was budgeted at construction, and the K-selection
  // loop (reset_k) reuses every one of them.
  (void)f; (void)keep;
}

void* Engine::inverse_buffer(int k, int i, long long* bytes) {
  SPB_REQUIRE(cache_inverses_, "the stage-1 inverse cache is disabled (not enough device memory)");
  SPB_REQUIRE(k >= 0 && k < M_, "fold index out of range");
  const int n1 = (int)tau1_.size();
  SPB_REQUIRE(i >= 1 && i < n1, "tau1 index out of range");
  FoldData& f = folds_[k];
  if ((int)f.inv_cache.size() != n1) {
    f.inv_cache.resize(n1);
    f.inv_valid.assign(n1, 0);
  }
  f.inv_cache[i].ensure(ld_ * (size_t)p_);
  if (bytes) *bytes = (long long)(ld_ * (size_t)p_ * sizeof(double));
  return f.inv_cache[i].get();
}

void Engine::build_inverse(int k, int i) {
  double* buf = static_cast<double*>(inverse_buffer(k, i, nullptr));
  FoldData& f = folds_[k];
  SPB_REQUIRE(f.prepared, "prepare_fold(k) must run before build_inverse(k, i)");
  if (f.inv_valid[i]) return;
  Lane& L = lane_of(k);
  ensure_gram(L, k);
  wait_omega(L);
  {
    Scoped sc(L.clock, PH_INVERSE, L.res->stream);
    spd_inverse(L.exec(ctx_.flag.get()), omega_u_.get(), ld_, L.gram.get(), ld_, p_, tau1_[i], f.rho, 2.0, buf, ld_, blocks_);
    stats_.n_inverses++;
    stats_.p3_flops += factor_flops(p_, blocks_);
  }
  f.inv_valid[i] = 1;
  L.dirty = true;
}

void Engine::mark_inverse(int k, int i) {
  inverse_buffer(k, i, nullptr);
  folds_[k].inv_valid[i] = 1;
}

void Engine::run_core3(Lane& L, int kind, int fold, int idx, const double* tinv, double tau2, double rho,
                       ChainState& st) {
  cudaStream_t s = L.res->stream;
  if (L.next_slot >= slot_cap_) flush_lane(L);
  AdmmCall c;
  c.Ainv = tinv; c.blocks = blocks_path_; c.ld = ld_; c.p = p_; c.K = K_; c.mode = 3; c.rho = rho; c.tau2 = tau2;
  c.maxit = maxit_; c.tol = tol_;
  c.Phi = st.Phi.get(); c.R = st.R.get(); c.C = st.C.get(); c.L1 = st.L1.get(); c.L2 = st.L2.get();
  c.stat_slot = L.stat_slots.get() + (size_t)L.next_slot * 4;
  c.err_slot = L.err_slots.get() + L.next_slot;
  {
    Scoped sc(L.clock, PH_ADMM, s);
    SPB_CUDA(cudaMemsetAsync(c.stat_slot, 0, 4 * sizeof(int), s));
    if (maxit_ > 0) launch_admm(c, L.admm_ws.get(), s);
  }
  L.pending.push_back({kind, fold, idx, L.next_slot, 0});
  L.next_slot++;
  L.dirty = true;
}

void Engine::launch_loss(Lane& L, const FoldData& f, const double* Phi, double* out_dev) {
  cudaStream_t s = L.res->stream;
  Scoped sc(L.clock, PH_CVLOSS, s);
  if (f.n_va == 0) {
    SPB_CUDA(cudaMemsetAsync(out_dev, 0, sizeof(double), s));
    return;
  }
  cv_loss(s, f.Yva.get(), f.n_va, p_, Phi, K_, L.lossW.get(), L.loss_scratch.get(), out_dev);
  L.dirty = true;
}

void Engine::flush_lane(Lane& L) {
  SPB_CUDA(cudaStreamSynchronize(L.res->stream));
  L.dirty = false;
  if (L.next_slot > 0) {
    std::vector<int> st((size_t)L.next_slot * 4);
    SPB_CUDA(cudaMemcpy(st.data(), L.stat_slots.get(), sizeof(int) * st.size(), cudaMemcpyDeviceToHost));
    bool polar_failed = false, cg_failed = false, not_posdef = false;
    for (auto& r : L.pending) {
      int iters = st[(size_t)r.slot * 4 + 0];
      const int status = st[(size_t)r.slot * 4 + 1];
      if (status == SPB_ERR_NOT_POSDEF) not_posdef = true;
      else if (status == SPB_ERR_INTERNAL) cg_failed = true;
      else if (status != 0) polar_failed = true;
      int kind = r.kind;
      const double pp8 = 8.0 * (double)p_ * (double)p_, pk8 = 8.0 * (double)p_ * (double)K_;
      if (kind >= 100) {                                     // matrix-free call: matrix passes in [3] ([2] from the DMMA kernel)
        const bool dmma = kind >= 200;
        kind -= dmma ? 200 : 100;
        const int passes = st[(size_t)r.slot * 4 + (dmma ? 2 : 3)];
        stats_.n_cg_calls++;
        stats_.cg_passes += passes;
        stats_.admm_bytes += passes * (pp8 + 16.0 * (double)r.n_tr * (double)p_) + iters * 5.0 * pk8;
      } else {
        stats_.admm_bytes += iters * (pp8 + (kind == 3 ? 9.0 : 5.0) * pk8);   // 8p^2 + 72pK / 40pK
      }
      log_.push_back(kind); log_.push_back(r.fold); log_.push_back(r.idx); log_.push_back(iters);
    }
    L.pending.clear();
    L.next_slot = 0;
    if (not_posdef) throw Error(SPB_ERR_NOT_POSDEF, "inv_sympd(): matrix is singular or not positive definite");
    if (cg_failed) throw Error(SPB_ERR_INTERNAL, "matrix-free Core2: conjugate gradients did not converge");
    if (polar_failed)
      throw Error(SPB_ERR_POLAR, "svd_econ(): the p x K iterate is rank deficient, Stiefel projection undefined");
  }
}

void Engine::flush() {
  finish_omega();
  for (auto& L : lanes_) flush_lane(*L);
  if (!batch_pending_.empty()) {
    std::vector<int> its(bcg_it_used_), sts(bcg_st_used_);
    SPB_CUDA(cudaMemcpy(its.data(), bcg_iters_.get(), sizeof(int) * its.size(), cudaMemcpyDeviceToHost));
    SPB_CUDA(cudaMemcpy(sts.data(), bcg_status_.get(), sizeof(int) * sts.size(), cudaMemcpyDeviceToHost));
    int bad = 0;
    const double pp8 = 8.0 * (double)p_ * (double)p_, pk8 = 8.0 * (double)p_ * (double)K_;
    for (const BatchRec& b : batch_pending_) {
      const int F = (int)b.folds.size(), ns = (int)b.idx.size();
      const int status = sts[b.st_off], passes = sts[b.st_off + 1];
      if (status != 0 && bad == 0) bad = status;
      long long iters_sum = 0;
      for (int i = 0; i < ns; ++i)
        for (int f = 0; f < F; ++f) {
          const int it = its[b.it_off + (size_t)i * F + f];
          const int kd = b.kinds.empty() ? b.kind : b.kinds[f];
          const int ix = (!b.idx_override.empty() && b.idx_override[f] >= 0) ? b.idx_override[f] : b.idx[i];
          log_.push_back(kd); log_.push_back(b.folds[f]); log_.push_back(ix); log_.push_back(it);
          iters_sum += it;
        }
      if (trace_on())
        fprintf(stderr, "[spb batch] kind %d: %d problems x %d steps, %lld ADMM iterations, %d matrix passes\n", b.kind, F, ns,
                iters_sum, passes);
      stats_.n_cg_calls += F * ns;
      stats_.cg_passes += passes;
      stats_.admm_bytes += passes * (pp8 + 16.0 * (double)b.n * (double)p_) + (double)iters_sum * (b.kind == 3 ? 9.0 : 5.0) * pk8;
    }
    batch_pending_.clear();
    bcg_it_used_ = bcg_st_used_ = 0;
    snap_in_use_ = false;
    if (bad == SPB_ERR_NOT_POSDEF) throw Error(SPB_ERR_NOT_POSDEF, "inv_sympd(): matrix is singular or not positive definite");
    if (bad == SPB_ERR_INTERNAL) throw Error(SPB_ERR_INTERNAL, "matrix-free Core2: conjugate gradients did not converge");
    if (bad != 0)
      throw Error(SPB_ERR_POLAR, "svd_econ(): the p x K iterate is rank deficient, Stiefel projection undefined");
  }
  ctx_.sync_and_check();
}

// ---- stage 1 (:592-650) -----------------------------------------------------------------
// Folds of a stage grouped into rounds of folds that live on DISTINCT lanes (folds that share a lane share
// its chain buffers and run one after the other).  Within a round the host feeds the lanes round-robin, one
// (fold, grid point) unit at a time: a unit is ~10^3 launches at p = 10 000, more than a stream's launch
// queue holds, so enqueueing a whole tau path of one fold before the next would block the host on the
// first lane and serialise the lanes (measured: 5 lanes = 1 lane at p = 10 000, profiles/r01_notes.md).
std::vector<std::vector<int>> Engine::lane_rounds(const int* folds, int nfolds) const {
  std::vector<std::vector<int>> per_lane(n_fold_lanes_);
  for (int f = 0; f < nfolds; ++f) per_lane[folds[f] % n_fold_lanes_].push_back(folds[f]);
  std::vector<std::vector<int>> rounds;
  for (size_t r = 0;; ++r) {
    std::vector<int> cur;
    for (auto& q : per_lane)
      if (r < q.size()) cur.push_back(q[r]);
    if (cur.empty()) break;
    rounds.push_back(cur);
  }
  return rounds;
}

// step 0: the fold's start (and the whole single-tau1 branch); step i >= 1: the i-th unit of the tau1 path
void Engine::enqueue_stage1(int k, int step) {
  FoldData& f = folds_[k];
  Lane& L = lane_of(k);
  cudaStream_t s = L.res->stream;
  ChainState& ch = L.chain;
  const int n1 = (int)tau1_.size();
  const double max_tau2 = *std::max_element(tau2_.begin(), tau2_.end());
  double* scores = fold_scores(k);
  if (step > 0) {                                                             // warm-started path :329-332
    const int i = step;
    double* buf = nullptr;
    bool have = false;
    const bool kept = cache_inverses_ && f.inv_valid[i] != 0;     // built by an earlier run on this engine (other K)
    if (cache_inverses_ && (kept || !core2_matrix_free(k, tau1_[i], f.rho, 3))) {
      f.inv_cache[i].ensure(ld_ * (size_t)p_);
      buf = f.inv_cache[i].get();
      have = kept;
      f.inv_valid[i] = 1;
    }
    run_core2(L, 2, k, i, tau1_[i], f.rho, ch, buf, have, 3);     // warm start: 1-5 iterations
    launch_loss(L, f, ch.Phi.get(), scores + i);
    return;
  }
  copy_pk(s, f.Phi0.get(), f.V.get());                                        // Phi_old.cols(0, K-1) (:324)
  if (n1 > 1) {                                                               // worker :312-334
    init_lambda(s, f.Phi0.get(), f.sv, f.rho, f.L20.get());
    copy_pk(s, ch.Phi.get(), f.Phi0.get());
    copy_pk(s, ch.C.get(), f.Phi0.get());
    copy_pk(s, ch.L2.get(), f.L20.get());
    launch_loss(L, f, f.Phi0.get(), scores + 0);                              // :327 (tau1[0] scored as PCA)
    if (cache_inverses_ && (int)f.inv_cache.size() != n1) {
      f.inv_cache.resize(n1);
      f.inv_valid.assign(n1, 0);
    }
    L.dirty = true;
    return;
  }
  const double sel = tau1_[0];
  if (sel != 0.0 && max_tau2 == 0.0) {                                        // :604-623
    init_lambda(s, f.Phi0.get(), f.sv, f.rho, f.L20.get());
    copy_pk(s, ch.Phi.get(), f.Phi0.get());
    copy_pk(s, ch.C.get(), f.Phi0.get());
    copy_pk(s, ch.L2.get(), f.L20.get());
    run_core2(L, 2, k, 0, sel, f.rho, ch);
    copy_pk(s, f.Phi0.get(), ch.Phi.get());                                   // Phi_cv.slice(k) (:619)
    copy_pk(s, f.L20.get(), ch.L2.get());                                     // Lambd2_cv.slice(k) (:620)
    // (:621 also forms tempinv_cv here; nothing reads it unless stage 2 runs, which recomputes it)
  } else if (sel == 0.0 && max_tau2 == 0.0) {                                 // :624-629, :576-590
    // Phi_cv.slice(k) = fold PCA start (already in Phi0); Lambd2_cv is never read on this path.
    SPB_CUDA(cudaMemsetAsync(f.L20.get(), 0, sizeof(double) * (size_t)p_ * K_, s));
  } else {                                                                    // :630-648
    init_lambda(s, f.Phi0.get(), sv_full_, f.rho, f.L20.get());               // quirk :642: FULL-data singular values
    if (sel != 0.0) {
      copy_pk(s, ch.Phi.get(), f.Phi0.get());
      copy_pk(s, ch.C.get(), f.Phi0.get());
      copy_pk(s, ch.L2.get(), f.L20.get());
      run_core2(L, 2, k, 0, sel, f.rho, ch);
      copy_pk(s, f.Phi0.get(), ch.Phi.get());
      copy_pk(s, f.L20.get(), ch.L2.get());
    }
  }
  L.dirty = true;
}

void Engine::stage1_folds(const int* folds, int nfolds, double* rows) {
  NvtxRange nvtx_range("spb::stage1_folds (tau1 paths)");
  SPB_REQUIRE(full_ready_, "prepare() must be called first");
  SPB_REQUIRE(folds != nullptr && nfolds >= 0, "invalid fold list");
  for (int f = 0; f < nfolds; ++f) SPB_REQUIRE(folds[f] >= 0 && folds[f] < M_, "fold index out of range");
  const int n1 = (int)tau1_.size();
  my_stage1_folds_ = nfolds;
  prepare_folds(folds, nfolds);                                               // SVD starts, one host thread per fold
  for (const auto& round : lane_rounds(folds, nfolds)) {                      // tau1 paths, concurrent lanes
    for (int k : round) enqueue_stage1(k, 0);
    bool batched = false;
    if (n1 > 1) {
      // all folds of the round in lock-step through ONE launch of the batched matrix-free kernel (warm-started
      // tau1 path, :329-332); the hold-out losses (:331) are taken from its per-step snapshots afterwards
      bool fresh = true;
      for (int k : round)
        for (int i = 1; i < n1; ++i)
          if (cache_inverses_ && i < (int)folds_[k].inv_valid.size() && folds_[k].inv_valid[i]) fresh = false;
      std::vector<double> taus(tau1_.begin() + 1, tau1_.end());
      std::vector<int> idx(n1 - 1);
      for (int i = 1; i < n1; ++i) idx[i - 1] = i;
      double* snap = nullptr;
      if (fresh && run_core2_batched(round, taus, idx, 2, 3, &snap)) {
        batched = true;
        const size_t pk = (size_t)p_ * K_;
        const int F = (int)round.size();
        for (int i = 1; i < n1; ++i)
          for (int f = 0; f < F; ++f) {
            const int k = round[f];
            launch_loss(lane_of(k), folds_[k], snap + ((size_t)(i - 1) * F + f) * pk, fold_scores(k) + i);
          }
      }
    }
    if (!batched)
      for (int i = 1; i < n1; ++i)
        for (int k : round) enqueue_stage1(k, i);
  }
  flush();
  if (n1 > 1) {
    SPB_REQUIRE(rows != nullptr, "cv rows pointer is NULL");
    for (int f = 0; f < nfolds; ++f)
      SPB_CUDA(cudaMemcpy(rows + (size_t)f * n1, fold_scores(folds[f]), sizeof(double) * n1, cudaMemcpyDeviceToHost));
    stats_.bytes_d2h += (long long)sizeof(double) * n1 * nfolds;
  }
}

void Engine::stage1_full(int index1) {
  NvtxRange nvtx_range("spb::stage1_full");
  SPB_REQUIRE(full_ready_, "prepare() must be called first");
  const int n1 = (int)tau1_.size();
  const double max_tau2 = *std::max_element(tau2_.begin(), tau2_.end());
  Lane& L = lane_of(-1);
  // enqueued on the full-data lane and NOT waited for: it overlaps with the folds' next stage
  if (n1 > 1) {
    SPB_REQUIRE(index1 >= 0 && index1 < n1, "index1 out of range");
    if (index1 > 0) {                                                         // :598-599
      double* buf = nullptr;
      bool have = false;
      const bool kept = cache_inverses_ && full_inv2_index_ == index1;
      if (!kept && my_stage1_folds_ > 0 && tau2_.size() > 1 && batch_eligible({-1}, 1, tau1_[index1], 8, 2)) {
        pend_core2p_ = index1;               // rides with the folds' cold Core2 batch of stage 2 (same tau1)
        const bool stage2_needs = tau2_.size() > 1 || tau2_[0] > 0.0;
        const int plen = tau2_.size() > 1 ? (int)tau2_.size() : (int)l2_.size();
        if (cache_inverses_ && stage2_needs && !batch_eligible({-1}, std::min(plen, 64), selected_tau1(index1), plen + 2, 3))
          full_tempinv(index1);
        return;
      }
      if (cache_inverses_ && (kept || !core2_matrix_free(-1, tau1_[index1], rho_full_, 8))) {
        full_inv2_.ensure(ld_ * (size_t)p_);
        buf = full_inv2_.get();
        have = kept;
        full_inv2_index_ = index1;
      }
      run_core2(L, 20, -1, index1, tau1_[index1], rho_full_, full_, buf, have);
    }
  } else {
    const double sel = tau1_[0];
    if (sel != 0.0 && max_tau2 == 0.0) run_core2(L, 20, -1, 0, sel, rho_full_, full_); // :605
  }
  const bool stage2_needs_tempinv = tau2_.size() > 1 || tau2_[0] > 0.0;       // the cases in which stage2_full() works
  // early, so that it overlaps with the folds' stage 2 -- unless stage2_full() will run matrix-free
  const int path_len = tau2_.size() > 1 ? (int)tau2_.size() : (int)l2_.size();
  if (cache_inverses_ && stage2_needs_tempinv &&
      !batch_eligible({-1}, std::min(path_len, 64), selected_tau1(index1), path_len + 2, 3))
    full_tempinv(index1);
}

// ---- stage 2 (:652-680) -----------------------------------------------------------------
// step 0: start of the fold's chain; step 1: Core2 at the selected tau1 (skipped when the round ran it batched);
// step 2: Phi_cv / Lambd2_cv and the fold's tempinv; step 3: the tau2 path
void Engine::enqueue_stage2(int k, int index1, int step) {
  FoldData& f = folds_[k];
  Lane& L = lane_of(k);
  cudaStream_t s = L.res->stream;
  ChainState& ch = L.chain;
  const int n2 = (int)tau2_.size();
  const double sel1 = selected_tau1(index1);
  double* scores = fold_scores(k);
  if (step == 3) {
    ensure_fold_tempinv(k, index1);
    for (int i = 0; i < n2; ++i) {                                            // :398-401
      run_core3(L, 3, k, i, f.tinv.get(), tau2_[i], f.rho, ch);
      launch_loss(L, f, ch.Phi.get(), scores + i);
    }
    return;
  }
  if (step == 0) {
    copy_pk(s, ch.Phi.get(), f.Phi0.get());                                   // :389
    copy_pk(s, ch.C.get(), f.Phi0.get());
    zero_pk(s, ch.L1.get());                                                  // :390
    copy_pk(s, ch.L2.get(), f.L20.get());                                     // :391
    drop_cached_inverses(f, index1);
    L.dirty = true;
    return;
  }
  if (step == 1) {
    if (sel1 != 0.0) {                                                        // :392-393
      double* inv = cached_inverse(f, index1);
      run_core2(L, 2, k, 0, sel1, f.rho, ch, inv, inv != nullptr);
    }
    return;
  }
  // (the kept inverse is in use on this lane's stream: it is released after stage 3's flush)
  copy_pk(s, ch.R.get(), ch.Phi.get());                                       // :394
  copy_pk(s, f.Phi0.get(), ch.Phi.get());                                     // :395
  copy_pk(s, f.L20.get(), ch.L2.get());                                       // :396
}

void Engine::ensure_fold_tempinv(int k, int index1) {
  FoldData& f = folds_[k];
  Lane& L = lane_of(k);
  cudaStream_t s = L.res->stream;
  const double sel1 = selected_tau1(index1);
  const int tag1 = ((int)tau1_.size() > 1) ? index1 : 0;
  if (f.tinv.get() != nullptr && f.tinv_index == tag1 && f.tinv_valid) return;
  ensure_gram(L, k);
  wait_omega(L);
  Scoped sc(L.clock, PH_INVERSE, s);
  f.tinv.ensure(ld_ * (size_t)p_);
  spd_inverse(L.exec(ctx_.flag.get()), omega_u_.get(), ld_, L.gram.get(), ld_, p_, sel1, f.rho, 1.0, f.tinv.get(),
              ld_, blocks_path_);                                             // :397
  stats_.n_inverses++;
  stats_.p3_flops += factor_flops(p_, blocks_path_);
  f.tinv_index = tag1;
  f.tinv_valid = true;
}

void Engine::stage2_folds(const int* folds, int nfolds, int index1, double* rows) {
  NvtxRange nvtx_range("spb::stage2_folds (tau2 paths)");
  const int n2 = (int)tau2_.size();
  if (n2 <= 1) return;
  SPB_REQUIRE(folds != nullptr && nfolds >= 0, "invalid fold list");
  for (int f = 0; f < nfolds; ++f) {
    SPB_REQUIRE(folds[f] >= 0 && folds[f] < M_, "fold index out of range");
    SPB_REQUIRE(folds_[folds[f]].prepared, "stage 1 must run before stage 2 for every fold");
  }
  const double sel1 = selected_tau1(index1);
  for (const auto& round : lane_rounds(folds, nfolds)) {
    for (int k : round) enqueue_stage2(k, index1, 0);
    bool batched = false;
    if (sel1 != 0.0) {
      bool have_inverse = false;
      for (int k : round) have_inverse |= (cached_inverse(folds_[k], index1) != nullptr);
      if (!have_inverse && pend_core2p_ == index1 && (int)tau1_.size() > 1 && tau1_[index1] == sel1) {
        // the deferred full-data Core2p (:598-599) as one more problem of this launch
        std::vector<int> with_full(round), kinds(round.size(), 2), ixo(round.size(), -1);
        with_full.push_back(-1); kinds.push_back(20); ixo.push_back(index1);
        batched = run_core2_batched(with_full, {sel1}, {0}, 2, 8, nullptr, 2, 0.0, &kinds, &ixo);
        if (batched) pend_core2p_ = -1;
      }
      if (!batched && !have_inverse) batched = run_core2_batched(round, {sel1}, {0}, 2, 8, nullptr);   // :392-393, all folds at once
    }
    if (pend_core2p_ >= 0) run_pending_full();
    if (!batched)
      for (int k : round) enqueue_stage2(k, index1, 1);
    for (int k : round) enqueue_stage2(k, index1, 2);
    // the tau2 path (:398-401) of all folds of the round in ONE launch of the matrix-free kernel -- no tempinv (:397)
    // is ever formed; the hold-out losses (:400) are taken from the per-step snapshots.  Otherwise: factor + Core3.
    std::vector<int> idx(n2);
    for (int i = 0; i < n2; ++i) idx[i] = i;
    double* snap = nullptr;
    if (run_core2_batched(round, tau2_, idx, 3, n2 + 2, &snap, 3, sel1)) {
      const size_t pk = (size_t)p_ * K_;
      const int F = (int)round.size();
      for (int i = 0; i < n2; ++i)
        for (int f = 0; f < F; ++f) {
          const int k = round[f];
          launch_loss(lane_of(k), folds_[k], snap + ((size_t)i * F + f) * pk, fold_scores(k) + i);
        }
    } else {
      for (int k : round) enqueue_stage2(k, index1, 3);
    }
  }
  flush();
  SPB_REQUIRE(rows != nullptr, "cv rows pointer is NULL");
  for (int f = 0; f < nfolds; ++f)
    SPB_CUDA(cudaMemcpy(rows + (size_t)f * n2, fold_scores(folds[f]), sizeof(double) * n2, cudaMemcpyDeviceToHost));
  stats_.bytes_d2h += (long long)sizeof(double) * n2 * nfolds;
}

// tempinv of the full-data fit (:661, :673): depends on the selected tau1 only, so stage1_full() already
// enqueues it (when it has its own buffer) and it overlaps with the folds' stage 2 instead of trailing the job
double* Engine::full_tempinv(int index1) {
  Lane& L = lane_of(-1);
  cudaStream_t s = L.res->stream;
  const double sel1 = selected_tau1(index1);
  const int tag1 = ((int)tau1_.size() > 1) ? index1 : 0;
  double* tbuf = nullptr;
  bool have = false;
  if (cache_inverses_) {
    full_tinv_.ensure(ld_ * (size_t)p_);
    tbuf = full_tinv_.get();
    have = (full_tinv_index_ == tag1);
    full_tinv_index_ = tag1;
  } else {
    L.work.ensure(ld_ * (size_t)p_);
    tbuf = L.work.get();
  }
  if (!have) {
    ensure_gram(L, -1);
    wait_omega(L);
    Scoped sc(L.clock, PH_INVERSE, s);
    spd_inverse(L.exec(ctx_.flag.get()), omega_u_.get(), ld_, L.gram.get(), ld_, p_, sel1, rho_full_, 1.0, tbuf,
                ld_, blocks_path_);                                           // :661, :673
    stats_.n_inverses++;
    stats_.p3_flops += factor_flops(p_, blocks_path_);
  }
  return tbuf;
}

void Engine::stage2_full(int index1, int index2) {
  NvtxRange nvtx_range("spb::stage2_full");
  const int n2 = (int)tau2_.size();
  const double sel1 = selected_tau1(index1);
  const std::vector<double>* path = nullptr;
  int upto = 0;
  if (n2 > 1) {
    SPB_REQUIRE(index2 >= 0 && index2 < n2, "index2 out of range");
    path = &tau2_; upto = index2 + 1;                                         // :665-666
  } else if (tau2_[0] > 0.0) {
    path = &l2_; upto = (int)l2_.size();                                      // :672-678
  }
  if (!path) return;
  if (pend_core2p_ >= 0) run_pending_full();                                  // (not consumed by a stage-2 batch)
  // the tau2 path can ride with the folds' refit batch of stage 3 (same tau1, same tau2 values up to index2)
  if (n2 > 1 && my_stage1_folds_ > 0 && batch_eligible({-1}, upto, sel1, upto + 2, 3)) {
    const int tag1 = ((int)tau1_.size() > 1) ? index1 : 0;
    const bool have_tinv = cache_inverses_ && full_tinv_.get() != nullptr && full_tinv_index_ == tag1;
    if (!have_tinv) {
      pend_path_ = true; pend_path_i1_ = index1; pend_path_upto_ = upto;
      return;
    }
  }
  run_full_path(index1, *path, upto);
}

void Engine::run_full_path(int index1, const std::vector<double>& path, int upto) {
  const double sel1 = selected_tau1(index1);
  Lane& L = lane_of(-1);
  cudaStream_t s = L.res->stream;
  copy_pk(s, full_.R.get(), full_.Phi.get());                                 // :662
  zero_pk(s, full_.L1.get());                                                 // :663
  // the whole path in one launch of the matrix-free kernel (one problem), unless a tempinv exists already
  const int tag1 = ((int)tau1_.size() > 1) ? index1 : 0;
  const bool have_tinv = cache_inverses_ && full_tinv_.get() != nullptr && full_tinv_index_ == tag1;
  if (!have_tinv) {
    std::vector<double> steps(path.begin(), path.begin() + upto);
    std::vector<int> idx(upto);
    for (int i = 0; i < upto; ++i) idx[i] = i;
    if (run_core2_batched({-1}, steps, idx, 3, upto + 2, nullptr, 3, sel1)) return;
  }
  double* tbuf = full_tempinv(index1);
  for (int i = 0; i < upto; ++i) run_core3(L, 3, -1, i, tbuf, path[i], rho_full_, full_);
}

void Engine::run_pending_full() {
  if (pend_core2p_ >= 0) {
    const int i1 = pend_core2p_;
    pend_core2p_ = -1;
    run_core2(lane_of(-1), 20, -1, i1, tau1_[i1], rho_full_, full_, nullptr, false);
  }
  if (pend_path_) {
    pend_path_ = false;
    run_full_path(pend_path_i1_, tau2_, pend_path_upto_);
  }
}

// ---- stage 3 (:682-692, worker :459-525) ---------------------------------------------------
// chain = false: only the start of the refit (copies); the caller runs the Core2 / Core3 calls batched
void Engine::enqueue_stage3_refit(int k, int index1, int index2, bool chain) {
  FoldData& f = folds_[k];
  Lane& L = lane_of(k);
  cudaStream_t s = L.res->stream;
  ChainState& ch = L.chain;
  const int n2 = (int)tau2_.size();
  const double sel1 = selected_tau1(index1);
  const double max_tau2 = *std::max_element(tau2_.begin(), tau2_.end());
  copy_pk(s, ch.Phi.get(), f.Phi0.get());                                     // :469
  copy_pk(s, ch.C.get(), f.Phi0.get());
  copy_pk(s, ch.L2.get(), f.L20.get());                                       // :471
  if (max_tau2 != 0.0 || sel1 != 0.0) {                                       // :473
    if (n2 == 1) {
      drop_cached_inverses(f, index1);
      double* inv = cached_inverse(f, index1);
      if (chain) run_core2(L, 2, k, 0, sel1, f.rho, ch, inv, inv != nullptr); // :475
    } else {
      copy_pk(s, ch.R.get(), ch.Phi.get());                                   // :478
      zero_pk(s, ch.L1.get());                                                // :479
      if (chain) {
        ensure_fold_tempinv(k, index1);
        for (int i = 0; i <= index2; ++i)                                     // :480-482
          run_core3(L, 3, k, i, f.tinv.get(), tau2_[i], f.rho, ch);
      }
    }
  }
  L.dirty = true;
}

void Engine::stage3_folds(const int* folds, int nfolds, int index1, int index2, double* rows) {
  NvtxRange nvtx_range("spb::stage3_folds (gamma losses)");
  SPB_REQUIRE(folds != nullptr && nfolds >= 0 && rows != nullptr, "invalid fold list");
  for (int f = 0; f < nfolds; ++f) {
    SPB_REQUIRE(folds[f] >= 0 && folds[f] < M_, "fold index out of range");
    SPB_REQUIRE(folds_[folds[f]].prepared, "stage 1 must run before stage 3 for every fold");
  }
  const int ng = (int)gamma_.size();
  // folds sharing a lane share its chain buffers: refit + loss one lane-round at a time.  The rounds come
  // from lane_rounds() -- a fold's lane is k % n_fold_lanes_, NOT its position in the caller's list, so a
  // strided list such as [0, 2, 4] with 2 lanes must not put folds 0 and 2 into the same round.
  std::map<int, int> row_of;
  for (int f = 0; f < nfolds; ++f) row_of[folds[f]] = f;
  if (gamma_dev_.count < (size_t)ng) {
    // stream-ordered copy + wait: a plain cudaMemcpy from pageable memory returns once the data is STAGED, and the
    // lanes' non-blocking streams do not synchronise with the legacy stream -- the one-warp eigen kernel then read
    // the recycled buffer's previous contents now and then (found as a flaky last-gamma score in test_cv_branches)
    gamma_dev_.alloc(ng);
    SPB_CUDA(cudaMemcpyAsync(gamma_dev_.get(), gamma_.data(), sizeof(double) * ng, cudaMemcpyHostToDevice, ctx_.stream));
    SPB_CUDA(cudaStreamSynchronize(ctx_.stream));
  }
  gamma_out_.ensure((size_t)M_ * ng);
  const int n2 = (int)tau2_.size();
  const double sel1 = selected_tau1(index1);
  const double max_tau2 = *std::max_element(tau2_.begin(), tau2_.end());
  if (pend_core2p_ >= 0) run_pending_full();
  for (const auto& round : lane_rounds(folds, nfolds)) {
    // the refits (:473-482) of the round's folds in one launch of the batched matrix-free kernel when it applies
    bool batched = false;
    if (max_tau2 != 0.0 || sel1 != 0.0) {
      if (n2 == 1) {
        bool have_inverse = false;
        for (int k : round) have_inverse |= (cached_inverse(folds_[k], index1) != nullptr);
        batched = !have_inverse && sel1 != 0.0 && batch_eligible(round, 1, sel1, 8, 2);
      } else {
        bool have_tinv = false;
        for (int k : round) have_tinv |= folds_[k].tinv_valid;
        batched = !have_tinv && batch_eligible(round, index2 + 1, sel1, index2 + 3, 3);
      }
    }
    for (int k : round) enqueue_stage3_refit(k, index1, index2, !batched);
    if (batched) {
      bool ok;
      if (n2 == 1) {
        ok = run_core2_batched(round, {sel1}, {0}, 2, 8, nullptr);
      } else {
        std::vector<double> path(tau2_.begin(), tau2_.begin() + index2 + 1);
        std::vector<int> idx(index2 + 1);
        for (int i = 0; i <= index2; ++i) idx[i] = i;
        ok = false;
        if (pend_path_ && pend_path_i1_ == index1 && pend_path_upto_ == index2 + 1) {
          std::vector<int> with_full(round);
          with_full.push_back(-1);
          if (batch_eligible(with_full, index2 + 1, sel1, index2 + 3, 3)) {
            cudaStream_t sf = lane_of(-1).res->stream;
            copy_pk(sf, full_.R.get(), full_.Phi.get());                      // :662
            zero_pk(sf, full_.L1.get());                                      // :663
            ok = run_core2_batched(with_full, path, idx, 3, index2 + 3, nullptr, 3, sel1);
            SPB_REQUIRE(ok, "internal: the batched refit was refused after batch_eligible() accepted it");
            pend_path_ = false;
          }
        }
        if (!ok) ok = run_core2_batched(round, path, idx, 3, index2 + 3, nullptr, 3, sel1);
      }
      SPB_REQUIRE(ok, "internal: the batched refit was refused after batch_eligible() accepted it");
    }
    for (int k : round) {
      // eigen step, water filling and the covariance losses follow the refit on the fold's own stream: no host
      // synchronisation per fold (round 1 synchronised twice per fold around a host Jacobi solve)
      Lane& L = lane_of(k);
      FoldData& fd = folds_[k];
      drop_cached_inverses(fd, -1);
      Scoped sc(L.clock, PH_GAMMA, L.res->stream);
      gamma_losses_async(L.res->stream, L.gwork, L.chain.Phi.get(), p_, K_, fd.Ytr.get(), fd.n_tr, fd.tv, fd.Yva.get(),
                         fd.n_va, gamma_dev_.get(), ng, gamma_out_.get() + (size_t)k * ng);
      L.dirty = true;
    }
  }
  run_pending_full();                       // (a full-data path that found no batch to ride with)
  flush();
  for (int f = 0; f < nfolds; ++f)
    SPB_CUDA(cudaMemcpy(rows + (size_t)f * ng, gamma_out_.get() + (size_t)folds[f] * ng, sizeof(double) * ng,
                        cudaMemcpyDeviceToHost));
  stats_.bytes_d2h += (long long)sizeof(double) * ng * nfolds;
}

void Engine::get_eigenfn(double* out) {
  SPB_REQUIRE(out != nullptr, "NULL output");
  auto t0 = std::chrono::steady_clock::now();
  run_pending_full();
  flush();
  SPB_CUDA(cudaMemcpy(out, full_.Phi.get(), sizeof(double) * (size_t)p_ * K_, cudaMemcpyDeviceToHost));
  stats_.bytes_d2h += (long long)sizeof(double) * p_ * K_;
  stats_.ms_d2h += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

void Engine::get_stats(spb_cv_stats* st) {
  run_pending_full();
  flush();
  double ms[PH_COUNT], sum[PH_COUNT] = {0, 0, 0, 0, 0, 0, 0};
  clock_.collect(ms);
  for (int i = 0; i < PH_COUNT; ++i) sum[i] = ms[i];
  for (auto& L : lanes_) {
    L->clock.collect(ms);
    for (int i = 0; i < PH_COUNT; ++i) sum[i] += ms[i];
  }
  // per-phase STREAM time summed over lanes: with concurrent lanes the sum exceeds the wall time
  stats_.ms_omega = sum[PH_OMEGA];
  stats_.ms_prologue = sum[PH_PROLOGUE] + prologue_wall_ms_;     // threaded fold starts: host wall time of the group
  stats_.ms_inverse = sum[PH_INVERSE];
  stats_.ms_admm = sum[PH_ADMM];
  stats_.ms_cv_loss = sum[PH_CVLOSS];
  stats_.ms_gamma = sum[PH_GAMMA];
  stats_.ms_cg = sum[PH_CG];
  stats_.n_admm_calls = (int)(log_.size() / 4);
  long long it = 0;
  for (size_t i = 0; i < log_.size(); i += 4) it += log_[i + 3];
  stats_.admm_iters = it;
  stats_.gpu_launches = (int)(g_launch_count - launches_at_start_);
  stats_.ms_total = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_created).count();
  if (st) *st = stats_;
}

int Engine::call_log(int* log4, int cap) {
  flush();
  int ncalls = (int)(log_.size() / 4);
  if (log4) {
    int ncopy = std::min(ncalls, cap);
    std::memcpy(log4, log_.data(), sizeof(int) * 4 * (size_t)ncopy);
  }
  return ncalls;
}

void Engine::release_fold(int k) {
  SPB_REQUIRE(k >= 0 && k < M_, "fold index out of range");
  flush();
  folds_[k].tinv.release();
}

}  // namespace spb
