// extern "C" boundary (include/spatpca_b200.h).  No exception leaves this file.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

#include "engine.cuh"

namespace spb {
static thread_local std::string g_last_error;
void set_last_error(const std::string& s) { g_last_error = s; }
}  // namespace spb

using namespace spb;

struct spb_engine {
  Engine* impl;
};

namespace {

template <typename F>
int guarded(F&& f) {
  try {
    f();
    return SPB_OK;
  } catch (const Error& e) {
    set_last_error(e.what());
    return e.code;
  } catch (const std::bad_alloc&) {
    set_last_error("out of host memory");
    return SPB_ERR_INTERNAL;
  } catch (const std::exception& e) {
    set_last_error(e.what());
    return SPB_ERR_INTERNAL;
  } catch (...) {
    set_last_error("unknown failure");
    return SPB_ERR_INTERNAL;
  }
}

// p x p host (ld = p) <-> device (ld = padded)
void upload_square(const double* host, int p, double* dev, size_t ld, cudaStream_t s) {
  SPB_CUDA(cudaMemsetAsync(dev, 0, sizeof(double) * ld * p, s));
  SPB_CUDA(cudaMemcpy2DAsync(dev, ld * sizeof(double), host, (size_t)p * sizeof(double),
                             (size_t)p * sizeof(double), p, cudaMemcpyHostToDevice, s));
}
void download_square(double* host, int p, const double* dev, size_t ld, cudaStream_t s) {
  SPB_CUDA(cudaMemcpy2DAsync(host, (size_t)p * sizeof(double), dev, ld * sizeof(double),
                             (size_t)p * sizeof(double), p, cudaMemcpyDeviceToHost, s));
}

int select_index_impl(const double* cv, int M, int n_grid, double* score_out) {
  // index_min(sum(cv, 0)) with Armadillo's accumulation order; first minimum wins (:596, 658, 686)
  int best = 0;
  double best_v = 0.0;
  for (int j = 0; j < n_grid; ++j) {
    double sm = arma_accumulate(cv + (size_t)j * M, M);
    if (score_out) score_out[j] = sm / (double)M;
    if (j == 0 || sm < best_v) { best_v = sm; best = j; }
  }
  return best;
}

}  // namespace

extern "C" {

int spb_abi_version(void) { return SPB_ABI_VERSION; }

const char* spb_last_error(void) { return g_last_error.c_str(); }

int spb_device_count(void) {
  int c = 0;
  cudaError_t e = cudaGetDeviceCount(&c);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_last_error(std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
    return SPB_ERR_NO_DEVICE;
  }
  return c;
}

int spb_set_device(int ordinal) {
  return guarded([&] {
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess || c <= 0) {
      cudaGetLastError();
      throw Error(SPB_ERR_NO_DEVICE, "no CUDA device available (spatpca_b200 has no CPU fallback)");
    }
    SPB_REQUIRE(ordinal >= 0 && ordinal < c, "device ordinal out of range");
    SPB_CUDA(cudaSetDevice(ordinal));
  });
}

int spb_trim_memory(void) {
  return guarded([&] {
    spb_session_clear();
    cudaDeviceSynchronize();
    cached_trim();
  });
}

int spb_select_index(const double* cv, int M, int n_grid, double* score_out) {
  if (!cv || M < 1 || n_grid < 1) {
    set_last_error("spb_select_index: invalid argument");
    return SPB_ERR_INVALID_ARG;
  }
  return select_index_impl(cv, M, n_grid, score_out);
}

// ---------------------------------------------------------------------------------------------
int spb_tps_matrix(const double* location, int p, int d, double* omega_out) {
  return guarded([&] {
    SPB_REQUIRE(location && omega_out, "NULL argument");
    SPB_REQUIRE(p >= 1, "p must be positive");
    SPB_REQUIRE(d >= 1 && d <= 3, "Dimension of locations must be less than 4.");
    DeviceCtx& ctx = DeviceCtx::get();
    const size_t ld = padded_ld(p);
    DevBuf<double> x((size_t)p * d), om(ld * (size_t)p);
    SPB_CUDA(cudaMemcpyAsync(x.get(), location, sizeof(double) * p * d, cudaMemcpyHostToDevice, ctx.stream));
    SPB_CUDA(cudaStreamSynchronize(ctx.stream));
    OmegaRes& R = ctx.omega();
    OmegaTemps tmp;
    Engine::build_omega_raw(R, ctx.flag.get(), x.get(), p, d, om.get(), ld, false, tmp);   // the full raw product (:75)
    SPB_CUDA(cudaStreamSynchronize(R.stream));
    double r0 = -1.0;
    const bool ok = !Engine::omega_refine_enabled() || omega_build_ok(tmp, &r0);
    if (std::getenv("SPB_TRACE")) fprintf(stderr, "[spb omega] p %d spd %d ok %d max|I - A X0| = %.3e\n", p, (int)tmp.used_spd, (int)ok, r0);
    if (!ok) {      // SPD route to X0 broke down: pivoted LU
      tmp.force_lu = true;
      Engine::build_omega_raw(R, ctx.flag.get(), x.get(), p, d, om.get(), ld, false, tmp);
      SPB_CUDA(cudaStreamSynchronize(R.stream));
    }
    download_square(omega_out, p, om.get(), ld, ctx.stream);
    ctx.sync_and_check();
  });
}

int spb_eigen_function(const double* new_location, int p_new, const double* original_location, int p,
                       int d, const double* Phi, int K, double* out) {
  return guarded([&] {
    // eigenFunction (:94-147): solve the (un-ridged) bordered TPS system for the spline
    // coefficients of each eigenfunction, then evaluate at the new sites.
    SPB_REQUIRE(new_location && original_location && Phi && out, "NULL argument");
    SPB_REQUIRE(p >= 1 && p_new >= 1 && K >= 1, "sizes must be positive");
    SPB_REQUIRE(d >= 1 && d <= 3, "Dimension of locations must be less than 4.");
    DeviceCtx& ctx = DeviceCtx::get();
    cudaStream_t s = ctx.stream;
    const int q = p + d + 1;
    const size_t ldq = padded_ld(q);
    DevBuf<double> xo((size_t)p * d), xn((size_t)p_new * d), L(ldq * (size_t)q), rhs(ldq * (size_t)K),
        res((size_t)p_new * K);
    SPB_CUDA(cudaMemcpyAsync(xo.get(), original_location, sizeof(double) * p * d, cudaMemcpyHostToDevice, s));
    SPB_CUDA(cudaMemcpyAsync(xn.get(), new_location, sizeof(double) * p_new * d, cudaMemcpyHostToDevice, s));
    launch_tps_fill(xo.get(), p, d, L.get(), ldq, 0.0, true, s);                      // :100-103
    SPB_CUDA(cudaMemsetAsync(rhs.get(), 0, sizeof(double) * ldq * K, s));
    SPB_CUDA(cudaMemcpy2DAsync(rhs.get(), ldq * sizeof(double), Phi, (size_t)p * sizeof(double),
                               (size_t)p * sizeof(double), K, cudaMemcpyHostToDevice, s));   // :105-106
    int lwork = 0;
    SPB_CUSOLVER(cusolverDnDgetrf_bufferSize(ctx.solver, q, q, L.get(), (int)ldq, &lwork));
    ctx.ipiv.ensure(q);
    double* work = ctx.work((size_t)lwork);
    SPB_CUSOLVER(cusolverDnDgetrf(ctx.solver, q, q, L.get(), (int)ldq, work, ctx.ipiv.get(), ctx.info.get()));
    launch_check_info(ctx.info.get(), ctx.flag.get(), SPB_ERR_SINGULAR, s);
    SPB_CUSOLVER(cusolverDnDgetrs(ctx.solver, CUBLAS_OP_N, q, K, L.get(), (int)ldq, ctx.ipiv.get(), rhs.get(),
                                  (int)ldq, ctx.info.get()));                           // :108
    launch_tps_eval(xn.get(), p_new, xo.get(), p, d, rhs.get(), ldq, K, res.get(), s); // :113-145
    SPB_CUDA(cudaMemcpyAsync(out, res.get(), sizeof(double) * (size_t)p_new * K, cudaMemcpyDeviceToHost, s));
    ctx.sync_and_check();
  });
}

int spb_spatial_prediction(const double* phi, int p, int K, const double* Y, int n, double gamma,
                           const double* predicted_phi, int p_new, double* prediction,
                           double* estimated_covariance, double* eigenvalue, double* error) {
  return guarded([&] {
    // spatialPrediction (:715-770).  phi' cov phi and trace(cov) come from Y phi and ||Y||_F^2,
    // so the p x p covariance of :728 is never formed.
    SPB_REQUIRE(phi && Y && predicted_phi, "NULL argument");
    SPB_REQUIRE(p >= 1 && K >= 1 && n >= 1 && p_new >= 1, "sizes must be positive");
    SPB_REQUIRE(K <= SPB_MAX_K, "K must be in 1..SPB_MAX_K");
    DeviceCtx& ctx = DeviceCtx::get();
    cudaStream_t s = ctx.stream;
    DevBuf<double> dphi((size_t)p * K), dY((size_t)n * p), dpp((size_t)p_new * K), W((size_t)n * K),
        scr(yphi_scratch_doubles(n, p, K) + 512), H((size_t)K * K), ss(1);
    SPB_CUDA(cudaMemcpyAsync(dphi.get(), phi, sizeof(double) * (size_t)p * K, cudaMemcpyHostToDevice, s));
    SPB_CUDA(cudaMemcpyAsync(dY.get(), Y, sizeof(double) * (size_t)n * p, cudaMemcpyHostToDevice, s));
    SPB_CUDA(cudaMemcpyAsync(dpp.get(), predicted_phi, sizeof(double) * (size_t)p_new * K, cudaMemcpyHostToDevice, s));
    launch_yphi(dY.get(), n, p, dphi.get(), K, W.get(), scr.get(), s);
    launch_small_gram(W.get(), n, K, 1.0 / (double)n, H.get(), s);                  // phi' cov phi (:729)
    launch_sumsq(dY.get(), (size_t)n * p, scr.get(), ss.get(), s);                  // trace(cov) * n (:731)
    double ssq = 0.0;
    SPB_CUDA(cudaMemcpyAsync(&ssq, ss.get(), sizeof(double), cudaMemcpyDeviceToHost, s));
    ctx.sync_and_check();
    const double tv = ssq / (double)n;
    // eig_sym + water filling on the device (the one-warp kernel of stage 3, :729-762)
    DevBuf<double> gdev(1), Qd_dev((size_t)K * K), lam_dev((size_t)K), err_dev(1);
    SPB_CUDA(cudaMemcpyAsync(gdev.get(), &gamma, sizeof(double), cudaMemcpyHostToDevice, s));
    launch_gamma_eigen(H.get(), K, tv, p, gdev.get(), 1, Qd_dev.get(), lam_dev.get(), err_dev.get(), nullptr, s);
    std::vector<double> ev(K), Qd((size_t)K * K), Qs((size_t)K * K);
    double err = 0.0;
    SPB_CUDA(cudaMemcpyAsync(ev.data(), lam_dev.get(), sizeof(double) * K, cudaMemcpyDeviceToHost, s));   // :762
    SPB_CUDA(cudaMemcpyAsync(Qd.data(), Qd_dev.get(), sizeof(double) * K * K, cudaMemcpyDeviceToHost, s));
    SPB_CUDA(cudaMemcpyAsync(&err, err_dev.get(), sizeof(double), cudaMemcpyDeviceToHost, s));
    ctx.sync_and_check();
    for (int k = 0; k < K; ++k) {
      const double ratio = ev[k] / (ev[k] + err);                                   // :763-764
      for (int i = 0; i < K; ++i) Qs[(size_t)k * K + i] = Qd[(size_t)k * K + i] * ratio;
    }
    if (eigenvalue) std::memcpy(eigenvalue, ev.data(), sizeof(double) * K);
    if (error) *error = err;
    if (prediction || estimated_covariance) {
      DevBuf<double> Qdev((size_t)K * K), Qsdev((size_t)K * K), A1((size_t)p * K), A2((size_t)p_new * K),
          cov((size_t)p * p_new);
      SPB_CUDA(cudaMemcpyAsync(Qdev.get(), Qd.data(), sizeof(double) * K * K, cudaMemcpyHostToDevice, s));
      SPB_CUDA(cudaMemcpyAsync(Qsdev.get(), Qs.data(), sizeof(double) * K * K, cudaMemcpyHostToDevice, s));
      launch_small_rmul(dphi.get(), p, K, Qsdev.get(), A1.get(), s);                // phi Q diag(ratio)
      launch_small_rmul(dpp.get(), p_new, K, Qdev.get(), A2.get(), s);              // predicted_phi Q
      const double one = 1.0, zero = 0.0;
      SPB_CUBLAS(cublasDgemm(ctx.blas, CUBLAS_OP_N, CUBLAS_OP_T, p, p_new, K, &one, A1.get(), p, A2.get(), p_new,
                             &zero, cov.get(), p));                                 // :764
      if (estimated_covariance)
        SPB_CUDA(cudaMemcpyAsync(estimated_covariance, cov.get(), sizeof(double) * (size_t)p * p_new,
                                 cudaMemcpyDeviceToHost, s));
      if (prediction) {
        DevBuf<double> pred((size_t)n * p_new);
        SPB_CUBLAS(cublasDgemm(ctx.blas, CUBLAS_OP_N, CUBLAS_OP_N, n, p_new, p, &one, dY.get(), n, cov.get(), p,
                               &zero, pred.get(), n));                              // :765
        SPB_CUDA(cudaMemcpyAsync(prediction, pred.get(), sizeof(double) * (size_t)n * p_new,
                                 cudaMemcpyDeviceToHost, s));
        ctx.sync_and_check();
      }
      ctx.sync_and_check();
    }
  });
}

// ---------------------------------------------------------------------------------------------
int spb_engine_create(spb_engine** out, int p, int d, int n, int M, int K, const double* tau1, int n_tau1,
                      const double* tau2, int n_tau2, const double* gamma, int n_gamma, int maxit,
                      double tol, const double* l2, int n_l2) {
  return guarded([&] {
    SPB_REQUIRE(out != nullptr, "NULL handle pointer");
    *out = nullptr;
    Engine* e = new Engine(p, d, n, M, K, tau1, n_tau1, tau2, n_tau2, gamma, n_gamma, maxit, tol, l2, n_l2);
    *out = new spb_engine{e};
  });
}

int spb_engine_destroy(spb_engine* e) {
  return guarded([&] {
    if (e) {
      delete e->impl;
      delete e;
    }
  });
}

#define SPB_ENGINE_CALL(body)                                   \
  return guarded([&] {                                           \
    SPB_REQUIRE(e != nullptr && e->impl != nullptr, "NULL engine"); \
    body;                                                        \
  })

int spb_engine_upload(spb_engine* e, const double* sxy, const double* Y, const double* nk) {
  SPB_ENGINE_CALL(e->impl->upload(sxy, Y, nk));
}
int spb_engine_set_omega(spb_engine* e, const double* omega) { SPB_ENGINE_CALL(e->impl->set_omega(omega)); }
int spb_engine_prepare(spb_engine* e) { SPB_ENGINE_CALL(e->impl->prepare()); }
int spb_engine_get_omega(spb_engine* e, double* omega_out) { SPB_ENGINE_CALL(e->impl->get_omega(omega_out)); }
int spb_engine_stage1_fold(spb_engine* e, int k, double* cv_row) { SPB_ENGINE_CALL(e->impl->stage1_fold(k, cv_row)); }
int spb_engine_stage1_full(spb_engine* e, int index1) { SPB_ENGINE_CALL(e->impl->stage1_full(index1)); }
int spb_engine_stage2_fold(spb_engine* e, int k, int index1, double* cv_row) {
  SPB_ENGINE_CALL(e->impl->stage2_fold(k, index1, cv_row));
}
int spb_engine_stage2_full(spb_engine* e, int index1, int index2) {
  SPB_ENGINE_CALL(e->impl->stage2_full(index1, index2));
}
int spb_engine_stage3_fold(spb_engine* e, int k, int index1, int index2, double* cv_row) {
  SPB_ENGINE_CALL(e->impl->stage3_fold(k, index1, index2, cv_row));
}
int spb_engine_stage1_folds(spb_engine* e, const int* folds, int nfolds, double* cv_rows) {
  SPB_ENGINE_CALL(e->impl->stage1_folds(folds, nfolds, cv_rows));
}
int spb_engine_stage2_folds(spb_engine* e, const int* folds, int nfolds, int index1, double* cv_rows) {
  SPB_ENGINE_CALL(e->impl->stage2_folds(folds, nfolds, index1, cv_rows));
}
int spb_engine_stage3_folds(spb_engine* e, const int* folds, int nfolds, int index1, int index2, double* cv_rows) {
  SPB_ENGINE_CALL(e->impl->stage3_folds(folds, nfolds, index1, index2, cv_rows));
}
int spb_engine_can_share_inverses(spb_engine* e) {
  int v = 0;
  int rc = guarded([&] {
    SPB_REQUIRE(e != nullptr && e->impl != nullptr, "NULL engine");
    v = e->impl->can_share_inverses() ? 1 : 0;
  });
  return rc == SPB_OK ? v : rc;
}
int spb_engine_prepare_fold(spb_engine* e, int k) { SPB_ENGINE_CALL(e->impl->prepare_fold(k)); }
int spb_engine_prepare_starts(spb_engine* e, const int* folds, int nfolds) {
  SPB_ENGINE_CALL(e->impl->prepare_starts(folds, nfolds));
}
int spb_engine_inverse_buffer(spb_engine* e, int k, int i, void** dev_ptr, long long* bytes) {
  SPB_ENGINE_CALL(SPB_REQUIRE(dev_ptr != nullptr, "NULL output"); *dev_ptr = e->impl->inverse_buffer(k, i, bytes));
}
int spb_engine_build_inverse(spb_engine* e, int k, int i) { SPB_ENGINE_CALL(e->impl->build_inverse(k, i)); }
int spb_engine_mark_inverse(spb_engine* e, int k, int i) { SPB_ENGINE_CALL(e->impl->mark_inverse(k, i)); }
int spb_engine_sync(spb_engine* e) { SPB_ENGINE_CALL(e->impl->sync()); }
int spb_engine_omega_begin(spb_engine* e, int* needed) {
  SPB_ENGINE_CALL(SPB_REQUIRE(needed != nullptr, "NULL output"); *needed = e->impl->omega_begin() ? 1 : 0);
}
int spb_engine_omega_solve_cols(spb_engine* e, int c0, int c1) { SPB_ENGINE_CALL(e->impl->omega_solve_cols(c0, c1)); }
int spb_engine_omega_product_cols(spb_engine* e, int c0, int c1) {
  SPB_ENGINE_CALL(e->impl->omega_product_cols(c0, c1));
}
int spb_engine_omega_refine_cols(spb_engine* e, int c0, int c1, int* active) {
  SPB_ENGINE_CALL(SPB_REQUIRE(active != nullptr, "NULL output"); *active = e->impl->omega_refine_cols(c0, c1) ? 1 : 0);
}
int spb_engine_omega_buffer(spb_engine* e, int which, int c0, int c1, void** dev_ptr, long long* bytes) {
  SPB_ENGINE_CALL(SPB_REQUIRE(dev_ptr != nullptr, "NULL output"); *dev_ptr = e->impl->omega_buffer(which, c0, c1, bytes));
}
int spb_engine_omega_sync(spb_engine* e) { SPB_ENGINE_CALL(e->impl->omega_stream_sync()); }
int spb_engine_omega_finish(spb_engine* e) { SPB_ENGINE_CALL(e->impl->omega_finish()); }
int spb_engine_reset_k(spb_engine* e, int K) { SPB_ENGINE_CALL(e->impl->reset_k(K)); }
int spb_engine_get_eigenfn(spb_engine* e, double* out) { SPB_ENGINE_CALL(e->impl->get_eigenfn(out)); }
int spb_engine_get_stats(spb_engine* e, spb_cv_stats* st) { SPB_ENGINE_CALL(e->impl->get_stats(st)); }
int spb_engine_release_fold(spb_engine* e, int k) { SPB_ENGINE_CALL(e->impl->release_fold(k)); }

int spb_engine_call_log(spb_engine* e, int* log4, int cap) {
  int n = 0;
  int rc = guarded([&] {
    SPB_REQUIRE(e != nullptr && e->impl != nullptr, "NULL engine");
    n = e->impl->call_log(log4, cap);
  });
  return rc == SPB_OK ? n : rc;
}

// ---------------------------------------------------------------------------------------------
namespace {

// stages of spatpcaCV on a prepared engine (:592-692); fills the caller's outputs
void run_cv_stages(Engine& eng, int M, const double* tau2, int n_tau1, int n_tau2, const double* gamma, int n_gamma,
                   double* cv_score_tau1, double* cv_score_tau2, double* cv_score_gamma, double* estimated_eigenfn,
                   double* selected, spb_cv_stats* stats) {
  // SPB_TRACE=1: host wall clock at every stage boundary (the *_full calls only enqueue)
  static const bool tr_on = [] { const char* e = std::getenv("SPB_TRACE"); return e && (e[0] == '1' || e[0] == '2'); }();
  auto tr_last = std::chrono::steady_clock::now();
  auto tr = [&](const char* label) {
    if (!tr_on) return;
    auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[spb stage] %-14s +%8.3f ms\n", label, std::chrono::duration<double, std::milli>(now - tr_last).count());
    tr_last = now;
  };
  // stage 1 (:592-650): all folds run concurrently on their lanes
  int index1 = 0, index2 = 0, index3 = 0;
  std::vector<int> all(M);
  for (int k = 0; k < M; ++k) all[k] = k;
  auto to_colmajor = [&](const std::vector<double>& rows, int width) {   // rows: M x width row-major
    std::vector<double> cv((size_t)M * width);
    for (int k = 0; k < M; ++k)
      for (int i = 0; i < width; ++i) cv[(size_t)i * M + k] = rows[(size_t)k * width + i];
    return cv;
  };
  {
    std::vector<double> rows((size_t)M * n_tau1, 0.0);
    eng.stage1_folds(all.data(), M, rows.data());
    if (n_tau1 > 1) {
      std::vector<double> cv1 = to_colmajor(rows, n_tau1);
      index1 = select_index_impl(cv1.data(), M, n_tau1, cv_score_tau1);
    } else {
      cv_score_tau1[0] = 0.0;                                             // :649
    }
  }
  tr("stage1_folds");
  eng.stage1_full(index1);                 // full-data fit: overlaps with the folds' stage 2
  tr("stage1_full");
  selected[0] = eng.selected_tau1(index1);
  // stage 2 (:652-680)
  if (n_tau2 > 1) {
    std::vector<double> rows((size_t)M * n_tau2, 0.0);
    eng.stage2_folds(all.data(), M, index1, rows.data());
    std::vector<double> cv2 = to_colmajor(rows, n_tau2);
    index2 = select_index_impl(cv2.data(), M, n_tau2, cv_score_tau2);
    selected[1] = tau2[index2];
  } else {
    cv_score_tau2[0] = 0.0;                                               // :679
    selected[1] = tau2[0];                                                // :671
  }
  tr("stage2_folds");
  eng.stage2_full(index1, index2);         // overlaps with the folds' stage 3
  tr("stage2_full");
  // stage 3 (:682-692)
  {
    std::vector<double> rows((size_t)M * n_gamma, 0.0);
    eng.stage3_folds(all.data(), M, index1, index2, rows.data());
    std::vector<double> cv3 = to_colmajor(rows, n_gamma);
    index3 = select_index_impl(cv3.data(), M, n_gamma, cv_score_gamma);
    selected[2] = (n_gamma > 1) ? gamma[index3] : gamma[0];               // :684-691
  }
  tr("stage3_folds");
  eng.get_eigenfn(estimated_eigenfn);
  eng.get_stats(stats);
  tr("results");
}

// 64-bit content hash (word-wise multiply-xorshift); only used to recognise "same inputs as the last call"
uint64_t hash_doubles(const double* x, size_t n, uint64_t h) {
  const uint64_t* w = reinterpret_cast<const uint64_t*>(x);
  for (size_t i = 0; i < n; ++i) {
    h ^= w[i] + 0x9e3779b97f4a7c15ULL + (h << 6) + (h >> 2);
    h *= 0xff51afd7ed558ccdULL;
    h ^= h >> 32;
  }
  return h;
}

struct Session {
  std::unique_ptr<Engine> eng;
  uint64_t key = 0;
  int device = -1;
};
thread_local Session g_session;

}  // namespace

int spb_cv(const double* sxy, int p, int d, const double* Y, int n, int M, int K, const double* tau1,
           int n_tau1, const double* tau2, int n_tau2, const double* gamma, int n_gamma, const double* nk,
           int maxit, double tol, const double* l2, int n_l2, double* cv_score_tau1, double* cv_score_tau2,
           double* cv_score_gamma, double* estimated_eigenfn, double* selected, spb_cv_stats* stats) {
  return guarded([&] {
    SPB_REQUIRE(cv_score_tau1 && cv_score_tau2 && cv_score_gamma && estimated_eigenfn && selected,
                "NULL output");
    // SPB_TRACE=2: host laps only (no extra synchronisation); SPB_TRACE=1 adds the synchronous prologue trace
    static const bool tr_on = [] { const char* e = std::getenv("SPB_TRACE"); return e && (e[0] == '1' || e[0] == '2'); }();
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* label) {
      if (!tr_on) return;
      auto now = std::chrono::steady_clock::now();
      fprintf(stderr, "[spb cv] %-14s +%8.3f ms\n", label, std::chrono::duration<double, std::milli>(now - t0).count());
      t0 = now;
    };
    {
      Engine eng(p, d, n, M, K, tau1, n_tau1, tau2, n_tau2, gamma, n_gamma, maxit, tol, l2, n_l2);
      lap("engine ctor");
      eng.upload(sxy, Y, nk);
      lap("upload");
      eng.prepare();
      lap("prepare");
      run_cv_stages(eng, M, tau2, n_tau1, n_tau2, gamma, n_gamma, cv_score_tau1, cv_score_tau2, cv_score_gamma,
                    estimated_eigenfn, selected, stats);
      lap("stages");
    }
    lap("engine dtor");
  });
}

int spb_cv_session(const double* sxy, int p, int d, const double* Y, int n, int M, int K, const double* tau1,
                   int n_tau1, const double* tau2, int n_tau2, const double* gamma, int n_gamma, const double* nk,
                   int maxit, double tol, const double* l2, int n_l2, double* cv_score_tau1,
                   double* cv_score_tau2, double* cv_score_gamma, double* estimated_eigenfn, double* selected,
                   spb_cv_stats* stats) {
  return guarded([&] {
    SPB_REQUIRE(cv_score_tau1 && cv_score_tau2 && cv_score_gamma && estimated_eigenfn && selected,
                "NULL output");
    SPB_REQUIRE(sxy && Y && nk && tau1 && tau2 && gamma, "NULL input");
    SPB_REQUIRE(p >= 1 && n >= 1 && d >= 1 && n_tau1 >= 1 && n_tau2 >= 1 && n_gamma >= 1, "invalid size");
    int dev = 0;
    SPB_CUDA(cudaGetDevice(&dev));
    uint64_t key = 0x243f6a8885a308d3ULL;
    // the grid lengths are part of the key: tau1 = [a, b], tau2 = [c] must not collide with tau1 = [a], tau2 = [b, c]
    const double hdr[10] = {(double)p, (double)d, (double)n, (double)M, (double)maxit, tol, (double)n_l2,
                            (double)n_tau1, (double)n_tau2, (double)n_gamma};
    key = hash_doubles(hdr, 10, key);
    key = hash_doubles(sxy, (size_t)p * d, key);
    key = hash_doubles(Y, (size_t)n * p, key);
    key = hash_doubles(nk, (size_t)n, key);
    key = hash_doubles(tau1, n_tau1, key);
    key = hash_doubles(tau2, n_tau2, key);
    key = hash_doubles(gamma, n_gamma, key);
    if (l2 && n_l2 > 0) key = hash_doubles(l2, n_l2, key);
    Session& ss = g_session;
    if (ss.eng && ss.key == key && ss.device == dev) {
      ss.eng->reset_k(K);                      // same data and grids: Omega, SVD starts, inverses are reused
    } else {
      ss.eng.reset();
      ss.eng.reset(new Engine(p, d, n, M, K, tau1, n_tau1, tau2, n_tau2, gamma, n_gamma, maxit, tol, l2, n_l2));
      ss.key = key;
      ss.device = dev;
      ss.eng->upload(sxy, Y, nk);
    }
    try {
      ss.eng->prepare();
      run_cv_stages(*ss.eng, M, tau2, n_tau1, n_tau2, gamma, n_gamma, cv_score_tau1, cv_score_tau2, cv_score_gamma,
                    estimated_eigenfn, selected, stats);
    } catch (...) {
      ss.eng.reset();                          // never keep a half-run engine
      ss.key = 0;
      throw;
    }
  });
}

int spb_session_clear(void) {
  return guarded([&] {
    g_session.eng.reset();
    g_session.key = 0;
  });
}

// ---------------------------------------------------------------------------------------------
// kernel-level entry points
// ---------------------------------------------------------------------------------------------
int spb_inv_sympd(const double* omega, const double* G, int p, double tau1, double rho, double scale,
                  double* out) {
  return guarded([&] {
    SPB_REQUIRE(omega && G && out && p >= 1, "invalid argument");
    DeviceCtx& ctx = DeviceCtx::get();
    const size_t ld = padded_ld(p);
    DevBuf<double> om(ld * (size_t)p), g(ld * (size_t)p), w(ld * (size_t)p);
    upload_square(omega, p, om.get(), ld, ctx.stream);
    upload_square(G, p, g.get(), ld, ctx.stream);
    launch_symmatu_inplace(om.get(), p, ld, 0.0, ctx.stream);
    Engine::spd_inverse(ctx.exec(), om.get(), ld, g.get(), ld, p, tau1, rho, scale, w.get(), ld);
    download_square(out, p, w.get(), ld, ctx.stream);
    ctx.sync_and_check();
  });
}

struct SystemSpec {            // build the system on the device instead of uploading an inverse
  const double* omega;
  const double* G;
  double tau1, scale;
  int blocks;
};

static Blocking blocking_for(int p, int blocks) {
  return blocks <= 0 ? spd_blocking(p) : spd_blocking_m(p, blocks);
}

static void run_admm_host(int mode, const double* tempinv, int p, int K, double tau2, double rho, int maxit,
                          double tol, double* Phi, double* R, double* C, double* L1, double* L2, int* iters,
                          const SystemSpec* sys = nullptr) {
  SPB_REQUIRE((tempinv || sys) && Phi && C && L2 && p >= 1, "invalid argument");
  SPB_REQUIRE(K >= 1 && K <= SPB_MAX_K && K <= p, "K must be in 1..min(SPB_MAX_K, p)");
  if (mode == 3) SPB_REQUIRE(R && L1, "invalid argument");
  DeviceCtx& ctx = DeviceCtx::get();
  cudaStream_t s = ctx.stream;
  const size_t ld = padded_ld(p), pk = (size_t)p * K;
  DevBuf<double> A(ld * (size_t)p), err(1);
  DevBuf<int> stat(4);
  DevBuf<char> ws(admm_workspace_bytes(p, K));
  ChainState st;
  st.alloc(pk);
  Blocking blk;
  if (sys != nullptr) {
    SPB_REQUIRE(sys->omega && sys->G, "invalid argument");
    blk = blocking_for(p, sys->blocks);
    DevBuf<double> om(ld * (size_t)p), g(ld * (size_t)p);
    upload_square(sys->omega, p, om.get(), ld, s);
    upload_square(sys->G, p, g.get(), ld, s);
    launch_symmatu_inplace(om.get(), p, ld, 0.0, s);
    Engine::spd_inverse(ctx.exec(), om.get(), ld, g.get(), ld, p, sys->tau1, rho, sys->scale, A.get(), ld, blk);
    SPB_CUDA(cudaStreamSynchronize(s));          // om / g are released on scope exit
  } else {
    upload_square(tempinv, p, A.get(), ld, s);
  }
  SPB_CUDA(cudaMemsetAsync(st.Phi.get(), 0, sizeof(double) * pk, s));
  SPB_CUDA(cudaMemcpyAsync(st.C.get(), C, sizeof(double) * pk, cudaMemcpyHostToDevice, s));
  SPB_CUDA(cudaMemcpyAsync(st.L2.get(), L2, sizeof(double) * pk, cudaMemcpyHostToDevice, s));
  if (mode == 3) {
    SPB_CUDA(cudaMemcpyAsync(st.R.get(), R, sizeof(double) * pk, cudaMemcpyHostToDevice, s));
    SPB_CUDA(cudaMemcpyAsync(st.L1.get(), L1, sizeof(double) * pk, cudaMemcpyHostToDevice, s));
  }
  SPB_CUDA(cudaMemsetAsync(stat.get(), 0, 4 * sizeof(int), s));
  AdmmCall c;
  c.Ainv = A.get(); c.blocks = blk; c.ld = ld; c.p = p; c.K = K; c.mode = mode; c.rho = rho; c.tau2 = tau2;
  c.maxit = maxit; c.tol = tol;
  c.Phi = st.Phi.get(); c.R = st.R.get(); c.C = st.C.get(); c.L1 = st.L1.get(); c.L2 = st.L2.get();
  c.stat_slot = stat.get(); c.err_slot = err.get();
  if (maxit > 0) launch_admm(c, ws.get(), s);
  int sth[4] = {0, 0, 0, 0};
  SPB_CUDA(cudaMemcpyAsync(sth, stat.get(), sizeof(sth), cudaMemcpyDeviceToHost, s));
  if (maxit > 0) SPB_CUDA(cudaMemcpyAsync(Phi, st.Phi.get(), sizeof(double) * pk, cudaMemcpyDeviceToHost, s));
  SPB_CUDA(cudaMemcpyAsync(C, st.C.get(), sizeof(double) * pk, cudaMemcpyDeviceToHost, s));
  SPB_CUDA(cudaMemcpyAsync(L2, st.L2.get(), sizeof(double) * pk, cudaMemcpyDeviceToHost, s));
  if (mode == 3) {
    SPB_CUDA(cudaMemcpyAsync(R, st.R.get(), sizeof(double) * pk, cudaMemcpyDeviceToHost, s));
    SPB_CUDA(cudaMemcpyAsync(L1, st.L1.get(), sizeof(double) * pk, cudaMemcpyDeviceToHost, s));
  }
  ctx.sync_and_check();
  if (iters) *iters = sth[0];
  if (sth[1] != 0)
    throw Error(SPB_ERR_POLAR, "svd_econ(): the p x K iterate is rank deficient, Stiefel projection undefined");
}

int spb_admm_core2(const double* tempinv, int p, int K, double rho, int maxit, double tol, double* Phi,
                   double* C, double* Lambda2, int* iters) {
  return guarded([&] {
    run_admm_host(2, tempinv, p, K, 0.0, rho, maxit, tol, Phi, nullptr, C, nullptr, Lambda2, iters);
  });
}

int spb_admm_core3(const double* tempinv, int p, int K, double tau2, double rho, int maxit, double tol,
                   double* Phi, double* R, double* C, double* Lambda1, double* Lambda2, int* iters) {
  return guarded([&] {
    run_admm_host(3, tempinv, p, K, tau2, rho, maxit, tol, Phi, R, C, Lambda1, Lambda2, iters);
  });
}

int spb_admm_system(const double* omega, const double* G, int p, double tau1, double rho, double scale, int blocks,
                    int mode, int K, double tau2, int maxit, double tol, double* Phi, double* R, double* C,
                    double* Lambda1, double* Lambda2, int* iters) {
  return guarded([&] {
    SPB_REQUIRE(mode == 2 || mode == 3, "mode must be 2 or 3");
    SPB_REQUIRE(blocks >= 0 && blocks <= kMaxBlocks, "blocks must be in 0..kMaxBlocks");
    SystemSpec sys{omega, G, tau1, scale, blocks};
    run_admm_host(mode, nullptr, p, K, tau2, rho, maxit, tol, Phi, R, C, Lambda1, Lambda2, iters, &sys);
  });
}

int spb_core2_matrix_free(const double* omega, const double* Ytr, int n_tr, int p, int K, double tau1, double rho,
                          int maxit, double tol, double cg_tol, double* Phi, double* C, double* Lambda2, int* iters,
                          int* passes) {
  return guarded([&] {
    SPB_REQUIRE(omega && Ytr && Phi && C && Lambda2 && p >= 1 && n_tr >= 1, "invalid argument");
    SPB_REQUIRE(K >= 1 && K <= kCgMaxK && K <= p, "K must be in 1..8");
    DeviceCtx& ctx = DeviceCtx::get();
    cudaStream_t s = ctx.stream;
    const size_t ld = padded_ld(p), pk = (size_t)p * K;
    DevBuf<double> om(ld * (size_t)p), yt((size_t)n_tr * p), ytt((size_t)n_tr * p), err(1);
    DevBuf<int> stat(4);
    DevBuf<char> ws(cg_workspace_bytes(p, K, n_tr));
    ChainState st;
    st.alloc(pk);
    upload_square(omega, p, om.get(), ld, s);
    launch_symmatu_inplace(om.get(), p, ld, 0.0, s);
    SPB_CUDA(cudaMemcpyAsync(yt.get(), Ytr, sizeof(double) * (size_t)n_tr * p, cudaMemcpyHostToDevice, s));
    launch_transpose(yt.get(), n_tr, p, ytt.get(), s);
    SPB_CUDA(cudaMemcpyAsync(st.Phi.get(), Phi, sizeof(double) * pk, cudaMemcpyHostToDevice, s));
    SPB_CUDA(cudaMemcpyAsync(st.C.get(), C, sizeof(double) * pk, cudaMemcpyHostToDevice, s));
    SPB_CUDA(cudaMemcpyAsync(st.L2.get(), Lambda2, sizeof(double) * pk, cudaMemcpyHostToDevice, s));
    SPB_CUDA(cudaMemsetAsync(stat.get(), 0, 4 * sizeof(int), s));
    CgCall c;
    c.omega_u = om.get(); c.ld = ld; c.p = p; c.K = K; c.Ytr = yt.get(); c.Yt = ytt.get(); c.n_tr = n_tr;
    c.c_omega = 2.0 * tau1; c.c_gram = 2.0; c.rho = rho; c.maxit = maxit; c.tol = tol;
    c.cg_tol = cg_tol > 0.0 ? cg_tol : 1e-14; c.cg_cap = 2000;
    c.Phi = st.Phi.get(); c.C = st.C.get(); c.L2 = st.L2.get();
    c.stat_slot = stat.get(); c.err_slot = err.get();
    if (maxit > 0) launch_core2_cg(c, ws.get(), s);
    int sth[4] = {0, 0, 0, 0};
    SPB_CUDA(cudaMemcpyAsync(sth, stat.get(), sizeof(sth), cudaMemcpyDeviceToHost, s));
    SPB_CUDA(cudaMemcpyAsync(Phi, st.Phi.get(), sizeof(double) * pk, cudaMemcpyDeviceToHost, s));
    SPB_CUDA(cudaMemcpyAsync(C, st.C.get(), sizeof(double) * pk, cudaMemcpyDeviceToHost, s));
    SPB_CUDA(cudaMemcpyAsync(Lambda2, st.L2.get(), sizeof(double) * pk, cudaMemcpyDeviceToHost, s));
    ctx.sync_and_check();
    if (iters) *iters = sth[0];
    if (passes) *passes = sth[3];
    if (sth[1] == SPB_ERR_NOT_POSDEF) throw Error(SPB_ERR_NOT_POSDEF, "inv_sympd(): matrix is singular or not positive definite");
    if (sth[1] == SPB_ERR_INTERNAL) throw Error(SPB_ERR_INTERNAL, "matrix-free Core2: conjugate gradients did not converge");
    if (sth[1] != 0)
      throw Error(SPB_ERR_POLAR, "svd_econ(): the p x K iterate is rank deficient, Stiefel projection undefined");
  });
}

// mode 2: steps over tau1[0..nsteps); mode 3: steps over tau2[0..nsteps) at tau1[0], with R and Lambda1
static int core_batched_impl(int mode, const double* omega, const double* Y, const double* nk, int p, int n, int K, int F,
                             const int* fold_ids, const double* rho, const double* tau1, const double* tau2, int nsteps,
                             int maxit, double tol, double cg_tol, double* Phi, double* Rm, double* C, double* Lambda1,
                             double* Lambda2, int* iters, double* snap, int* passes) {
  return guarded([&] {
    SPB_REQUIRE(omega && Y && nk && fold_ids && rho && tau1 && Phi && C && Lambda2 && iters, "NULL argument");
    SPB_REQUIRE(mode == 2 || (tau2 && Rm && Lambda1), "NULL argument");
    SPB_REQUIRE(p >= 1 && n >= 1 && K >= 1 && F >= 1 && nsteps >= 1 && maxit >= 1, "invalid size");
    SPB_REQUIRE(batched_core2_supported(p, n, K, F), "batched Core2: unsupported problem shape on this device");
    DeviceCtx& ctx = DeviceCtx::get();
    cudaStream_t s = ctx.stream;
    const size_t ld = padded_ld(p), pk = (size_t)p * K;
    DevBuf<double> om(ld * (size_t)p), dY((size_t)n * p), dYt((size_t)n * p), st(5 * pk * F),
        dsnap(snap ? (size_t)nsteps * F * pk : 1);
    DevBuf<int> dnk(n), dit((size_t)nsteps * F), dstat(2);
    DevBuf<char> ws(batched_core2_workspace_bytes(p, n, K, F));
    upload_square(omega, p, om.get(), ld, s);
    launch_symmatu_inplace(om.get(), p, ld, 0.0, s);
    SPB_CUDA(cudaMemcpyAsync(dY.get(), Y, sizeof(double) * (size_t)n * p, cudaMemcpyHostToDevice, s));
    launch_transpose(dY.get(), n, p, dYt.get(), s);
    std::vector<int> nki(n);
    for (int r = 0; r < n; ++r) nki[r] = (int)nk[r];
    SPB_CUDA(cudaMemcpyAsync(dnk.get(), nki.data(), sizeof(int) * n, cudaMemcpyHostToDevice, s));
    SPB_CUDA(cudaMemcpyAsync(st.get(), Phi, sizeof(double) * pk * F, cudaMemcpyHostToDevice, s));
    SPB_CUDA(cudaMemcpyAsync(st.get() + pk * F, C, sizeof(double) * pk * F, cudaMemcpyHostToDevice, s));
    SPB_CUDA(cudaMemcpyAsync(st.get() + 2 * pk * F, Lambda2, sizeof(double) * pk * F, cudaMemcpyHostToDevice, s));
    if (mode == 3) {
      SPB_CUDA(cudaMemcpyAsync(st.get() + 3 * pk * F, Rm, sizeof(double) * pk * F, cudaMemcpyHostToDevice, s));
      SPB_CUDA(cudaMemcpyAsync(st.get() + 4 * pk * F, Lambda1, sizeof(double) * pk * F, cudaMemcpyHostToDevice, s));
    }
    SPB_CUDA(cudaMemsetAsync(dit.get(), 0, sizeof(int) * nsteps * F, s));
    SPB_CUDA(cudaMemsetAsync(dstat.get(), 0, 2 * sizeof(int), s));
    BatchedCore2Call c;
    c.omega_u = om.get(); c.ld = ld; c.p = p; c.n = n; c.K = K; c.F = F;
    c.Y = dY.get(); c.Yt = dYt.get(); c.nk = dnk.get();
    for (int f = 0; f < 8; ++f) {
      c.fold_id[f] = f < F ? fold_ids[f] : 0;
      c.rho[f] = f < F ? rho[f] : 1.0;
      c.Phi[f] = f < F ? st.get() + (size_t)f * pk : nullptr;
      c.C[f] = f < F ? st.get() + pk * F + (size_t)f * pk : nullptr;
      c.L2[f] = f < F ? st.get() + 2 * pk * F + (size_t)f * pk : nullptr;
      c.R[f] = (mode == 3 && f < F) ? st.get() + 3 * pk * F + (size_t)f * pk : nullptr;
      c.L1[f] = (mode == 3 && f < F) ? st.get() + 4 * pk * F + (size_t)f * pk : nullptr;
    }
    std::vector<double> t1rep(nsteps, tau1[0]);
    c.mode = mode;
    c.nsteps = nsteps; c.tau1 = mode == 3 ? t1rep.data() : tau1; c.tau2 = mode == 3 ? tau2 : nullptr;
    c.maxit = maxit; c.tol = tol;
    c.cg_tol = cg_tol > 0.0 ? cg_tol : 1e-14; c.cg_cap = 2000;
    c.snap = snap ? dsnap.get() : nullptr; c.iters_out = dit.get(); c.status_out = dstat.get();
    const bool dbg_on = std::getenv("SPB_BCG_DBG") != nullptr;
    DevBuf<unsigned long long> dbg(400);
    SPB_CUDA(cudaMemsetAsync(dbg.get(), 0, sizeof(unsigned long long) * 400, s));
    if (dbg_on) c.dbg = dbg.get();
    launch_batched_core2(c, ws.get(), s);
    if (dbg_on) {
      unsigned long long h[400];
      SPB_CUDA(cudaStreamSynchronize(s));
      SPB_CUDA(cudaMemcpy(h, dbg.get(), sizeof(h), cudaMemcpyDeviceToHost));
      const char* names[8] = {"stream", "r.u", "barrier1", "S+z", "update", "Spartial", "barrier2", "reduce"};
      for (int itx = 1; itx < 6; ++itx) {
        const unsigned long long* q = h + 9 * itx;
        if (q[8] == 0) break;
        fprintf(stderr, "[bcg dbg] p=%d F=%d K=%d CG step %d:", p, F, K, itx);
        for (int k = 0; k < 8; ++k) fprintf(stderr, " %s=%.1f", names[k], (double)(q[k + 1] - q[k]) * 1e-3);
        fprintf(stderr, " total=%.1f us\n", (double)(q[8] - q[0]) * 1e-3);
      }
    }
    int sth[2] = {0, 0};
    SPB_CUDA(cudaMemcpyAsync(sth, dstat.get(), sizeof(sth), cudaMemcpyDeviceToHost, s));
    SPB_CUDA(cudaMemcpyAsync(iters, dit.get(), sizeof(int) * nsteps * F, cudaMemcpyDeviceToHost, s));
    SPB_CUDA(cudaMemcpyAsync(Phi, st.get(), sizeof(double) * pk * F, cudaMemcpyDeviceToHost, s));
    SPB_CUDA(cudaMemcpyAsync(C, st.get() + pk * F, sizeof(double) * pk * F, cudaMemcpyDeviceToHost, s));
    SPB_CUDA(cudaMemcpyAsync(Lambda2, st.get() + 2 * pk * F, sizeof(double) * pk * F, cudaMemcpyDeviceToHost, s));
    if (mode == 3) {
      SPB_CUDA(cudaMemcpyAsync(Rm, st.get() + 3 * pk * F, sizeof(double) * pk * F, cudaMemcpyDeviceToHost, s));
      SPB_CUDA(cudaMemcpyAsync(Lambda1, st.get() + 4 * pk * F, sizeof(double) * pk * F, cudaMemcpyDeviceToHost, s));
    }
    if (snap)
      SPB_CUDA(cudaMemcpyAsync(snap, dsnap.get(), sizeof(double) * (size_t)nsteps * F * pk, cudaMemcpyDeviceToHost, s));
    ctx.sync_and_check();
    if (passes) *passes = sth[1];
    if (sth[0] == SPB_ERR_NOT_POSDEF) throw Error(SPB_ERR_NOT_POSDEF, "inv_sympd(): matrix is singular or not positive definite");
    if (sth[0] == SPB_ERR_INTERNAL) throw Error(SPB_ERR_INTERNAL, "matrix-free Core2: conjugate gradients did not converge");
    if (sth[0] != 0)
      throw Error(SPB_ERR_POLAR, "svd_econ(): the p x K iterate is rank deficient, Stiefel projection undefined");
  });
}

int spb_core2_batched(const double* omega, const double* Y, const double* nk, int p, int n, int K, int F,
                      const int* fold_ids, const double* rho, const double* tau1, int nsteps, int maxit, double tol,
                      double cg_tol, double* Phi, double* C, double* Lambda2, int* iters, double* snap, int* passes) {
  return core_batched_impl(2, omega, Y, nk, p, n, K, F, fold_ids, rho, tau1, nullptr, nsteps, maxit, tol, cg_tol, Phi,
                           nullptr, C, nullptr, Lambda2, iters, snap, passes);
}

int spb_core3_batched(const double* omega, const double* Y, const double* nk, int p, int n, int K, int F,
                      const int* fold_ids, const double* rho, double tau1, const double* tau2, int nsteps, int maxit,
                      double tol, double cg_tol, double* Phi, double* R, double* C, double* Lambda1, double* Lambda2,
                      int* iters, double* snap, int* passes) {
  return core_batched_impl(3, omega, Y, nk, p, n, K, F, fold_ids, rho, &tau1, tau2, nsteps, maxit, tol, cg_tol, Phi, R, C,
                           Lambda1, Lambda2, iters, snap, passes);
}

int spb_core2_batched_supported(int p, int n, int K, int F) {
  int v = 0;
  int rc = guarded([&] { DeviceCtx::get(); v = batched_core2_supported(p, n, K, F) ? 1 : 0; });
  return rc < 0 ? rc : v;
}

int spb_default_blocks(int p) {
  int m = -1;
  int rc = guarded([&] {
    SPB_REQUIRE(p >= 1, "invalid argument");
    m = spd_blocking(p).m;
  });
  return rc < 0 ? rc : m;
}

int spb_block_offsets(int p, int blocks, int* off, int cap) {
  int m = -1;
  int rc = guarded([&] {
    SPB_REQUIRE(p >= 1 && off != nullptr && blocks >= 0 && blocks <= kMaxBlocks, "invalid argument");
    const Blocking b = blocking_for(p, blocks);
    SPB_REQUIRE(cap >= b.m + 1, "offset buffer too small");
    for (int j = 0; j <= b.m; ++j) off[j] = j == 0 ? 0 : b.off[j];
    if (b.m == 1) off[1] = p;
    m = b.m;
  });
  return rc < 0 ? rc : m;
}

int spb_admm_phase_plan(int p, int K, int blocks, int* plan8, int cap) {
  int n = -1;
  int rc = guarded([&] {
    SPB_REQUIRE(p >= 1 && K >= 1 && K <= SPB_MAX_K && plan8 != nullptr && cap >= 1, "invalid argument");
    SPB_REQUIRE(blocks >= 0 && blocks <= kMaxBlocks, "blocks must be in 0..kMaxBlocks");
    n = admm_phase_plan(p, K, blocking_for(p, blocks), plan8, cap);
  });
  return rc < 0 ? rc : n;
}

int spb_default_path_blocks(int p) {
  int m = -1;
  int rc = guarded([&] {
    SPB_REQUIRE(p >= 1, "invalid argument");
    m = spd_blocking_path(p).m;
  });
  return rc < 0 ? rc : m;
}

int spb_cv_loss(const double* Yva, int n_va, int p, const double* Phi, int K, double* loss) {
  return guarded([&] {
    SPB_REQUIRE(Yva && Phi && loss && n_va >= 1 && p >= 1 && K >= 1, "invalid argument");
    DeviceCtx& ctx = DeviceCtx::get();
    cudaStream_t s = ctx.stream;
    DevBuf<double> dY((size_t)n_va * p), dP((size_t)p * K), W((size_t)n_va * K),
        scr(yphi_scratch_doubles(n_va, p, K) + 16), out(1);
    SPB_CUDA(cudaMemcpyAsync(dY.get(), Yva, sizeof(double) * (size_t)n_va * p, cudaMemcpyHostToDevice, s));
    SPB_CUDA(cudaMemcpyAsync(dP.get(), Phi, sizeof(double) * (size_t)p * K, cudaMemcpyHostToDevice, s));
    Engine::cv_loss(s, dY.get(), n_va, p, dP.get(), K, W.get(), scr.get(), out.get());
    SPB_CUDA(cudaMemcpyAsync(loss, out.get(), sizeof(double), cudaMemcpyDeviceToHost, s));
    ctx.sync_and_check();
  });
}

int spb_svd_right(const double* Y, int m, int p, int K, double* sv_out, double* V_out) {
  return guarded([&] {
    SPB_REQUIRE(Y && sv_out && m >= 1 && p >= 1, "invalid argument");
    const int r = m < p ? m : p;
    SPB_REQUIRE(K >= 1 && K <= r, "K must be in 1 .. min(m, p)");
    DeviceCtx& ctx = DeviceCtx::get();
    cudaStream_t s = ctx.stream;
    DevBuf<double> dY((size_t)m * p), dV((size_t)p * K);
    SPB_CUDA(cudaMemcpyAsync(dY.get(), Y, sizeof(double) * (size_t)m * p, cudaMemcpyHostToDevice, s));
    std::vector<double> sv;
    Engine::svd_right(ctx, dY.get(), m, p, K, sv, dV.get());          // synchronises the stream, checks the flag
    for (int i = 0; i < r; ++i) sv_out[i] = sv[i];
    if (V_out) {
      SPB_CUDA(cudaMemcpyAsync(V_out, dV.get(), sizeof(double) * (size_t)p * K, cudaMemcpyDeviceToHost, s));
      ctx.sync_and_check();
    }
  });
}

int spb_gamma_loss(const double* Phi, int p, int K, const double* Ytr, int n_tr, const double* Yva, int n_va,
                   const double* gamma, int n_gamma, double* out) {
  return guarded([&] {
    SPB_REQUIRE(Phi && Ytr && Yva && gamma && out, "NULL argument");
    SPB_REQUIRE(p >= 1 && K >= 1 && K <= SPB_MAX_K && n_tr >= 1 && n_va >= 1 && n_gamma >= 1, "invalid size");
    DeviceCtx& ctx = DeviceCtx::get();
    cudaStream_t s = ctx.stream;
    DevBuf<double> dP((size_t)p * K), dtr((size_t)n_tr * p), dva((size_t)n_va * p), scr(512), ss(1);
    SPB_CUDA(cudaMemcpyAsync(dP.get(), Phi, sizeof(double) * (size_t)p * K, cudaMemcpyHostToDevice, s));
    SPB_CUDA(cudaMemcpyAsync(dtr.get(), Ytr, sizeof(double) * (size_t)n_tr * p, cudaMemcpyHostToDevice, s));
    SPB_CUDA(cudaMemcpyAsync(dva.get(), Yva, sizeof(double) * (size_t)n_va * p, cudaMemcpyHostToDevice, s));
    launch_sumsq(dtr.get(), (size_t)n_tr * p, scr.get(), ss.get(), s);
    double ssq = 0.0;
    SPB_CUDA(cudaMemcpyAsync(&ssq, ss.get(), sizeof(double), cudaMemcpyDeviceToHost, s));
    ctx.sync_and_check();
    std::vector<double> g(gamma, gamma + n_gamma);
    Engine::gamma_losses(ctx, s, dP.get(), p, K, dtr.get(), n_tr, ssq / (double)n_tr, dva.get(), n_va, g, out);
  });
}

// ---------------------------------------------------------------------------------------------
// device-resident benchmark hooks
// ---------------------------------------------------------------------------------------------
int spb_bench_admm(int p, int K, int mode, int iters, int repeats, double* ms_per_iter) {
  return spb_bench_admm_blocks(p, K, mode, iters, repeats, 1, ms_per_iter);
}

int spb_bench_admm_blocks(int p, int K, int mode, int iters, int repeats, int blocks, double* ms_per_iter) {
  return guarded([&] {
    SPB_REQUIRE(blocks >= 0 && blocks <= kMaxBlocks, "blocks must be in 0..kMaxBlocks");
    const Blocking blk = blocking_for(p, blocks);
    SPB_REQUIRE(p >= 1 && K >= 1 && K <= SPB_MAX_K && K <= p && iters >= 1 && repeats >= 1 && ms_per_iter,
                "invalid argument");
    SPB_REQUIRE(mode == 2 || mode == 3, "mode must be 2 or 3");
    DeviceCtx& ctx = DeviceCtx::get();
    cudaStream_t s = ctx.stream;
    const size_t ld = padded_ld(p), pk = (size_t)p * K;
    DevBuf<double> A(ld * (size_t)p), err(1);
    DevBuf<int> stat(4);
    DevBuf<char> ws(admm_workspace_bytes(p, K));
    ChainState st;
    st.alloc(pk);
    // a well-posed synthetic chain: Ainv = I (also a valid block factor: D_j^-1 = I, U = 0),
    // C = R = first K unit vectors, duals = 0, rho = 1
    SPB_CUDA(cudaMemsetAsync(A.get(), 0, sizeof(double) * ld * p, s));
    launch_set_identity(A.get(), p, p, ld, s);
    cudaEvent_t e0, e1;
    SPB_CUDA(cudaEventCreate(&e0));
    SPB_CUDA(cudaEventCreate(&e1));
    double sum_ms = 0.0;
    const bool dbg_on = std::getenv("SPB_ADMM_DBG") != nullptr;
    DevBuf<unsigned long long> dbg(2048);
    SPB_CUDA(cudaMemsetAsync(dbg.get(), 0, sizeof(unsigned long long) * 2048, s));
    // start: C = R = Q with NON-orthogonal columns (e_k + 0.15 e_{k+1} + 0.1 e_{k+2}), so that T'T is not
    // diagonal and the Jacobi step of every iteration does real sweeps, as in a job (SPB_BENCH_DIAG=1: Q = E,
    // the zero-sweep best case)
    std::vector<double> q0(pk, 0.0);
    const bool diag_start = std::getenv("SPB_BENCH_DIAG") != nullptr;
    for (int k = 0; k < K; ++k) {
      q0[(size_t)k * p + k] = 1.0;
      if (!diag_start && k + 1 < p) q0[(size_t)k * p + k + 1] = 0.15;
      if (!diag_start && k + 2 < p) q0[(size_t)k * p + k + 2] = 0.1;
    }
    DevBuf<double> Q0(pk);
    SPB_CUDA(cudaMemcpyAsync(Q0.get(), q0.data(), sizeof(double) * pk, cudaMemcpyHostToDevice, s));
    SPB_CUDA(cudaStreamSynchronize(s));
    int last_sweeps = 0;
    for (int r = 0; r < repeats + 1; ++r) {
      SPB_CUDA(cudaMemcpyAsync(st.C.get(), Q0.get(), sizeof(double) * pk, cudaMemcpyDeviceToDevice, s));
      SPB_CUDA(cudaMemcpyAsync(st.R.get(), Q0.get(), sizeof(double) * pk, cudaMemcpyDeviceToDevice, s));
      SPB_CUDA(cudaMemsetAsync(st.L1.get(), 0, sizeof(double) * pk, s));
      SPB_CUDA(cudaMemsetAsync(st.L2.get(), 0, sizeof(double) * pk, s));
      AdmmCall c;
      c.Ainv = A.get(); c.blocks = blk; c.ld = ld; c.p = p; c.K = K; c.mode = mode; c.rho = 1.0; c.tau2 = 1e-3;
      c.maxit = iters; c.tol = -1.0;          // never satisfied: every call runs `iters` iterations
      c.Phi = st.Phi.get(); c.R = st.R.get(); c.C = st.C.get(); c.L1 = st.L1.get(); c.L2 = st.L2.get();
      c.stat_slot = stat.get(); c.err_slot = err.get();
      c.dbg = dbg_on ? dbg.get() : nullptr;
      SPB_CUDA(cudaEventRecord(e0, s));
      launch_admm(c, ws.get(), s);
      SPB_CUDA(cudaEventRecord(e1, s));
      SPB_CUDA(cudaStreamSynchronize(s));
      int sth[4];
      SPB_CUDA(cudaMemcpy(sth, stat.get(), sizeof(sth), cudaMemcpyDeviceToHost));
      if (sth[0] != iters || sth[1] != 0) throw Error(SPB_ERR_INTERNAL, "bench chain did not run all iterations");
      last_sweeps = sth[2];
      float ms = 0.f;
      SPB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      if (r > 0) sum_ms += (double)ms / iters;   // r == 0 is the warm-up; the rest are averaged
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (dbg_on) {
      unsigned long long h[2048];
      SPB_CUDA(cudaMemcpy(h, dbg.get(), sizeof(h), cudaMemcpyDeviceToHost));
      if (std::getenv("SPB_ADMM_DBG_CTAS")) {
        fprintf(stderr, "[admm dbg] per-CTA stream end (us after CTA0's iteration-5 start):");
        for (int b = 0; b < 296; ++b) fprintf(stderr, " %.0f", ((double)h[256 + b] - (double)h[5 * 8]) * 1e-3);
        fprintf(stderr, "\n");
      }
      for (int b = 0; b < 4 && blk.m > 1; ++b) {
        fprintf(stderr, "[admm dbg] CTA %d, iteration 3, per phase (us): busy / wait-at-barrier:", b);
        const unsigned long long* q = h + 1024 + b * 64;
        for (int f = 0; f < 2 * blk.m - 1; ++f)
          fprintf(stderr, " %.1f/%.1f", (double)(q[2 * f + 1] - q[2 * f]) * 1e-3,
                  f + 1 < 2 * blk.m - 1 ? (double)(q[2 * f + 2] - q[2 * f + 1]) * 1e-3 : 0.0);
        fprintf(stderr, "\n");
      }
      const char* names[8] = {"stream+reduce", "barrier1", "update", "barrier2", "(loop)", "-", "-", "-"};
      for (int it = 2; it < std::min(iters, 6); ++it) {
        fprintf(stderr, "[admm dbg] p=%d K=%d it=%d:", p, K, it);
        for (int q = 0; q < 4; ++q) fprintf(stderr, " %s=%.1fus", names[q], (double)(h[it * 8 + q + 1] - h[it * 8 + q]) * 1e-3);
        fprintf(stderr, " total=%.1fus\n", (double)(h[(it + 1) * 8] - h[it * 8]) * 1e-3);
      }
      fprintf(stderr, "[admm dbg] last iteration: gram-reduce=%.1fus jacobi=%.1fus (%llu sweeps; init %.1f, sweeps %.1f, epilogue %.1f) tile0-reduce=%.1fus\n",
              (double)(h[121] - h[120]) * 1e-3, (double)(h[122] - h[121]) * 1e-3, h[125], (double)(h[126] - h[121]) * 1e-3,
              (double)(h[127] - h[126]) * 1e-3, (double)(h[122] - h[127]) * 1e-3, (double)(h[124] - h[123]) * 1e-3);
    }
    if (dbg_on) fprintf(stderr, "[admm dbg] Jacobi sweeps in the last iteration: %d\n", last_sweeps);
    *ms_per_iter = sum_ms / repeats;
  });
}

namespace {
__global__ void bench_fill_kernel(double* A, size_t ld, int p, double diag, double off) {
  // banded symmetric matrix: diag on the diagonal, off / (1 + |i - j|) within 8 of it -- SPD and non-trivial
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (size_t)p * p) return;
  const int i = (int)(e % p), j = (int)(e / p);
  const int d = i > j ? i - j : j - i;
  A[(size_t)j * ld + i] = d == 0 ? diag : (d <= 8 ? off / (1.0 + d) : 0.0);
}
__global__ void bench_sign_kernel(double* Y, size_t n, double scale) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  unsigned int h = (unsigned int)e * 2654435761u + 12345u;
  h ^= h >> 15; h *= 0x85ebca6bu; h ^= h >> 13;
  Y[e] = (h & 1u) ? scale : -scale;
}
}  // namespace

int spb_bench_core2_cg(int p, int K, int n_tr, int iters, int repeats, double* ms_per_launch, int* passes_per_launch) {
  return guarded([&] {
    SPB_REQUIRE(p >= 1 && K >= 1 && K <= kCgMaxK && K <= p && n_tr >= 1 && iters >= 1 && repeats >= 1 && ms_per_launch,
                "invalid argument");
    DeviceCtx& ctx = DeviceCtx::get();
    cudaStream_t s = ctx.stream;
    const size_t ld = padded_ld(p), pk = (size_t)p * K;
    DevBuf<double> om(ld * (size_t)p), yt((size_t)n_tr * p), ytt((size_t)n_tr * p), err(1), Q0(pk);
    DevBuf<int> stat(4);
    DevBuf<char> ws(cg_workspace_bytes(p, K, n_tr));
    ChainState st;
    st.alloc(pk);
    // rho = 1; Omega_u has lambda_max ~ 0.25 / c_omega-scaled, the Gram part ~0.1: cond(A) ~ 1.6, as on the C4 job
    SPB_CUDA(cudaMemsetAsync(om.get(), 0, sizeof(double) * ld * p, s));
    bench_fill_kernel<<<ceil_div((long long)p * p, 256), 256, 0, s>>>(om.get(), ld, p, 0.1, 0.03);
    bench_sign_kernel<<<ceil_div((long long)n_tr * p, 256), 256, 0, s>>>(yt.get(), (size_t)n_tr * p,
                                                                         std::sqrt(0.05 / ((double)n_tr * p)));
    count_launch(2);
    launch_transpose(yt.get(), n_tr, p, ytt.get(), s);
    std::vector<double> q0(pk, 0.0);
    for (int k = 0; k < K; ++k) {
      q0[(size_t)k * p + k] = 1.0;
      if (k + 1 < p) q0[(size_t)k * p + k + 1] = 0.15;
      if (k + 2 < p) q0[(size_t)k * p + k + 2] = 0.1;
    }
    SPB_CUDA(cudaMemcpyAsync(Q0.get(), q0.data(), sizeof(double) * pk, cudaMemcpyHostToDevice, s));
    SPB_CUDA(cudaStreamSynchronize(s));
    cudaEvent_t e0, e1;
    SPB_CUDA(cudaEventCreate(&e0));
    SPB_CUDA(cudaEventCreate(&e1));
    double sum_ms = 0.0;
    int passes = 0;
    for (int r = 0; r < repeats + 1; ++r) {
      SPB_CUDA(cudaMemcpyAsync(st.C.get(), Q0.get(), sizeof(double) * pk, cudaMemcpyDeviceToDevice, s));
      SPB_CUDA(cudaMemcpyAsync(st.Phi.get(), Q0.get(), sizeof(double) * pk, cudaMemcpyDeviceToDevice, s));
      SPB_CUDA(cudaMemsetAsync(st.L2.get(), 0, sizeof(double) * pk, s));
      CgCall c;
      c.omega_u = om.get(); c.ld = ld; c.p = p; c.K = K; c.Ytr = yt.get(); c.Yt = ytt.get(); c.n_tr = n_tr;
      c.c_omega = 2.0; c.c_gram = 2.0; c.rho = 1.0; c.maxit = iters; c.tol = -1.0;
      c.cg_tol = 1e-14; c.cg_cap = 2000;
      c.Phi = st.Phi.get(); c.C = st.C.get(); c.L2 = st.L2.get();
      c.stat_slot = stat.get(); c.err_slot = err.get();
      SPB_CUDA(cudaEventRecord(e0, s));
      launch_core2_cg(c, ws.get(), s);
      SPB_CUDA(cudaEventRecord(e1, s));
      SPB_CUDA(cudaStreamSynchronize(s));
      int sth[4];
      SPB_CUDA(cudaMemcpy(sth, stat.get(), sizeof(sth), cudaMemcpyDeviceToHost));
      if (sth[0] != iters || sth[1] != 0) throw Error(SPB_ERR_INTERNAL, "bench chain did not run all iterations");
      passes = sth[3];
      float ms = 0.f;
      SPB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      if (r > 0) sum_ms += (double)ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *ms_per_launch = sum_ms / repeats;
    if (passes_per_launch) *passes_per_launch = passes;
  });
}

int spb_dmma_matvec(const double* A, int p, const double* V, int nc, double* U) {
  return guarded([&] {
    SPB_REQUIRE(A && V && U && p >= 1 && nc >= 1 && nc <= 32, "invalid argument");
    DeviceCtx& ctx = DeviceCtx::get();
    cudaStream_t s = ctx.stream;
    const size_t ld = padded_ld(p);
    DevBuf<double> dA(ld * (size_t)p), dV(ld * (size_t)nc), dU((size_t)p * nc);
    upload_square(A, p, dA.get(), ld, s);
    SPB_CUDA(cudaMemsetAsync(dV.get(), 0, sizeof(double) * ld * nc, s));
    SPB_CUDA(cudaMemcpy2DAsync(dV.get(), ld * sizeof(double), V, (size_t)p * sizeof(double), (size_t)p * sizeof(double), nc,
                               cudaMemcpyHostToDevice, s));
    launch_dmma_matvec(dA.get(), ld, p, dV.get(), ld, nc, dU.get(), s);
    SPB_CUDA(cudaMemcpyAsync(U, dU.get(), sizeof(double) * (size_t)p * nc, cudaMemcpyDeviceToHost, s));
    ctx.sync_and_check();
  });
}

int spb_bench_dmma_matvec(int p, int nc, int repeats, double* ms_per_pass) {
  return guarded([&] {
    SPB_REQUIRE(p >= 1 && nc >= 1 && nc <= 32 && repeats >= 1 && ms_per_pass, "invalid argument");
    DeviceCtx& ctx = DeviceCtx::get();
    cudaStream_t s = ctx.stream;
    const size_t ld = padded_ld(p);
    DevBuf<double> dA(ld * (size_t)p), dV(ld * (size_t)nc), dU((size_t)p * nc);
    SPB_CUDA(cudaMemsetAsync(dA.get(), 0, sizeof(double) * ld * p, s));
    bench_fill_kernel<<<ceil_div((long long)p * p, 256), 256, 0, s>>>(dA.get(), ld, p, 0.1, 0.03);
    bench_sign_kernel<<<ceil_div((long long)ld * nc, 256), 256, 0, s>>>(dV.get(), ld * (size_t)nc, 1.0);
    count_launch(2);
    cudaEvent_t e0, e1;
    SPB_CUDA(cudaEventCreate(&e0));
    SPB_CUDA(cudaEventCreate(&e1));
    double sum = 0.0;
    for (int r = 0; r < repeats + 2; ++r) {
      SPB_CUDA(cudaEventRecord(e0, s));
      launch_dmma_matvec(dA.get(), ld, p, dV.get(), ld, nc, dU.get(), s);
      SPB_CUDA(cudaEventRecord(e1, s));
      ctx.sync_and_check();
      float ms = 0.f;
      SPB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      if (r > 1) sum += (double)ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *ms_per_pass = sum / repeats;
  });
}

int spb_bench_core2_batched(int p, int n, int K, int F, int iters, int repeats, double* ms_per_launch,
                            int* passes_per_launch) {
  return guarded([&] {
    SPB_REQUIRE(p >= 1 && n >= 2 && K >= 1 && F >= 1 && iters >= 1 && repeats >= 1 && ms_per_launch, "invalid argument");
    SPB_REQUIRE(batched_core2_supported(p, n, K, F), "batched Core2: unsupported problem shape on this device");
    DeviceCtx& ctx = DeviceCtx::get();
    cudaStream_t s = ctx.stream;
    const size_t ld = padded_ld(p), pk = (size_t)p * K;
    DevBuf<double> om(ld * (size_t)p), dY((size_t)n * p), dYt((size_t)n * p), st(3 * pk * F), Q0(pk);
    DevBuf<int> dnk(n), dit(F), dstat(2);
    DevBuf<char> ws(batched_core2_workspace_bytes(p, n, K, F));
    // the synthetic system of spb_bench_core2_cg: rho = 1, banded Omega_u, random-sign data, cond(A) ~ 1.6
    SPB_CUDA(cudaMemsetAsync(om.get(), 0, sizeof(double) * ld * p, s));
    bench_fill_kernel<<<ceil_div((long long)p * p, 256), 256, 0, s>>>(om.get(), ld, p, 0.1, 0.03);
    bench_sign_kernel<<<ceil_div((long long)n * p, 256), 256, 0, s>>>(dY.get(), (size_t)n * p, std::sqrt(0.05 / ((double)n * p)));
    count_launch(2);
    launch_transpose(dY.get(), n, p, dYt.get(), s);
    std::vector<int> nki(n);
    for (int r = 0; r < n; ++r) nki[r] = 1 + r % (F > 1 ? F : 2);
    SPB_CUDA(cudaMemcpyAsync(dnk.get(), nki.data(), sizeof(int) * n, cudaMemcpyHostToDevice, s));
    std::vector<double> q0(pk, 0.0);
    for (int k = 0; k < K; ++k) {
      q0[(size_t)k * p + k] = 1.0;
      if (k + 1 < p) q0[(size_t)k * p + k + 1] = 0.15;
      if (k + 2 < p) q0[(size_t)k * p + k + 2] = 0.1;
    }
    SPB_CUDA(cudaMemcpyAsync(Q0.get(), q0.data(), sizeof(double) * pk, cudaMemcpyHostToDevice, s));
    SPB_CUDA(cudaStreamSynchronize(s));
    cudaEvent_t e0, e1;
    SPB_CUDA(cudaEventCreate(&e0));
    SPB_CUDA(cudaEventCreate(&e1));
    double sum_ms = 0.0;
    int passes = 0;
    const double tau = 1.0;
    for (int r = 0; r < repeats + 1; ++r) {
      for (int f = 0; f < F; ++f) {
        SPB_CUDA(cudaMemcpyAsync(st.get() + (size_t)f * pk, Q0.get(), sizeof(double) * pk, cudaMemcpyDeviceToDevice, s));
        SPB_CUDA(cudaMemcpyAsync(st.get() + pk * F + (size_t)f * pk, Q0.get(), sizeof(double) * pk, cudaMemcpyDeviceToDevice, s));
      }
      SPB_CUDA(cudaMemsetAsync(st.get() + 2 * pk * F, 0, sizeof(double) * pk * F, s));
      BatchedCore2Call c;
      c.omega_u = om.get(); c.ld = ld; c.p = p; c.n = n; c.K = K; c.F = F;
      c.Y = dY.get(); c.Yt = dYt.get(); c.nk = dnk.get();
      for (int f = 0; f < 8; ++f) {
        c.fold_id[f] = (f < F && F > 1) ? f + 1 : 0;
        c.rho[f] = 1.0;
        c.Phi[f] = f < F ? st.get() + (size_t)f * pk : nullptr;
        c.C[f] = f < F ? st.get() + pk * F + (size_t)f * pk : nullptr;
        c.L2[f] = f < F ? st.get() + 2 * pk * F + (size_t)f * pk : nullptr;
      }
      c.nsteps = 1; c.tau1 = &tau; c.maxit = iters; c.tol = -1.0; c.cg_tol = 1e-14; c.cg_cap = 2000;
      c.snap = nullptr; c.iters_out = dit.get(); c.status_out = dstat.get();
      SPB_CUDA(cudaEventRecord(e0, s));
      launch_batched_core2(c, ws.get(), s);
      SPB_CUDA(cudaEventRecord(e1, s));
      SPB_CUDA(cudaStreamSynchronize(s));
      int sth[2], ith[8];
      SPB_CUDA(cudaMemcpy(sth, dstat.get(), sizeof(sth), cudaMemcpyDeviceToHost));
      SPB_CUDA(cudaMemcpy(ith, dit.get(), sizeof(int) * F, cudaMemcpyDeviceToHost));
      if (sth[0] != 0 || ith[0] != iters) throw Error(SPB_ERR_INTERNAL, "bench chain did not run all iterations");
      passes = sth[1];
      float ms = 0.f;
      SPB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      if (r > 0) sum_ms += (double)ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *ms_per_launch = sum_ms / repeats;
    if (passes_per_launch) *passes_per_launch = passes;
  });
}

int spb_bench_dgemm(int m, int repeats, double* tflops) {
  return guarded([&] {
    SPB_REQUIRE(m >= 1 && repeats >= 1 && tflops, "invalid argument");
    DeviceCtx& ctx = DeviceCtx::get();
    cudaStream_t s = ctx.stream;
    const size_t mm = (size_t)m * m;
    DevBuf<double> A(mm), B(mm), C(mm);
    launch_set_identity(A.get(), m, m, m, s);
    launch_set_identity(B.get(), m, m, m, s);
    const double one = 1.0, zero = 0.0;
    cudaEvent_t e0, e1;
    SPB_CUDA(cudaEventCreate(&e0));
    SPB_CUDA(cudaEventCreate(&e1));
    double best = 1e300;
    for (int r = 0; r < repeats + 1; ++r) {
      SPB_CUDA(cudaEventRecord(e0, s));
      SPB_CUBLAS(cublasDgemm(ctx.blas, CUBLAS_OP_N, CUBLAS_OP_N, m, m, m, &one, A.get(), m, B.get(), m, &zero,
                             C.get(), m));
      SPB_CUDA(cudaEventRecord(e1, s));
      SPB_CUDA(cudaStreamSynchronize(s));
      float ms = 0.f;
      SPB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      if (r > 0) best = std::min(best, (double)ms);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tflops = 2.0 * (double)m * m * m / (best * 1e-3) / 1e12;
  });
}

int spb_bench_inverse(int p, int repeats, double* ms_out) { return spb_bench_factor(p, repeats, 1, ms_out); }

int spb_bench_factor(int p, int repeats, int blocks, double* ms_out) {
  return guarded([&] {
    SPB_REQUIRE(p >= 1 && repeats >= 1 && ms_out, "invalid argument");
    SPB_REQUIRE(blocks >= 0 && blocks <= kMaxBlocks, "blocks must be in 0..kMaxBlocks");
    const Blocking blk = blocking_for(p, blocks);
    DeviceCtx& ctx = DeviceCtx::get();
    cudaStream_t s = ctx.stream;
    const size_t ld = padded_ld(p);
    DevBuf<double> om(ld * (size_t)p), g(ld * (size_t)p), w(ld * (size_t)p);
    SPB_CUDA(cudaMemsetAsync(om.get(), 0, sizeof(double) * ld * p, s));
    SPB_CUDA(cudaMemsetAsync(g.get(), 0, sizeof(double) * ld * p, s));
    launch_set_identity(om.get(), p, p, ld, s);
    cudaEvent_t e0, e1;
    SPB_CUDA(cudaEventCreate(&e0));
    SPB_CUDA(cudaEventCreate(&e1));
    double sum = 0.0;
    for (int r = 0; r < repeats + 1; ++r) {
      SPB_CUDA(cudaEventRecord(e0, s));
      Engine::spd_inverse(ctx.exec(), om.get(), ld, g.get(), ld, p, 0.5, 10.0, 2.0, w.get(), ld, blk);
      SPB_CUDA(cudaEventRecord(e1, s));
      ctx.sync_and_check();
      float ms = 0.f;
      SPB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      if (r > 0) sum += (double)ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *ms_out = sum / repeats;
  });
}

}  // extern "C"
