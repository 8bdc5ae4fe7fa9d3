// Fold-batched matrix-free spatpcaCore2 on the TMA-fed DMMA stream (tma_dmma.cuh).  Step 1 of this file: the
// stream itself as a stand-alone product  U = Omega_u V  (kernel-level parity test and roofline probe).
#include <cudaTypedefs.h>

#include <cmath>
#include <cstdlib>
#include <map>
#include <mutex>

#include "engine.cuh"
#include "tma_dmma.cuh"

namespace spb {

namespace {

using namespace tmx;

PFN_cuTensorMapEncodeTiled_v12000 tensor_map_encoder() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess || f == nullptr) {
      cudaGetLastError();
      throw Error(SPB_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    }
    return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(f);
  }();
  return fn;
}

// column-major FP64 matrix (rows x cols, leading dimension ld) -> 2-D tensor map with boxes of 16 rows x box_cols
CUtensorMap make_map(const double* base, int rows, int cols, size_t ld, int box_cols) {
  CUtensorMap m;
  const cuuint64_t gdim[2] = {(cuuint64_t)rows, (cuuint64_t)cols};
  const cuuint64_t gstr[1] = {(cuuint64_t)ld * sizeof(double)};
  const cuuint32_t box[2] = {16u, (cuuint32_t)box_cols};
  const cuuint32_t estr[2] = {1u, 1u};
  const CUresult r = tensor_map_encoder()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), gdim, gstr,
                                          box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw Error(SPB_ERR_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
  return m;
}

int pick_mt(int nc) { return nc <= 8 ? 1 : (nc <= 16 ? 2 : 4); }

struct HostGeom {
  StreamGeom g;
  int MT, NCP, nblk;
  size_t smem_stream;     // stage rings + barriers
};

// NT: the smallest column-tile count with ceil(p / (8 NT)) <= SMs; stages: as many as fit beside `other_smem`
HostGeom make_stream_geom(int p, int nc, int sms, size_t other_smem, size_t smem_cap) {
  HostGeom h;
  h.MT = pick_mt(nc);
  h.NCP = 8 * h.MT;
  const int CR = 64;
  int NT = ceil_div(ceil_div(p, 8), sms);
  if (NT < 1) NT = 1;
  SPB_REQUIRE(NT <= (h.MT == 4 ? 15 : 16), "p is too large for the DMMA stream on this device");
  h.g.p = p; h.g.NT = NT; h.g.J = 8 * NT;
  h.nblk = ceil_div(p, h.g.J);
  h.g.nchunks = ceil_div(p, CR);
  h.g.stageA = (uint32_t)(CR * h.g.J * 8);
  h.g.stageV = (uint32_t)(CR * h.NCP * 8);
  const size_t per_stage = (size_t)h.g.stageA + h.g.stageV;
  const size_t fixed = other_smem + 2048 /* alignment slack + barriers */;
  int ns = (int)((smem_cap - fixed) / per_stage);
  if (ns > 12) ns = 12;
  SPB_REQUIRE(ns >= 3, "not enough shared memory for the DMMA stream pipeline");
  h.g.NS = ns;
  h.g.D = ns - 1;
  if (const char* e = std::getenv("SPB_TMA_AHEAD")) { const int d = std::atoi(e); if (d >= 1 && d <= ns - 1) h.g.D = d; }
  h.smem_stream = (size_t)ns * per_stage;
  return h;
}

struct MvParams {
  StreamGeom g;
  int nc;
  double* U;        // p x nc, ld = p
};

template <int MT, int NTM>
__global__ void __launch_bounds__(tmx::Shape<MT>::kThreads, 1)
dmma_matvec_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapV, MvParams P) {
  using S = Shape<MT>;
  extern __shared__ unsigned char dyn_raw[];
  unsigned char* dyn = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(dyn_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sA = dyn;
  unsigned char* sV = sA + (size_t)P.g.NS * P.g.stageA;
  uint64_t* full = reinterpret_cast<uint64_t*>(sV + (size_t)P.g.NS * P.g.stageV);
  uint64_t* empty = full + P.g.NS;
  double* us = reinterpret_cast<double*>(sA);          // aliases the ring: written after the last chunk is consumed
  if (threadIdx.x == 0) {
    for (int s = 0; s < P.g.NS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], S::kConsumers); }
    fence_barrier_init();
  }
  __syncthreads();
  unsigned int it = 0;
  const int j0 = blockIdx.x * P.g.J;
  dmma_stream_pass<MT, NTM>(&mapA, &mapV, P.g, j0, sA, sV, full, empty, it, us);
  for (int e = threadIdx.x; e < P.g.J * P.nc; e += S::kThreads) {
    const int j = e % P.g.J, v = e / P.g.J;
    if (j0 + j < P.g.p) P.U[(size_t)v * P.g.p + j0 + j] = us[(size_t)v * P.g.J + j];
  }

}

size_t smem_cap_for_device() {
  int dev = 0, cap = 0;
  SPB_CUDA(cudaGetDevice(&dev));
  SPB_CUDA(cudaDeviceGetAttribute(&cap, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  return (size_t)cap;
}

int sm_count() {
  int dev = 0, sms = 0;
  SPB_CUDA(cudaGetDevice(&dev));
  SPB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  return sms;
}

template <int MT, int NTM>
void launch_matvec_instance2(const CUtensorMap& mA, const CUtensorMap& mV, const MvParams& P, int nblk, size_t smem,
                             cudaStream_t s) {
  SPB_CUDA(cudaFuncSetAttribute(dmma_matvec_kernel<MT, NTM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dmma_matvec_kernel<MT, NTM><<<nblk, tmx::Shape<MT>::kThreads, smem, s>>>(mA, mV, P);
  count_launch();
  SPB_CUDA(cudaGetLastError());
}
template <int MT>
void launch_matvec_instance(const CUtensorMap& mA, const CUtensorMap& mV, const MvParams& P, int nblk, size_t smem,
                            cudaStream_t s) {
  if constexpr (MT == 4) {     // NTM = column tiles per warp set (3 sets)
    if (P.g.NT <= 9) launch_matvec_instance2<MT, 3>(mA, mV, P, nblk, smem, s);
    else launch_matvec_instance2<MT, 5>(mA, mV, P, nblk, smem, s);
  } else {
    if (P.g.NT <= 10) launch_matvec_instance2<MT, 10>(mA, mV, P, nblk, smem, s);
    else launch_matvec_instance2<MT, 16>(mA, mV, P, nblk, smem, s);
  }
}

// =============================================================================================
// Fold-batched matrix-free spatpcaCore2 (reference src/RcppSpatPCA.cpp:150-222, called along the tau1 path of
// :329-332 for every fold, and cold at :393 / :475 / :599): F problems (the folds of a stage, or the full-data
// fit) advance in lock-step through one persistent cooperative kernel.  Every conjugate-gradient step streams
// Omega_u ONCE for all F K <= 32 columns on the TMA-fed DMMA stream above; problem f solves with
//     A_f = c Omega_u + rho_f I - c' Ytr_f' Ytr_f,     Ytr_f = the rows of Y outside validation fold f,
// the rank-n part applied through Y and its transpose (both L2-resident) with the fold's rows masked out of
// S = Y r.  CTA b owns rows [b J, (b + 1) J) of every p x K vector for the whole kernel; sums over CTAs go through
// per-CTA partials + a grid barrier and are added in CTA order -- the small ones redundantly by every CTA (so all
// CTAs take the same decisions without a broadcast), S = Y r in slices.  A whole tau1 path (all folds, all grid
// values, warm starts carried) is ONE launch; the hold-out losses are taken afterwards from per-step snapshots.
// =============================================================================================
constexpr int kMaxF = 8;
constexpr int kMaxSteps = 64;

struct BcgParams {
  StreamGeom g;                    // the Omega_u stream
  StreamGeom gy;                   // the same column block of Y (n rows): z = Y[:, J]' S on the same DMMA pipeline
  int p, n, K, F, NC;
  size_t ldv;                      // leading dimension of the p x NCP work vectors
  int nblk;
  int nsteps;
  int mode;                        // 2: spatpcaCore2 / 2p (steps over tau1); 3: spatpcaCore3 (steps over tau2, tau1 fixed)
  double tau1[kMaxSteps];
  double tau2[kMaxSteps];          // mode 3
  double c_gram;
  int maxit;
  double tol, cg_tol2;
  int cg_cap;
  int refresh;                     // recompute the true residual b - A x every `refresh` solves on an unchanged matrix (0: always)
  int fold_id[kMaxF];              // value of nk[] that marks problem f's validation rows (0: none, full data)
  double rho[kMaxF];
  const int* nk;                   // n fold ids
  const double* Y;                 // n x p (ld n)
  const double* Yt;                // p x n (ld p)
  double* Phi[kMaxF];              // per problem p x K (ld p): in = warm start, out = last iterate
  double* C[kMaxF];
  double* L2[kMaxF];
  double* R[kMaxF];                // mode 3
  double* L1[kMaxF];               // mode 3
  double* snap;                    // nsteps x F x (p x K) iterates after every step (may be nullptr)
  int* iters_out;                  // nsteps x F
  int* status_out;                 // [0] status, [1] matrix passes
  double *Xg, *Rg, *Pg, *Qg, *Bg, *Cg, *L2g;
  double *Radg, *L1g;              // mode 3: the ADMM variables R and Lambda1 (Rg is the CG residual)
  double *part_ru, *part_gam, *Spart, *Sred, *gram_part, *norm_part;
  unsigned int* bar;
  unsigned long long* dbg;         // optional globaltimer stamps of CTA 0 (performance debugging), or nullptr
};

__device__ __forceinline__ unsigned long long gtimer_b() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ void grid_barrier_b(unsigned int* counter, unsigned int& gen) {
  __syncthreads();
  if (threadIdx.x == 0) {
    ++gen;
    const unsigned int target = gen * gridDim.x;
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned int v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    } while ((int)(v - target) < 0);
    __threadfence();
  }
  __syncthreads();
}

__device__ __forceinline__ double div_rho_b(double a, double rho, double rinv) {   // correctly rounded a / rho
  const double q = __dmul_rn(a, rinv);
  const double r = fma(-q, rho, a);
  return fma(r, rinv, q);
}

// K x K inverse square root in one warp: coupled Newton-Schulz (as admm_kernels.cu), serial Jacobi as the fallback.
// G, Y, Z, M, Pout: K*K doubles each in shared memory.  Returns 0, or 1 if G is not numerically positive definite.
__device__ int warp_inv_sqrt(const double* G, double* Y, double* Z, double* M, double* Pout, int K) {
  const int lane = threadIdx.x & 31;
  const int KK = K * K;
  double sc = 0.0;
  for (int i = 0; i < K; ++i) {
    double r = 0.0;
    for (int j = 0; j < K; ++j) r += fabs(G[i * K + j]);
    sc = fmax(sc, r);
  }
  bool ok = (sc > 0.0) && (sc < 1e300);
  if (ok) {
    const double sinv = fast_rcp(sc);
    for (int e = lane; e < KK; e += 32) { Y[e] = G[e] * sinv; Z[e] = ((e / K) == (e % K)) ? 1.0 : 0.0; }
    __syncwarp();
    ok = false;
    for (int it = 0; it < 24; ++it) {
      double emax = 0.0;
      for (int e = lane; e < KK; e += 32) {
        const int a = e / K, b = e - a * K;
        double acc = 0.0;
        for (int k = 0; k < K; ++k) acc = fma(Z[a * K + k], Y[k * K + b], acc);
        const double m = ((a == b) ? 3.0 : 0.0) - acc;
        emax = fmax(emax, fabs(m - ((a == b) ? 2.0 : 0.0)));
        M[e] = m;
      }
      const bool big = __any_sync(0xffffffffu, !(emax < 3e-9));
      __syncwarp();
      double yn[2], zn[2];
      int q = 0;
      for (int e = lane; e < KK; e += 32, ++q) {
        const int a = e / K, b = e - a * K;
        double ay = 0.0, az = 0.0;
        for (int k = 0; k < K; ++k) {
          ay = fma(Y[a * K + k], M[k * K + b], ay);
          az = fma(M[a * K + k], Z[k * K + b], az);
        }
        yn[q] = 0.5 * ay; zn[q] = 0.5 * az;
      }
      __syncwarp();
      q = 0;
      for (int e = lane; e < KK; e += 32, ++q) { Y[e] = yn[q]; Z[e] = zn[q]; }
      __syncwarp();
      if (!big) { ok = true; break; }
    }
    if (ok) {
      const double rs = fast_rsqrt(sc);
      for (int e = lane; e < KK; e += 32) {
        const int a = e / K, b = e - a * K;
        Pout[e] = (0.5 * (Z[a * K + b] + Z[b * K + a])) * rs;
      }
      __syncwarp();
      return 0;
    }
  }
  // fallback: cyclic Jacobi by lane 0 (ill-conditioned or indefinite G: rare, and classified here)
  int bad = 0;
  if (lane == 0) {
    double* A = Y;       // working copy
    double* V = Z;
    for (int e = 0; e < KK; ++e) { A[e] = G[e]; V[e] = ((e / K) == (e % K)) ? 1.0 : 0.0; }
    for (int sweep = 0; sweep < 64; ++sweep) {
      bool rotated = false;
      for (int pp = 0; pp < K - 1; ++pp)
        for (int qq = pp + 1; qq < K; ++qq) {
          const double apq = A[pp * K + qq], app = A[pp * K + pp], aqq = A[qq * K + qq];
          if (apq == 0.0 || apq * apq <= 1e-32 * fabs(app * aqq)) continue;
          rotated = true;
          const double theta = (aqq - app) / (2.0 * apq);
          double t = 1.0 / (fabs(theta) + sqrt(theta * theta + 1.0));
          if (theta < 0.0) t = -t;
          const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c, tau = sn / (1.0 + c);
          for (int i = 0; i < K; ++i) {
            if (i != pp && i != qq) {
              const double aip = A[i * K + pp], aiq = A[i * K + qq];
              const double nip = aip - sn * (aiq + tau * aip), niq = aiq + sn * (aip - tau * aiq);
              A[i * K + pp] = nip; A[pp * K + i] = nip;
              A[i * K + qq] = niq; A[qq * K + i] = niq;
            }
            const double vip = V[i * K + pp], viq = V[i * K + qq];
            V[i * K + pp] = vip - sn * (viq + tau * vip);
            V[i * K + qq] = viq + sn * (vip - tau * viq);
          }
          A[pp * K + pp] = app - t * apq;
          A[qq * K + qq] = aqq + t * apq;
          A[pp * K + qq] = 0.0; A[qq * K + pp] = 0.0;
        }
      if (!rotated) break;
    }
    double lmax = 0.0;
    for (int i = 0; i < K; ++i) lmax = fmax(lmax, A[i * K + i]);
    for (int i = 0; i < K; ++i) {
      const double lam = A[i * K + i];
      if (!(lam > 1e-28 * lmax && lam < 1e300)) bad = 1;
      M[i] = 1.0 / sqrt(lam);
    }
    for (int a = 0; a < K; ++a)
      for (int b = 0; b < K; ++b) {
        double acc = 0.0;
        for (int k = 0; k < K; ++k) acc = fma(V[a * K + k] * M[k], V[b * K + k], acc);
        Pout[a * K + b] = acc;
      }
  }
  bad = __shfl_sync(0xffffffffu, bad, 0);
  __syncwarp();
  return bad;
}

// sum over CTAs of per-CTA partials part[b * stride + e], e < count, in CTA order: warp per entry, lanes over b
__device__ __forceinline__ void reduce_partials(const double* part, int nblk, int stride, int count, double* out,
                                                int nthreads) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = nthreads >> 5;
  for (int e = warp; e < count; e += nw) {
    double t = 0.0;
    for (int b = lane; b < nblk; b += 32) t += __ldcg(part + (size_t)b * stride + e);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) out[e] = t;
  }
}

// column sums of tmp[c * J + j], j < Jv: warp per column, lanes over j, fixed xor tree
__device__ __forceinline__ void column_sums(const double* tmp, int J, int Jv, int ncols, double* out, int nthreads) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = nthreads >> 5;
  for (int c = warp; c < ncols; c += nw) {
    double t = 0.0;
    for (int j = lane; j < Jv; j += 32) t += tmp[(size_t)c * J + j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) out[c] = t;
  }
}

struct BcgState {                 // identical in every CTA (computed redundantly from the reduced partials)
  double gam[32], gam_old[32], alpha[32], alpha_old[32], beta[32], bn[32], ss[32], ru[32];
  double rho[kMaxF], rinv[kMaxF], err[kMaxF];
  double rsys[kMaxF], thr[kMaxF];  // diagonal shift of the system (rho, or 2 rho in mode 3: Phi = (2 tau1 Omega - 2 G + 2 rho I)^-1 b); tau2 / rho
  double red[64];
  int frozen[32], fresh[32];
  int active[kMaxF], its[kMaxF];
  int cg_done, status, cg_its;
};

template <int MT, int NTM>
__global__ void __launch_bounds__(tmx::Shape<MT>::kThreads, 1)
batched_core2_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapX,
                     const __grid_constant__ CUtensorMap mapR, const __grid_constant__ CUtensorMap mapY,
                     const __grid_constant__ CUtensorMap mapS, const __grid_constant__ BcgParams P) {
  using S = Shape<MT>;
  constexpr int NTH = S::kThreads;
  extern __shared__ unsigned char dyn_raw[];
  unsigned char* dyn = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(dyn_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sA = dyn;
  unsigned char* sV = sA + (size_t)P.g.NS * P.g.stageA;
  unsigned char* ring_end = sV + (size_t)P.g.NS * P.g.stageV;
  uint64_t* full = reinterpret_cast<uint64_t*>(ring_end);
  uint64_t* empty = full + P.g.NS;
  BcgState& st = *reinterpret_cast<BcgState*>(empty + P.g.NS);
  const int J = P.g.J, NC = P.NC, K = P.K, F = P.F, n = P.n, p = P.p;
  // U = Omega_u v must survive the DMMA mini-pass that forms z, so it has its own tile; everything else that the row
  // phases use aliases the ring (no TMA is in flight while it is live)
  double* us = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(&st) + ((sizeof(BcgState) + 15) & ~size_t(15)));  // [NCP][J]
  double* zsm = reinterpret_cast<double*>(sA);                   // [NCP][J]  Y[:, J]' S (output of the mini-pass)
  double* vsm = zsm + (size_t)S::NCP * J;                        // [NC][J]   rows of the vector whose S partial is formed
  double* tmp0 = vsm + (size_t)NC * J;                           // [NC][J]
  double* tmp1 = tmp0 + (size_t)NC * J;                          // [NC][J]
  double* Pm = tmp1 + (size_t)NC * J;                            // [F][64]     (T'T)^(-1/2) per problem
  double* polw = Pm + (size_t)kMaxF * 64;                        // [F][4][64]  scratch of the polar step: G, Y, Z, M
  const int tid = threadIdx.x;
  const int j0 = blockIdx.x * J;
  const int Jv = min(J, p - j0);
  const size_t ldv = P.ldv;
  const double rdenom = rsqrt((double)p * (double)K);

  if (tid == 0) {
    for (int s = 0; s < P.g.NS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], S::kConsumers); }
    fence_barrier_init();
    st.status = 0; st.cg_done = 0; st.cg_its = 0;
  }
  if (tid < kMaxF) {
    st.rho[tid] = tid < F ? P.rho[tid] : 1.0;
    st.rinv[tid] = tid < F ? 1.0 / P.rho[tid] : 1.0;
    st.rsys[tid] = (P.mode == 3 ? 2.0 : 1.0) * st.rho[tid];
    st.thr[tid] = 0.0;
    st.active[tid] = 0; st.its[tid] = 0; st.err[tid] = 0.0;
  }
  __syncthreads();
  // gather this CTA's rows of the per-problem iterates into the p x NC work vectors; b = rho C - Lambda2 (:168)
  for (int e = tid; e < Jv * NC; e += NTH) {
    const int j = e % Jv, c = e / Jv, f = c / K, k = c - f * K;
    const size_t src = (size_t)k * p + j0 + j, dst = (size_t)c * ldv + j0 + j;
    const double cv = P.C[f][src], l2 = P.L2[f][src];
    P.Xg[dst] = P.Phi[f][src];
    P.Cg[dst] = cv;
    P.L2g[dst] = l2;
    if (P.mode == 3) {                                   // b = rho (R + C) - (Lambda1 + Lambda2) (:247)
      const double rv = P.R[f][src], l1 = P.L1[f][src];
      P.Radg[dst] = rv;
      P.L1g[dst] = l1;
      P.Bg[dst] = __dsub_rn(__dmul_rn(st.rho[f], __dadd_rn(rv, cv)), __dadd_rn(l1, l2));
    } else {
      P.Bg[dst] = __dsub_rn(__dmul_rn(st.rho[f], cv), l2);
    }
    P.Rg[dst] = 0.0;
  }
  unsigned int it = 0, bar_gen = 0;
  int passes = 0, dbg_n = 0;
  auto stamp = [&]() {
    if (P.dbg != nullptr && blockIdx.x == 0 && tid == 0 && dbg_n < 400) P.dbg[dbg_n] = gtimer_b();
    ++dbg_n;
  };
  fence_proxy_async();
  grid_barrier_b(P.bar, bar_gen);

  // S partial of vsm (this CTA's rows of a p x NC vector): Spart[cta][c][r] = sum_j Y[r, j0 + j] v[j, c]
  auto s_partial = [&]() {
    double* out = P.Spart + (size_t)blockIdx.x * n * NC;
    for (int e = tid; e < n * F; e += NTH) {
      const int r = e % n, f = e / n;
      double acc[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = 0.0;
      const double* yp = P.Y + (size_t)j0 * n + r;
      const double* vp = vsm + (size_t)f * K * J;
      int j = 0;
      for (; j + 8 <= Jv; j += 8) {                     // 8 independent L2 loads in flight per thread
        double y[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) y[u] = __ldg(yp + (size_t)(j + u) * n);
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (k < K) acc[k] = fma(y[u], vp[(size_t)k * J + j + u], acc[k]);
      }
      for (; j < Jv; ++j) {
        const double y = __ldg(yp + (size_t)j * n);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (k < K) acc[k] = fma(y, vp[(size_t)k * J + j], acc[k]);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (k < K) __stcg(out + (size_t)(f * K + k) * n + r, acc[k]);
    }
  };
  // this CTA's slice of S = sum over CTAs of Spart, validation rows of each problem masked out -> Sred
  auto s_slice_reduce = [&]() {
    const int total = n * NC;
    const int per = (total + (int)gridDim.x - 1) / (int)gridDim.x;
    const int e0 = blockIdx.x * per, e1 = min(total, e0 + per);
    const int warp = tid >> 5, lane = tid & 31;
    for (int e = e0 + warp; e < e1; e += NTH / 32) {             // warp per entry, lanes over the CTAs' partials
      double t = 0.0;
      for (int b = lane; b < P.nblk; b += 32) t += __ldcg(P.Spart + (size_t)b * total + e);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      if (lane == 0) {
        const int c = e / n, r = e - c * n;
        const int fid = P.fold_id[c / K];
        if (fid != 0 && __ldg(P.nk + r) == fid) t = 0.0;
        __stcg(P.Sred + e, t);
      }
    }
  };
  // |S_c|^2 from the reduced S, and z = Y[:, J]' S on the DMMA pipeline: the same pass as the Omega_u stream with
  // Y (n x p) as the matrix and S (n x NC) as the vectors -- the p x n contraction of the north-star
  auto load_s_and_z = [&]() {
    {
      const int warp = tid >> 5, lane = tid & 31;
      for (int c = warp; c < NC; c += NTH / 32) {
        double t = 0.0;
        for (int r = lane; r < n; r += 32) { const double v = __ldcg(P.Sred + (size_t)c * n + r); t = fma(v, v, t); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) st.ss[c] = t;
      }
    }
    dmma_stream_pass<MT, NTM>(&mapY, &mapS, P.gy, j0, sA, sV, full, empty, it, zsm);
  };

  // Residual recycling: after a solve r = b - A x is known, and the next solve on the SAME matrix starts from
  // r + (b_new - b_old) -- the update phase applies that difference -- instead of spending a matrix pass (plus the
  // S / z phases and two grid barriers) on A x.  A warm-started solve takes ~1.2 CG steps, so this removes ~45 % of
  // the passes of a tau2 path.  The recursion drifts from the true residual by ~eps * cond per update; it is
  // recomputed every P.refresh solves and whenever the matrix changes (a new tau1).
  bool have_r = false;
  int since_refresh = 0;
  double c_prev = 0.0;
  for (int step = 0; step < P.nsteps; ++step) {
    const double c_omega = 2.0 * P.tau1[step];
    if (tid < F) { st.active[tid] = 1; st.its[tid] = 0; st.thr[tid] = (P.mode == 3) ? P.tau2[step] / st.rho[tid] : 0.0; }
    __syncthreads();
    for (int admm = 0; admm < P.maxit; ++admm) {
      int any = 0;
      for (int f = 0; f < F; ++f) any |= st.active[f];
      if (!any || st.status != 0) break;
      // ================= CG solve  A_f Phi_f = b_f  for the active problems, warm start =================
      const bool recycle = have_r && P.refresh > 0 && since_refresh < P.refresh && c_omega == c_prev;
      if (!recycle) {
        for (int e = tid; e < Jv * NC; e += NTH) {
          const int j = e % Jv, c = e / Jv;
          vsm[(size_t)c * J + j] = P.Xg[(size_t)c * ldv + j0 + j];
        }
        __syncthreads();
        s_partial();
        grid_barrier_b(P.bar, bar_gen);
        s_slice_reduce();
        dmma_stream_pass<MT, NTM>(&mapA, &mapX, P.g, j0, sA, sV, full, empty, it, us);    // U = Omega_u Phi
        ++passes;
        // us lives outside the ring: the next grid barrier cannot be passed by a TMA write of another pass (this CTA
        // issues its own loads)
        grid_barrier_b(P.bar, bar_gen);
        load_s_and_z();                                   // keeps U: overwrites only zsm
        since_refresh = 0;
      } else {
        ++since_refresh;
      }
      c_prev = c_omega;
      {
        for (int e = tid; e < Jv * NC; e += NTH) {
          const int j = e % Jv, c = e / Jv, f = c / K;
          const size_t off = (size_t)c * ldv + j0 + j;
          double r = 0.0, b = 0.0;
          if (st.active[f]) {
            b = P.Bg[off];
            if (recycle) {
              r = P.Rg[off];                              // maintained by the update phase
            } else {
              const double x = P.Xg[off];
              const double w = fma(c_omega, us[(size_t)c * J + j], fma(st.rsys[f], x, -P.c_gram * zsm[(size_t)c * J + j]));
              r = b - w;
            }
          }
          // an inactive problem keeps its residual in Rg when recycling is on (it may be reactivated by the next step)
          if (!recycle && (st.active[f] || P.refresh <= 0)) P.Rg[off] = r;
          vsm[(size_t)c * J + j] = r;
          tmp0[(size_t)c * J + j] = r * r;
          tmp1[(size_t)c * J + j] = b * b;
        }
        __syncthreads();
        column_sums(tmp0, J, Jv, NC, P.part_gam + (size_t)blockIdx.x * 64, NTH);
        column_sums(tmp1, J, Jv, NC, P.part_gam + (size_t)blockIdx.x * 64 + 32, NTH);
        s_partial();
        fence_proxy_async();
      }
      grid_barrier_b(P.bar, bar_gen);
      reduce_partials(P.part_gam, P.nblk, 64, NC, st.gam, NTH);
      reduce_partials(P.part_gam + 32, P.nblk, 64, NC, st.bn, NTH);
      __syncthreads();
      if (tid < NC) {
        const int f = tid / K;
        st.fresh[tid] = 1; st.alpha_old[tid] = 1.0; st.gam_old[tid] = 1.0;
        st.frozen[tid] = (!st.active[f] || st.gam[tid] <= P.cg_tol2 * st.bn[tid]) ? 1 : 0;
      }
      __syncthreads();
      if (tid == 0) {
        int done = 1;
        for (int c = 0; c < NC; ++c) done &= st.frozen[c];
        st.cg_done = done; st.cg_its = 0;
      }
      s_slice_reduce();
      __syncthreads();
      while (!st.cg_done) {
        stamp();                                                                        // 0: iteration start
        dmma_stream_pass<MT, NTM>(&mapA, &mapR, P.g, j0, sA, sV, full, empty, it, us);  // U = Omega_u r
        ++passes;
        stamp();                                                                        // 1: stream done
        for (int e = tid; e < Jv * NC; e += NTH) {
          const int j = e % Jv, c = e / Jv;
          tmp0[(size_t)c * J + j] = P.Rg[(size_t)c * ldv + j0 + j] * us[(size_t)c * J + j];
        }
        __syncthreads();
        column_sums(tmp0, J, Jv, NC, P.part_ru + (size_t)blockIdx.x * 32, NTH);
        stamp();                                                                        // 2: r.u partials
        grid_barrier_b(P.bar, bar_gen);
        stamp();                                                                        // 3: barrier 1
        reduce_partials(P.part_ru, P.nblk, 32, NC, st.ru, NTH);
        load_s_and_z();                                   // also publishes st.ru (barriers inside)
        stamp();                                                                        // 4: S, z
        if (tid < NC) {
          const int c = tid, f = c / K;
          double alpha = 0.0, beta = 0.0;
          if (!st.frozen[c]) {
            const double gam = st.gam[c];
            const double delta = fma(c_omega, st.ru[c], fma(st.rsys[f], gam, -P.c_gram * st.ss[c]));     // r . A r
            double den = delta;
            if (!st.fresh[c]) {
              beta = gam * fast_rcp(st.gam_old[c]);
              den = delta - beta * gam * fast_rcp(st.alpha_old[c]);
            }
            if (!(delta > 0.0) || !(den > 0.0) || !(den < 1e300)) { st.status = SPB_ERR_NOT_POSDEF; den = 1.0; }
            alpha = gam * fast_rcp(den);
            st.alpha_old[c] = alpha;
          }
          st.alpha[c] = alpha; st.beta[c] = beta;
        }
        __syncthreads();
        for (int e = tid; e < Jv * NC; e += NTH) {
          const int j = e % Jv, c = e / Jv, f = c / K;
          const size_t off = (size_t)c * ldv + j0 + j;
          const double r = P.Rg[off];
          double rn = r;
          if (!st.frozen[c]) {
            const double w = fma(c_omega, us[(size_t)c * J + j], fma(st.rsys[f], r, -P.c_gram * zsm[(size_t)c * J + j]));
            double pn = r, qn = w;
            if (!st.fresh[c]) {
              pn = fma(st.beta[c], P.Pg[off], r);
              qn = fma(st.beta[c], P.Qg[off], w);
            }
            P.Pg[off] = pn;
            P.Qg[off] = qn;
            P.Xg[off] = fma(st.alpha[c], pn, P.Xg[off]);
            rn = fma(-st.alpha[c], qn, r);
            P.Rg[off] = rn;
          }
          vsm[(size_t)c * J + j] = rn;
          tmp0[(size_t)c * J + j] = rn * rn;
        }
        __syncthreads();
        stamp();                                                                        // 5: update
        column_sums(tmp0, J, Jv, NC, P.part_gam + (size_t)blockIdx.x * 64, NTH);
        s_partial();
        fence_proxy_async();
        stamp();                                                                        // 6: S partial
        grid_barrier_b(P.bar, bar_gen);
        stamp();                                                                        // 7: barrier 2
        reduce_partials(P.part_gam, P.nblk, 64, NC, st.red, NTH);
        __syncthreads();
        if (tid < NC && !st.frozen[tid]) {
          st.gam_old[tid] = st.gam[tid];
          st.gam[tid] = st.red[tid];
          st.fresh[tid] = 0;
          if (st.red[tid] <= P.cg_tol2 * st.bn[tid] || !(st.red[tid] < 1e300)) st.frozen[tid] = 1;
        }
        __syncthreads();
        if (tid == 0) {
          int done = 1;
          for (int c = 0; c < NC; ++c) done &= st.frozen[c];
          st.cg_its++;
          if (!done && st.cg_its >= P.cg_cap) { st.status = SPB_ERR_INTERNAL; done = 1; }
          if (st.status != 0) done = 1;
          st.cg_done = done;
        }
        s_slice_reduce();
        __syncthreads();
        stamp();                                                                        // 8: reductions
      }
      if (st.status != 0) break;
      have_r = true;
      // ================= Stiefel projection, dual update, stopping rule (:169-179) =================
      for (int e = tid; e < Jv * NC; e += NTH) {
        const int j = e % Jv, c = e / Jv, f = c / K;
        const size_t off = (size_t)c * ldv + j0 + j;
        tmp0[(size_t)c * J + j] = __dadd_rn(P.Xg[off], div_rho_b(P.L2g[off], st.rho[f], st.rinv[f]));   // T = Phi + Lambda2 / rho
      }
      __syncthreads();
      {
        const int KG = K * (K + 1) / 2;
        const int warp = tid >> 5, lane = tid & 31;
        for (int e = warp; e < F * KG; e += NTH / 32) {
          const int f = e / KG;
          int q = e - f * KG, a = 0;
          while (q >= K - a) { q -= K - a; ++a; }
          const int b = a + q;
          const double* ta = tmp0 + (size_t)(f * K + a) * J;
          const double* tb = tmp0 + (size_t)(f * K + b) * J;
          double t = 0.0;
          for (int j = lane; j < Jv; j += 32) t = fma(ta[j], tb[j], t);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
          if (lane == 0) __stcg(P.gram_part + (size_t)blockIdx.x * (kMaxF * 36) + e, t);
        }
      }
      grid_barrier_b(P.bar, bar_gen);
      {
        const int KG = K * (K + 1) / 2;
        double* gfull = tmp1;                                          // F * KG sums
        reduce_partials(P.gram_part, P.nblk, kMaxF * 36, F * KG, gfull, NTH);
        __syncthreads();
        const int warp = tid >> 5, lane = tid & 31;
        if (warp < F && st.active[warp]) {
          const int f = warp;
          double* G = polw + (size_t)f * 256;
          for (int e = lane; e < K * K; e += 32) {
            int a = e / K, b = e - a * K;
            if (a > b) { const int t2 = a; a = b; b = t2; }
            int idx = 0;
            for (int aa = 0; aa < a; ++aa) idx += K - aa;
            G[e] = gfull[f * KG + idx + (b - a)];
          }
          __syncwarp();
          const int bad = warp_inv_sqrt(G, G + 64, G + 128, G + 192, Pm + (size_t)f * 64, K);
          if (bad && lane == 0) st.status = 1;
        }
        __syncthreads();
      }
      for (int e = tid; e < Jv * NC; e += NTH) {
        const int j = e % Jv, c = e / Jv, f = c / K, k = c - f * K;
        double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
        if (st.active[f]) {
          const size_t off = (size_t)c * ldv + j0 + j;
          double cn = 0.0;
          for (int a = 0; a < K; ++a) cn = fma(tmp0[(size_t)(f * K + a) * J + j], Pm[(size_t)f * 64 + a * K + k], cn);
          const double x = P.Xg[off], cold = P.Cg[off], l2old = P.L2g[off];
          const double dC = __dsub_rn(x, cn);
          const double l2n = __dadd_rn(l2old, __dmul_rn(st.rho[f], dC));          // :173, :258
          const double ec = __dsub_rn(cn, cold);
          P.Cg[off] = cn;
          P.L2g[off] = l2n;
          d0 = dC * dC; d1 = ec * ec;
          if (P.mode == 3) {
            const double rold = P.Radg[off], l1old = P.L1g[off];
            const double a1 = __dadd_rn(div_rho_b(l1old, st.rho[f], st.rinv[f]), x);                    // :248
            const double mag = fmax(0.0, __dsub_rn(fabs(a1), st.thr[f]));
            const double sg = (a1 > 0.0) ? 1.0 : ((a1 < 0.0) ? -1.0 : 0.0);
            const double rn = sg * mag;                                                                  // :249
            const double dR = __dsub_rn(x, rn);
            const double l1n = __dadd_rn(l1old, __dmul_rn(st.rho[f], dR));                               // :257
            const double er = __dsub_rn(rn, rold);
            P.Radg[off] = rn;
            P.L1g[off] = l1n;
            const double bnew = __dsub_rn(__dmul_rn(st.rho[f], __dadd_rn(rn, cn)), __dadd_rn(l1n, l2n));
            P.Rg[off] = P.Rg[off] + (bnew - P.Bg[off]);                                                  // residual of the next solve
            P.Bg[off] = bnew;
            d2 = dR * dR; d3 = er * er;
          } else {
            const double bnew = __dsub_rn(__dmul_rn(st.rho[f], cn), l2n);
            P.Rg[off] = P.Rg[off] + (bnew - P.Bg[off]);
            P.Bg[off] = bnew;
          }
        }
        vsm[(size_t)c * J + j] = d0;
        zsm[(size_t)c * J + j] = d1;
        if (P.mode == 3) { us[(size_t)c * J + j] = d2; tmp1[(size_t)c * J + j] = d3; }
      }
      __syncthreads();
      {
        // per problem: sums over its K columns and this CTA's rows, fixed order (warp per (problem, norm))
        const int warp = tid >> 5, lane = tid & 31;
        const int NN = (P.mode == 3) ? 4 : 2;                // norms per problem (:175-176; :261-264)
        for (int e = warp; e < NN * F; e += NTH / 32) {
          const int f = e / NN, v = e - f * NN;
          const double* src = v == 0 ? vsm : (v == 1 ? zsm : (v == 2 ? us : tmp1));
          double t = 0.0;
          for (int k = 0; k < K; ++k)
            for (int j = lane; j < Jv; j += 32) t += src[(size_t)(f * K + k) * J + j];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
          if (lane == 0) __stcg(P.norm_part + (size_t)blockIdx.x * (4 * kMaxF) + e, t);
        }
      }
      grid_barrier_b(P.bar, bar_gen);
      const int NN = (P.mode == 3) ? 4 : 2;
      reduce_partials(P.norm_part, P.nblk, 4 * kMaxF, NN * F, st.red, NTH);
      __syncthreads();
      if (tid < F && st.active[tid]) {
        double emax = 0.0;
        for (int v = 0; v < NN; ++v) {
          const double a = st.red[NN * tid + v];
          emax = fmax(emax, (a > 0.0) ? (a * fast_rsqrt(a)) * rdenom : 0.0);
        }
        st.err[tid] = emax;
        st.its[tid] = st.its[tid] + 1;
        if (st.err[tid] <= P.tol) st.active[tid] = 0;                          // stopping rule (:178)
      }
      __syncthreads();
    }
    // ---- end of the step: iteration counts, snapshot of the iterates (the hold-out loss is taken from it) ----
    if (blockIdx.x == 0 && tid < F) P.iters_out[step * F + tid] = st.its[tid];
    if (P.snap != nullptr) {
      for (int e = tid; e < Jv * NC; e += NTH) {
        const int j = e % Jv, c = e / Jv, f = c / K, k = c - f * K;
        P.snap[((size_t)(step * F + f) * K + k) * p + j0 + j] = P.Xg[(size_t)c * ldv + j0 + j];
      }
    }
    __syncthreads();
    if (st.status != 0) break;
  }
  // scatter the final iterates back to the per-problem arrays
  for (int e = tid; e < Jv * NC; e += NTH) {
    const int j = e % Jv, c = e / Jv, f = c / K, k = c - f * K;
    const size_t dst = (size_t)k * p + j0 + j, src = (size_t)c * ldv + j0 + j;
    P.Phi[f][dst] = P.Xg[src];
    P.C[f][dst] = P.Cg[src];
    P.L2[f][dst] = P.L2g[src];
    if (P.mode == 3) { P.R[f][dst] = P.Radg[src]; P.L1[f][dst] = P.L1g[src]; }
  }
  if (blockIdx.x == 0 && tid == 0) { P.status_out[0] = st.status; P.status_out[1] = passes; }
}

}  // namespace

// U (p x nc, ld = p) = A (p x p symmetric, full storage, ld) * V (p x nc, ldv); all device pointers; nc <= 32
void launch_dmma_matvec(const double* A, size_t ld, int p, const double* V, size_t ldv, int nc, double* U,
                        cudaStream_t s) {
  SPB_REQUIRE(nc >= 1 && nc <= 32, "the DMMA stream takes 1..32 vectors");
  SPB_REQUIRE(ld % 2 == 0 && ldv % 2 == 0, "leading dimensions must be multiples of 2 (16-byte TMA strides)");
  const size_t cap = smem_cap_for_device();
  HostGeom h = make_stream_geom(p, nc, sm_count(), 16 * 12 * 2, cap);
  const size_t smem = 1024 + h.smem_stream + (size_t)2 * h.g.NS * sizeof(uint64_t);
  SPB_REQUIRE(smem <= cap, "DMMA stream: shared memory request too large");
  const CUtensorMap mA = make_map(A, p, p, ld, h.g.J);
  const CUtensorMap mV = make_map(V, p, h.NCP <= nc ? h.NCP : nc, ldv, h.NCP);
  MvParams P;
  P.g = h.g; P.nc = nc; P.U = U;
  switch (h.MT) {
    case 1: launch_matvec_instance<1>(mA, mV, P, h.nblk, smem, s); break;
    case 2: launch_matvec_instance<2>(mA, mV, P, h.nblk, smem, s); break;
    default: launch_matvec_instance<4>(mA, mV, P, h.nblk, smem, s); break;
  }
}

namespace {

struct BcgHost {
  HostGeom h;
  size_t smem, alias_need, ring;
  size_t vec, part_ru, part_gam, spart, sred, gram, norm;      // doubles
};

bool bcg_geometry(int p, int n, int K, int F, BcgHost& o) {
  const int NC = K * F;
  if (K < 1 || K > 8 || F < 1 || F > kMaxF || NC > 32) return false;
  if (n % 2 != 0) return false;              // TMA strides are multiples of 16 bytes: Y (ld = n) needs an even n
  const int sms = sm_count();
  const int MT = pick_mt(NC);
  const int NT = ceil_div(ceil_div(p, 8), sms);
  if (NT > (MT == 4 ? 15 : 16)) return false;
  const size_t cap = smem_cap_for_device();
  const int NTg = ceil_div(ceil_div(p, 8), sms);
  const size_t us_bytes = (size_t)8 * MT * 8 * NTg * sizeof(double);
  const size_t state = sizeof(BcgState) + 16 + us_bytes + 2 * 16 * sizeof(uint64_t) + 1024;
  try {
    o.h = make_stream_geom(p, NC, sms, state, cap);
  } catch (const Error&) {
    return false;
  }
  o.ring = o.h.smem_stream;
  o.alias_need = ((size_t)o.h.NCP * o.h.g.J + 3 * (size_t)NC * o.h.g.J + (size_t)kMaxF * 320) * sizeof(double);
  if (o.alias_need > o.ring) return false;
  o.smem = 1024 + o.ring + (size_t)2 * o.h.g.NS * sizeof(uint64_t) + sizeof(BcgState) + 16 + us_bytes + 64;
  if (o.smem > cap) return false;
  const size_t ldv = padded_ld(p);
  o.vec = ldv * (size_t)o.h.NCP;
  o.part_ru = round_up((size_t)o.h.nblk * 32, 16);
  o.part_gam = round_up((size_t)o.h.nblk * 64, 16);
  o.spart = round_up((size_t)o.h.nblk * n * NC, 16);
  o.sred = round_up((size_t)n * NC, 16);
  o.gram = round_up((size_t)o.h.nblk * kMaxF * 36, 16);
  o.norm = round_up((size_t)o.h.nblk * 4 * kMaxF, 16);
  return true;
}

template <int MT, int NTM>
void launch_bcg_instance(const CUtensorMap& mA, const CUtensorMap& mX, const CUtensorMap& mR, const CUtensorMap& mY,
                         const CUtensorMap& mS, const BcgParams& P, int nblk, size_t smem, cudaStream_t s) {
  int dev = 0, per_sm = 0;
  SPB_CUDA(cudaGetDevice(&dev));
  SPB_CUDA(cudaFuncSetAttribute(batched_core2_kernel<MT, NTM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  SPB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, batched_core2_kernel<MT, NTM>,
                                                         tmx::Shape<MT>::kThreads, smem));
  if (per_sm < 1 || nblk > per_sm * sm_count())
    throw Error(SPB_ERR_INTERNAL, "batched Core2 kernel: " + std::to_string(nblk) + " CTAs of " + std::to_string(smem) +
                                      " B shared memory are not co-resident (" + std::to_string(per_sm) + " per SM)");
  CUtensorMap a = mA, x = mX, r = mR, y = mY, sm = mS;
  BcgParams Pc = P;
  void* args[] = {&a, &x, &r, &y, &sm, &Pc};
  SPB_CUDA(cudaLaunchCooperativeKernel((void*)batched_core2_kernel<MT, NTM>, dim3(nblk), dim3(tmx::Shape<MT>::kThreads),
                                       args, smem, s));
  count_launch();
}

}  // namespace

bool batched_core2_supported(int p, int n, int K, int F) {
  BcgHost o;
  return bcg_geometry(p, n, K, F, o);
}

size_t batched_core2_workspace_bytes(int p, int n, int K, int F) {
  BcgHost o;
  if (!bcg_geometry(p, n, K, F, o)) return 0;
  return (9 * o.vec + o.part_ru + o.part_gam + o.spart + o.sred + o.gram + o.norm) * sizeof(double) + 256;
}

void launch_batched_core2(const BatchedCore2Call& c, void* ws, cudaStream_t s) {
  BcgHost o;
  SPB_REQUIRE(bcg_geometry(c.p, c.n, c.K, c.F, o), "batched Core2: unsupported problem shape");
  SPB_REQUIRE(c.nsteps >= 1 && c.nsteps <= kMaxSteps, "batched Core2: 1..64 steps per launch");
  const size_t ldv = padded_ld(c.p);
  BcgParams P;
  P.g = o.h.g;
  P.gy = o.h.g;
  P.gy.p = c.n;
  P.gy.nchunks = ceil_div(c.n, 64);
  P.p = c.p; P.n = c.n; P.K = c.K; P.F = c.F; P.NC = c.K * c.F; P.ldv = ldv; P.nblk = o.h.nblk;
  P.nsteps = c.nsteps;
  SPB_REQUIRE(c.mode == 2 || c.mode == 3, "batched kernel: mode must be 2 (Core2) or 3 (Core3)");
  SPB_REQUIRE(c.mode == 2 || c.tau2 != nullptr, "batched Core3: tau2 path missing");
  P.mode = c.mode;
  for (int i = 0; i < kMaxSteps; ++i) P.tau1[i] = i < c.nsteps ? c.tau1[i] : 0.0;
  for (int i = 0; i < kMaxSteps; ++i) P.tau2[i] = (c.mode == 3 && i < c.nsteps) ? c.tau2[i] : 0.0;
  P.c_gram = 2.0; P.maxit = c.maxit; P.tol = c.tol; P.cg_tol2 = c.cg_tol * c.cg_tol; P.cg_cap = c.cg_cap;
  static const int refresh = [] { const char* e = std::getenv("SPB_CG_REFRESH"); return e ? std::max(0, std::atoi(e)) : 8; }();
  P.refresh = refresh;
  for (int f = 0; f < kMaxF; ++f) {
    P.fold_id[f] = f < c.F ? c.fold_id[f] : 0;
    P.rho[f] = f < c.F ? c.rho[f] : 1.0;
    P.Phi[f] = f < c.F ? c.Phi[f] : nullptr;
    P.C[f] = f < c.F ? c.C[f] : nullptr;
    P.L2[f] = f < c.F ? c.L2[f] : nullptr;
    P.R[f] = (c.mode == 3 && f < c.F) ? c.R[f] : nullptr;
    P.L1[f] = (c.mode == 3 && f < c.F) ? c.L1[f] : nullptr;
  }
  P.nk = c.nk; P.Y = c.Y; P.Yt = c.Yt;
  P.snap = c.snap; P.iters_out = c.iters_out; P.status_out = c.status_out;
  double* w = reinterpret_cast<double*>(ws);
  P.Xg = w; w += o.vec;
  P.Rg = w; w += o.vec;
  P.Pg = w; w += o.vec;
  P.Qg = w; w += o.vec;
  P.Bg = w; w += o.vec;
  P.Cg = w; w += o.vec;
  P.L2g = w; w += o.vec;
  P.Radg = w; w += o.vec;
  P.L1g = w; w += o.vec;
  P.part_ru = w; w += o.part_ru;
  P.part_gam = w; w += o.part_gam;
  P.Spart = w; w += o.spart;
  P.Sred = w; w += o.sred;
  P.gram_part = w; w += o.gram;
  P.norm_part = w; w += o.norm;
  P.bar = reinterpret_cast<unsigned int*>(w);
  SPB_CUDA(cudaMemsetAsync(P.bar, 0, 64, s));
  P.dbg = c.dbg;
  const CUtensorMap mA = make_map(c.omega_u, c.p, c.p, c.ld, o.h.g.J);
  const CUtensorMap mX = make_map(P.Xg, c.p, P.NC, ldv, o.h.NCP);
  const CUtensorMap mR = make_map(P.Rg, c.p, P.NC, ldv, o.h.NCP);
  const CUtensorMap mY = make_map(c.Y, c.n, c.p, (size_t)c.n, o.h.g.J);
  const CUtensorMap mS = make_map(P.Sred, c.n, P.NC, (size_t)c.n, o.h.NCP);
  const int NT = o.h.g.NT;
  switch (o.h.MT) {
    case 1:
      if (NT <= 10) launch_bcg_instance<1, 10>(mA, mX, mR, mY, mS, P, o.h.nblk, o.smem, s);
      else launch_bcg_instance<1, 16>(mA, mX, mR, mY, mS, P, o.h.nblk, o.smem, s);
      break;
    case 2:
      if (NT <= 10) launch_bcg_instance<2, 10>(mA, mX, mR, mY, mS, P, o.h.nblk, o.smem, s);
      else launch_bcg_instance<2, 16>(mA, mX, mR, mY, mS, P, o.h.nblk, o.smem, s);
      break;
    default:
      if (NT <= 9) launch_bcg_instance<4, 3>(mA, mX, mR, mY, mS, P, o.h.nblk, o.smem, s);
      else launch_bcg_instance<4, 5>(mA, mX, mR, mY, mS, P, o.h.nblk, o.smem, s);
      break;
  }
}

}  // namespace spb
