// Persistent cooperative ADMM kernel (sm_100a): one launch runs a whole
// spatpcaCore2 / spatpcaCore2p / spatpcaCore3 call (reference
// src/RcppSpatPCA.cpp:150-270) including the stopping rule, so the host never
// synchronises inside the iteration loop.
//
// Per iteration (three grid-wide barriers):
//   phase 1  Phi_partial = Ainv[:, cols] * X[cols, :]      -- the p x p matrix stream.
//            HBM-bound: 8 p^2 bytes, 2 p^2 K flop (K/4 flop/B).  The (row tile x column)
//            space is cut into S equal segments ("stream-K") so that every CTA streams the
//            same number of bytes; a thread owns RPT consecutive rows, loads them with one
//            16-byte non-allocating load per column (a warp reads 512 contiguous bytes), and
//            keeps RPT x K accumulators in registers; X rows are broadcast from shared memory.
//   phase 2  Phi = sum of partials (fixed order); T = Phi + Lambda2 / rho; per-row-block
//            partial Gram matrices of T.
//   phase 3  G = T'T (fixed order) -> K x K Jacobi eigen (one warp) -> P = G^{-1/2};
//            C = T P (polar factor = U V' of the reference's svd_econ), R = soft threshold,
//            dual updates, the four squared error norms, and X for the next iteration --
//            one fused, coalesced pass over the five p x K arrays.
// Every reduction has a fixed shape that depends only on (p, K), never on the grid size or
// on atomics, so results are bitwise reproducible across launches and devices.
#include <cooperative_groups.h>

#include <map>
#include <mutex>

#include "kernels.cuh"

namespace cg = cooperative_groups;

namespace spb {

namespace {

constexpr int kThreads = 256;
constexpr int kSegTarget = 296;      // 2 CTAs x 148 SMs (B200); fixed so results do not depend on the device
constexpr int kMinSegCols = 16;
constexpr int kRowBlock = kThreads;  // phases 2/3: one row per thread
constexpr int kXsBytes = 56 * 1024;  // shared-memory window for X rows in phase 1

template <int KP>
struct Cfg {
  static constexpr int RPT = (KP <= 8) ? 2 : 1;        // rows per thread in phase 1
  static constexpr int TR = kThreads * RPT;            // rows per row tile
  static constexpr int KX = (KP + 1) & ~1;             // X row stride in smem (even -> 16B aligned)
  static constexpr int KG = KP * (KP + 1) / 2;         // packed Gram entries
};

struct Params {
  const double* A;
  size_t ld;
  int p, K, mode, maxit;
  double rho, thr, tol, scale;
  double *Phi, *R, *C, *L1, *L2;
  double *X, *part, *gpart, *npart;
  int* stat;
  double* err;
  int nrt, S, nrb, xcap;
  long long U;
};

__device__ __forceinline__ double2 ld_stream2(const double* ptr) {
  double2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];"
               : "=d"(v.x), "=d"(v.y) : "l"(ptr));
  return v;
}
__device__ __forceinline__ double ld_stream1(const double* ptr) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(ptr));
  return v;
}

__device__ __forceinline__ long long seg_begin(long long s, long long U, int S) { return s * U / S; }
// index of the segment that contains unit u
__device__ __forceinline__ int seg_of(long long u, long long U, int S) {
  return (int)(((u + 1) * S + U - 1) / U - 1);
}

// Fixed-shape block reduction of NV values per thread: warp shuffles, then the 8 warp
// partials are summed in warp order by thread v.  `red` holds 8 * NV doubles.
template <int NV>
__device__ __forceinline__ void block_reduce_store(double (&v)[NV], double* red, double* out) {
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
#pragma unroll
  for (int q = 0; q < NV; ++q) {
    double x = v[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    if (l == 0) red[w * NV + q] = x;
  }
  __syncthreads();
  for (int q = threadIdx.x; q < NV; q += kThreads) {
    double t = 0.0;
#pragma unroll
    for (int ww = 0; ww < kThreads / 32; ++ww) t += red[ww * NV + q];
    out[q] = t;
  }
  __syncthreads();
}

// ---- phase 0: X from the incoming state -----------------------------------------
__device__ __forceinline__ double make_x(int mode, double rho, double r, double c, double l1, double l2) {
  // mode 2: rho*C - Lambda2 (:168);  mode 3: rho*(R+C) - (Lambda1+Lambda2) (:247)
  if (mode == 2) return __dsub_rn(__dmul_rn(rho, c), l2);
  return __dsub_rn(__dmul_rn(rho, __dadd_rn(r, c)), __dadd_rn(l1, l2));
}

// ---- phase 1: matrix stream --------------------------------------------------------
template <int KP>
__device__ __forceinline__ void phase1(const Params& P, double* xs) {
  using C = Cfg<KP>;
  constexpr int RPT = C::RPT, TR = C::TR, KX = C::KX;
  const int tid = threadIdx.x;
  const int p = P.p, K = P.K;
  for (int s = blockIdx.x; s < P.S; s += gridDim.x) {
    const long long u0 = seg_begin(s, P.U, P.S), u1 = seg_begin(s + 1, P.U, P.S);
    const int t0 = (int)(u0 / p);
    for (int w = 0; w < 2; ++w) {
      const int t = t0 + w;
      const long long a = max(u0, (long long)t * p), b = min(u1, (long long)(t + 1) * p);
      if (a >= b) continue;
      const int c0 = (int)(a - (long long)t * p), c1 = (int)(b - (long long)t * p);
      const int r0 = t * TR + RPT * tid;
      double acc[RPT][KP];
#pragma unroll
      for (int r = 0; r < RPT; ++r)
#pragma unroll
        for (int k = 0; k < KP; ++k) acc[r][k] = 0.0;

      for (int cb = c0; cb < c1; cb += P.xcap) {
        const int ncols = min(c1 - cb, P.xcap);
        __syncthreads();                       // previous window fully consumed
        for (int idx = tid; idx < ncols * KX; idx += kThreads) {
          const int jj = idx / KX, k = idx - jj * KX;
          xs[idx] = (k < K) ? __ldcg(P.X + (size_t)k * p + cb + jj) : 0.0;
        }
        __syncthreads();
        if (r0 < p) {
          const double* ap = P.A + (size_t)cb * P.ld + r0;
          int jj = 0;
          constexpr int UN = 8;
          for (; jj + UN <= ncols; jj += UN) {
            if (RPT == 2) {
              double2 av[UN];
#pragma unroll
              for (int u = 0; u < UN; ++u) av[u] = ld_stream2(ap + (size_t)(jj + u) * P.ld);
#pragma unroll
              for (int u = 0; u < UN; ++u) {
                const double* xv = xs + (jj + u) * KX;
#pragma unroll
                for (int k = 0; k < KP; ++k) {
                  const double x = xv[k];
                  acc[0][k] = fma(av[u].x, x, acc[0][k]);
                  acc[RPT - 1][k] = fma(av[u].y, x, acc[RPT - 1][k]);
                }
              }
            } else {
              double av[UN];
#pragma unroll
              for (int u = 0; u < UN; ++u) av[u] = ld_stream1(ap + (size_t)(jj + u) * P.ld);
#pragma unroll
              for (int u = 0; u < UN; ++u) {
                const double* xv = xs + (jj + u) * KX;
#pragma unroll
                for (int k = 0; k < KP; ++k) acc[0][k] = fma(av[u], xv[k], acc[0][k]);
              }
            }
          }
          for (; jj < ncols; ++jj) {
            const double* xv = xs + jj * KX;
            if (RPT == 2) {
              const double2 a2 = ld_stream2(ap + (size_t)jj * P.ld);
#pragma unroll
              for (int k = 0; k < KP; ++k) {
                const double x = xv[k];
                acc[0][k] = fma(a2.x, x, acc[0][k]);
                acc[RPT - 1][k] = fma(a2.y, x, acc[RPT - 1][k]);
              }
            } else {
              const double a1 = ld_stream1(ap + (size_t)jj * P.ld);
#pragma unroll
              for (int k = 0; k < KP; ++k) acc[0][k] = fma(a1, xv[k], acc[0][k]);
            }
          }
        }
      }
      // partial slot of (segment, piece): [k][row in tile]
      double* slot = P.part + ((size_t)(2 * s + w) * KP) * TR;
      if (r0 < p) {
#pragma unroll
        for (int k = 0; k < KP; ++k) {
          if (k < K) {
            if (RPT == 2) {
              double2 o; o.x = acc[0][k]; o.y = acc[RPT - 1][k];
              __stcg(reinterpret_cast<double2*>(slot + (size_t)k * TR + RPT * tid), o);
            } else {
              __stcg(slot + (size_t)k * TR + tid, acc[0][k]);
            }
          }
        }
      }
    }
  }
}

// ---- phase 2: reduce partials, Gram partials -----------------------------------------
template <int KP>
__device__ __forceinline__ void phase2(const Params& P, double* red) {
  using C = Cfg<KP>;
  constexpr int TR = C::TR, KG = C::KG;
  const int tid = threadIdx.x;
  const int p = P.p, K = P.K;
  for (int rb = blockIdx.x; rb < P.nrb; rb += gridDim.x) {
    const int i = rb * kRowBlock + tid;
    double g[KG];
#pragma unroll
    for (int q = 0; q < KG; ++q) g[q] = 0.0;
    if (i < p) {
      const int t = i / TR, li = i - t * TR;
      const int s_lo = seg_of((long long)t * p, P.U, P.S);
      const int s_hi = seg_of((long long)(t + 1) * p - 1, P.U, P.S);
      double phi[KP];
#pragma unroll
      for (int k = 0; k < KP; ++k) phi[k] = 0.0;
      for (int s = s_lo; s <= s_hi; ++s) {
        const int w = ((int)(seg_begin(s, P.U, P.S) / p) == t) ? 0 : 1;
        const double* slot = P.part + ((size_t)(2 * s + w) * KP) * TR + li;
#pragma unroll
        for (int k = 0; k < KP; ++k)
          if (k < K) phi[k] += __ldcg(slot + (size_t)k * TR);
      }
      double tt[KP];
#pragma unroll
      for (int k = 0; k < KP; ++k) {
        tt[k] = 0.0;
        if (k < K) {
          const double ph = __dmul_rn(P.scale, phi[k]);          // 0.5 * (...) in Core3 (:247)
          P.Phi[(size_t)k * p + i] = ph;
          tt[k] = __dadd_rn(ph, __ddiv_rn(__ldcg(P.L2 + (size_t)k * p + i), P.rho));   // :169, :251
        }
      }
      int q = 0;
#pragma unroll
      for (int a = 0; a < KP; ++a)
#pragma unroll
        for (int b = a; b < KP; ++b) g[q++] = tt[a] * tt[b];
    }
    block_reduce_store<KG>(g, red, P.gpart + (size_t)rb * KG);
  }
}

// ---- K x K Jacobi eigen-decomposition + inverse square root, one warp --------------------
// Am (K x K symmetric, ld K) is destroyed; Pm <- Am^{-1/2}.  Returns 0, or 1 if Am is not
// numerically positive definite (rank-deficient T).  All lanes execute the same scalar
// control flow; lanes 0..K-1 apply the rotations.
__device__ int jacobi_inv_sqrt_warp(double* Am, double* Vm, double* Pm, int K, int* sweeps_out) {
  const int lane = threadIdx.x & 31;
  for (int idx = lane; idx < K * K; idx += 32) Vm[idx] = ((idx / K) == (idx % K)) ? 1.0 : 0.0;
  __syncwarp();
  int sweeps = 0;
  for (; sweeps < 64; ++sweeps) {
    bool rotated = false;
    for (int pp = 0; pp < K - 1; ++pp) {
      for (int qq = pp + 1; qq < K; ++qq) {
        const double apq = Am[pp * K + qq];
        const double app = Am[pp * K + pp];
        const double aqq = Am[qq * K + qq];
        if (apq == 0.0 || fabs(apq) <= 1e-16 * sqrt(fabs(app * aqq))) continue;
        rotated = true;
        const double theta = (aqq - app) / (2.0 * apq);
        double t = 1.0 / (fabs(theta) + sqrt(theta * theta + 1.0));
        if (theta < 0.0) t = -t;
        const double c = 1.0 / sqrt(t * t + 1.0);
        const double sn = t * c;
        const double tau = sn / (1.0 + c);
        double aip = 0.0, aiq = 0.0, vip = 0.0, viq = 0.0;
        if (lane < K) {
          aip = Am[lane * K + pp];
          aiq = Am[lane * K + qq];
          vip = Vm[lane * K + pp];
          viq = Vm[lane * K + qq];
        }
        __syncwarp();
        if (lane < K) {
          if (lane != pp && lane != qq) {
            const double nip = aip - sn * (aiq + tau * aip);
            const double niq = aiq + sn * (aip - tau * aiq);
            Am[lane * K + pp] = nip; Am[pp * K + lane] = nip;
            Am[lane * K + qq] = niq; Am[qq * K + lane] = niq;
          }
          Vm[lane * K + pp] = vip - sn * (viq + tau * vip);
          Vm[lane * K + qq] = viq + sn * (vip - tau * viq);
        }
        if (lane == 0) {
          Am[pp * K + pp] = app - t * apq;
          Am[qq * K + qq] = aqq + t * apq;
          Am[pp * K + qq] = 0.0;
          Am[qq * K + pp] = 0.0;
        }
        __syncwarp();
      }
    }
    if (!rotated) break;
  }
  *sweeps_out = sweeps;
  int bad = 0;
  double lmax = 0.0;
  for (int k = 0; k < K; ++k) lmax = fmax(lmax, Am[k * K + k]);
  for (int k = 0; k < K; ++k) {
    const double l = Am[k * K + k];
    if (!(l > 1e-28 * lmax) || !isfinite(l)) bad = 1;
  }
  // P = V diag(lam^-1/2) V'
  for (int idx = lane; idx < K * K; idx += 32) {
    const int a = idx / K, b = idx - a * K;
    double acc = 0.0;
    for (int k = 0; k < K; ++k) acc = fma(Vm[a * K + k] / sqrt(Am[k * K + k]), Vm[b * K + k], acc);
    Pm[idx] = acc;
  }
  __syncwarp();
  return bad;
}

// ---- phase 3: polar factor, soft threshold, dual updates, norms, next X ------------------
template <int KP>
__device__ __forceinline__ int phase3(const Params& P, double* red, double* jA, double* jV, double* jP,
                                      int* sh_flag) {
  using C = Cfg<KP>;
  constexpr int KG = C::KG;
  const int tid = threadIdx.x;
  const int p = P.p, K = P.K;
  // G = sum over row blocks, fixed order; scatter to the symmetric K x K matrix
  for (int q = tid; q < KG; q += kThreads) {
    double t = 0.0;
    for (int rb = 0; rb < P.nrb; ++rb) t += __ldcg(P.gpart + (size_t)rb * KG + q);
    // unpack q -> (a, b), a <= b, in the KP-packed order used by phase 2
    int a = 0, rem = q;
    while (rem >= KP - a) { rem -= KP - a; ++a; }
    const int b = a + rem;
    if (a < K && b < K) { jA[a * K + b] = t; jA[b * K + a] = t; }
  }
  __syncthreads();
  if (tid < 32) {
    int sweeps = 0;
    int bad = jacobi_inv_sqrt_warp(jA, jV, jP, K, &sweeps);
    if (tid == 0) { sh_flag[0] = bad; sh_flag[1] = sweeps; }
  }
  __syncthreads();
  const int bad = sh_flag[0];

  for (int rb = blockIdx.x; rb < P.nrb; rb += gridDim.x) {
    const int i = rb * kRowBlock + tid;
    double nv[4] = {0.0, 0.0, 0.0, 0.0};
    if (i < p) {
      double phi[KP], tt[KP];
#pragma unroll
      for (int k = 0; k < KP; ++k) {
        phi[k] = 0.0; tt[k] = 0.0;
        if (k < K) {
          phi[k] = P.Phi[(size_t)k * p + i];
          tt[k] = __dadd_rn(phi[k], __ddiv_rn(P.L2[(size_t)k * p + i], P.rho));
        }
      }
#pragma unroll
      for (int k = 0; k < KP; ++k) {
        if (k < K) {
          const size_t off = (size_t)k * p + i;
          double cn = 0.0;
#pragma unroll
          for (int a = 0; a < KP; ++a)
            if (a < K) cn = fma(tt[a], jP[a * K + k], cn);
          const double cold = P.C[off];
          const double l2old = P.L2[off];
          const double dC = __dsub_rn(phi[k], cn);
          const double l2n = __dadd_rn(l2old, __dmul_rn(P.rho, dC));        // :173, :258
          const double ec = __dsub_rn(cn, cold);
          nv[2] = fma(dC, dC, nv[2]);
          nv[3] = fma(ec, ec, nv[3]);
          double xn;
          if (P.mode == 3) {
            const double rold = P.R[off];
            const double l1old = P.L1[off];
            const double a1 = __dadd_rn(__ddiv_rn(l1old, P.rho), phi[k]);   // :248
            const double mag = fmax(0.0, __dsub_rn(fabs(a1), P.thr));
            const double sg = (a1 > 0.0) ? 1.0 : ((a1 < 0.0) ? -1.0 : 0.0);
            const double rn = sg * mag;
            const double dR = __dsub_rn(phi[k], rn);
            const double l1n = __dadd_rn(l1old, __dmul_rn(P.rho, dR));      // :257
            const double er = __dsub_rn(rn, rold);
            nv[0] = fma(dR, dR, nv[0]);
            nv[1] = fma(er, er, nv[1]);
            P.R[off] = rn;
            P.L1[off] = l1n;
            xn = make_x(3, P.rho, rn, cn, l1n, l2n);
          } else {
            xn = make_x(2, P.rho, 0.0, cn, 0.0, l2n);
          }
          P.C[off] = cn;
          P.L2[off] = l2n;
          P.X[off] = xn;
        }
      }
    }
    block_reduce_store<4>(nv, red, P.npart + (size_t)rb * 4);
  }
  return bad;
}

template <int KP>
__global__ void __launch_bounds__(kThreads, (KP <= 8) ? 2 : 1)
admm_persistent_kernel(Params P) {
  using C = Cfg<KP>;
  extern __shared__ __align__(16) double dyn[];
  double* xs = dyn;                       // phase 1 window
  double* red = dyn;                      // phases 2/3 reduction scratch (aliases xs)
  __shared__ double jA[KP * KP], jV[KP * KP], jP[KP * KP];
  __shared__ double sh_err[4];
  __shared__ int sh_flag[2];

  cg::grid_group grid = cg::this_grid();
  const int tid = threadIdx.x;
  const size_t pk = (size_t)P.p * P.K;

  // phase 0
  for (size_t idx = (size_t)blockIdx.x * kThreads + tid; idx < pk; idx += (size_t)gridDim.x * kThreads) {
    const double r = (P.mode == 3) ? P.R[idx] : 0.0;
    const double l1 = (P.mode == 3) ? P.L1[idx] : 0.0;
    P.X[idx] = make_x(P.mode, P.rho, r, P.C[idx], l1, P.L2[idx]);
  }
  grid.sync();

  const double denom = sqrt((double)P.p * (double)P.K);
  int iters = 0, status = 0, sweeps = 0;
  double errmax = 0.0;
  for (int it = 0; it < P.maxit; ++it) {
    iters = it + 1;
    phase1<KP>(P, xs);
    grid.sync();
    phase2<KP>(P, red);
    grid.sync();
    status = phase3<KP>(P, red, jA, jV, jP, sh_flag);
    sweeps = sh_flag[1];
    grid.sync();
    // stopping rule (:175-179, :260-266): fixed-order sum of the per-row-block partials
    if (tid < 4) {
      double t = 0.0;
      for (int rb = 0; rb < P.nrb; ++rb) t += __ldcg(P.npart + (size_t)rb * 4 + tid);
      sh_err[tid] = sqrt(t) / denom;
    }
    __syncthreads();
    errmax = fmax(fmax(sh_err[0], sh_err[1]), fmax(sh_err[2], sh_err[3]));
    __syncthreads();
    if (status != 0) break;
    if (errmax <= P.tol) break;
  }
  if (blockIdx.x == 0 && tid == 0) {
    P.stat[0] = iters;
    P.stat[1] = status;
    P.stat[2] = sweeps;
    P.err[0] = errmax;
  }
}

struct Geom {
  int KP, RPT, TR, KX, KG;
  int nrt, S, nrb, xcap;
  long long U;
  size_t part_doubles, gpart_doubles, npart_doubles, x_doubles;
  size_t smem_bytes;
};

int pick_kp(int K) {
  if (K <= 8) return K;
  if (K <= 16) return 16;
  return 32;
}

Geom make_geom(int p, int K) {
  Geom g;
  g.KP = pick_kp(K);
  g.RPT = (g.KP <= 8) ? 2 : 1;
  g.TR = kThreads * g.RPT;
  g.KX = (g.KP + 1) & ~1;
  g.KG = g.KP * (g.KP + 1) / 2;
  g.nrt = ceil_div(p, g.TR);
  g.U = (long long)g.nrt * p;
  long long s = g.U / kMinSegCols;
  if (s > kSegTarget) s = kSegTarget;
  if (s < g.nrt) s = g.nrt;
  if (s < 1) s = 1;
  g.S = (int)s;
  g.nrb = ceil_div(p, kRowBlock);
  int maxcols = (int)((g.U + g.S - 1) / g.S) + 1;
  int cap = kXsBytes / (g.KX * (int)sizeof(double));
  g.xcap = maxcols < cap ? maxcols : cap;
  if (g.xcap < 1) g.xcap = 1;
  g.part_doubles = (size_t)2 * g.S * g.TR * g.KP;
  g.gpart_doubles = (size_t)g.nrb * g.KG;
  g.npart_doubles = (size_t)g.nrb * 4;
  g.x_doubles = (size_t)p * K;
  size_t xs_bytes = (size_t)g.xcap * g.KX * sizeof(double);
  size_t red_bytes = (size_t)(kThreads / 32) * (g.KG > 4 ? g.KG : 4) * sizeof(double);
  g.smem_bytes = xs_bytes > red_bytes ? xs_bytes : red_bytes;
  return g;
}

template <int KP>
void launch_instance(const Params& P, const Geom& g, cudaStream_t s) {
  static std::mutex mu;
  static std::map<std::pair<int, size_t>, int> occ_cache;   // (device, smem) -> co-resident CTAs
  int dev = 0;
  SPB_CUDA(cudaGetDevice(&dev));
  int max_ctas = 0;
  {
    std::lock_guard<std::mutex> lock(mu);
    auto key = std::make_pair(dev, g.smem_bytes);
    auto it = occ_cache.find(key);
    if (it == occ_cache.end()) {
      SPB_CUDA(cudaFuncSetAttribute(admm_persistent_kernel<KP>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kXsBytes));
      int per_sm = 0, sms = 0;
      SPB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, admm_persistent_kernel<KP>,
                                                             kThreads, g.smem_bytes));
      SPB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
      if (per_sm < 1) throw Error(SPB_ERR_INTERNAL, "ADMM kernel does not fit on an SM");
      max_ctas = per_sm * sms;
      occ_cache[key] = max_ctas;
    } else {
      max_ctas = it->second;
    }
  }
  int want = g.S > g.nrb ? g.S : g.nrb;
  int grid = want < max_ctas ? want : max_ctas;
  Params Pc = P;
  void* args[] = {&Pc};
  SPB_CUDA(cudaLaunchCooperativeKernel((void*)admm_persistent_kernel<KP>, dim3(grid), dim3(kThreads),
                                       args, g.smem_bytes, s));
  count_launch();
}

}  // namespace

size_t admm_workspace_bytes(int p, int K) {
  Geom g = make_geom(p, K);
  return (g.x_doubles + g.part_doubles + g.gpart_doubles + g.npart_doubles) * sizeof(double);
}

int admm_num_segments(int p, int K) { return make_geom(p, K).S; }

void launch_admm(const AdmmCall& c, void* ws, cudaStream_t s) {
  SPB_REQUIRE(c.K >= 1 && c.K <= SPB_MAX_K, "K must be in 1..SPB_MAX_K");
  SPB_REQUIRE(c.mode == 2 || c.mode == 3, "ADMM mode must be 2 or 3");
  Geom g = make_geom(c.p, c.K);
  Params P;
  P.A = c.Ainv; P.ld = c.ld; P.p = c.p; P.K = c.K; P.mode = c.mode; P.maxit = c.maxit;
  P.rho = c.rho; P.thr = c.tau2 / c.rho; P.tol = c.tol;
  P.scale = (c.mode == 3) ? 0.5 : 1.0;
  P.Phi = c.Phi; P.R = c.R; P.C = c.C; P.L1 = c.L1; P.L2 = c.L2;
  double* w = reinterpret_cast<double*>(ws);
  P.X = w; w += g.x_doubles;
  P.part = w; w += g.part_doubles;
  P.gpart = w; w += g.gpart_doubles;
  P.npart = w;
  P.stat = c.stat_slot; P.err = c.err_slot;
  P.nrt = g.nrt; P.S = g.S; P.nrb = g.nrb; P.xcap = g.xcap; P.U = g.U;
  switch (g.KP) {
    case 1: launch_instance<1>(P, g, s); break;
    case 2: launch_instance<2>(P, g, s); break;
    case 3: launch_instance<3>(P, g, s); break;
    case 4: launch_instance<4>(P, g, s); break;
    case 5: launch_instance<5>(P, g, s); break;
    case 6: launch_instance<6>(P, g, s); break;
    case 7: launch_instance<7>(P, g, s); break;
    case 8: launch_instance<8>(P, g, s); break;
    case 16: launch_instance<16>(P, g, s); break;
    case 32: launch_instance<32>(P, g, s); break;
    default: throw Error(SPB_ERR_INTERNAL, "no ADMM kernel instance for this K");
  }
}

}  // namespace spb
