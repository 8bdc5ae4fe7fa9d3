// Persistent cooperative ADMM kernel (sm_100a): one launch runs a whole
// spatpcaCore2 / spatpcaCore2p / spatpcaCore3 call (reference
// src/RcppSpatPCA.cpp:150-270) including the stopping rule, so the host never
// synchronises inside the iteration loop.
//
// Per iteration (two grid-wide barriers):
//   stream   Phi_partial = Ainv[:, cols] * X[cols, :]      -- the p x p matrix stream.
//            HBM-bound: 8 p^2 bytes, 2 p^2 K flop (K/4 flop/B).  The (row tile x column)
//            space is cut into S equal segments ("stream-K") so that every CTA streams the
//            same number of bytes; a thread owns RPT consecutive rows, loads them with one
//            16-byte non-allocating load per column (a warp reads 512 contiguous bytes), and
//            keeps RPT x K accumulators in registers; X rows are broadcast from shared memory.
//   reduce   the CTA that delivers the LAST partial of a row tile (atomic ticket) sums the
//            tile's partials in fixed slot order -> Phi, forms T = Phi + Lambda2 / rho and the
//            tile's partial Gram matrix.  The CTA that finishes the last tile sums the tile
//            Grams in tile order, runs the K x K Jacobi eigen-solve in one warp and publishes
//            P = (T'T)^(-1/2).  Who does the work is decided at run time; what is computed,
//            and in which order, is not -- so the result is bitwise reproducible.
//   barrier
//   update   C = T P (polar factor = U V' of the reference's svd_econ), R = soft threshold,
//            dual updates, the squared error norms and X for the next iteration: one fused,
//            coalesced pass over the five p x K arrays.  The CTA that delivers the last norm
//            partial sums them in row-block order and publishes the stopping decision.
//   barrier
// Every reduction has a fixed shape that depends only on (p, K), never on the grid size or on
// floating-point atomics.
//
// Factored systems.  Most systems of a spatpcaCV job are used for one to five iterations, so the
// p^3 explicit inverse is the job's cost, not the stream.  With a Blocking of m > 1 blocks the
// matrix holds the block factor F of spd_inverse.cu (A = U'DU: D_j^-1 on the diagonal blocks,
// -U_jk above, mirrored below) and `stream` becomes 2m - 1 dependent phases over disjoint
// rectangles of F -- together still exactly p^2 entries, 8 p^2 bytes:
//   FWD(j), j = 0..m-1:   rows [o_j, p) x cols block j  times  z[block j]
//                         -> w[block j] = D_j^-1 z_j,   z[rows below] += -U_j.' z_j
//   BWD(k), k = m-1..1:   rows [0, o_k) x cols block k  times  w[block k]  -> w[rows above] += -U_.k w_k
// (z starts as X, w ends as A^-1 X), separated by grid barriers.  Block m-1 of w is final after FWD(m-1),
// block k-1 after BWD(k): the reducer that finalises a row tile forms its Phi, T and Gram partial on the
// spot, so the polar step follows the last phase without an extra pass.  m = 1 is the single fused phase
// described above.
#include <cooperative_groups.h>

#include <cmath>
#include <map>
#include <mutex>
#include <string>
#include <cstdlib>

#include "kernels.cuh"

namespace cg = cooperative_groups;

namespace spb {

namespace {

constexpr int kThreads = 256;
constexpr int kSegTarget = 296;      // 2 CTAs x 148 SMs (B200); fixed so results do not depend on the device
constexpr int kMinSegCols = 64;
constexpr int kRowBlock = kThreads;  // phases 2/3: one row per thread
constexpr int kXsBytes = 56 * 1024;  // shared-memory window for X rows in phase 1

template <int KP>
struct Cfg {
  static constexpr int RPT = (KP <= 8) ? 2 : 1;        // rows per thread in phase 1
  static constexpr int TR = kThreads * RPT;            // rows per row tile
  static constexpr int KX = (KP + 1) & ~1;             // X row stride in smem (even -> 16B aligned)
  static constexpr int KG = KP * (KP + 1) / 2;         // packed Gram entries
};

// One matrix-stream phase: the rectangle rows [row0, row0 + nrows) x cols [col0, col0 + ncols) of the
// matrix times the rows [col0, col0 + ncols) of the input vector.
enum PhaseKind { PH_FINAL = 0, PH_FWD = 1, PH_BWD = 2 };
struct PhaseDesc {
  int row0, nrows, col0, ncols;
  int kind;
  int nrt;            // row tiles of this phase (row0 is a multiple of the tile height)
  int S;              // segments
  int tile0;          // first entry of this phase in the tile table / ticket counters
  int fin0, fin1;     // rows [fin0, fin1) of A^-1 X become final in this phase: their reducer runs the tile epilogue
  long long U;        // nrt * ncols units
};
constexpr int kMaxPhases = 2 * kMaxBlocks - 1;

struct Params {
  const double* A;
  size_t ld;
  int p, K, mode, maxit;
  double rdenom;                              // 1 / sqrt(p K)
  double rho, rinv_rho, thr, tol, scale;     // rinv_rho = RN(1 / rho), computed on the host
  double inv_scale;                           // 1 / scale (scale is 1 or 0.5, so this is exact)
  int force_jacobi;                           // SPB_POLAR=jacobi: skip Newton-Schulz (cross-checks)
  double *Phi, *R, *C, *L1, *L2;
  double *X, *part, *gpart, *npart;
  double* ctrl;        // [0] = max error, [1] = status, [2] = sweeps, [8 ..] = P (K x K)
  int* counters;       // [0] = final slices done, [1] = row blocks done, [2 .. 2 + ntiles) = per-tile arrival counters (only grow),
                       // [2 + ntiles] = arrival counter of grid_barrier
  int* stat;
  double* err;
  int nrt, nrb, xcap;  // nrt: row tiles of the whole p rows
  int nphases, ntiles; // ntiles: tile-table entries over all phases
  int total_slices;    // final row slices per iteration (the last one to finish runs the polar step)
  PhaseDesc ph[kMaxPhases];
  unsigned long long* dbg;   // optional: globaltimer stamps of CTA 0 (8 per iteration), or nullptr
};

// Extra state of the matrix-free Core2 kernel (conjugate gradients on the ADMM system, see cg_core2_kernel).
constexpr int kCgStride = SPB_MAX_K;
enum CgSlot { CG_GAMMA = 0, CG_GAMMA_OLD, CG_ALPHA, CG_ALPHA_OLD, CG_BETA, CG_BN, CG_SS, CG_FROZEN, CG_FRESH, CG_MISC,
              CG_SLOTS };
// CG_MISC: [0] done, [1] matrix passes of the whole call, [2] status, [3] iterations of the running solve
struct CgExtra {
  const double* Ytr;          // n_tr x p (ld n_tr): training rows of the fold (or all of Y)
  const double* Yt;           // p x n_tr (ld p): its transpose
  int n_tr;
  double c_omega, c_gram;     // operator A = c_omega * Omega_u + rho I - c_gram * Ytr'Ytr
  double tol2;                // squared relative residual at which a column is frozen
  int cg_cap;                 // iteration cap of one solve
  int scap;                   // rows of S staged in shared memory at a time
  double *Rv, *Pv, *Qv, *Uv;  // p x K work vectors
  double* Spart;              // nrb x (n_tr x K): row-block partials of Ytr * v
  double* Sg;                 // n_tr x K: Ytr * (current residual / iterate)
  double* rupart;             // nrt * kMaxSlices x KP: slice partials of r . (Omega_u r)
  double* npart2;             // nrb x 2 KP: row-block partials of |r|^2 and |b|^2
  double* cgc;                // CG_SLOTS x kCgStride scalars
};

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define SPB_STAMP(slot)                                                         \
  do {                                                                          \
    if (P.dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 0 && it < 16)     \
      P.dbg[it * 8 + (slot)] = gtimer();                                        \
  } while (0)

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

// Grid-wide barrier of the (co-resident) cooperative grid: one arrival counter that only grows -- the
// barrier of generation g opens when it reaches g * gridDim.x -- so there is nothing to reset and no second
// flag.  Same guarantees as cooperative_groups' grid.sync() for this use (all global writes before the
// barrier are visible after it) at roughly half its latency.
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int& gen) {
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

__host__ __device__ __forceinline__ long long seg_begin(long long s, long long U, int S) { return s * U / S; }
// index of the segment that contains unit u
__host__ __device__ __forceinline__ int seg_of(long long u, long long U, int S) {
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

// a / rho, correctly rounded, in three dependent FP64 operations instead of the ~40 of an IEEE division
// (Markstein: with y = RN(1/rho), q = RN(a y), r = a - rho q exactly by FMA, RN(q + r y) is the
// correctly rounded quotient).  A dependent FP64 operation costs ~55 cycles on this part, and these
// divisions sit on the serial tail of every iteration.
__device__ __forceinline__ double div_rho(double a, double rho, double rinv) {
  const double q = __dmul_rn(a, rinv);
  const double r = fma(-q, rho, a);
  return fma(r, rinv, q);
}

// ---- phase 0: X from the incoming state -----------------------------------------
__device__ __forceinline__ double make_x(int mode, double rho, double r, double c, double l1, double l2) {
  // mode 2: rho*C - Lambda2 (:168);  mode 3: rho*(R+C) - (Lambda1+Lambda2) (:247)
  if (mode == 2) return __dsub_rn(__dmul_rn(rho, c), l2);
  return __dsub_rn(__dmul_rn(rho, __dadd_rn(r, c)), __dadd_rn(l1, l2));
}

// per-tile segment bookkeeping, built once per launch in shared memory
struct TileInfo { int s_lo, s_hi, first_w; };

__device__ __forceinline__ void build_tile_table(const Params& P, TileInfo* tiles) {
  for (int f = 0; f < P.nphases; ++f) {
    const PhaseDesc& ph = P.ph[f];
    for (int t = threadIdx.x; t < ph.nrt; t += kThreads) {
      TileInfo ti;
      ti.s_lo = seg_of((long long)t * ph.ncols, ph.U, ph.S);
      ti.s_hi = seg_of((long long)(t + 1) * ph.ncols - 1, ph.U, ph.S);
      ti.first_w = (seg_begin(ti.s_lo, ph.U, ph.S) < (long long)t * ph.ncols) ? 1 : 0;   // first piece started in tile t-1
      tiles[ph.tile0 + t] = ti;
    }
  }
}

// ---- K x K Jacobi eigen-decomposition + inverse square root, one warp --------------------
// Am (K x K symmetric, ld K) is destroyed; Pm <- Am^{-1/2}.  Returns 0, or 1 if Am is not
// numerically positive definite (rank-deficient T).  All lanes execute the same scalar
// control flow; lanes 0..K-1 apply the rotations.
__device__ int jacobi_inv_sqrt_warp(double* Am, double* Vm, double* Pm, int K, int* sweeps_out,
                                    unsigned long long* dbg) {
  const int lane = threadIdx.x & 31;
  for (int idx = lane; idx < K * K; idx += 32) Vm[idx] = ((idx / K) == (idx % K)) ? 1.0 : 0.0;
  __syncwarp();
  if (dbg != nullptr && lane == 0) dbg[126] = gtimer();
  int sweeps = 0;
  for (; sweeps < 64; ++sweeps) {
    bool rotated = false;
    for (int pp = 0; pp < K - 1; ++pp) {
      for (int qq = pp + 1; qq < K; ++qq) {
        const double apq = Am[pp * K + qq];
        const double app = Am[pp * K + pp];
        const double aqq = Am[qq * K + qq];
        if (apq == 0.0 || apq * apq <= 1e-32 * fabs(app * aqq)) continue;
        rotated = true;
        // rotation angle from short reciprocal / rsqrt chains (any accurate-to-ulps angle does;
        // c^2 + s^2 = 1 holds to ~2 ulp because s = t c and c = rsqrt(t^2 + 1))
        const double theta = (aqq - app) * fast_rcp(2.0 * apq);
        const double th2 = fma(theta, theta, 1.0);
        double t = fast_rcp(fabs(theta) + th2 * fast_rsqrt(th2));
        if (theta < 0.0) t = -t;
        const double c = fast_rsqrt(fma(t, t, 1.0));
        const double sn = t * c;
        const double tau = sn * fast_rcp(1.0 + c);
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
  if (dbg != nullptr && lane == 0) dbg[127] = gtimer();
  // eigenvalues -> lam^-1/2 (one lane each), positivity check, P = V diag(lam^-1/2) V'
  double lam = 1.0;
  if (lane < K) lam = Am[lane * K + lane];
  double lmax = lam;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lmax = fmax(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
  const bool mybad = (lane < K) && !(lam > 1e-28 * lmax && lam < 1e300);   // also false for NaN
  const int bad = __any_sync(0xffffffffu, mybad) ? 1 : 0;
  __syncwarp();
  if (lane < K) Am[lane] = fast_rsqrt(lam);       // row 0 of Am reused as scratch
  __syncwarp();
  {
    int a = 0, b = lane;                           // lane -> (a, b) without an integer division
    while (b >= K) { b -= K; ++a; }
    for (int idx = lane; idx < K * K; idx += 32) {
      double acc = 0.0;
      for (int k = 0; k < K; ++k) acc = fma(Vm[a * K + k] * Am[k], Vm[b * K + k], acc);
      Pm[idx] = acc;
      b += 32;
      while (b >= K) { b -= K; ++a; }
    }
  }
  __syncwarp();
  return bad;
}


// ---- fix-up of one row tile, split over the CTAs that contributed to it ---------------------------
// The pieces of a row tile (one partial per segment that crosses it, 12-30 of them) used to be summed by
// the CTA that delivered the last one: ~10 us of dependent L2 round trips on every phase's critical path.
// Now the tile is cut into ns = min(pieces, kMaxSlices) row slices and the CTA of piece q < ns reduces slice
// q as soon as the tile's ticket counter shows that all pieces have arrived: every thread has at most a few
// 16-byte loads, all in flight together.  Thread (pair, group) sums the pieces group, group + G, ... of one
// row pair in increasing order; the G group sums are combined through shared memory in group order -- the
// shape depends on (p, K, blocking) only, never on arrival order, so results stay bitwise reproducible.
// The slice of a tile whose rows become final in this phase also forms Phi = scale * (A^-1 X) and the slice's
// T'T partial (T = Phi + Lambda2 / rho, :169, :251).  Returns true if the slice was a final one.
constexpr int kMaxSlices = 8;

__host__ __device__ inline int slice_rows(int TR, int ns) { return (((TR + ns - 1) / ns) + 1) & ~1; }

template <int KP, int EPI>
__device__ __forceinline__ bool reduce_slice(const Params& P, const PhaseDesc& ph, int t, const TileInfo& ti, int q,
                                             double* scratch, double* red, const double* cg_vin, const CgExtra* cx) {
  using C = Cfg<KP>;
  constexpr int TR = C::TR, KG = C::KG;
  constexpr int KC = (KP <= 8) ? KP : 8;                 // columns per pass (bounds the shared scratch)
  const int tid = threadIdx.x;
  const int p = P.p, K = P.K;
  const int np = ti.s_hi - ti.s_lo + 1;
  const int ns = np < kMaxSlices ? np : kMaxSlices;
  const int SR = slice_rows(TR, ns);
  const int l0 = q * SR;                                 // first row of the slice inside the tile (even)
  const int nrows = max(0, min(TR, l0 + SR) - l0);       // even: TR and SR are
  const int npairs = nrows >> 1;
  const int tile_row0 = ph.row0 + t * TR;
  const int row_end = ph.row0 + ph.nrows;
  const bool final_tile = (ph.kind == PH_FINAL) || (tile_row0 >= ph.fin0 && tile_row0 < ph.fin1);
  const int diag_end = ph.col0 + ph.ncols;               // FWD: rows of the diagonal block D_j^-1
  int NP2 = 1, sh = 0;
  while (NP2 < npairs) { NP2 <<= 1; ++sh; }
  const int G = kThreads >> sh;                          // piece groups
  const int pi = tid & (NP2 - 1), g = tid >> sh;
  if (npairs > 0) {
    for (int kc0 = 0; kc0 < K; kc0 += KC) {
      double a0[KC], a1[KC];
#pragma unroll
      for (int k = 0; k < KC; ++k) { a0[k] = 0.0; a1[k] = 0.0; }
      if (pi < npairs) {
        constexpr int kB = 4;                            // pieces whose loads are in flight together
        for (int s0 = g; s0 < np; s0 += kB * G) {
          double2 v[kB][KC];
#pragma unroll
          for (int u = 0; u < kB; ++u) {
            const int sp = s0 + u * G;
            const bool live = sp < np;
            const int w = (sp == 0) ? ti.first_w : 0;
            const double* slot = P.part + ((size_t)(2 * (ti.s_lo + (live ? sp : 0)) + w) * KP) * TR + l0 + 2 * pi;
#pragma unroll
            for (int k = 0; k < KC; ++k) {
              if (live && kc0 + k < K) v[u][k] = __ldcg(reinterpret_cast<const double2*>(slot + (size_t)(kc0 + k) * TR));
              else v[u][k] = make_double2(0.0, 0.0);
            }
          }
#pragma unroll
          for (int u = 0; u < kB; ++u)
#pragma unroll
            for (int k = 0; k < KC; ++k) { a0[k] += v[u][k].x; a1[k] += v[u][k].y; }   // + 0.0 for dead slots is exact
        }
      }
      __syncthreads();                                   // scratch free (previous pass / the stream window)
      if (pi < npairs) {
#pragma unroll
        for (int k = 0; k < KC; ++k) {
          double2 o; o.x = a0[k]; o.y = a1[k];
          reinterpret_cast<double2*>(scratch)[(size_t)(g * NP2 + pi) * KC + k] = o;
        }
      }
      __syncthreads();
      for (int e = tid; e < NP2 * KC; e += kThreads) {
        const int pj = e & (NP2 - 1), k = e >> sh;
        const int kk = kc0 + k;
        if (pj >= npairs || kk >= K) continue;
        double2 v = reinterpret_cast<const double2*>(scratch)[(size_t)pj * KC + k];
        for (int gg = 1; gg < G; ++gg) {
          const double2 x = reinterpret_cast<const double2*>(scratch)[(size_t)(gg * NP2 + pj) * KC + k];
          v.x += x.x; v.y += x.y;
        }
        const int i = tile_row0 + l0 + 2 * pj;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int ir = i + r;
          if (ir >= row_end) continue;
          double val = r == 0 ? v.x : v.y;
          const size_t off = (size_t)kk * p + ir;
          if (EPI == 1) { __stcg(cx->Uv + off, val); continue; }                               // U = Omega_u v
          if (ph.kind == PH_FWD) {
            if (ir >= diag_end) { __stcg(P.X + off, __ldcg(P.X + off) + val); continue; }   // z_k -= U_jk' z_j
          } else if (ph.kind == PH_BWD) {
            val = __ldcg(P.Phi + off) + val;                                               // w_j -= U_jk w_k
          }
          // FINAL / FWD diagonal rows / BWD: the running w; a final tile gets Phi = scale * w (0.5 * ... in Core3, :247)
          __stcg(P.Phi + off, final_tile ? __dmul_rn(P.scale, val) : val);
        }
      }
    }
  }
  if (EPI == 1) {
    // slice partial of v . (Omega_u v) per column (the only inner product of a CG step that needs the stream)
    __syncthreads();
    double dq[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) dq[k] = 0.0;
    for (int r = tid; r < nrows; r += kThreads) {
      const int i = tile_row0 + l0 + r;
      if (i >= row_end) continue;
#pragma unroll
      for (int k = 0; k < KP; ++k)
        if (k < K) {
          const size_t off = (size_t)k * p + i;
          dq[k] = fma(__ldcg(cg_vin + off), __ldcg(cx->Uv + off), dq[k]);
        }
    }
    block_reduce_store<KP>(dq, red, cx->rupart + ((size_t)(tile_row0 / TR) * kMaxSlices + q) * KP);
    return true;
  }
  if (!final_tile) return false;
  // T'T partial of the slice: rows re-read from L2 (this CTA wrote them just above)
  __syncthreads();
  double gq[KG];
#pragma unroll
  for (int e = 0; e < KG; ++e) gq[e] = 0.0;
  for (int r = tid; r < nrows; r += kThreads) {
    const int i = tile_row0 + l0 + r;
    if (i >= row_end) continue;
    double tt[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) {
      tt[k] = 0.0;
      if (k < K) {
        const size_t off = (size_t)k * p + i;
        tt[k] = __dadd_rn(__ldcg(P.Phi + off), div_rho(__ldcg(P.L2 + off), P.rho, P.rinv_rho));   // :169, :251
      }
    }
    int e = 0;
#pragma unroll
    for (int a = 0; a < KP; ++a)
#pragma unroll
      for (int b = a; b < KP; ++b) { gq[e] = fma(tt[a], tt[b], gq[e]); ++e; }
  }
  block_reduce_store<KG>(gq, red, P.gpart + ((size_t)(tile_row0 / TR) * kMaxSlices + q) * KG);
  return true;
}

// ---- K x K inverse square root by the coupled Newton-Schulz iteration, whole CTA (K*K <= kThreads) -------
//   Y0 = G / s, Z0 = I;   M = 3I - Z Y;   Y <- Y M / 2,   Z <- M Z / 2;   Z -> (G / s)^(-1/2)
// with s = max row sum >= lambda_max, so the spectrum of Y0 is in (0, 1] and the iteration converges for
// every SPD G, quadratically once ||I - Z Y|| is small.  On this path G = T'T has cond <= ~1.3:
// 4-6 iterations of three K x K products, ~2 us, against 5 Jacobi sweeps of 10 dependent rotations each
// (~45 us at K = 5: every rotation is a chain of ~35 dependent FP64 operations at ~55 cycles).  Thread
// (a, b) owns one entry; the error is tested BEFORE a step, so the step after ||.|| < 3e-9 leaves ~1e-17.
// Returns false (G untouched) when it does not converge in 24 steps -- an ill-conditioned, indefinite or
// non-finite G -- and the caller falls back to the Jacobi eigen-solve, which also classifies the failure.
__device__ bool ns_inv_sqrt_cta(const double* G, double* Y, double* Z, double* M, double* Pout, int K,
                                int* iters_out) {
  __shared__ double rowsum[32];
  const int tid = threadIdx.x;
  const int a = tid / K, b = tid - a * K;
  const bool act = tid < K * K;
  if (tid < K) {
    double r = 0.0;
    for (int j = 0; j < K; ++j) r += fabs(G[tid * K + j]);
    rowsum[tid] = r;
  }
  __syncthreads();
  double sc = 0.0;
  for (int i = 0; i < K; ++i) sc = fmax(sc, rowsum[i]);
  if (!(sc > 0.0) || !(sc < 1e300)) return false;          // uniform: every thread sees the same sc
  const double sinv = fast_rcp(sc);
  if (act) { Y[tid] = G[tid] * sinv; Z[tid] = (a == b) ? 1.0 : 0.0; }
  __syncthreads();
  bool ok = false;
  int it = 0;
  for (; it < 24; ++it) {
    double e = 0.0;
    if (act) {
      double acc = 0.0;
      for (int k = 0; k < K; ++k) acc = fma(Z[a * K + k], Y[k * K + b], acc);
      const double m = ((a == b) ? 3.0 : 0.0) - acc;
      e = fabs(m - ((a == b) ? 2.0 : 0.0));
      M[tid] = m;
    }
    const int big = __syncthreads_or((act && !(e < 3e-9)) ? 1 : 0);     // also publishes M
    double yn = 0.0, zn = 0.0;
    if (act) {
      double ay = 0.0, az = 0.0;
      for (int k = 0; k < K; ++k) {
        ay = fma(Y[a * K + k], M[k * K + b], ay);
        az = fma(M[a * K + k], Z[k * K + b], az);
      }
      yn = 0.5 * ay; zn = 0.5 * az;
    }
    __syncthreads();
    if (act) { Y[tid] = yn; Z[tid] = zn; }
    __syncthreads();
    if (!big) { ok = true; ++it; break; }
  }
  if (!ok) return false;
  const double rs = fast_rsqrt(sc);
  if (act) Pout[tid] = (0.5 * (Z[a * K + b] + Z[b * K + a])) * rs;      // exactly symmetric
  *iters_out = it;
  __syncthreads();
  return true;
}

// ---- Gram reduce + polar factor by the CTA that finished the last slice ------------------------------
template <int KP>
__device__ __forceinline__ void finish_polar(const Params& P, double* jA, double* jV, double* jP, double* scratch) {
  using C = Cfg<KP>;
  constexpr int KG = C::KG;
  const int tid = threadIdx.x;
  const int K = P.K;
  if (P.dbg != nullptr && tid == 0) P.dbg[120] = gtimer();
  // fixed-shape reduction over tiles: lane j of a 16-lane group sums tiles j, j+16, ... in order,
  // then a xor-shuffle tree combines the 16 lanes (the shape depends only on nrt)
  for (int q0 = 0; q0 < KG; q0 += kThreads / 16) {
    const int q = q0 + (tid >> 4), j = tid & 15;
    double t = 0.0;
    if (q < KG) {
      const int nslots = P.nrt * kMaxSlices;           // unused slots hold zeros
#pragma unroll 8
      for (int tb = j; tb < nslots; tb += 16) t += __ldcg(P.gpart + (size_t)tb * KG + q);
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (q < KG && j == 0) {
      int a = 0, rem = q;                          // unpack q -> (a, b), a <= b, KP-packed order
      while (rem >= KP - a) { rem -= KP - a; ++a; }
      const int b = a + rem;
      if (a < K && b < K) { jA[a * K + b] = t; jA[b * K + a] = t; }
    }
  }
  __syncthreads();
  if (P.dbg != nullptr && tid == 0) P.dbg[121] = gtimer();
  // P = (T'T)^(-1/2): Newton-Schulz on the whole CTA; Jacobi eigen-solve in one warp as the fallback (and for
  // K > 16, where K*K entries no longer map to one thread each)
  int steps = 0;
  bool done = false;
  if (K * K <= kThreads && !P.force_jacobi)
    done = ns_inv_sqrt_cta(jA, scratch, scratch + K * K, scratch + 2 * K * K, jP, K, &steps);
  if (done) {
    if (P.dbg != nullptr && tid == 0) { P.dbg[122] = gtimer(); P.dbg[125] = (unsigned long long)steps; P.dbg[126] = P.dbg[127] = P.dbg[121]; }
    for (int idx = tid; idx < K * K; idx += kThreads) __stcg(P.ctrl + 8 + idx, jP[idx]);
    if (tid == 0) { __stcg(P.ctrl + 1, 0.0); __stcg(P.ctrl + 2, (double)steps); }
  } else if (tid < 32) {
    int sweeps = 0;
    const int bad = jacobi_inv_sqrt_warp(jA, jV, jP, K, &sweeps, P.dbg);
    if (P.dbg != nullptr && tid == 0) { P.dbg[122] = gtimer(); P.dbg[125] = (unsigned long long)sweeps; }
    for (int idx = tid; idx < K * K; idx += 32) __stcg(P.ctrl + 8 + idx, jP[idx]);
    if (tid == 0) { __stcg(P.ctrl + 1, (double)bad); __stcg(P.ctrl + 2, (double)(100 + sweeps)); }
  }
  __syncthreads();
}

// ---- stream: matrix-stream partials of one phase, then ticketed tile (/ polar) reduction ---------
template <int KP>
__device__ void cg_finish_stream(const Params& P, const CgExtra& X, int stage);

template <int KP, int EPI = 0>
__device__ __forceinline__ void stream_phase(const Params& P, const PhaseDesc& ph, double* xs, double* red,
                                             const TileInfo* tiles_all, double* jA, double* jV, double* jP,
                                             int* sh_i, bool rev, int dbg_it, const double* cg_vin = nullptr,
                                             const CgExtra* cx = nullptr, int cg_stage = 0) {
  using C = Cfg<KP>;
  constexpr int RPT = C::RPT, TR = C::TR, KX = C::KX;
  const int tid = threadIdx.x;
  const int p = P.p, K = P.K;
  const int nc = ph.ncols;
  const TileInfo* tiles = tiles_all + ph.tile0;
  int* tickets = P.counters + 2 + ph.tile0;
  const double* vin = (EPI == 1) ? cg_vin : ((ph.kind == PH_BWD) ? P.Phi : P.X);     // input vector of the phase
  // the input block of a BWD phase was finalised earlier, i.e. it already holds Phi = scale * w with
  // scale = 1 or 0.5: undo the (exact, power-of-two) scaling on the way into shared memory
  const double vscale = (ph.kind == PH_BWD) ? P.inv_scale : 1.0;
  for (int s = blockIdx.x; s < ph.S; s += gridDim.x) {
    const long long u0 = seg_begin(s, ph.U, ph.S), u1 = seg_begin(s + 1, ph.U, ph.S);
    const int t0 = (int)(u0 / nc);
    for (int w = 0; w < 2; ++w) {
      const int t = t0 + w;
      const long long a = max(u0, (long long)t * nc), b = min(u1, (long long)(t + 1) * nc);
      if (a >= b) continue;
      const int c0 = (int)(a - (long long)t * nc), c1 = (int)(b - (long long)t * nc);
      const int r0 = t * TR + RPT * tid;          // row inside the phase's rectangle
      double acc[RPT][KP];
#pragma unroll
      for (int r = 0; r < RPT; ++r)
#pragma unroll
        for (int k = 0; k < KP; ++k) acc[r][k] = 0.0;

      // Boustrophedon: odd iterations walk the segment's columns backwards, so the ~100 MB of the
      // matrix that the previous iteration left in the 126 MB L2 (its tail) is what this iteration
      // reads first.  The order inside a segment depends only on the iteration parity, so results
      // stay bitwise reproducible.
      const int nwin = (c1 - c0 + P.xcap - 1) / P.xcap;
      for (int wi = 0; wi < nwin; ++wi) {
        const int cb = c0 + (rev ? (nwin - 1 - wi) : wi) * P.xcap;
        const int ncols = min(c1 - cb, P.xcap);
        __syncthreads();                       // previous window fully consumed
        for (int idx = tid; idx < ncols * KX; idx += kThreads) {
          const int jj = idx / KX, k = idx - jj * KX;
          xs[idx] = (k < K) ? vscale * __ldcg(vin + (size_t)k * p + ph.col0 + cb + jj) : 0.0;
        }
        __syncthreads();
        if (r0 < ph.nrows) {
          const long long cs = rev ? -(long long)P.ld : (long long)P.ld;
          const int xstep = rev ? -KX : KX;
          const double* ap = P.A + (size_t)(ph.col0 + (rev ? cb + ncols - 1 : cb)) * P.ld + ph.row0 + r0;
          const double* xp = xs + (rev ? (ncols - 1) * KX : 0);
          int jj = 0;
          constexpr int UN = 8;
          for (; jj + UN <= ncols; jj += UN) {
            if (RPT == 2) {
              double2 av[UN];
#pragma unroll
              for (int u = 0; u < UN; ++u) av[u] = ld_stream2(ap + (long long)(jj + u) * cs);
#pragma unroll
              for (int u = 0; u < UN; ++u) {
                const double* xv = xp + (jj + u) * xstep;
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
              for (int u = 0; u < UN; ++u) av[u] = ld_stream1(ap + (long long)(jj + u) * cs);
#pragma unroll
              for (int u = 0; u < UN; ++u) {
                const double* xv = xp + (jj + u) * xstep;
#pragma unroll
                for (int k = 0; k < KP; ++k) acc[0][k] = fma(av[u], xv[k], acc[0][k]);
              }
            }
          }
          for (; jj < ncols; ++jj) {
            const double* xv = xp + jj * xstep;
            if (RPT == 2) {
              const double2 a2 = ld_stream2(ap + (long long)jj * cs);
#pragma unroll
              for (int k = 0; k < KP; ++k) {
                const double x = xv[k];
                acc[0][k] = fma(a2.x, x, acc[0][k]);
                acc[RPT - 1][k] = fma(a2.y, x, acc[RPT - 1][k]);
              }
            } else {
              const double a1 = ld_stream1(ap + (long long)jj * cs);
#pragma unroll
              for (int k = 0; k < KP; ++k) acc[0][k] = fma(a1, xv[k], acc[0][k]);
            }
          }
        }
      }
      // partial slot of (segment, piece): [k][row in tile]; thread tid holds rows RPT*tid .. +RPT-1
      double* slot = P.part + ((size_t)(2 * s + w) * KP) * TR;
      if (r0 < ph.nrows) {
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
      if (P.dbg != nullptr && tid == 0 && dbg_it == 5 && ph.kind == PH_FINAL) { P.dbg[256 + blockIdx.x] = gtimer(); P.dbg[256 + 512 + blockIdx.x] = P.dbg[5 * 8]; }
      // ticket: announce the piece; if this piece's index is below the tile's slice count this CTA reduces
      // that slice later -- after it has streamed all of its own pieces, so its fix-up never delays a stream
      __threadfence();
      __syncthreads();
      if (tid == 0) {
        atomicAdd(tickets + t, 1);
        const int npieces = tiles[t].s_hi - tiles[t].s_lo + 1;
        const int q = s - tiles[t].s_lo;
        if (q < (npieces < kMaxSlices ? npieces : kMaxSlices) && sh_i[2] < 8) {
          sh_i[3 + sh_i[2]] = t | (q << 16);
          sh_i[2] = sh_i[2] + 1;
        }
      }
    }
  }
  __syncthreads();
  const int nred = sh_i[2];
  for (int r = 0; r < nred; ++r) {
    const int t = sh_i[3 + r] & 0xffff, q = sh_i[3 + r] >> 16;
    if (tid == 0) {
      // the counters only grow: after this phase's pass of iteration `it` tile t has seen (it + 1) * pieces arrivals
      const unsigned int target = (unsigned int)(tiles[t].s_hi - tiles[t].s_lo + 1) * (unsigned int)(dbg_it + 1);
      unsigned int v;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(tickets + t) : "memory");
      } while ((int)(v - target) < 0);
    }
    __syncthreads();
    if (P.dbg != nullptr && tid == 0 && t == 0) P.dbg[123] = gtimer();
    const bool fin = reduce_slice<KP, EPI>(P, ph, t, tiles[t], q, xs, red, cg_vin, cx);
    if (P.dbg != nullptr && tid == 0 && t == 0) P.dbg[124] = gtimer();
    if (!fin) continue;
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      const int old = atomicAdd(P.counters + 0, 1);
      sh_i[1] = (old == P.total_slices - 1) ? 1 : 0;
    }
    __syncthreads();
    if (sh_i[1]) {
      __threadfence();
      if (tid == 0) P.counters[0] = 0;
      if (EPI == 1) cg_finish_stream<KP>(P, *cx, cg_stage);
      else finish_polar<KP>(P, jA, jV, jP, xs);
    }
  }
  __syncthreads();
  if (tid == 0) sh_i[2] = 0;
}

// ---- update: polar factor applied, soft threshold, dual updates, norms, next X ------------------
template <int KP>
__device__ __forceinline__ void update_phase(const Params& P, double* red, double* jP, double* sh_d, int* sh_i) {
  const int tid = threadIdx.x;
  const int p = P.p, K = P.K;
  for (int idx = tid; idx < K * K; idx += kThreads) jP[idx] = __ldcg(P.ctrl + 8 + idx);
  __syncthreads();
  const double rdenom = P.rdenom;
  for (int rb = blockIdx.x; rb < P.nrb; rb += gridDim.x) {
    const int i = rb * kRowBlock + tid;
    double nv[4] = {0.0, 0.0, 0.0, 0.0};
    if (i < p) {
      double phi[KP], tt[KP];
#pragma unroll
      for (int k = 0; k < KP; ++k) {
        phi[k] = 0.0; tt[k] = 0.0;
        if (k < K) {
          phi[k] = __ldcg(P.Phi + (size_t)k * p + i);
          tt[k] = __dadd_rn(phi[k], div_rho(P.L2[(size_t)k * p + i], P.rho, P.rinv_rho));
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
            const double a1 = __dadd_rn(div_rho(l1old, P.rho, P.rinv_rho), phi[k]);   // :248
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
          __stcg(P.X + off, xn);
        }
      }
    }
    block_reduce_store<4>(nv, red, P.npart + (size_t)rb * 4);
    // ticket: the CTA that delivers the last norm partial publishes the stopping decision
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      const int old = atomicAdd(P.counters + 1, 1);
      sh_i[0] = (old == P.nrb - 1) ? 1 : 0;
    }
    __syncthreads();
    if (sh_i[0]) {
      __threadfence();
      if (tid < 128) {                            // 4 norms x 32 lanes, fixed tree over row blocks
        const int v = tid >> 5, j = tid & 31;
        double t = 0.0;
        for (int b = j; b < P.nrb; b += 32) t += __ldcg(P.npart + (size_t)b * 4 + v);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        // norm(., "fro") / sqrt(p K)  (:175-176, :261-264); sqrt(t) = t * rsqrt(t) to ~1 ulp
        if (j == 0) sh_d[v] = (t > 0.0) ? (t * fast_rsqrt(t)) * rdenom : 0.0;
      }
      __syncthreads();
      if (tid == 0) {
        __stcg(P.ctrl + 0, fmax(fmax(sh_d[0], sh_d[1]), fmax(sh_d[2], sh_d[3])));
        P.counters[1] = 0;
      }
      __syncthreads();
    }
  }
}

template <int KP>
__global__ void __launch_bounds__(kThreads, (KP <= 8) ? 2 : 1)
admm_persistent_kernel(Params P) {
  extern __shared__ __align__(16) double dyn[];
  double* xs = dyn;                       // stream window
  double* red = dyn + P.xcap * Cfg<KP>::KX;   // reduction scratch (8 KG + 32 doubles)
  __shared__ double jA[KP * KP], jV[KP * KP], jP[KP * KP];
  __shared__ double sh_d[4];
  __shared__ int sh_i[12];   // [0],[1]: ticket results; [2]: number of tiles to reduce; [3..10]: their ids
  TileInfo* tiles = reinterpret_cast<TileInfo*>(dyn + P.xcap * Cfg<KP>::KX + 8 * Cfg<KP>::KG + 32);

  cg::grid_group grid = cg::this_grid();
  const int tid = threadIdx.x;
  const size_t pk = (size_t)P.p * P.K;

  // set-up: tile table, tickets, X from the incoming state
  build_tile_table(P, tiles);
  if (threadIdx.x == 0) sh_i[2] = 0;
  if (blockIdx.x == 0) {
    for (int i = tid; i < P.ntiles + 3; i += kThreads) P.counters[i] = 0;
    for (int i = tid; i < P.nrt * kMaxSlices * Cfg<KP>::KG; i += kThreads) P.gpart[i] = 0.0;   // unused slice slots stay 0
    if (tid == 0) { __stcg(P.ctrl + 0, 0.0); __stcg(P.ctrl + 1, 0.0); __stcg(P.ctrl + 2, 0.0); }
  }
  for (size_t idx = (size_t)blockIdx.x * kThreads + tid; idx < pk; idx += (size_t)gridDim.x * kThreads) {
    const double r = (P.mode == 3) ? P.R[idx] : 0.0;
    const double l1 = (P.mode == 3) ? P.L1[idx] : 0.0;
    __stcg(P.X + idx, make_x(P.mode, P.rho, r, P.C[idx], l1, P.L2[idx]));
  }
  // the first barrier is cooperative_groups' own: it also orders the counter reset above
  __threadfence();
  grid.sync();
  unsigned int* bar = reinterpret_cast<unsigned int*>(P.counters + 2 + P.ntiles);
  unsigned int bar_gen = 0;

  int iters = 0, status = 0;
  double errmax = 0.0, sweeps = 0.0;
  for (int it = 0; it < P.maxit; ++it) {
    iters = it + 1;
    SPB_STAMP(0);
    for (int f = 0; f < P.nphases; ++f) {
      if (f > 0) grid_barrier(bar, bar_gen);
      if (P.dbg != nullptr && it == 3 && tid == 0 && blockIdx.x < 4) P.dbg[1024 + blockIdx.x * 64 + f * 2] = gtimer();
      stream_phase<KP>(P, P.ph[f], xs, red, tiles, jA, jV, jP, sh_i, (it & 1) != 0, it);
      if (P.dbg != nullptr && it == 3 && tid == 0 && blockIdx.x < 4) P.dbg[1024 + blockIdx.x * 64 + f * 2 + 1] = gtimer();
    }
    SPB_STAMP(1);
    grid_barrier(bar, bar_gen);
    SPB_STAMP(2);
    status = (int)__ldcg(P.ctrl + 1);
    sweeps = __ldcg(P.ctrl + 2);
    update_phase<KP>(P, red, jP, sh_d, sh_i);
    SPB_STAMP(3);
    grid_barrier(bar, bar_gen);
    SPB_STAMP(4);
    errmax = __ldcg(P.ctrl + 0);
    if (status != 0) break;
    if (errmax <= P.tol) break;              // stopping rule (:178, :266)
  }
  if (blockIdx.x == 0 && tid == 0) {
    P.stat[0] = iters;
    P.stat[1] = status;
    P.stat[2] = (int)sweeps;
    P.err[0] = errmax;
  }
}

// =============================================================================================
// Matrix-free Core2: the ADMM Phi-update  Phi = inv_sympd(2 (tau1 Omega - G) + rho I) (rho C - Lambda2)
// (reference src/RcppSpatPCA.cpp:164-168) WITHOUT the p^3 inverse.
//
// Most systems of a spatpcaCV job serve one warm-started spatpcaCore2 call of 1-5 iterations (:329-332), and
// the system  A = c Omega_u + rho I - c' Ytr'Ytr  is very well conditioned on this path: rho = 10 sigma_1^2
// dominates Ytr'Ytr (<= sigma_1^2) and the thin-plate-spline matrix is bounded by its own ridge
// (lambda_max(Omega) <= 1 / (4e-8)), so cond(A) is 1.25 .. ~10 and conjugate gradients reach 1e-15 in 5-40
// steps.  A step streams Omega_u once (8 p^2 bytes, the same HBM-bound stream as an ADMM iteration on an explicit
// inverse) and applies the rank-n_tr part through two skinny products with the L2-resident data block -- the
// p x n contractions of the north-star -- so a whole Core2 call costs ~10^2 matrix passes instead of a
// 0.4 p^3 factorisation (~150 passes at p = 10 000) plus its multi-phase applies.
//
// Chronopoulos-Gear CG (one global reduction point per step), K independent columns sharing the stream:
//     w = A r,  gamma = r.r,  delta = r.w = c (r.Omega_u r) + rho gamma - c' |Ytr r|^2
//     beta = gamma / gamma_old,  alpha = gamma / (delta - beta gamma / alpha_old)
//     p = r + beta p,  q = w + beta q,  x += alpha p,  r -= alpha q
// so only r.(Omega_u r) needs the stream's output: it is formed per row slice in the fix-up, the CTA finishing
// the last slice publishes alpha / beta, and w itself is assembled row by row in the update phase.  A column whose
// relative residual falls below cg_tol is frozen (alpha = beta = 0).  Every reduction has a fixed shape.
// The host decides per call between this kernel and the factored path from a condition-number bound
// (Engine::use_cg); the cap on the CG steps is a safety net that raises an error, never a silent fallback.
// =============================================================================================
__device__ __forceinline__ double* cg_slot(const CgExtra& X, int slot) { return X.cgc + (size_t)slot * kCgStride; }

template <int KP>
__device__ void cg_finish_stream(const Params& P, const CgExtra& X, int stage) {
  if (stage == 0) return;                    // the stream of the initial residual: only U = Omega_u x is needed
  __shared__ double ru_s[KP];
  const int tid = threadIdx.x, K = P.K;
  const int nslots = P.nrt * kMaxSlices;     // unused slots hold zeros
  for (int q0 = 0; q0 < K; q0 += kThreads / 16) {
    const int q = q0 + (tid >> 4), j = tid & 15;
    double t = 0.0;
    if (q < K) {
#pragma unroll 8
      for (int tb = j; tb < nslots; tb += 16) t += __ldcg(X.rupart + (size_t)tb * KP + q);
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (q < K && j == 0) ru_s[q] = t;
  }
  __syncthreads();
  if (tid < K) {
    const int k = tid;
    const double gam = __ldcg(cg_slot(X, CG_GAMMA) + k);
    const double ss = __ldcg(cg_slot(X, CG_SS) + k);
    const bool frozen = __ldcg(cg_slot(X, CG_FROZEN) + k) != 0.0;
    const bool fresh = __ldcg(cg_slot(X, CG_FRESH) + k) != 0.0;
    double alpha = 0.0, beta = 0.0;
    if (!frozen) {
      const double delta = fma(X.c_omega, ru_s[k], fma(P.rho, gam, -X.c_gram * ss));      // r . A r
      double den = delta;
      if (!fresh) {
        beta = gam * fast_rcp(__ldcg(cg_slot(X, CG_GAMMA_OLD) + k));
        den = delta - beta * gam * fast_rcp(__ldcg(cg_slot(X, CG_ALPHA_OLD) + k));
      }
      if (!(delta > 0.0) || !(den > 0.0) || !(den < 1e300)) {
        __stcg(cg_slot(X, CG_MISC) + 2, (double)SPB_ERR_NOT_POSDEF);   // r.Ar <= 0: the system is not positive definite
        den = 1.0;
      }
      alpha = gam * fast_rcp(den);
    }
    __stcg(cg_slot(X, CG_ALPHA) + k, alpha);
    __stcg(cg_slot(X, CG_BETA) + k, beta);
    if (!frozen) __stcg(cg_slot(X, CG_ALPHA_OLD) + k, alpha);
  }
  __syncthreads();
}

// Row-block phases of the CG solve (one row per thread, kRowBlock rows per block, fixed block -> CTA map):
//   STAGE 0   S = Ytr Phi                      (warm start of the solve)
//   STAGE 1   r = b - A Phi,  gamma, |b|^2, S = Ytr r
//   STAGE 2   w = A r;  p, q, Phi, r updates;  gamma, S = Ytr r
// The block partials of S (n_tr x K) and of the norms are summed in block order by the CTA that delivers the
// last partial (ticket), which also takes the per-column decisions.
template <int KP, int STAGE>
__device__ __forceinline__ void cg_row_phase(const Params& P, const CgExtra& X, double* sm, double* red, int* sh_i) {
  const int tid = threadIdx.x, p = P.p, K = P.K, n_tr = X.n_tr;
  const int scap = X.scap;
  double* s_sm = sm;                          // [scap][KP]     chunk of S
  double* v_sm = sm + (size_t)scap * KP;      // [kRowBlock][KP] rows of the vector whose S partial is formed
  const size_t nS = (size_t)n_tr * K;
  double al[KP], be[KP];
  bool fr[KP];
#pragma unroll
  for (int k = 0; k < KP; ++k) { al[k] = 0.0; be[k] = 0.0; fr[k] = true; }
  if (STAGE == 2) {
#pragma unroll
    for (int k = 0; k < KP; ++k)
      if (k < K) {
        al[k] = __ldcg(cg_slot(X, CG_ALPHA) + k);
        be[k] = __ldcg(cg_slot(X, CG_BETA) + k);
        fr[k] = __ldcg(cg_slot(X, CG_FRESH) + k) != 0.0 || __ldcg(cg_slot(X, CG_FROZEN) + k) != 0.0;
      }
  }
  for (int rb = blockIdx.x; rb < P.nrb; rb += gridDim.x) {
    const int i = rb * kRowBlock + tid;
    const bool live = i < p;
    double z[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) z[k] = 0.0;
    if (STAGE != 0) {
      // z = (Ytr' S)(i, :): the transposed block is read row-contiguously across threads, S from shared memory
      for (int r0 = 0; r0 < n_tr; r0 += scap) {
        const int nr = min(scap, n_tr - r0);
        __syncthreads();
        for (int idx = tid; idx < nr * KP; idx += kThreads) {
          const int r = idx / KP, k = idx - r * KP;
          s_sm[idx] = (k < K) ? __ldcg(X.Sg + (size_t)k * n_tr + r0 + r) : 0.0;
        }
        __syncthreads();
        if (live) {
          const double* yp = X.Yt + (size_t)r0 * p + i;
          int r = 0;
          for (; r + 4 <= nr; r += 4) {
            double y[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) y[u] = __ldcg(yp + (size_t)(r + u) * p);
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
              for (int k = 0; k < KP; ++k) z[k] = fma(y[u], s_sm[(r + u) * KP + k], z[k]);
          }
          for (; r < nr; ++r) {
            const double y = __ldcg(yp + (size_t)r * p);
#pragma unroll
            for (int k = 0; k < KP; ++k) z[k] = fma(y, s_sm[r * KP + k], z[k]);
          }
        }
      }
    }
    double vnew[KP], nv[2 * KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) { vnew[k] = 0.0; nv[k] = 0.0; nv[KP + k] = 0.0; }
    if (live) {
#pragma unroll
      for (int k = 0; k < KP; ++k) {
        if (k < K) {
          const size_t off = (size_t)k * p + i;
          if (STAGE == 0) {
            vnew[k] = __ldcg(P.Phi + off);
          } else if (STAGE == 1) {
            const double x = __ldcg(P.Phi + off);
            const double w = fma(X.c_omega, __ldcg(X.Uv + off), fma(P.rho, x, -X.c_gram * z[k]));
            const double b = __ldcg(P.X + off);
            const double r = b - w;
            __stcg(X.Rv + off, r);
            vnew[k] = r;
            nv[k] = r * r;
            nv[KP + k] = b * b;
          } else {
            const double r = __ldcg(X.Rv + off);
            const double w = fma(X.c_omega, __ldcg(X.Uv + off), fma(P.rho, r, -X.c_gram * z[k]));
            double pn = r, qn = w;
            if (!fr[k]) {
              pn = fma(be[k], __ldcg(X.Pv + off), r);
              qn = fma(be[k], __ldcg(X.Qv + off), w);
            }
            __stcg(X.Pv + off, pn);
            __stcg(X.Qv + off, qn);
            __stcg(P.Phi + off, fma(al[k], pn, __ldcg(P.Phi + off)));
            const double rn = fma(-al[k], qn, r);
            __stcg(X.Rv + off, rn);
            vnew[k] = rn;
            nv[k] = rn * rn;
          }
        }
      }
    }
    // S partial of the block: thread r owns row r of Ytr[:, block] * v[block, :]
    __syncthreads();
#pragma unroll
    for (int k = 0; k < KP; ++k) v_sm[tid * KP + k] = vnew[k];
    __syncthreads();
    {
      const int rows = min(kRowBlock, p - rb * kRowBlock);
      for (int r = tid; r < n_tr; r += kThreads) {
        double acc[KP];
#pragma unroll
        for (int k = 0; k < KP; ++k) acc[k] = 0.0;
        const double* yp = X.Ytr + (size_t)rb * kRowBlock * n_tr + r;
        int ii = 0;
        for (; ii + 8 <= rows; ii += 8) {
          double y[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) y[u] = __ldcg(yp + (size_t)(ii + u) * n_tr);
#pragma unroll
          for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int k = 0; k < KP; ++k) acc[k] = fma(y[u], v_sm[(ii + u) * KP + k], acc[k]);
        }
        for (; ii < rows; ++ii) {
          const double y = __ldcg(yp + (size_t)ii * n_tr);
#pragma unroll
          for (int k = 0; k < KP; ++k) acc[k] = fma(y, v_sm[ii * KP + k], acc[k]);
        }
#pragma unroll
        for (int k = 0; k < KP; ++k)
          if (k < K) __stcg(X.Spart + (size_t)rb * nS + (size_t)k * n_tr + r, acc[k]);
      }
    }
    block_reduce_store<2 * KP>(nv, red, X.npart2 + (size_t)rb * 2 * KP);
    // ticket: the CTA that delivers the last block sums the partials in block order and decides
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      const int old = atomicAdd(P.counters + 1, 1);
      sh_i[0] = (old == P.nrb - 1) ? 1 : 0;
    }
    __syncthreads();
    if (sh_i[0]) {
      __threadfence();
      for (size_t e = tid; e < nS; e += kThreads) {
        double t = 0.0;
#pragma unroll 8
        for (int b = 0; b < P.nrb; ++b) t += __ldcg(X.Spart + (size_t)b * nS + e);
        __stcg(X.Sg + e, t);
      }
      __syncthreads();
      // per column: |S_k|^2, gamma_k (and |b_k|^2): one warp per column, lane-strided sums + xor tree
      const int w = tid >> 5, l = tid & 31;
      for (int k = w; k < K; k += kThreads / 32) {
        double ss = 0.0, g = 0.0, bn = 0.0;
        for (int r = l; r < n_tr; r += 32) { const double v = __ldcg(X.Sg + (size_t)k * n_tr + r); ss = fma(v, v, ss); }
        for (int b = l; b < P.nrb; b += 32) {
          g += __ldcg(X.npart2 + (size_t)b * 2 * KP + k);
          bn += __ldcg(X.npart2 + (size_t)b * 2 * KP + KP + k);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          ss += __shfl_xor_sync(0xffffffffu, ss, o);
          g += __shfl_xor_sync(0xffffffffu, g, o);
          bn += __shfl_xor_sync(0xffffffffu, bn, o);
        }
        if (l == 0 && STAGE != 0) {
          __stcg(cg_slot(X, CG_SS) + k, ss);
          if (STAGE == 1) {
            __stcg(cg_slot(X, CG_GAMMA) + k, g);
            __stcg(cg_slot(X, CG_BN) + k, bn);
            __stcg(cg_slot(X, CG_FRESH) + k, 1.0);
            __stcg(cg_slot(X, CG_ALPHA_OLD) + k, 1.0);
            __stcg(cg_slot(X, CG_GAMMA_OLD) + k, 1.0);
            __stcg(cg_slot(X, CG_FROZEN) + k, (g <= X.tol2 * bn) ? 1.0 : 0.0);
          } else {
            const double bnk = __ldcg(cg_slot(X, CG_BN) + k);
            const bool was = __ldcg(cg_slot(X, CG_FROZEN) + k) != 0.0;
            if (!was) {
              __stcg(cg_slot(X, CG_GAMMA_OLD) + k, __ldcg(cg_slot(X, CG_GAMMA) + k));
              __stcg(cg_slot(X, CG_GAMMA) + k, g);
              __stcg(cg_slot(X, CG_FRESH) + k, 0.0);
              if (g <= X.tol2 * bnk || !(g < 1e300)) __stcg(cg_slot(X, CG_FROZEN) + k, 1.0);
            }
          }
        }
      }
      __syncthreads();
      if (STAGE != 0 && tid == 0) {
        int done = 1;
        for (int k = 0; k < K; ++k) done &= (__ldcg(cg_slot(X, CG_FROZEN) + k) != 0.0) ? 1 : 0;
        double* misc = cg_slot(X, CG_MISC);
        if (STAGE == 1) {
          __stcg(misc + 3, 0.0);
        } else {
          const double its = __ldcg(misc + 3) + 1.0;
          __stcg(misc + 3, its);
          if (!done && its >= (double)X.cg_cap) { __stcg(misc + 2, (double)SPB_ERR_INTERNAL); done = 1; }
        }
        __stcg(misc + 0, (double)done);
      }
      if (tid == 0) P.counters[1] = 0;
      __syncthreads();
    }
  }
}

// Gram partials of T = Phi + Lambda2 / rho per row block; the CTA delivering the last one runs the polar step
template <int KP>
__device__ __forceinline__ void cg_gram_phase(const Params& P, double* sm, double* red, double* jA, double* jV,
                                              double* jP, int* sh_i) {
  using C = Cfg<KP>;
  constexpr int KG = C::KG;
  const int tid = threadIdx.x, p = P.p, K = P.K;
  for (int rb = blockIdx.x; rb < P.nrb; rb += gridDim.x) {
    const int i = rb * kRowBlock + tid;
    double gq[KG];
#pragma unroll
    for (int e = 0; e < KG; ++e) gq[e] = 0.0;
    if (i < p) {
      double tt[KP];
#pragma unroll
      for (int k = 0; k < KP; ++k) {
        tt[k] = 0.0;
        if (k < K) {
          const size_t off = (size_t)k * p + i;
          tt[k] = __dadd_rn(__ldcg(P.Phi + off), div_rho(__ldcg(P.L2 + off), P.rho, P.rinv_rho));   // :169
        }
      }
      int e = 0;
#pragma unroll
      for (int a = 0; a < KP; ++a)
#pragma unroll
        for (int b = a; b < KP; ++b) { gq[e] = fma(tt[a], tt[b], gq[e]); ++e; }
    }
    block_reduce_store<KG>(gq, red, P.gpart + (size_t)rb * KG);
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      const int old = atomicAdd(P.counters + 1, 1);
      sh_i[0] = (old == P.nrb - 1) ? 1 : 0;
    }
    __syncthreads();
    if (sh_i[0]) {
      __threadfence();
      if (tid == 0) P.counters[1] = 0;
      finish_polar<KP>(P, jA, jV, jP, sm);
    }
  }
}

template <int KP>
__global__ void __launch_bounds__(kThreads, 2)
cg_core2_kernel(Params P, CgExtra X) {
  extern __shared__ __align__(16) double dyn[];
  double* xs = dyn;
  double* red = dyn + P.xcap * Cfg<KP>::KX;
  __shared__ double jA[KP * KP], jV[KP * KP], jP[KP * KP];
  __shared__ double sh_d[4];
  __shared__ int sh_i[12];
  TileInfo* tiles = reinterpret_cast<TileInfo*>(dyn + P.xcap * Cfg<KP>::KX + 8 * Cfg<KP>::KG + 32);

  cg::grid_group grid = cg::this_grid();
  const int tid = threadIdx.x;
  const size_t pk = (size_t)P.p * P.K;

  build_tile_table(P, tiles);
  if (threadIdx.x == 0) sh_i[2] = 0;
  if (blockIdx.x == 0) {
    for (int i = tid; i < P.ntiles + 3; i += kThreads) P.counters[i] = 0;
    for (int i = tid; i < P.nrt * kMaxSlices * Cfg<KP>::KG; i += kThreads) P.gpart[i] = 0.0;
    for (int i = tid; i < P.nrt * kMaxSlices * KP; i += kThreads) X.rupart[i] = 0.0;
    for (int i = tid; i < CG_SLOTS * kCgStride; i += kThreads) X.cgc[i] = 0.0;
    if (tid == 0) { __stcg(P.ctrl + 0, 0.0); __stcg(P.ctrl + 1, 0.0); __stcg(P.ctrl + 2, 0.0); }
  }
  for (size_t idx = (size_t)blockIdx.x * kThreads + tid; idx < pk; idx += (size_t)gridDim.x * kThreads)
    __stcg(P.X + idx, make_x(2, P.rho, 0.0, P.C[idx], 0.0, P.L2[idx]));        // b = rho C - Lambda2 (:168)
  __threadfence();
  grid.sync();
  unsigned int* bar = reinterpret_cast<unsigned int*>(P.counters + 2 + P.ntiles);
  unsigned int bar_gen = 0;

  int iters = 0, status = 0, pass = 0;
  double errmax = 0.0, sweeps = 0.0;
  for (int it = 0; it < P.maxit; ++it) {
    iters = it + 1;
    // ---- solve A Phi = b by CG, warm-started from the incoming Phi ----
    cg_row_phase<KP, 0>(P, X, xs, red, sh_i);
    grid_barrier(bar, bar_gen);
    stream_phase<KP, 1>(P, P.ph[0], xs, red, tiles, jA, jV, jP, sh_i, (pass & 1) != 0, pass, P.Phi, &X, 0);
    ++pass;
    grid_barrier(bar, bar_gen);
    cg_row_phase<KP, 1>(P, X, xs, red, sh_i);
    grid_barrier(bar, bar_gen);
    int done = (int)__ldcg(cg_slot(X, CG_MISC) + 0);
    while (!done) {
      stream_phase<KP, 1>(P, P.ph[0], xs, red, tiles, jA, jV, jP, sh_i, (pass & 1) != 0, pass, X.Rv, &X, 1);
      ++pass;
      grid_barrier(bar, bar_gen);
      cg_row_phase<KP, 2>(P, X, xs, red, sh_i);
      grid_barrier(bar, bar_gen);
      done = (int)__ldcg(cg_slot(X, CG_MISC) + 0);
    }
    status = (int)__ldcg(cg_slot(X, CG_MISC) + 2);
    if (status != 0) break;
    // ---- Stiefel projection, dual update, stopping rule (:169-179) ----
    cg_gram_phase<KP>(P, xs, red, jA, jV, jP, sh_i);
    grid_barrier(bar, bar_gen);
    status = (int)__ldcg(P.ctrl + 1);
    sweeps = __ldcg(P.ctrl + 2);
    update_phase<KP>(P, red, jP, sh_d, sh_i);
    grid_barrier(bar, bar_gen);
    errmax = __ldcg(P.ctrl + 0);
    if (status != 0) break;
    if (errmax <= P.tol) break;
  }
  if (blockIdx.x == 0 && tid == 0) {
    P.stat[0] = iters;
    P.stat[1] = status;
    P.stat[2] = (int)sweeps;
    P.stat[3] = pass;
    P.err[0] = errmax;
  }
}

struct Geom {
  int KP, RPT, TR, KX, KG;
  int nrt, nrb, xcap;
  int nphases, ntiles, smax, total_slices;
  PhaseDesc ph[kMaxPhases];
  size_t part_doubles, gpart_doubles, npart_doubles, x_doubles, ctrl_doubles, counter_ints;
  size_t smem_bytes;
};

int pick_kp(int K) {
  if (K <= 8) return K;
  if (K <= 16) return 16;
  return 32;
}

// One phase over rows [row0, row0 + nrows) x cols [col0, col0 + ncols): stream-K geometry
void add_phase(Geom& g, int kind, int row0, int nrows, int col0, int ncols) {
  PhaseDesc& ph = g.ph[g.nphases++];
  ph.kind = kind; ph.row0 = row0; ph.nrows = nrows; ph.col0 = col0; ph.ncols = ncols;
  ph.nrt = ceil_div(nrows, g.TR);
  ph.U = (long long)ph.nrt * ncols;
  long long s = ph.U / kMinSegCols;
  if (s > kSegTarget) s = kSegTarget;
  if (s < ph.nrt) s = ph.nrt;
  if (s < 1) s = 1;
  ph.S = (int)s;
  ph.tile0 = g.ntiles;
  ph.fin0 = ph.fin1 = 0;
  g.ntiles += ph.nrt;
  if (ph.S > g.smax) g.smax = ph.S;
}

Geom make_geom(int p, int K, const Blocking& blk) {
  Geom g;
  g.KP = pick_kp(K);
  g.RPT = (g.KP <= 8) ? 2 : 1;
  g.TR = kThreads * g.RPT;
  g.KX = (g.KP + 1) & ~1;
  g.KG = g.KP * (g.KP + 1) / 2;
  g.nrt = ceil_div(p, g.TR);
  g.nphases = 0; g.ntiles = 0; g.smax = 1;
  const int m = blk.m < 1 ? 1 : blk.m;
  if (m == 1) {
    add_phase(g, PH_FINAL, 0, p, 0, p);
  } else {
    SPB_REQUIRE(m <= kMaxBlocks && blk.off[0] == 0 && blk.off[m] == p, "invalid block partition");
    for (int j = 0; j < m; ++j) {
      SPB_REQUIRE(blk.off[j] % g.TR == 0 && blk.off[j + 1] > blk.off[j], "block offsets must be multiples of the row tile");
      add_phase(g, PH_FWD, blk.off[j], p - blk.off[j], blk.off[j], blk.off[j + 1] - blk.off[j]);
    }
    g.ph[m - 1].fin0 = blk.off[m - 1]; g.ph[m - 1].fin1 = p;                // w_{m-1} = D^-1 z is final
    for (int k = m - 1; k >= 1; --k) {
      add_phase(g, PH_BWD, 0, blk.off[k], blk.off[k], blk.off[k + 1] - blk.off[k]);
      g.ph[g.nphases - 1].fin0 = blk.off[k - 1]; g.ph[g.nphases - 1].fin1 = blk.off[k];   // block k-1 is final
    }
  }
  // final row slices per iteration: for every tile of the rows a phase finalises, min(pieces, kMaxSlices)
  g.total_slices = 0;
  for (int f = 0; f < g.nphases; ++f) {
    const PhaseDesc& ph = g.ph[f];
    for (int t = 0; t < ph.nrt; ++t) {
      const int row0 = ph.row0 + t * g.TR;
      if (!(ph.kind == PH_FINAL || (row0 >= ph.fin0 && row0 < ph.fin1))) continue;
      const int np = seg_of((long long)(t + 1) * ph.ncols - 1, ph.U, ph.S) - seg_of((long long)t * ph.ncols, ph.U, ph.S) + 1;
      g.total_slices += np < kMaxSlices ? np : kMaxSlices;
    }
  }
  g.nrb = ceil_div(p, kRowBlock);
  int maxcols = 1;
  for (int f = 0; f < g.nphases; ++f) {
    const int mc = (int)((g.ph[f].U + g.ph[f].S - 1) / g.ph[f].S) + 1;
    if (mc > maxcols) maxcols = mc;
  }
  // leave room in the 56 KB window budget for the reduction scratch and the tile table
  int budget = kXsBytes - (int)((8 * (size_t)g.KG + 32) * sizeof(double)) - g.ntiles * 12;
  if (budget < 4096) budget = 4096;
  int cap = budget / (g.KX * (int)sizeof(double));
  g.xcap = maxcols < cap ? maxcols : cap;
  if (g.xcap < 1) g.xcap = 1;
  // every sub-buffer starts 128-byte aligned (the partial slots are read / written with 16-byte accesses)
  g.part_doubles = round_up((size_t)2 * g.smax * g.TR * g.KP, 16);
  g.gpart_doubles = round_up((size_t)g.nrt * kMaxSlices * g.KG, 16);
  g.npart_doubles = round_up((size_t)g.nrb * 4, 16);
  g.x_doubles = round_up((size_t)p * K, 16);
  g.ctrl_doubles = 8 + (size_t)SPB_MAX_K * SPB_MAX_K;
  g.counter_ints = (size_t)g.ntiles + 3;
  // the stream window doubles as the scratch of the slice fix-up (kThreads x min(KP, 8) double2)
  const int kc = g.KP <= 8 ? g.KP : 8;
  while ((size_t)g.xcap * g.KX < (size_t)2 * kThreads * kc) ++g.xcap;
  // dynamic shared memory: [stream window | reduction scratch (8 KG + 32) | tile table]
  g.smem_bytes = ((size_t)g.xcap * g.KX + 8 * (size_t)g.KG + 32) * sizeof(double) + (size_t)g.ntiles * 3 * sizeof(int);
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
      // 2 CTAs per SM for K <= 8 (56 KB each); the padded instances run 1 CTA per SM and their reduction
      // scratch alone is 34 KB at KP = 32, so they get a larger cap
      constexpr int kSmemCap = (KP <= 8) ? kXsBytes : 100 * 1024;
      if (g.smem_bytes > (size_t)kSmemCap) throw Error(SPB_ERR_INTERNAL, "ADMM kernel: shared memory request too large");
      SPB_CUDA(cudaFuncSetAttribute(admm_persistent_kernel<KP>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemCap));
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
  int want = g.smax > g.nrb ? g.smax : g.nrb;
  int grid = want < max_ctas ? want : max_ctas;
  // a CTA streams ceil(S / grid) segments of up to two pieces each and remembers at most 8 (tile, slice)
  // fix-ups in shared memory: true on any device with >= 74 co-resident CTAs (B200: 148 or 296)
  if (2 * ceil_div(g.smax, grid) > 8)
    throw Error(SPB_ERR_INTERNAL, "ADMM kernel: too few co-resident CTAs on this device for the segment count");
  Params Pc = P;
  void* args[] = {&Pc};
  SPB_CUDA(cudaLaunchCooperativeKernel((void*)admm_persistent_kernel<KP>, dim3(grid), dim3(kThreads),
                                       args, g.smem_bytes, s));
  count_launch();
}

// ---- matrix-free Core2: geometry, workspace, launch ---------------------------------------------------
struct CgGeom {
  Geom g;
  int scap;
  size_t vec_doubles, spart_doubles, sg_doubles, rupart_doubles, npart2_doubles, cgc_doubles;
};

CgGeom make_cg_geom(int p, int K, int n_tr) {
  CgGeom c;
  c.g = make_geom(p, K, Blocking());
  Geom& g = c.g;
  c.scap = n_tr < 256 ? n_tr : 256;
  if (c.scap < 1) c.scap = 1;
  // the stream window doubles as the scratch of the row phases: [scap][KP] of S + [kRowBlock][KP] of v
  const size_t need = (size_t)(c.scap + kRowBlock) * g.KP;
  while ((size_t)g.xcap * g.KX < need) ++g.xcap;
  g.smem_bytes = ((size_t)g.xcap * g.KX + 8 * (size_t)g.KG + 32) * sizeof(double) + (size_t)g.ntiles * 3 * sizeof(int);
  c.vec_doubles = round_up((size_t)p * K, 16);
  c.spart_doubles = round_up((size_t)g.nrb * n_tr * K, 16);
  c.sg_doubles = round_up((size_t)n_tr * K, 16);
  c.rupart_doubles = round_up((size_t)g.nrt * kMaxSlices * g.KP, 16);
  c.npart2_doubles = round_up((size_t)g.nrb * 2 * g.KP, 16);
  c.cgc_doubles = round_up((size_t)CG_SLOTS * kCgStride, 16);
  return c;
}

template <int KP>
void launch_cg_instance(const Params& P, const CgExtra& X, const Geom& g, cudaStream_t s) {
  static std::mutex mu;
  static std::map<std::pair<int, size_t>, int> occ_cache;
  int dev = 0;
  SPB_CUDA(cudaGetDevice(&dev));
  int max_ctas = 0;
  {
    std::lock_guard<std::mutex> lock(mu);
    auto key = std::make_pair(dev, g.smem_bytes);
    auto it = occ_cache.find(key);
    if (it == occ_cache.end()) {
      if (g.smem_bytes > (size_t)kXsBytes) throw Error(SPB_ERR_INTERNAL, "CG kernel: shared memory request too large");
      SPB_CUDA(cudaFuncSetAttribute(cg_core2_kernel<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, kXsBytes));
      int per_sm = 0, sms = 0;
      SPB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cg_core2_kernel<KP>, kThreads, g.smem_bytes));
      SPB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
      if (per_sm < 1) throw Error(SPB_ERR_INTERNAL, "CG kernel does not fit on an SM");
      max_ctas = per_sm * sms;
      occ_cache[key] = max_ctas;
    } else {
      max_ctas = it->second;
    }
  }
  int want = g.smax > g.nrb ? g.smax : g.nrb;
  int grid = want < max_ctas ? want : max_ctas;
  if (2 * ceil_div(g.smax, grid) > 8)
    throw Error(SPB_ERR_INTERNAL, "CG kernel: too few co-resident CTAs on this device for the segment count");
  Params Pc = P;
  CgExtra Xc = X;
  void* args[] = {&Pc, &Xc};
  SPB_CUDA(cudaLaunchCooperativeKernel((void*)cg_core2_kernel<KP>, dim3(grid), dim3(kThreads), args, g.smem_bytes, s));
  count_launch();
}

}  // namespace

size_t admm_workspace_bytes(int p, int K) {
  // sized for any Blocking of p: the partial slots are bounded by kSegTarget segments, the ticket
  // counters by (2 kMaxBlocks - 1) phases of at most nrt tiles
  Geom g = make_geom(p, K, Blocking());
  const size_t part = round_up((size_t)2 * kSegTarget * g.TR * g.KP, 16);
  const size_t counters = (size_t)kMaxPhases * g.nrt + 3;
  return (g.x_doubles + (part > g.part_doubles ? part : g.part_doubles) + g.gpart_doubles + g.npart_doubles +
          g.ctrl_doubles) * sizeof(double) + counters * sizeof(int) + 64;
}

int admm_num_segments(int p, int K) { return make_geom(p, K, Blocking()).smax; }

int admm_phase_plan(int p, int K, const Blocking& blk, int* out, int cap) {
  const Geom g = make_geom(p, K, blk);
  for (int f = 0; f < g.nphases && f < cap; ++f) {
    const PhaseDesc& ph = g.ph[f];
    int* o = out + 8 * f;
    o[0] = ph.kind; o[1] = ph.row0; o[2] = ph.nrows; o[3] = ph.col0; o[4] = ph.ncols;
    // a single-phase (explicit inverse) stream finalises every row
    o[5] = ph.kind == PH_FINAL ? 0 : ph.fin0; o[6] = ph.kind == PH_FINAL ? p : ph.fin1; o[7] = ph.S;
  }
  return g.nphases;
}

void launch_admm(const AdmmCall& c, void* ws, cudaStream_t s) {
  SPB_REQUIRE(c.K >= 1 && c.K <= SPB_MAX_K, "K must be in 1..SPB_MAX_K");
  SPB_REQUIRE(c.mode == 2 || c.mode == 3, "ADMM mode must be 2 or 3");
  Geom g = make_geom(c.p, c.K, c.blocks);
  Params P;
  P.A = c.Ainv; P.ld = c.ld; P.p = c.p; P.K = c.K; P.mode = c.mode; P.maxit = c.maxit;
  P.rdenom = 1.0 / std::sqrt((double)c.p * (double)c.K);
  P.rho = c.rho; P.rinv_rho = 1.0 / c.rho; P.thr = c.tau2 / c.rho; P.tol = c.tol;
  P.scale = (c.mode == 3) ? 0.5 : 1.0;
  P.inv_scale = 1.0 / P.scale;
  static const int force_jacobi = [] {
    const char* e = std::getenv("SPB_POLAR");
    return (e != nullptr && std::string(e) == "jacobi") ? 1 : 0;
  }();
  P.force_jacobi = force_jacobi;
  P.Phi = c.Phi; P.R = c.R; P.C = c.C; P.L1 = c.L1; P.L2 = c.L2;
  double* w = reinterpret_cast<double*>(ws);
  P.X = w; w += g.x_doubles;
  P.part = w; w += g.part_doubles;
  P.gpart = w; w += g.gpart_doubles;
  P.npart = w; w += g.npart_doubles;
  P.ctrl = w; w += g.ctrl_doubles;
  P.counters = reinterpret_cast<int*>(w);
  SPB_REQUIRE(g.ntiles <= 4096, "p is too large for the ADMM kernel's tile table");
  P.stat = c.stat_slot; P.err = c.err_slot;
  P.dbg = c.dbg;
  P.nrt = g.nrt; P.nrb = g.nrb; P.xcap = g.xcap;
  P.nphases = g.nphases; P.ntiles = g.ntiles; P.total_slices = g.total_slices;
  for (int f = 0; f < kMaxPhases; ++f) P.ph[f] = g.ph[f < g.nphases ? f : 0];
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

size_t cg_workspace_bytes(int p, int K, int n_tr) {
  if (K > kCgMaxK) return 0;
  const CgGeom c = make_cg_geom(p, K, n_tr);
  return admm_workspace_bytes(p, K) +
         (4 * c.vec_doubles + c.spart_doubles + c.sg_doubles + c.rupart_doubles + c.npart2_doubles + c.cgc_doubles) *
             sizeof(double) + 256;
}

void launch_core2_cg(const CgCall& c, void* ws, cudaStream_t s) {
  SPB_REQUIRE(c.K >= 1 && c.K <= kCgMaxK, "the matrix-free Core2 kernel is built for K <= 8");
  SPB_REQUIRE(c.n_tr >= 1 && c.Ytr && c.Yt && c.omega_u, "invalid argument");
  CgGeom cgm = make_cg_geom(c.p, c.K, c.n_tr);
  const Geom& g = cgm.g;
  Params P;
  P.A = c.omega_u; P.ld = c.ld; P.p = c.p; P.K = c.K; P.mode = 2; P.maxit = c.maxit;
  P.rdenom = 1.0 / std::sqrt((double)c.p * (double)c.K);
  P.rho = c.rho; P.rinv_rho = 1.0 / c.rho; P.thr = 0.0; P.tol = c.tol;
  P.scale = 1.0; P.inv_scale = 1.0;
  static const int force_jacobi = [] {
    const char* e = std::getenv("SPB_POLAR");
    return (e != nullptr && std::string(e) == "jacobi") ? 1 : 0;
  }();
  P.force_jacobi = force_jacobi;
  P.Phi = c.Phi; P.R = nullptr; P.C = c.C; P.L1 = nullptr; P.L2 = c.L2;
  // base workspace exactly as launch_admm lays it out (sized by admm_workspace_bytes), CG extras behind it
  double* w = reinterpret_cast<double*>(ws);
  double* const w_end_base = reinterpret_cast<double*>(reinterpret_cast<char*>(ws) + admm_workspace_bytes(c.p, c.K));
  P.X = w; w += g.x_doubles;
  P.part = w; w += g.part_doubles;
  P.gpart = w; w += g.gpart_doubles;
  P.npart = w; w += g.npart_doubles;
  P.ctrl = w; w += g.ctrl_doubles;
  P.counters = reinterpret_cast<int*>(w);
  SPB_REQUIRE(g.ntiles <= 4096, "p is too large for the ADMM kernel's tile table");
  P.stat = c.stat_slot; P.err = c.err_slot;
  P.dbg = nullptr;
  P.nrt = g.nrt; P.nrb = g.nrb; P.xcap = g.xcap;
  P.nphases = g.nphases; P.ntiles = g.ntiles; P.total_slices = g.total_slices;
  for (int f = 0; f < kMaxPhases; ++f) P.ph[f] = g.ph[f < g.nphases ? f : 0];
  CgExtra X;
  X.Ytr = c.Ytr; X.Yt = c.Yt; X.n_tr = c.n_tr;
  X.c_omega = c.c_omega; X.c_gram = c.c_gram;
  X.tol2 = c.cg_tol * c.cg_tol; X.cg_cap = c.cg_cap; X.scap = cgm.scap;
  double* e = reinterpret_cast<double*>(round_up(reinterpret_cast<size_t>(w_end_base), 128));
  X.Rv = e; e += cgm.vec_doubles;
  X.Pv = e; e += cgm.vec_doubles;
  X.Qv = e; e += cgm.vec_doubles;
  X.Uv = e; e += cgm.vec_doubles;
  X.Spart = e; e += cgm.spart_doubles;
  X.Sg = e; e += cgm.sg_doubles;
  X.rupart = e; e += cgm.rupart_doubles;
  X.npart2 = e; e += cgm.npart2_doubles;
  X.cgc = e; e += cgm.cgc_doubles;
  switch (g.KP) {
    case 1: launch_cg_instance<1>(P, X, g, s); break;
    case 2: launch_cg_instance<2>(P, X, g, s); break;
    case 3: launch_cg_instance<3>(P, X, g, s); break;
    case 4: launch_cg_instance<4>(P, X, g, s); break;
    case 5: launch_cg_instance<5>(P, X, g, s); break;
    case 6: launch_cg_instance<6>(P, X, g, s); break;
    case 7: launch_cg_instance<7>(P, X, g, s); break;
    case 8: launch_cg_instance<8>(P, X, g, s); break;
    default: throw Error(SPB_ERR_INTERNAL, "no CG kernel instance for this K");
  }
}

}  // namespace spb
