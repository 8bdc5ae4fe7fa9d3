// Hold-out losses of the cross-validation (sm_100a).
//
//  * cv loss   ||Yva (I - Phi Phi')||_F^2  (reference src/RcppSpatPCA.cpp:327, 331, 400) is
//    evaluated as ||Yva - (Yva Phi) Phi'||_F^2: two coalesced passes over the m x p hold-out
//    block (8 m p bytes each) instead of a p x p projector and an m x p x p product.
//  * gamma loss ||S_va - Sigma_hat - err I||_F^2 (:520-522) is evaluated tile by tile from the
//    low-rank factors (Yva: m x p, B = Phi Q: p x K) without ever forming a p x p matrix:
//    2 p^2 (m + K n_gamma) flop over 8 p (m + K) bytes -- FP64-pipe-bound, 64 x 64 tiles,
//    4 x 4 register blocking, only tiles on or above the diagonal (x2 weight off-diagonal).
// All reductions are fixed-shape (block tree, then per-block partials summed in block order).
#include "kernels.cuh"

namespace spb {

namespace {

constexpr int kT = 256;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// ---- W = Ymat * Phi -------------------------------------------------------------------
// Block b handles columns j in [b*JB, (b+1)*JB).  Thread (r-lane, js): rows r, r+64, ...;
// columns js, js+4, ... of the block.  Partial W per block, then an ordered sum over blocks.
constexpr int kJB = 128;        // columns (locations) per block
constexpr int kRL = 64;         // row lanes
constexpr int kJS = kT / kRL;   // column splits

__global__ void __launch_bounds__(kT)
yphi_partial_kernel(const double* __restrict__ Ym, int m, int p, const double* __restrict__ Phi, int K,
                    double* __restrict__ wpart) {
  extern __shared__ double sh[];            // Phi block: kJB x K
  double* phis = sh;
  double* red = sh + kJB * K;               // kJS x (kRL) staging for the js reduction
  const int tid = threadIdx.x;
  const int j0 = blockIdx.x * kJB;
  const int nj = min(kJB, p - j0);
  for (int idx = tid; idx < nj * K; idx += kT) {
    const int jj = idx % nj, k = idx / nj;
    phis[jj * K + k] = Phi[(size_t)k * p + j0 + jj];
  }
  __syncthreads();
  const int rl = tid % kRL, js = tid / kRL;
  double* out = wpart + (size_t)blockIdx.x * m * K;
  for (int r0 = 0; r0 < m; r0 += kRL) {
    const int r = r0 + rl;
    for (int k0 = 0; k0 < K; k0 += 8) {
      double acc[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = 0.0;
      if (r < m) {
        for (int jj = js; jj < nj; jj += kJS) {
          const double y = Ym[(size_t)(j0 + jj) * m + r];
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (k0 + k < K) acc[k] = fma(y, phis[jj * K + k0 + k], acc[k]);
        }
      }
      // reduce over js (fixed order) through shared memory
      for (int k = 0; k < 8 && k0 + k < K; ++k) {
        __syncthreads();
        red[js * kRL + rl] = acc[k];
        __syncthreads();
        if (js == 0 && r < m) {
          double t = 0.0;
#pragma unroll
          for (int q = 0; q < kJS; ++q) t += red[q * kRL + rl];
          out[(size_t)(k0 + k) * m + r] = t;
        }
      }
    }
  }
}

__global__ void ordered_sum_kernel(const double* __restrict__ part, int nblocks, size_t n,
                                   double* __restrict__ out) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  double t = 0.0;
  for (int b = 0; b < nblocks; ++b) t += part[(size_t)b * n + idx];
  out[idx] = t;
}

// ---- || Yva - W Phi' ||_F^2 ----------------------------------------------------------
__global__ void __launch_bounds__(kT)
residual_partial_kernel(const double* __restrict__ Ym, int m, int p, const double* __restrict__ Phi,
                        int K, const double* __restrict__ W, double* __restrict__ lpart) {
  extern __shared__ double sh[];
  double* phis = sh;                      // kJB x K
  __shared__ double wred[kT / 32];
  const int tid = threadIdx.x;
  const int j0 = blockIdx.x * kJB;
  const int nj = min(kJB, p - j0);
  for (int idx = tid; idx < nj * K; idx += kT) {
    const int jj = idx % nj, k = idx / nj;
    phis[jj * K + k] = Phi[(size_t)k * p + j0 + jj];
  }
  __syncthreads();
  double acc = 0.0;
  const size_t total = (size_t)nj * m;
  for (size_t e = tid; e < total; e += kT) {
    const int jj = (int)(e / m), r = (int)(e - (size_t)jj * m);
    double fit = 0.0;
    for (int k = 0; k < K; ++k) fit = fma(__ldg(W + (size_t)k * m + r), phis[jj * K + k], fit);
    const double d = Ym[(size_t)(j0 + jj) * m + r] - fit;
    acc = fma(d, d, acc);
  }
  acc = warp_sum(acc);
  if ((tid & 31) == 0) wred[tid >> 5] = acc;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int w = 0; w < kT / 32; ++w) t += wred[w];
    lpart[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(kT)
final_sum_kernel(const double* __restrict__ part, int n, double* __restrict__ out) {
  __shared__ double wred[kT / 32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += kT) acc += part[i];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) wred[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kT / 32; ++w) t += wred[w];
    out[0] = t;
  }
}

// ---- H = alpha * W' W  (K x K, one block) -------------------------------------------------
__global__ void __launch_bounds__(kT)
small_gram_kernel(const double* __restrict__ W, int m, int K, double alpha, double* __restrict__ H) {
  __shared__ double wred[kT / 32];
  for (int a = 0; a < K; ++a) {
    for (int b = a; b < K; ++b) {
      double acc = 0.0;
      for (int r = threadIdx.x; r < m; r += kT) acc = fma(W[(size_t)a * m + r], W[(size_t)b * m + r], acc);
      acc = warp_sum(acc);
      __syncthreads();
      if ((threadIdx.x & 31) == 0) wred[threadIdx.x >> 5] = acc;
      __syncthreads();
      if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < kT / 32; ++w) t += wred[w];
        t *= alpha;
        H[(size_t)b * K + a] = t;
        H[(size_t)a * K + b] = t;
      }
    }
  }
}

// ---- stage-3 covariance loss ----------------------------------------------------------------
constexpr int kGT = 64;     // tile edge
constexpr int kGR = 32;     // hold-out rows staged per step

template <int NG>
__global__ void __launch_bounds__(kT)
gamma_loss_kernel(const double* __restrict__ Yva, int m, int p, const double* __restrict__ B, int K,
                  const double* __restrict__ lam, const double* __restrict__ errv, int ng,
                  int ntile, double* __restrict__ part) {
  // linear tile index -> (bi <= bj)
  int bi = 0, rem = blockIdx.x;
  while (rem >= ntile - bi) { rem -= ntile - bi; ++bi; }
  const int bj = bi + rem;
  const int i0 = bi * kGT, j0 = bj * kGT;
  const int tid = threadIdx.x;
  const int ti = tid & 15, tj = tid >> 4;          // 16 x 16 threads, 4 x 4 outputs each

  __shared__ double ya[kGR][kGT + 4];
  __shared__ double yb[kGR][kGT + 4];
  extern __shared__ double dsh[];                  // Bi: kGT x K, Bj: kGT x K, lam: ng x K, err: ng
  double* Bi = dsh;
  double* Bj = Bi + kGT * K;
  double* lams = Bj + kGT * K;
  double* errs = lams + NG * K;
  __shared__ double wred[kT / 32][NG];

  for (int idx = tid; idx < kGT * K; idx += kT) {
    const int ii = idx % kGT, k = idx / kGT;
    Bi[ii * K + k] = (i0 + ii < p) ? B[(size_t)k * p + i0 + ii] : 0.0;
    Bj[ii * K + k] = (j0 + ii < p) ? B[(size_t)k * p + j0 + ii] : 0.0;
  }
  for (int idx = tid; idx < NG * K; idx += kT) lams[idx] = (idx < ng * K) ? lam[idx] : 0.0;
  for (int idx = tid; idx < NG; idx += kT) errs[idx] = (idx < ng) ? errv[idx] : 0.0;

  double s[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) s[a][b] = 0.0;

  for (int r0 = 0; r0 < m; r0 += kGR) {
    __syncthreads();
    for (int idx = tid; idx < kGR * kGT; idx += kT) {
      const int rr = idx % kGR, cc = idx / kGR;
      const int r = r0 + rr;
      ya[rr][cc] = (r < m && i0 + cc < p) ? Yva[(size_t)(i0 + cc) * m + r] : 0.0;
      yb[rr][cc] = (r < m && j0 + cc < p) ? Yva[(size_t)(j0 + cc) * m + r] : 0.0;
    }
    __syncthreads();
#pragma unroll 4
    for (int rr = 0; rr < kGR; ++rr) {
      double av[4], bv[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) av[a] = ya[rr][ti * 4 + a];
#pragma unroll
      for (int b = 0; b < 4; ++b) bv[b] = yb[rr][tj * 4 + b];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) s[a][b] = fma(av[a], bv[b], s[a][b]);
    }
  }
  __syncthreads();
  const double dm = (double)m;
  double acc[NG];
#pragma unroll
  for (int g = 0; g < NG; ++g) acc[g] = 0.0;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int gi = i0 + ti * 4 + a, gj = j0 + tj * 4 + b;
      if (gi < p && gj < p) {
        const double sv = s[a][b] / dm;              // S_va = Yva'Yva / n_va (:490)
        const double* bi_row = Bi + (ti * 4 + a) * K;
        const double* bj_row = Bj + (tj * 4 + b) * K;
#pragma unroll
        for (int g = 0; g < NG; ++g) {
          if (g < ng) {
            double sig = 0.0;
            for (int k = 0; k < K; ++k) sig = fma(bi_row[k] * lams[g * K + k], bj_row[k], sig);
            double dlt = sv - sig;
            if (gi == gj) dlt -= errs[g];
            acc[g] = fma(dlt, dlt, acc[g]);
          }
        }
      }
    }
  }
  const double wgt = (bi == bj) ? 1.0 : 2.0;         // symmetric residual: count (bj, bi) too
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    double v = warp_sum(acc[g]);
    if ((tid & 31) == 0) wred[tid >> 5][g] = v;
  }
  __syncthreads();
  if (tid < NG) {
    double t = 0.0;
    for (int w = 0; w < kT / 32; ++w) t += wred[w][tid];
    part[(size_t)tid * gridDim.x + blockIdx.x] = wgt * t;
  }
}

// out[g] = ordered sum over tiles of part[g][tile]
__global__ void __launch_bounds__(kT)
gamma_final_kernel(const double* __restrict__ part, int ntiles, int ng, double* __restrict__ out) {
  __shared__ double chunk[kT];
  const int g = blockIdx.x;
  if (g >= ng) return;
  // fixed shape: thread t sums tiles t, t+256, ...; then thread 0 sums the 256 partials in order
  double acc = 0.0;
  for (int i = threadIdx.x; i < ntiles; i += kT) acc += part[(size_t)g * ntiles + i];
  chunk[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < kT; ++i) t += chunk[i];
    out[g] = t;
  }
}

}  // namespace

size_t yphi_scratch_doubles(int m, int p, int K) {
  return (size_t)ceil_div(p, kJB) * m * K + (size_t)ceil_div(p, kJB);
}

void launch_yphi(const double* Ymat, int m, int p, const double* Phi, int K, double* W,
                 double* scratch, cudaStream_t s) {
  if (m == 0) return;
  const int nb = ceil_div(p, kJB);
  size_t smem = ((size_t)kJB * K + (size_t)kJS * kRL) * sizeof(double);
  yphi_partial_kernel<<<nb, kT, smem, s>>>(Ymat, m, p, Phi, K, scratch);
  size_t n = (size_t)m * K;
  ordered_sum_kernel<<<ceil_div(n, 256), 256, 0, s>>>(scratch, nb, n, W);
  count_launch(2);
  SPB_CUDA(cudaGetLastError());
}

void launch_residual_sumsq(const double* Yva, int m, int p, const double* Phi, int K,
                           const double* W, double* scratch, double* out, cudaStream_t s) {
  const int nb = ceil_div(p, kJB);
  size_t smem = (size_t)kJB * K * sizeof(double);
  residual_partial_kernel<<<nb, kT, smem, s>>>(Yva, m, p, Phi, K, W, scratch);
  final_sum_kernel<<<1, kT, 0, s>>>(scratch, nb, out);
  count_launch(2);
  SPB_CUDA(cudaGetLastError());
}

// ---- stage-3 eigen step on the device (reference src/RcppSpatPCA.cpp:492-520, :729-762) ------------------------
// One warp: cyclic Jacobi eigen-decomposition of H = Phi' S_tr Phi (K x K, upper triangle valid), eigenvalues
// sorted as the reference does (eig_sym ascending, then reversed), accu() with Armadillo's two interleaved
// accumulators, and per gamma the water-filling noise level `err` and the kept eigenvalues max(d - (err + gamma), 0).
// Rotations are applied by lanes i < K; the scalar control flow (thresholds, angles) is uniform.
namespace {
__global__ void __launch_bounds__(32)
gamma_eigen_kernel(const double* __restrict__ H, int K, double tv, int p, const double* __restrict__ gamma, int ng,
                   double* __restrict__ Qd /* K x K, column c = eigenvector of the c-th LARGEST eigenvalue */,
                   double* __restrict__ lam /* ng x K */, double* __restrict__ err /* ng */,
                   double* __restrict__ dvals_out /* K, descending (may be nullptr) */) {
  __shared__ double A[SPB_MAX_K * SPB_MAX_K], V[SPB_MAX_K * SPB_MAX_K], d[SPB_MAX_K];
  __shared__ int order[SPB_MAX_K];
  const int lane = threadIdx.x;
  for (int e = lane; e < K * K; e += 32) {
    const int i = e % K, j = e / K;                       // col-major (i, j); the upper triangle is the source
    A[e] = (i <= j) ? H[(size_t)j * K + i] : H[(size_t)i * K + j];
    V[e] = (i == j) ? 1.0 : 0.0;
  }
  __syncwarp();
  for (int sweep = 0; sweep < 100; ++sweep) {
    bool rotated = false;
    for (int pp = 0; pp < K - 1; ++pp) {
      for (int qq = pp + 1; qq < K; ++qq) {
        const double apq = A[qq * K + pp], app = A[pp * K + pp], aqq = A[qq * K + qq];
        if (apq == 0.0 || fabs(apq) <= 1e-17 * sqrt(fabs(app * aqq))) continue;
        rotated = true;
        const double theta = (aqq - app) / (2.0 * apq);
        double t = 1.0 / (fabs(theta) + sqrt(theta * theta + 1.0));
        if (theta < 0.0) t = -t;
        const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c, tau = sn / (1.0 + c);
        double aip = 0.0, aiq = 0.0, vip = 0.0, viq = 0.0;
        if (lane < K) {
          aip = A[pp * K + lane]; aiq = A[qq * K + lane];
          vip = V[pp * K + lane]; viq = V[qq * K + lane];
        }
        __syncwarp();
        if (lane < K) {
          if (lane != pp && lane != qq) {
            const double nip = aip - sn * (aiq + tau * aip), niq = aiq + sn * (aip - tau * aiq);
            A[pp * K + lane] = nip; A[lane * K + pp] = nip;
            A[qq * K + lane] = niq; A[lane * K + qq] = niq;
          }
          V[pp * K + lane] = vip - sn * (viq + tau * vip);
          V[qq * K + lane] = viq + sn * (vip - tau * viq);
        }
        if (lane == 0) {
          A[pp * K + pp] = app - t * apq;
          A[qq * K + qq] = aqq + t * apq;
          A[qq * K + pp] = 0.0; A[pp * K + qq] = 0.0;
        }
        __syncwarp();
      }
    }
    if (!rotated) break;
  }
  // stable ascending order of the eigenvalues, then reversed (:494-495)
  if (lane == 0) {
    for (int i = 0; i < K; ++i) order[i] = i;
    for (int i = 1; i < K; ++i) {                          // insertion sort: stable
      const int o = order[i];
      const double v = A[o * K + o];
      int j = i - 1;
      while (j >= 0 && A[order[j] * K + order[j]] > v) { order[j + 1] = order[j]; --j; }
      order[j + 1] = o;
    }
  }
  __syncwarp();
  if (lane < K) d[lane] = A[order[K - 1 - lane] * K + order[K - 1 - lane]];          // descending
  for (int e = lane; e < K * K; e += 32) {
    const int i = e % K, c = e / K;
    Qd[e] = V[order[K - 1 - c] * K + i];
  }
  __syncwarp();
  if (dvals_out != nullptr && lane < K) dvals_out[lane] = d[lane];
  // accu(transformed_eigenvalues) over the ASCENDING vector with two interleaved accumulators (:493)
  double total;
  {
    double a1 = 0.0, a2 = 0.0;
    int i = 0, j = 1;
    for (; j < K; j += 2, i += 2) { a1 += d[K - 1 - i]; a2 += d[K - 1 - (i + 1)]; }
    if ((j - 1) < K) a1 += d[K - 1 - i];
    total = a1 + a2;
  }
  for (int g = lane; g < ng; g += 32) {
    const double gm = gamma[g];
    // water filling (:499-519 / :743-760)
    int tempL = K;
    double tot = total, e;
    if (d[0] > gm) {
      e = (tv - tot + K * gm) / (p - tempL);
      double temp = d[tempL - 1];
      while (temp - gm < e) {
        if (tempL == 1) { e = (tv - d[0] + gm) / (p - 1); break; }
        tot -= d[tempL - 1];
        tempL--;
        e = (tv - tot + tempL * gm) / (p - tempL);
        temp = d[tempL - 1];
      }
      if (d[0] - gm < e) e = tv / p;
    } else {
      e = tv / p;
    }
    err[g] = e;
    for (int k = 0; k < K; ++k) lam[(size_t)g * K + k] = fmax(d[k] - (e + gm), 0.0);                  // :520
  }
}
}  // namespace

void launch_gamma_eigen(const double* H, int K, double tv, int p, const double* gamma_dev, int ng, double* Qd,
                        double* lam, double* err, double* dvals_out, cudaStream_t s) {
  SPB_REQUIRE(K >= 1 && K <= SPB_MAX_K, "K must be in 1..SPB_MAX_K");
  gamma_eigen_kernel<<<1, 32, 0, s>>>(H, K, tv, p, gamma_dev, ng, Qd, lam, err, dvals_out);
  count_launch();
  SPB_CUDA(cudaGetLastError());
}

void launch_small_gram(const double* W, int m, int K, double alpha, double* H, cudaStream_t s) {
  small_gram_kernel<<<1, kT, 0, s>>>(W, m, K, alpha, H);
  count_launch();
  SPB_CUDA(cudaGetLastError());
}

namespace {
// out[g] = |S_va - B diag(lam_g) B' - err_g I|_F^2 from O(K^2) sufficient statistics (see launch_gamma_combine)
__global__ void gamma_combine_kernel(const double* __restrict__ cst, const double* __restrict__ Cv,
                                     const double* __restrict__ BtB, const double* __restrict__ lam,
                                     const double* __restrict__ errv, int ng, int K, int p, double inv_m,
                                     double* __restrict__ out) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ng) return;
  const double s2 = cst[0] * inv_m * inv_m;       // |S_va|_F^2 = |Yva Yva'|_F^2 / m^2
  const double trS = cst[1] * inv_m;              // tr(S_va) = |Yva|_F^2 / m
  const double e = errv[g];
  const double* l = lam + (size_t)g * K;
  double quad = 0.0, cross = 0.0, trB = 0.0;
  for (int a = 0; a < K; ++a) {
    for (int b = 0; b < K; ++b) {
      const double t = BtB[(size_t)b * K + a];
      quad = fma(l[a] * l[b], t * t, quad);       // |B L B'|_F^2 = sum_ab l_a l_b (B'B)_ab^2
    }
    cross = fma(l[a], Cv[(size_t)a * K + a], cross);      // <S_va, B L B'> = sum_a l_a (B' S_va B)_aa
    trB = fma(l[a], BtB[(size_t)a * K + a], trB);         // tr(B L B')
  }
  out[g] = ((s2 + quad) + e * e * (double)p) - 2.0 * cross - 2.0 * e * trS + 2.0 * e * trB;
}
}  // namespace

// The stage-3 loss (:521-523) without its p x p residual:
//   |S - B L B' - e I|_F^2 = |S|_F^2 + |B L B'|_F^2 + e^2 p - 2 <S, B L B'> - 2 e tr S + 2 e tr(B L B'),
// every term a function of the m x m Gram matrix Yva Yva', of Wv = Yva B (m x K) and of B'B (K x K): O(p m^2)
// work instead of O(p^2 (m + n_gamma K)).  No cancellation to speak of: the loss is 0.1 .. 1.0 of |S|_F^2 on every
// configuration measured (S is a rank-m sample covariance with noise, the fit has rank K), and the parity tests
// compare with the oracle's tile-by-tile evaluation.
void launch_gamma_combine(const double* cst, const double* Cv, const double* BtB, const double* lam, const double* err,
                          int ng, int K, int p, int m, double* out, cudaStream_t s) {
  gamma_combine_kernel<<<ceil_div(ng, 64), 64, 0, s>>>(cst, Cv, BtB, lam, err, ng, K, p, 1.0 / (double)m, out);
  count_launch();
  SPB_CUDA(cudaGetLastError());
}

size_t gamma_scratch_doubles(int p) {
  size_t nt = ceil_div(p, kGT);
  return (size_t)kGammaBatch * (nt * (nt + 1) / 2);
}

void launch_gamma_loss(const double* Yva, int m, int p, const double* B, int K, const double* lam_dev,
                       const double* err_dev, int ng, double* scratch, double* out, cudaStream_t s) {
  SPB_REQUIRE(ng >= 1 && ng <= kGammaBatch, "gamma batch too large");
  const int nt = ceil_div(p, kGT);
  const int ntiles = nt * (nt + 1) / 2;
  size_t smem = ((size_t)2 * kGT * K + (size_t)kGammaBatch * K + kGammaBatch) * sizeof(double);
  gamma_loss_kernel<kGammaBatch><<<ntiles, kT, smem, s>>>(Yva, m, p, B, K, lam_dev, err_dev, ng, nt,
                                                          scratch);
  gamma_final_kernel<<<ng, kT, 0, s>>>(scratch, ntiles, ng, out);
  count_launch(2);
  SPB_CUDA(cudaGetLastError());
}

}  // namespace spb
