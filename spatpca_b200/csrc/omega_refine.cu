// thinPlateSplineMatrix (reference src/RcppSpatPCA.cpp:58-76) to the accuracy of the INPUT, not of the LU.
//
// The bordered system  A = [[E + 1e-8 I, T], [T', 1e-8 I]]  has cond ~1e8 (p = 1000) .. 1e10 (p = 10 000) on
// irregular sites, so any FP64 LU -- LAPACK's dgetrf/dgetri in the reference, cuSOLVER's here -- returns
// Lp = (A^-1)[0:p, 0:p] with a forward error of a few 1e-11 .. 1e-9 relative, and the reference's product
// Omega = Lp' (E Lp) adds as much again through cancellation (E Lp = I - 1e-8 Lp - T M').  Two such pipelines
// differ from each other by MORE than either differs from the exact result; measured against an 80-bit
// refinement (profiles/r02_omega_threeway.md): cuSOLVER + cuBLAS 1.1e-10, the oracle's LAPACK 5.8e-11, each other
// 9.5e-11 at p = 2000 -- and at p = 10 000 that difference alone pushed the eigenfunctions to 1.07e-8 of the
// oracle's, over the 1e-8 parity bar.  The only way under the bar is to remove this side's share of the error:
//
//   1. X0 = A^-1 by LU with partial pivoting (the reference's algorithm class), symmetrised.
//   2. One Newton-Schulz step  X1 = X0 + X0 (I - A X0)  with the residual formed EXACTLY where it matters:
//      A and X0 are split into a head of beta = (53 - ceil(log2 q)) / 2 bits per row / column exponent and a
//      tail (Ozaki splitting), so that the head product A1 X1 is free of rounding in FP64 DMMA arithmetic
//      (every partial sum is representable) and the two tail products carry 2^-beta of the usual GEMM error:
//          I - A X0 = (I - A1 X1) - (A1 X2 + A2 X0)        three DGEMMs, error ~1e-6 of a plain FP64 residual.
//   3. Omega from the identities of the exact inverse  (E + eps I) Lp + T M' = I,  T' Lp + eps M' = 0:
//          Lp' E Lp = Lp - eps Lp Lp + eps M M'             one symmetric product, no cancellation,
//      instead of the reference's two products (mathematically the same matrix; with Lp accurate to ~1e-15 it is
//      accurate to ~1e-14, while Lp' (E Lp) in FP64 keeps ~3e-11 even for an exactly rounded Lp).
//
// Flop count: LU 2/3 + solve 2 + residual 6 + correction 1 + product 1 = 10.7 p^3 (the unrefined pipeline: 5.7).
// What remains between this path and the oracle is the ORACLE's own rounding error.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <string>

#include "engine.cuh"

namespace spb {

namespace {

bool trace_on() {
  static const bool on = [] { const char* e = std::getenv("SPB_TRACE"); return e && e[0] == '1'; }();
  return on;
}

// column-wise max |a| of a (rows x cols) matrix -> out[cols]; one block per column
__global__ void __launch_bounds__(256) col_absmax_kernel(const double* __restrict__ A, size_t ld, int rows,
                                                         double* __restrict__ out) {
  __shared__ double red[8];
  const double* col = A + (size_t)blockIdx.x * ld;
  double m = 0.0;
  for (int i = threadIdx.x; i < rows; i += 256) m = fmax(m, fabs(col[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) m = fmax(m, red[w]);
    out[blockIdx.x] = m;
  }
}

// sigma[t] = 1.5 * 2^(52 + e - beta) with max[t] < 2^e: (a + sigma) - sigma rounds a to a multiple of 2^(e - beta)
__global__ void split_sigma_kernel(const double* __restrict__ mx, int n, int beta, double* __restrict__ sigma) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const double m = mx[t];
  if (!(m > 0.0) || !(m < 1e300)) { sigma[t] = 0.0; return; }
  int e;
  frexp(m, &e);                                   // m = f * 2^e, f in [0.5, 1)
  sigma[t] = ldexp(1.5, 52 + e - beta);
}

// out = head(src) (tail == 0) or src - head(src) (tail == 1); the exponent is per row (by_row) or per column.
// out may alias src.
__global__ void __launch_bounds__(128) split_kernel(const double* __restrict__ src, size_t lds, int rows, int cols,
                                                    const double* __restrict__ sigma, int by_row, int tail,
                                                    double* __restrict__ out, size_t ldo) {
  const int i = blockIdx.x * 128 + threadIdx.x;
  const int j0 = blockIdx.y * 8;
  if (i >= rows) return;
  const double srow = by_row ? sigma[i] : 0.0;
#pragma unroll
  for (int jj = 0; jj < 8; ++jj) {
    const int j = j0 + jj;
    if (j >= cols) break;
    const double s = by_row ? srow : sigma[j];
    const double a = src[(size_t)j * lds + i];
    const double hi = (s == 0.0) ? 0.0 : __dsub_rn(__dadd_rn(a, s), s);
    out[(size_t)j * ldo + i] = tail ? __dsub_rn(a, hi) : hi;
  }
}

// C1 <- C1 - C2
__global__ void __launch_bounds__(128) sub_inplace_kernel(double* __restrict__ C1, const double* __restrict__ C2,
                                                          size_t ld, int rows, int cols) {
  const int i = blockIdx.x * 128 + threadIdx.x;
  const int j0 = blockIdx.y * 8;
  if (i >= rows) return;
#pragma unroll
  for (int jj = 0; jj < 8; ++jj) {
    const int j = j0 + jj;
    if (j >= cols) break;
    const size_t o = (size_t)j * ld + i;
    C1[o] = __dsub_rn(C1[o], C2[o]);
  }
}

// C[:, c0 + j] <- e_{c0 + j} for j in [0, nc)
__global__ void __launch_bounds__(128) unit_cols_kernel(double* __restrict__ C, size_t ld, int rows, int c0, int nc) {
  const int i = blockIdx.x * 128 + threadIdx.x;
  const int j0 = blockIdx.y * 8;
  if (i >= rows) return;
#pragma unroll
  for (int jj = 0; jj < 8; ++jj) {
    const int j = j0 + jj;
    if (j >= nc) break;
    C[(size_t)(c0 + j) * ld + i] = (i == c0 + j) ? 1.0 : 0.0;
  }
}

// SPB_OMEGA_SPDINV=schur: the SPD route inverts its block with the recursive Schur-complement inverse (round 2a)
bool omega_spd_schur() {
  static const bool v = [] { const char* e = std::getenv("SPB_OMEGA_SPDINV"); return e && std::string(e) == "schur"; }();
  return v;
}

// strictly upper triangle of the leading n x n block <- 0 (a lower-triangular factor that DGEMM may read as dense)
__global__ void __launch_bounds__(128) zero_strict_upper_kernel(double* __restrict__ A, size_t ld, int n) {
  const int i = blockIdx.x * 128 + threadIdx.x;
  const int j0 = blockIdx.y * 8;
  if (i >= n) return;
#pragma unroll
  for (int jj = 0; jj < 8; ++jj) {
    const int j = j0 + jj;
    if (j >= n) break;
    if (i < j) A[(size_t)j * ld + i] = 0.0;
  }
}

// Strip width of the triangular-aware products below: a product with a triangular operand is cut into strips of
// kTriStrip rows / columns, and every strip only runs over the part of the contraction range where the triangular
// operand is non-zero (plain cuBLAS DGEMMs on sub-blocks, pointer arithmetic only).  10 strips at p = 10 000:
// 0.55 of the dense flop count instead of 1 (the ideal, 0.5, needs infinitely thin strips).
constexpr int kTriStrip = 1024;

// V = L^-1 in place for a lower-triangular L (n x n, its strict upper triangle stored as zeros): recursive,
//   [L11 0; L21 L22]^-1 = [V11 0; -V22 L21 V11  V22],
// every flop in a cuBLAS DGEMM at the DGEMM rate (instead of cuSOLVER's trtri rate); both products have a triangular
// operand and are cut into strips (see kTriStrip); 512-row leaves by TRSM on an identity.
// T: scratch of >= n^2 / 4 + 512^2.
void trtri_lower_gemm(cublasHandle_t blas, cudaStream_t s, double* L, int n, size_t ld, double* T) {
  const double one = 1.0, zero = 0.0, mone = -1.0;
  if (n <= 512) {
    launch_set_identity(T, n, n, (size_t)n, s);
    SPB_CUBLAS(cublasDtrsm(blas, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, n, n, &one, L,
                           (int)ld, T, n));
    SPB_CUDA(cudaMemcpy2DAsync(L, ld * sizeof(double), T, (size_t)n * sizeof(double), (size_t)n * sizeof(double), n,
                               cudaMemcpyDeviceToDevice, s));
    return;
  }
  const int h = (int)(((size_t)(n / 2) + 63) / 64 * 64);
  const int r = n - h;
  double* L21 = L + h;
  double* L22 = L + (size_t)h * ld + h;
  trtri_lower_gemm(blas, s, L, h, ld, T);
  trtri_lower_gemm(blas, s, L22, r, ld, T);
  // T = L21 V11 (r x h): column strip [j0, j0 + nb) of the lower-triangular V11 is zero above row j0
  for (int j0 = 0; j0 < h; j0 += kTriStrip) {
    const int nb = std::min(kTriStrip, h - j0);
    SPB_CUBLAS(cublasDgemm(blas, CUBLAS_OP_N, CUBLAS_OP_N, r, nb, h - j0, &one, L21 + (size_t)j0 * ld, (int)ld,
                           L + (size_t)j0 * ld + j0, (int)ld, &zero, T + (size_t)j0 * r, r));
  }
  // V21 = -V22 T (r x h): row strip [i0, i0 + nb) of the lower-triangular V22 is zero right of column i0 + nb
  for (int i0 = 0; i0 < r; i0 += kTriStrip) {
    const int nb = std::min(kTriStrip, r - i0);
    SPB_CUBLAS(cublasDgemm(blas, CUBLAS_OP_N, CUBLAS_OP_N, nb, h, i0 + nb, &mone, L22 + i0, (int)ld, T, r, &zero,
                           L21 + i0, (int)ld));
  }
}

// upper(C) = V'V for a lower-triangular V (n x n, strict upper triangle stored as zeros).  Column strip [j0, j0 + nb)
// of C needs the rows 0 .. j0 + nb - 1 (on / above the diagonal block) and, because V[k, j] = 0 for k < j, only the
// contraction range k >= j0: 0.44 n^3 flop in 10 strips instead of the n^3 of a product that treats V as dense.
// The part of C below the diagonal blocks is not written.
void vtv_upper_strips(cublasHandle_t blas, const double* V, int n, size_t ld, double* C, size_t ldc) {
  const double one = 1.0, zero = 0.0;
  for (int j0 = 0; j0 < n; j0 += kTriStrip) {
    const int nb = std::min(kTriStrip, n - j0);
    SPB_CUBLAS(cublasDgemm(blas, CUBLAS_OP_T, CUBLAS_OP_N, j0 + nb, nb, n - j0, &one, V + j0, (int)ld,
                           V + (size_t)j0 * ld + j0, (int)ld, &zero, C + (size_t)j0 * ldc, (int)ldc));
  }
}

// B (n x n, SPD, upper triangle valid) -> B^-1 (full storage) through its Cholesky factor: B = L L', V = L^-1,
// B^-1 = V'V.  Backward stable where the recursive Schur-complement inverse (spd_inverse_recursive) is not: for the
// rotated TPS block (cond ~1e8) the start X0 built from it has |I - A X0| = 3e-8 at p = 10 000 instead of 5e-4
// (measured), so ONE Newton-Schulz step is enough.  potrf: cuSOLVER; everything else GEMM-rate.
// C: n x n scratch (ldc), T: scratch of >= n^2 / 4 + 512^2 doubles.
void spd_inverse_cholesky(OmegaRes& R, int* flag, double* B, int n, size_t ld, double* C, size_t ldc, double* T) {
  cudaStream_t s = R.stream;
  launch_symmatu_inplace(B, n, ld, 0.0, s);                                   // potrf reads the lower triangle
  int lw = 0;
  SPB_CUSOLVER(cusolverDnDpotrf_bufferSize(R.solver, CUBLAS_FILL_MODE_LOWER, n, B, (int)ld, &lw));
  R.work.ensure((size_t)lw);
  SPB_CUSOLVER(cusolverDnDpotrf(R.solver, CUBLAS_FILL_MODE_LOWER, n, B, (int)ld, R.work.get(), lw, R.info.get()));
  launch_check_info(R.info.get(), flag, SPB_ERR_NOT_POSDEF, s);
  zero_strict_upper_kernel<<<dim3(ceil_div(n, 128), ceil_div(n, 8)), 128, 0, s>>>(B, ld, n);
  count_launch();
  trtri_lower_gemm(R.blas, s, B, n, ld, T);
  SPB_CUDA(cudaMemset2DAsync(C, ldc * sizeof(double), 0, (size_t)n * sizeof(double), n, s));
  vtv_upper_strips(R.blas, B, n, ld, C, ldc);                                 // upper(V'V)
  SPB_CUDA(cudaMemcpy2DAsync(B, ld * sizeof(double), C, ldc * sizeof(double), (size_t)n * sizeof(double), n,
                             cudaMemcpyDeviceToDevice, s));
  launch_symmatu_inplace(B, n, ld, 0.0, s);
  SPB_CUDA(cudaGetLastError());
}

// ---- X0 through the symmetric positive definite part of the bordered system ---------------------------------
// v = Householder vector j of a geqrf-factored p x m matrix: zeros above row j, 1 at row j, the stored tail below
__global__ void hh_vector_kernel(const double* __restrict__ V, size_t ldv, int p, int j, double* __restrict__ v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p) return;
  v[i] = i < j ? 0.0 : (i == j ? 1.0 : V[(size_t)j * ldv + i]);
}

// Two-sided reflector H = I - tau v v' on a symmetric matrix P:  H P H = P - tau (v u' + u v'),  u = w - (tau alpha / 2) v,
// w = P v, alpha = v'w.  One CTA: alpha, then u and vt = -tau v  (so that P += vt u' + u vt' is one dsyr2 with alpha = 1).
__global__ void __launch_bounds__(1024) hh_prep_kernel(const double* __restrict__ v, const double* __restrict__ w,
                                                       const double* __restrict__ tau, int p, double* __restrict__ u,
                                                       double* __restrict__ vt) {
  __shared__ double red[32];
  __shared__ double alpha_s;
  double t = 0.0;
  for (int i = threadIdx.x; i < p; i += 1024) t = fma(v[i], w[i], t);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x < 32) {
    double x = red[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (threadIdx.x == 0) alpha_s = x;
  }
  __syncthreads();
  const double tj = *tau;
  const double c = 0.5 * tj * alpha_s;
  for (int i = threadIdx.x; i < p; i += 1024) {
    u[i] = w[i] - c * v[i];
    vt[i] = -tj * v[i];
  }
}

// The 2m x 2m Schur complement of the rotated bordered system and its inverse (one thread; m <= 4):
//   S = [[F11 - C, R], [R', eps I]],  F11 = rotated block (upper valid, ridge included), C = F12 B^-1 F21, R = upper
//   triangular factor of T.  Xs = S^-1 by Gauss-Jordan with partial pivoting.  Xs: 2m x 2m, column-major, ld 2m.
__global__ void schur_small_kernel(const double* __restrict__ F, size_t ldf, const double* __restrict__ Cm,
                                   const double* __restrict__ Rm, size_t ldr, int m, double eps, double* __restrict__ Xs,
                                   int* __restrict__ flag, int code) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int n = 2 * m;
  double A[8][16];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < 2 * n; ++j) A[i][j] = (j == n + i) ? 1.0 : 0.0;
  for (int a = 0; a < m; ++a)
    for (int b = 0; b < m; ++b) {
      const double f = (a <= b) ? F[(size_t)b * ldf + a] : F[(size_t)a * ldf + b];
      A[a][b] = f - Cm[(size_t)b * m + a];
      const double r = (a <= b) ? Rm[(size_t)b * ldr + a] : 0.0;       // R upper triangular
      A[a][m + b] = r;
      A[m + b][a] = r;
      A[m + a][m + b] = (a == b) ? eps : 0.0;
    }
  bool bad = false;
  for (int c = 0; c < n; ++c) {
    int piv = c;
    for (int r = c + 1; r < n; ++r)
      if (fabs(A[r][c]) > fabs(A[piv][c])) piv = r;
    if (!(fabs(A[piv][c]) > 0.0)) { bad = true; break; }
    if (piv != c)
      for (int j = 0; j < 2 * n; ++j) { const double t = A[c][j]; A[c][j] = A[piv][j]; A[piv][j] = t; }
    const double inv = 1.0 / A[c][c];
    for (int j = 0; j < 2 * n; ++j) A[c][j] *= inv;
    for (int r = 0; r < n; ++r) {
      if (r == c) continue;
      const double f = A[r][c];
      if (f != 0.0)
        for (int j = 0; j < 2 * n; ++j) A[r][j] = fma(-f, A[c][j], A[r][j]);
    }
  }
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) Xs[(size_t)j * n + i] = A[i][n + j];
  if (bad && *flag == 0) *flag = code;
}

// dst (rows x cols, ldd) <- alpha * src (rows x cols, lds)
__global__ void __launch_bounds__(128) scaled_copy2d_kernel(const double* __restrict__ src, size_t lds, int rows, int cols,
                                                            double alpha, double* __restrict__ dst, size_t ldd) {
  const int i = blockIdx.x * 128 + threadIdx.x;
  if (i >= rows) return;
  for (int j = blockIdx.y; j < cols; j += gridDim.y) dst[(size_t)j * ldd + i] = alpha * src[(size_t)j * lds + i];
}

int split_bits(int q) {
  int lg = 0;
  while ((1 << lg) < q) ++lg;
  return (53 - lg) / 2;
}

}  // namespace

// X0 = A^-1 (q x q, symmetric, full storage) WITHOUT a pivoted LU.  The TPS kernels are conditionally positive definite
// (order 2 for r^3 and r^2 log r, order 1 for -r; T = [1 X] spans the polynomials they exclude), so in an orthonormal
// basis Q = [Q1 Q2] with span(Q1) = span(T) the block B = Q2'(E + eps I)Q2 is symmetric POSITIVE DEFINITE and carries
// all the p^3 work: it goes through the GEMM-rate recursive SPD inverse of spd_inverse.cu (p^3 flop at ~24 TFLOP/s:
// 42 ms at p = 10 000) instead of cuSOLVER getrf + getrs (2.7 p^3 at 10 / 25 TFLOP/s: 145 ms).  Q is three Householder
// reflectors (QR of T), applied to the symmetric matrix as rank-2 updates (O(p^2) each); the rest of the bordered
// system is a 2m x 2m Schur complement (m = d + 1), solved in a one-thread kernel.  cond(B) ~ |E| / eps, the same class
// as the bordered system, so X0 is as good a starting point for the Newton-Schulz step as the LU's (the step, with its
// exactly formed residual against A ITSELF, is what fixes the accuracy either way).  Pbuf: q x q work (destroyed).
void omega_x0_spd(OmegaRes& R, const double* x_dev, int p, int d, double* Pbuf, double* X0, double* inv_ws,
                  OmegaTemps& tmp) {
  const int m = d + 1, q = p + m, pm = p - m, n2 = 2 * m;
  // failures of this route (a non-positive pivot: B is positive definite in exact arithmetic, but with cond(B) ~ 1e11
  // -- one-dimensional sites -- rounding can break it) go to a flag of their own: the caller then falls back to the LU
  tmp.spdflag.ensure(4);
  int* flag = tmp.spdflag.get();
  SPB_CUDA(cudaMemsetAsync(flag, 0, 4 * sizeof(int), R.stream));
  const size_t ldq = padded_ld(q);
  cudaStream_t s = R.stream;
  const double one = 1.0, zero = 0.0, mone = -1.0;
  // small buffers: T / V (p x m, ld ldq), tau (m), v, w, u, vt (p each), Gt (m x pm), Ht (m x pm), Cm (m x m), Xs (2m x 2m),
  // Mrot (p x m, ld ldq)
  tmp.hh.ensure((size_t)2 * m * ldq + 4 * ldq + (size_t)2 * m * ldq + 256);
  double* Tm = tmp.hh.get();
  double* Mrot = Tm + (size_t)m * ldq;
  double* v = Mrot + (size_t)m * ldq;
  double* w = v + ldq;
  double* u = w + ldq;
  double* vt = u + ldq;
  double* Gt = vt + ldq;                       // m x pm, ld m
  double* Ht = Gt + (size_t)m * ldq;           // m x pm, ld m
  double* tau = Ht + (size_t)m * ldq;
  double* Cm = tau + 16;
  double* Xs = Cm + 32;
  // 1. bordered matrix; T out of its border columns; QR of T
  launch_tps_fill(x_dev, p, d, Pbuf, ldq, 1e-8, true, s);
  SPB_CUDA(cudaMemcpy2DAsync(Tm, ldq * sizeof(double), Pbuf + (size_t)p * ldq, ldq * sizeof(double), (size_t)p * sizeof(double),
                             m, cudaMemcpyDeviceToDevice, s));
  int lw1 = 0, lw2 = 0;
  SPB_CUSOLVER(cusolverDnDgeqrf_bufferSize(R.solver, p, m, Tm, (int)ldq, &lw1));
  SPB_CUSOLVER(cusolverDnDormqr_bufferSize(R.solver, CUBLAS_SIDE_LEFT, CUBLAS_OP_N, p, m, m, Tm, (int)ldq, tau, Mrot, (int)ldq,
                                           &lw2));
  int lw3 = 0;                               // potrf of the rotated block: sized here so that the workspace never moves
  SPB_CUSOLVER(cusolverDnDpotrf_bufferSize(R.solver, CUBLAS_FILL_MODE_LOWER, pm, Pbuf + (size_t)m * ldq + m, (int)ldq, &lw3));
  R.work.ensure((size_t)std::max(std::max(lw1, lw2), lw3));
  SPB_CUSOLVER(cusolverDnDgeqrf(R.solver, p, m, Tm, (int)ldq, tau, R.work.get(), lw1, R.info.get()));
  // 2. P <- Q'(E + eps I)Q on the upper triangle: reflectors 0 .. m-1
  auto two_sided = [&](int j) {
    hh_vector_kernel<<<ceil_div(p, 256), 256, 0, s>>>(Tm, ldq, p, j, v);
    SPB_CUBLAS(cublasDsymv(R.blas, CUBLAS_FILL_MODE_UPPER, p, &one, Pbuf, (int)ldq, v, 1, &zero, w, 1));
    hh_prep_kernel<<<1, 1024, 0, s>>>(v, w, tau + j, p, u, vt);
    SPB_CUBLAS(cublasDsyr2(R.blas, CUBLAS_FILL_MODE_UPPER, p, &one, vt, 1, u, 1, Pbuf, (int)ldq));
    count_launch(2);
  };
  for (int j = 0; j < m; ++j) two_sided(j);
  // 3. B^-1 in place (B = rows / columns m .. p-1), GEMM-rate recursive SPD inverse
  double* B = Pbuf + (size_t)m * ldq + m;
  // SPB_OMEGA_SPDINV = schur: the round-2a recursive Schur-complement inverse (35 ms, but a start 1e4 times further off)
  if (!omega_spd_schur()) {
    spd_inverse_cholesky(R, flag, B, pm, ldq, X0, ldq, inv_ws);               // X0 is free until step 7
  } else {
    Exec ex{s, R.blas, flag, nullptr};
    spd_inverse_recursive(ex, B, pm, ldq, inv_ws);
  }
  // 4. Gt = F12 B^-1 (m x pm), C = F12 B^-1 F21, Schur complement and its inverse
  const double* F12 = Pbuf + (size_t)m * ldq;                      // rows 0 .. m-1, columns m .. p-1
  SPB_CUBLAS(cublasDgemm(R.blas, CUBLAS_OP_N, CUBLAS_OP_N, m, pm, pm, &one, F12, (int)ldq, B, (int)ldq, &zero, Gt, m));
  SPB_CUBLAS(cublasDgemm(R.blas, CUBLAS_OP_N, CUBLAS_OP_T, m, m, pm, &one, Gt, m, F12, (int)ldq, &zero, Cm, m));
  schur_small_kernel<<<1, 1, 0, s>>>(Pbuf, ldq, Cm, Tm, ldq, m, 1e-8, Xs, flag, SPB_ERR_SINGULAR);
  count_launch();
  // 5. rotated inverse in place: X22 = B^-1 + G Xuu G', X12 = -Xuu G', X11 = Xuu;  Mrot = [Xul; -G Xul]
  SPB_CUBLAS(cublasDgemm(R.blas, CUBLAS_OP_N, CUBLAS_OP_N, m, pm, m, &one, Xs, n2, Gt, m, &zero, Ht, m));      // Xuu G'
  SPB_CUBLAS(cublasDgemm(R.blas, CUBLAS_OP_T, CUBLAS_OP_N, pm, pm, m, &one, Gt, m, Ht, m, &one, B, (int)ldq));
  scaled_copy2d_kernel<<<dim3(1, pm < 1024 ? pm : 1024), 128, 0, s>>>(Ht, m, m, pm, -1.0, Pbuf + (size_t)m * ldq, ldq);
  scaled_copy2d_kernel<<<dim3(1, m), 128, 0, s>>>(Xs, n2, m, m, 1.0, Pbuf, ldq);
  scaled_copy2d_kernel<<<dim3(1, m), 128, 0, s>>>(Xs + (size_t)m * n2, n2, m, m, 1.0, Mrot, ldq);            // Xul
  SPB_CUBLAS(cublasDgemm(R.blas, CUBLAS_OP_T, CUBLAS_OP_N, pm, m, m, &mone, Gt, m, Xs + (size_t)m * n2, n2, &zero,
                         Mrot + m, (int)ldq));                                                                 // -G Xul
  count_launch(3);
  // 6. back to the original basis: Lp0 = Q X Q' (reflectors m-1 .. 0), M0 = Q Mrot
  for (int j = m - 1; j >= 0; --j) two_sided(j);
  SPB_CUSOLVER(cusolverDnDormqr(R.solver, CUBLAS_SIDE_LEFT, CUBLAS_OP_N, p, m, m, Tm, (int)ldq, tau, Mrot, (int)ldq,
                                R.work.get(), lw2, R.info.get()));
  // 7. X0 = [[Lp0, M0], [M0', Xll]] (upper triangle, then mirrored)
  SPB_CUDA(cudaMemcpy2DAsync(X0, ldq * sizeof(double), Pbuf, ldq * sizeof(double), (size_t)p * sizeof(double), p,
                             cudaMemcpyDeviceToDevice, s));
  SPB_CUDA(cudaMemcpy2DAsync(X0 + (size_t)p * ldq, ldq * sizeof(double), Mrot, ldq * sizeof(double), (size_t)p * sizeof(double),
                             m, cudaMemcpyDeviceToDevice, s));
  scaled_copy2d_kernel<<<dim3(1, m), 128, 0, s>>>(Xs + (size_t)m * n2 + m, n2, m, m, 1.0, X0 + (size_t)p * ldq + p, ldq);
  count_launch();
  launch_symmatu_inplace(X0, q, ldq, 0.0, s);
  SPB_CUDA(cudaGetLastError());
}


// SPB_OMEGA_RESID=dgemm keeps the refinement products in FP64 (three head/tail DGEMMs + one DGEMM); the default
// runs them on the INT8 tensor cores (ozaki_int8.cu)
bool omega_int8_enabled() {
  static const bool fp64 = [] { const char* e = std::getenv("SPB_OMEGA_RESID"); return e && std::string(e) == "dgemm"; }();
  return !fp64;
}

// Columns [c0, c0 + nc) of one Newton-Schulz step on INT8 digit planes:
//   C1[:, cols]   = I - A X0[:, cols]              (S planes per operand, ~2^(-7 S) of rowmax * colmax)
//   Xout[:, cols] = X0[:, cols] + X0 C1[:, cols]   (Sc planes)
// A = T.L (the bordered matrix, symmetric, full); T.T1 is the FP64 accumulator.  X0 need NOT be symmetric: an LU
// inverse is accurate column by column (|I - A X0| ~ 1e-8 .. 1e-6), mirroring its upper triangle mixes the columns'
// forward errors and lifts the residual to ~cond * eps * sqrt(q) (0.1 at p = 10 000, measured) -- and the step only
// contracts the error by the size of that residual.  The rows of X0 (left operand of the correction) are sliced from
// a transposed copy parked in Xout (x0_symmetric: X0 itself).
// upper_result: only the part of Xout on / above the diagonal is read afterwards (the one-piece build mirrors the upper
// triangle of its last iterate): the correction then runs in column strips over the rows above each strip's end.
constexpr int kOzakiStrip = 2048;
static void refine_cols_int8(OmegaRes& R, OmegaTemps& T, int q, size_t ldq, int c0, int nc, const double* X0,
                             double* Xout, int S, bool x0_symmetric, bool x0_lu_accurate, bool upper_result = false) {
  if (nc <= 0) return;
  // correction planes: the step's |I - A X0| is at most 1e-5 (omega_build_ok), so 5 planes (2^-33 of the term sizes)
  // leave < 1e-15 |X0|; measured at p = 10 000 (tools/omega_planes_probe.py): 6 -> 5 -> 4 planes move Omega by
  // 0 / 3e-15 / 1e-14, while one residual plane less (12 -> 11) moves it by 1e-12 and one product plane less by 3e-11
  const int Sc = ozaki_env_slices("SPB_OZAKI_CORR_SLICES", x0_lu_accurate ? 5 : 8, 2, 16);
  const int Sm = std::max(S, Sc);
  const OzakiGeom g = ozaki_geom(q, nc);
  cudaStream_t s = R.stream;
  T.oz_left.ensure((size_t)Sm * g.kq * g.mp);
  T.oz_right.ensure((size_t)Sm * g.kq * g.np);
  T.oz_c.ensure((size_t)g.mp * g.np);
  T.oz_ea.ensure(g.mp);
  T.oz_eb.ensure(g.np);
  T.oz_bad.ensure(1);
  SPB_CUDA(cudaMemsetAsync(T.oz_bad.get(), 0, sizeof(int), s));
  const size_t o = (size_t)c0 * ldq;
  const double* A = T.L.get();
  double *acc = T.T1.get() + o, *C1 = T.C1.get() + o;
  int8_t *left = T.oz_left.get(), *right = T.oz_right.get();
  int *ea = T.oz_ea.get(), *eb = T.oz_eb.get(), *bad = T.oz_bad.get();
  ozaki_slice(s, A, ldq, q, q, g.mp, S, false, left, g.kq, ea, bad);              // rows of A = its columns
  ozaki_slice(s, X0 + o, ldq, q, nc, g.np, S, true, right, g.kq, eb, bad);
  ozaki_product(R.blas, s, g, left, S, right, S, S, q, nc, T.oz_c.get(), acc, ldq, ea, eb, 0, c0, 0.0, nullptr, 0, C1, ldq, bad);
  const double* X0rows = X0;
  if (!x0_symmetric) {
    ozaki_transpose(s, X0, ldq, q, Xout, ldq);                                    // overwritten by the epilogue below
    X0rows = Xout;
  }
  ozaki_slice(s, X0rows, ldq, q, q, g.mp, Sc, false, left, g.kq, ea, bad);        // rows of X0
  ozaki_slice(s, C1, ldq, q, nc, g.np, Sc, true, right, g.kq, eb, bad);
  if (upper_result && c0 == 0 && nc == q) {
    ozaki_product_upper(R.blas, s, g, left, Sc, right, Sc, Sc, q, nc, T.oz_c.get(), acc, ldq, ea, eb, 1, c0, 1.0, X0 + o, ldq,
                        Xout + o, ldq, bad, kOzakiStrip);
  } else {
    ozaki_product(R.blas, s, g, left, Sc, right, Sc, Sc, q, nc, T.oz_c.get(), acc, ldq, ea, eb, 1, c0, 1.0, X0 + o, ldq,
                  Xout + o, ldq, bad);
  }
}

// Newton-Schulz steps of the build: ONE.  The step takes the forward error E0 of X0 to E0 A E0; a second step cannot
// improve on that, because X1 rounded to FP64 already has a residual of ~cond * eps * sqrt(q) (measured: 0.5 at
// p = 10 000).  SPB_OMEGA_STEPS = 2..4 (INT8 planes only, earlier steps with fewer planes) is a debugging aid.
int omega_refine_steps(bool spd_start) {
  static const int n = [] {
    const char* e = std::getenv("SPB_OMEGA_STEPS");
    return e ? std::min(4, std::max(1, std::atoi(e))) : 0;
  }();
  return n > 0 ? n : ((spd_start && omega_spd_schur()) ? 2 : 1);
}
static int refine_planes(int step, int nsteps) {
  const int S = ozaki_env_slices("SPB_OZAKI_SLICES", 12, 4, 16);
  const int S1 = ozaki_env_slices("SPB_OZAKI_FIRST_SLICES", 9, 4, 16);
  return step + 1 < nsteps ? std::min(S, S1) : S;
}

// max |R| of the residual columns [0, ncols) -> resid[2 step] (columns < p: what the Lp block of the next iterate
// depends on) and resid[2 step + 1] (all columns; the border columns carry entries ~1 / 1e-8 and are much larger)
static void record_residual(OmegaTemps& T, cudaStream_t s, const double* C1, size_t ldq, int q, int p, int step, double* mx) {
  T.resid.ensure(8);
  col_absmax_kernel<<<q, 256, 0, s>>>(C1, ldq, q, mx);
  col_absmax_kernel<<<1, 256, 0, s>>>(mx, (size_t)p, p, T.resid.get() + 2 * step);
  col_absmax_kernel<<<1, 256, 0, s>>>(mx, (size_t)q, q, T.resid.get() + 2 * step + 1);
  count_launch(3);
}

// After the build's stream has been waited for: did the SPD route hold (no bad pivot) and is |I - A X0| small enough
// for one Newton-Schulz step (error after the step ~ |I - A X0|^2)?
bool omega_build_ok(OmegaTemps& tmp, double* resid_out) {
  if (resid_out) *resid_out = -1.0;
  double r[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const int steps = std::max(1, tmp.steps_done);
  if (tmp.resid.get() != nullptr) {
    SPB_CUDA(cudaMemcpy(r, tmp.resid.get(), sizeof(double) * 2 * steps, cudaMemcpyDeviceToHost));
    if (resid_out) *resid_out = r[0];
    if (trace_on())
      for (int k = 0; k < steps; ++k)
        fprintf(stderr, "[spb omega] step %d: max|I - A X| = %.3e over the Lp columns, %.3e over all\n", k, r[2 * k], r[2 * k + 1]);
  }
  if (tmp.spdflag.get() != nullptr && tmp.used_spd) {
    int hf[4] = {0, 0, 0, 0};
    SPB_CUDA(cudaMemcpy(hf, tmp.spdflag.get(), sizeof(hf), cudaMemcpyDeviceToHost));
    if (hf[0] != 0) return false;
  }
  if (tmp.resid.get() != nullptr && tmp.used_spd) {
    // the start must be close enough for its step(s): 2.3e-8 (Cholesky inverse) / 4.9e-4 (Schur inverse) at p = 10 000
    if (!(r[0] < (omega_spd_schur() ? 5e-3 : 1e-5))) return false;
  }
  return true;
}

// Which start?  Measured on B200 at p = 10 000 (q = 10 003), checked against an extended-precision ground truth at
// p = 4000 (tools/omega_truth.py) and against each other:
//   LU (cuSOLVER getrf + getrs on I)             149 ms, |I - A X0| = 1.2e-7 -> one Newton-Schulz step -> 1e-11 in Omega
//   SPD route, Cholesky inverse (default)          80 ms, |I - A X0| = 2.3e-8 -> one step, 4e-12 from the LU result
//   SPD route, recursive Schur-complement inverse  49 ms, |I - A X0| = 4.9e-4 -> needs two steps (141 ms) and tuned
//                                                  plane counts; kept as SPB_OMEGA_SPDINV=schur
// The SPD route is taken from p = 2048 on (below that the LU is as fast); finish_omega() / spb_tps_matrix() repeat the
// build with the LU when the Cholesky factorisation reports a non-positive pivot (one-dimensional sites at cond ~1e11)
// or the start is further off than 1e-5.  SPB_OMEGA_X0 = lu | spd forces one.
bool omega_x0_spd_enabled(int p, int d) {
  static const int forced = [] {
    const char* e = std::getenv("SPB_OMEGA_X0");
    return !e ? 0 : (std::string(e) == "spd" ? 1 : (std::string(e) == "lu" ? 2 : 0));
  }();
  if (forced == 2 || !omega_int8_enabled() || p < 64 || p <= 4 * (d + 1)) return false;
  return forced == 1 || p >= 2048;
}

double omega_refined_flops(int p) {          // X0 1 (SPD inverse; LU + solve: 2.7) + residual 6 + correction 1 + product 1
  return 9.0 * (double)p * p * p;
}

void Engine::build_omega_refined(OmegaRes& R, int* flag, const double* x_dev, int p, int d, double* omega_dev,
                                 size_t ld, bool upper_only, OmegaTemps& tmp) {
  const int q = p + d + 1;
  const size_t ldq = padded_ld(q);
  const size_t qq = ldq * (size_t)q;
  tmp.L.ensure(qq);        // A -> LU factors -> head / tail of A
  tmp.Bm.ensure(qq);       // X0 (symmetrised)
  tmp.T1.ensure(qq);       // head / tail of X0
  tmp.C1.ensure(qq);       // I - A1 X1 -> residual
  tmp.C2.ensure(qq);       // A1 X2 + A2 X0 -> X1
  tmp.vec.ensure(4 * (size_t)q);
  double *A = tmp.L.get(), *X0 = tmp.Bm.get(), *S2 = tmp.T1.get(), *C1 = tmp.C1.get(), *C2 = tmp.C2.get();
  double *mxA = tmp.vec.get(), *sgA = mxA + q, *mxX = sgA + q, *sgX = mxX + q;
  cudaStream_t s = R.stream;
  const double one = 1.0, zero = 0.0, mone = -1.0;
  const dim3 g2(ceil_div(q, 128), ceil_div(q, 8));

  // 1. X0 = A^-1: through the SPD part of the system (omega_x0_spd), or by LU with partial pivoting (small p,
  //    SPB_OMEGA_X0=lu); all q columns (the border block M is needed by step 4), symmetric
  tmp.used_spd = !tmp.force_lu && omega_x0_spd_enabled(p, d);
  if (tmp.used_spd) {
    omega_x0_spd(R, x_dev, p, d, A, X0, C1, tmp);
  } else {
    launch_tps_fill(x_dev, p, d, A, ldq, 1e-8, true, s);
    launch_set_identity(X0, q, q, ldq, s);
    int lwork = 0;
    SPB_CUSOLVER(cusolverDnDgetrf_bufferSize(R.solver, q, q, A, (int)ldq, &lwork));
    R.ipiv.ensure(q);
    R.work.ensure((size_t)lwork);
    SPB_CUSOLVER(cusolverDnDgetrf(R.solver, q, q, A, (int)ldq, R.work.get(), R.ipiv.get(), R.info.get()));
    launch_check_info(R.info.get(), flag, SPB_ERR_SINGULAR, s);
    SPB_CUSOLVER(cusolverDnDgetrs(R.solver, CUBLAS_OP_N, q, q, A, (int)ldq, R.ipiv.get(), X0, (int)ldq, R.info.get()));
    // the FP64 products below read X0 through its upper triangle; the INT8 step takes it as it is (see refine_cols_int8)
    if (!omega_int8_enabled()) launch_symmatu_inplace(X0, q, ldq, 0.0, s);
  }

  auto lap = [&](const char* label) {               // SPB_TRACE=1: synchronous per-step timing
    if (!trace_on()) return;
    static auto last = std::chrono::steady_clock::now();
    cudaStreamSynchronize(s);
    auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[spb omega] %-12s +%8.3f ms\n", label, std::chrono::duration<double, std::milli>(now - last).count());
    last = now;
  };
  lap("X0");
  // 2 + 3. residual of the symmetrised X0 and the Newton-Schulz correction
  const bool int8 = omega_int8_enabled();
  const int nsteps = int8 ? omega_refine_steps(tmp.used_spd) : 1;
  tmp.steps_done = nsteps;
  if (int8) {
    launch_tps_fill(x_dev, p, d, A, ldq, 1e-8, true, s);       // A again (the LU overwrote it)
    for (int step = 0; step < nsteps; ++step) {
      // the iterate moves back into X0's buffer AS IT IS: mirroring its upper triangle would put its asymmetry (the
      // size of its forward error, ~1e-9) into the next residual as A * 1e-9 |X| ~ 1, and the step would not contract
      if (step > 0) std::swap(X0, C2);
      // the last iterate is only read through its upper triangle (mirrored below)
      refine_cols_int8(R, tmp, q, ldq, 0, q, X0, C2, refine_planes(step, nsteps),     // C1 = I - A X0, C2 = X0 + X0 C1
                       tmp.used_spd && step == 0, !(tmp.used_spd && omega_spd_schur()), step + 1 == nsteps);
      record_residual(tmp, s, C1, ldq, q, p, step, mxX);
    }
  } else {
    // 2. residual of the symmetrised X0 with an exact head product
    const int beta = split_bits(q);
    launch_tps_fill(x_dev, p, d, A, ldq, 1e-8, true, s);         // A again (the LU overwrote it)
    col_absmax_kernel<<<q, 256, 0, s>>>(A, ldq, q, mxA);         // A is symmetric: column maxima = row maxima
    col_absmax_kernel<<<q, 256, 0, s>>>(X0, ldq, q, mxX);
    split_sigma_kernel<<<ceil_div(q, 256), 256, 0, s>>>(mxA, q, beta, sgA);
    split_sigma_kernel<<<ceil_div(q, 256), 256, 0, s>>>(mxX, q, beta, sgX);
    split_kernel<<<g2, 128, 0, s>>>(A, ldq, q, q, sgA, 1, 0, A, ldq);      // A  <- A1 (head, per row)
    split_kernel<<<g2, 128, 0, s>>>(X0, ldq, q, q, sgX, 0, 0, S2, ldq);    // S2 <- X1 (head, per column)
    launch_set_identity(C1, q, q, ldq, s);
    count_launch(6);
    SPB_CUBLAS(cublasDgemm(R.blas, CUBLAS_OP_N, CUBLAS_OP_N, q, q, q, &mone, A, (int)ldq, S2, (int)ldq, &one, C1,
                           (int)ldq));                                     // C1 = I - A1 X1   (exact)
    split_kernel<<<g2, 128, 0, s>>>(X0, ldq, q, q, sgX, 0, 1, S2, ldq);    // S2 <- X2 = X0 - X1
    SPB_CUBLAS(cublasDgemm(R.blas, CUBLAS_OP_N, CUBLAS_OP_N, q, q, q, &one, A, (int)ldq, S2, (int)ldq, &zero, C2,
                           (int)ldq));                                     // C2 = A1 X2
    launch_tps_fill(x_dev, p, d, A, ldq, 1e-8, true, s);
    split_kernel<<<g2, 128, 0, s>>>(A, ldq, q, q, sgA, 1, 1, A, ldq);      // A  <- A2 = A - A1
    SPB_CUBLAS(cublasDgemm(R.blas, CUBLAS_OP_N, CUBLAS_OP_N, q, q, q, &one, A, (int)ldq, X0, (int)ldq, &one, C2,
                           (int)ldq));                                     // C2 += A2 X0
    sub_inplace_kernel<<<g2, 128, 0, s>>>(C1, C2, ldq, q, q);              // C1 = I - A X0
  }
  // max |I - A X0|: the Newton-Schulz step squares it; checked on the host when the build is waited for
  if (!int8) record_residual(tmp, s, C1, ldq, q, p, 0, mxX);

  if (!int8) {
    // 3. X1 = X0 + X0 (I - A X0): tiles on / above the diagonal only (X0 symmetric => X0 R = X0' R is symmetric)
    SPB_CUDA(cudaMemcpyAsync(C2, X0, sizeof(double) * qq, cudaMemcpyDeviceToDevice, s));
    gemm_upper_tn(R.blas, q, q, 1.0, X0, ldq, C1, ldq, C2, ldq);
  }
  launch_symmatu_inplace(C2, q, ldq, 0.0, s);                            // full Lp for the product below
  lap(int8 ? "refine int8" : "refine fp64");

  // 4. Omega = Lp - eps Lp Lp + eps M M'  (upper tiles; Lp = C2[0:p, 0:p], M = C2[0:p, p:q])
  if (int8) {
    // the product only enters with the weight eps: 7 planes (2^-49 of rowmax * colmax) leave < 1e-14 of Omega, and the
    // 28 plane products take 22 ms at p = 10 000 where the FP64 upper-tile product takes 40
    const int Sp = ozaki_env_slices("SPB_OZAKI_PROD_SLICES", 7, 3, 16);
    const OzakiGeom g = ozaki_geom(p, p);
    ozaki_slice(s, C2, ldq, p, p, g.mp, Sp, false, tmp.oz_left.get(), g.kq, tmp.oz_ea.get(), tmp.oz_bad.get());
    ozaki_slice(s, C2, ldq, p, p, g.np, Sp, true, tmp.oz_right.get(), g.kq, tmp.oz_eb.get(), tmp.oz_bad.get());
    // Omega is symmetric and only its upper triangle is read (symmatu, here or in prepare()): upper strips
    ozaki_product_upper(R.blas, s, g, tmp.oz_left.get(), Sp, tmp.oz_right.get(), Sp, Sp, p, p, tmp.oz_c.get(), S2, ldq,
                        tmp.oz_ea.get(), tmp.oz_eb.get(), 1, 0, -1e-8, C2, ldq, omega_dev, ld, tmp.oz_bad.get(), kOzakiStrip);
  } else {
    SPB_CUDA(cudaMemcpy2DAsync(omega_dev, ld * sizeof(double), C2, ldq * sizeof(double), (size_t)p * sizeof(double), p,
                               cudaMemcpyDeviceToDevice, s));
    gemm_upper_tn(R.blas, p, p, -1e-8, C2, ldq, C2, ldq, omega_dev, ld);
  }
  const double eps = 1e-8;
  SPB_CUBLAS(cublasDsyrk(R.blas, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_N, p, d + 1, &eps, C2 + (size_t)p * ldq, (int)ldq,
                         &one, omega_dev, (int)ld));
  if (!upper_only) launch_symmatu_inplace(omega_dev, p, ld, 0.0, s);     // the exported matrix: symmetric
  lap("product");
  SPB_CUDA(cudaGetLastError());
}

// ---- the same build in column blocks, for the ranks of a multi-GPU job ----------------
// Replicated: the LU (2/3 p^3) and the O(p^2) passes (symmetrisation, splitting).  Split by columns: the solves,
// the three residual products, the correction and the final product -- 10 p^3 / N per rank.  Three exchanges of
// p^2 doubles (X0, X1, Omega), each one NCCL broadcast per owner straight from / into these buffers.
bool Engine::omega_begin_refined() {
  OmegaRes& R = ctx_.omega();
  cudaStream_t s = R.stream;
  SPB_CUDA(cudaStreamSynchronize(ctx_.stream));
  const int q = p_ + d_ + 1;
  const size_t ldq = padded_ld(q), qq = ldq * (size_t)q;
  omega_u_.alloc(ld_ * (size_t)p_);
  omega_diag_raw_.alloc(p_);
  omega_tmp_.reset(new OmegaTemps());
  OmegaTemps& T = *omega_tmp_;
  T.L.ensure(qq); T.Bm.ensure(qq); T.T1.ensure(qq); T.C1.ensure(qq); T.C2.ensure(qq); T.vec.ensure(4 * (size_t)q);
  clock_.begin(PH_OMEGA, s);
  omega_x0_is_spd_ = omega_x0_spd_enabled(p_, d_);
  if (omega_x0_is_spd_) {
    // X0 is built whole on every rank (p^3 flop at GEMM rate, deterministic: the ranks hold identical copies, so the
    // first exchange is not needed); only the refinement and the final product are split by columns.  Enqueued only:
    // whether the SPD route held is read by omega_x0_resolve() at the next omega_* call, so that the caller can run
    // the SVD starts (prepare_starts) under it.
    omega_x0_spd(R, dX_.get(), p_, d_, T.L.get(), T.Bm.get(), T.C1.get(), T);
    omega_x0_check_pending_ = true;
  } else {
    omega_lu_factor();
  }
  clock_.end(s);
  stats_.p3_flops += omega_refined_flops(p_);
  omega_cols_open_ = true;
  return true;
}

// LU of the bordered system for the column-split build (the start when the SPD route is off or broke down)
void Engine::omega_lu_factor() {
  OmegaRes& R = ctx_.omega();
  cudaStream_t s = R.stream;
  OmegaTemps& T = *omega_tmp_;
  const int q = p_ + d_ + 1;
  const size_t ldq = padded_ld(q);
  launch_tps_fill(dX_.get(), p_, d_, T.L.get(), ldq, 1e-8, true, s);
  launch_set_identity(T.Bm.get(), q, q, ldq, s);
  int lwork = 0;
  SPB_CUSOLVER(cusolverDnDgetrf_bufferSize(R.solver, q, q, T.L.get(), (int)ldq, &lwork));
  R.ipiv.ensure(q);
  R.work.ensure((size_t)lwork);
  SPB_CUSOLVER(cusolverDnDgetrf(R.solver, q, q, T.L.get(), (int)ldq, R.work.get(), R.ipiv.get(), R.info.get()));
  launch_check_info(R.info.get(), ctx_.flag.get(), SPB_ERR_SINGULAR, s);
}

// Waits for the X0 enqueued by omega_begin_refined() and falls back to the LU when the Cholesky factorisation of the
// rotated block reported a non-positive pivot (not positive definite in FP64).  Every rank sees the same flag: the
// build is deterministic and replicated.
void Engine::omega_x0_resolve() {
  if (!omega_x0_check_pending_) return;
  omega_x0_check_pending_ = false;
  OmegaRes& R = ctx_.omega();
  int hf[4] = {0, 0, 0, 0};
  SPB_CUDA(cudaStreamSynchronize(R.stream));
  SPB_CUDA(cudaMemcpy(hf, omega_tmp_->spdflag.get(), sizeof(hf), cudaMemcpyDeviceToHost));
  if (hf[0] != 0) {
    omega_x0_is_spd_ = false;
    clock_.begin(PH_OMEGA, R.stream);
    omega_lu_factor();
    clock_.end(R.stream);
  }
}

bool Engine::omega_refine_cols(int c0, int c1) {
  SPB_REQUIRE(omega_cols_open_, "omega_begin() must be called first");
  SPB_REQUIRE(c0 >= 0 && c0 <= c1 && c1 <= p_, "column range out of bounds");
  if (!omega_refine_enabled()) return false;
  omega_x0_resolve();
  OmegaRes& R = ctx_.omega();
  cudaStream_t s = R.stream;
  const int q = p_ + d_ + 1;
  const size_t ldq = padded_ld(q);
  const int c1e = (c1 == p_) ? q : c1;                 // the owner of the last block also owns the border columns
  const int nc = c1e - c0;
  OmegaTemps& T = *omega_tmp_;
  // one Newton-Schulz step per call; false once all steps are done (the caller exchanges the columns after each)
  const bool int8 = omega_int8_enabled();
  const int nsteps = int8 ? omega_refine_steps(omega_x0_is_spd_) : 1;
  if (T.col_step >= nsteps) return false;
  if (T.col_step > 0) std::swap(T.Bm, T.C2);           // the exchanged iterate X1 becomes this step's X0
  double *A = T.L.get(), *X0 = T.Bm.get(), *S2 = T.T1.get(), *C1 = T.C1.get(), *C2 = T.C2.get();
  double *mxA = T.vec.get(), *sgA = mxA + q, *mxX = sgA + q, *sgX = mxX + q;
  const double one = 1.0, zero = 0.0, mone = -1.0;
  const dim3 gfull(ceil_div(q, 128), ceil_div(q, 8));
  clock_.begin(PH_OMEGA, s);
  const bool x0_symmetric = !int8 || (omega_x0_is_spd_ && T.col_step == 0);      // see build_omega_refined
  if (x0_symmetric) launch_symmatu_inplace(X0, q, ldq, 0.0, s);
  if (int8) {
    launch_tps_fill(dX_.get(), p_, d_, A, ldq, 1e-8, true, s);
    refine_cols_int8(R, T, q, ldq, c0, nc, X0, C2, refine_planes(T.col_step, nsteps), x0_symmetric, !(omega_x0_is_spd_ && omega_spd_schur()));
    ++T.col_step;
    clock_.end(s);
    SPB_CUDA(cudaGetLastError());
    return true;
  }
  ++T.col_step;
  const int beta = split_bits(q);
  launch_tps_fill(dX_.get(), p_, d_, A, ldq, 1e-8, true, s);
  col_absmax_kernel<<<q, 256, 0, s>>>(A, ldq, q, mxA);
  col_absmax_kernel<<<q, 256, 0, s>>>(X0, ldq, q, mxX);
  split_sigma_kernel<<<ceil_div(q, 256), 256, 0, s>>>(mxA, q, beta, sgA);
  split_sigma_kernel<<<ceil_div(q, 256), 256, 0, s>>>(mxX, q, beta, sgX);
  split_kernel<<<gfull, 128, 0, s>>>(A, ldq, q, q, sgA, 1, 0, A, ldq);                       // A <- A1
  count_launch(5);
  if (nc > 0) {
    const dim3 gc(ceil_div(q, 128), ceil_div(nc, 8));
    const size_t o = (size_t)c0 * ldq;
    split_kernel<<<gc, 128, 0, s>>>(X0 + o, ldq, q, nc, sgX + c0, 0, 0, S2 + o, ldq);        // X1 columns
    unit_cols_kernel<<<gc, 128, 0, s>>>(C1, ldq, q, c0, nc);
    SPB_CUBLAS(cublasDgemm(R.blas, CUBLAS_OP_N, CUBLAS_OP_N, q, nc, q, &mone, A, (int)ldq, S2 + o, (int)ldq, &one,
                           C1 + o, (int)ldq));                                                // I - A1 X1 (exact)
    split_kernel<<<gc, 128, 0, s>>>(X0 + o, ldq, q, nc, sgX + c0, 0, 1, S2 + o, ldq);        // X2 columns
    SPB_CUBLAS(cublasDgemm(R.blas, CUBLAS_OP_N, CUBLAS_OP_N, q, nc, q, &one, A, (int)ldq, S2 + o, (int)ldq, &zero,
                           C2 + o, (int)ldq));                                                // A1 X2
    launch_tps_fill(dX_.get(), p_, d_, A, ldq, 1e-8, true, s);
    split_kernel<<<gfull, 128, 0, s>>>(A, ldq, q, q, sgA, 1, 1, A, ldq);                     // A <- A2
    SPB_CUBLAS(cublasDgemm(R.blas, CUBLAS_OP_N, CUBLAS_OP_N, q, nc, q, &one, A, (int)ldq, X0 + o, (int)ldq, &one,
                           C2 + o, (int)ldq));                                                // + A2 X0
    sub_inplace_kernel<<<gc, 128, 0, s>>>(C1 + o, C2 + o, ldq, q, nc);                       // residual columns
    count_launch(5);
    SPB_CUDA(cudaMemcpyAsync(C2 + o, X0 + o, sizeof(double) * ldq * nc, cudaMemcpyDeviceToDevice, s));
    SPB_CUBLAS(cublasDgemm(R.blas, CUBLAS_OP_N, CUBLAS_OP_N, q, nc, q, &one, X0, (int)ldq, C1 + o, (int)ldq, &one,
                           C2 + o, (int)ldq));                                                // X1 = X0 + X0 R
  }
  clock_.end(s);
  SPB_CUDA(cudaGetLastError());
  return true;
}

void Engine::omega_product_cols_refined(int c0, int c1) {
  OmegaRes& R = ctx_.omega();
  cudaStream_t s = R.stream;
  const int q = p_ + d_ + 1;
  const size_t ldq = padded_ld(q);
  OmegaTemps& T = *omega_tmp_;
  double* X1 = T.C2.get();
  clock_.begin(PH_OMEGA, s);
  launch_symmatu_inplace(X1, q, ldq, 0.0, s);
  if (c1 > c0) {
    const int nc = c1 - c0;
    const double one = 1.0, eps = 1e-8, meps = -1e-8;
    double* out = omega_u_.get() + (size_t)c0 * ld_;
    // rows 0 .. c1-1 of these columns cover the upper triangle, which is all that is used downstream
    if (omega_int8_enabled()) {
      // Lp - eps Lp Lp on INT8 digit planes, as in the one-piece build: rows [0, c1) of Lp (its columns: Lp is
      // symmetric after the mirror above) against this rank's columns
      const int Sp = ozaki_env_slices("SPB_OZAKI_PROD_SLICES", 7, 3, 16);
      OzakiGeom g = ozaki_geom(p_, nc);
      g.mp = (int)(((size_t)c1 + 15) / 16 * 16);
      T.oz_left.ensure((size_t)Sp * g.kq * g.mp);
      T.oz_right.ensure((size_t)Sp * g.kq * g.np);
      T.oz_c.ensure((size_t)g.mp * g.np);
      T.oz_ea.ensure(g.mp);
      T.oz_eb.ensure(g.np);
      T.oz_bad.ensure(1);
      SPB_CUDA(cudaMemsetAsync(T.oz_bad.get(), 0, sizeof(int), s));
      ozaki_slice(s, X1, ldq, p_, c1, g.mp, Sp, false, T.oz_left.get(), g.kq, T.oz_ea.get(), T.oz_bad.get());
      ozaki_slice(s, X1 + (size_t)c0 * ldq, ldq, p_, nc, g.np, Sp, true, T.oz_right.get(), g.kq, T.oz_eb.get(),
                  T.oz_bad.get());
      ozaki_product(R.blas, s, g, T.oz_left.get(), Sp, T.oz_right.get(), Sp, Sp, c1, nc, T.oz_c.get(),
                    T.T1.get() + (size_t)c0 * ldq, ldq, T.oz_ea.get(), T.oz_eb.get(), 1, 0, -1e-8, X1 + (size_t)c0 * ldq, ldq,
                    out, ld_, T.oz_bad.get());
    } else {
      SPB_CUDA(cudaMemcpy2DAsync(out, ld_ * sizeof(double), X1 + (size_t)c0 * ldq, ldq * sizeof(double),
                                 (size_t)c1 * sizeof(double), nc, cudaMemcpyDeviceToDevice, s));
      SPB_CUBLAS(cublasDgemm(R.blas, CUBLAS_OP_T, CUBLAS_OP_N, c1, nc, p_, &meps, X1, (int)ldq, X1 + (size_t)c0 * ldq,
                             (int)ldq, &one, out, (int)ld_));                                   // - eps Lp Lp
    }
    SPB_CUBLAS(cublasDgemm(R.blas, CUBLAS_OP_N, CUBLAS_OP_T, c1, nc, d_ + 1, &eps, X1 + (size_t)p_ * ldq, (int)ldq,
                           X1 + (size_t)p_ * ldq + c0, (int)ldq, &one, out, (int)ld_));       // + eps M M'
  }
  clock_.end(s);
}

}  // namespace spb
