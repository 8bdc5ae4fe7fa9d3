// SPD inverse for the ADMM systems (reference: inv_sympd at src/RcppSpatPCA.cpp:165, 203, 397,
// 621, 661, 673), built for B200's FP64 GEMM rate.
//
// cuSOLVER's potrf + potri reach 13.7 and 6.9 TFLOP/s at p = 10 000 on B200 (measured,
// profiles/r01_inverse_blocks.txt) while cuBLAS DGEMM sustains 35.6 TFLOP/s, so the inverse is
// re-expressed as a recursive Schur-complement (block Gauss-Jordan) inversion whose flops are
// all GEMMs:
//
//     A = [A11 A12; A12' A22]     W = A11^-1 A12      S = A22 - A12' W
//     A^-1 = [A11^-1 + W S^-1 W'   -W S^-1 ;  .   S^-1]
//
// Symmetric products only compute the tiles on or above the diagonal.  Leaves (n <= 128) are
// inverted by one CTA in shared memory with a Gauss-Jordan sweep; a non-positive pivot there is
// exactly the "not positive definite" condition of a Cholesky factorisation and raises the same
// status.  For the matrices of this path (rho = 10 sigma_1^2 dominates, cond ~ 1..1e3) the
// forward error is ~cond * eps, the same class as potrf/potri.
#include <cstdlib>
#include <string>

#include "kernels.cuh"
#include "engine.cuh"

namespace spb {

namespace {

constexpr int kLeaf = 128;
constexpr int kLeafThreads = 512;
constexpr int kSymGemmMin = 1024;      // below this a symmetric product is one plain GEMM

// One CTA (32 x 16 threads): A (n x n, n <= 128, upper triangle valid) -> A^-1 (full storage) by a
// BLOCKED symmetric sweep: for each group B of 8 pivots
//     A_BB <- -A_BB^-1,   A_RB = A_BR' <- A_RB A_BB^-1,   A_RR <- A_RR - A_RB A_BB^-1 A_BR
// which keeps the matrix symmetric and ends at -A^-1.  The matrix lives in registers (thread
// (tx, ty) owns the 4 x 8 elements i = tx + 32 a, j = ty + 16 b).  Per group: the 8 owner warps
// publish the 128 x 8 panel, warp 0 sweeps the 8 x 8 pivot block, every thread forms one panel
// row of P A_BB^-1, and the rank-8 update runs from double-buffered shared memory (q-major, so
// the i-indexed reads are conflict-free and the j-indexed reads are broadcasts).  Compared with
// one barrier + one division chain per pivot this shortens the serial path ~8x
// (profiles/r01_notes.md).
constexpr int kPB = 8;
__global__ void __launch_bounds__(kLeafThreads)
leaf_inverse_kernel(double* __restrict__ A, int n, size_t ld, int* __restrict__ flag, int code) {
  __shared__ double P[2][kPB][kLeaf];     // raw panel, P[q][i] = a(i, k0 + q)
  __shared__ double S[2][kPB][kLeaf];     // P * A_BB^-1
  __shared__ double Dm[kPB][kPB + 1];     // pivot block, swept in place to -A_BB^-1
  __shared__ int bad;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int tid = ty * 32 + tx;
  if (tid == 0) bad = 0;
  double v[4][8];
#pragma unroll
  for (int b = 0; b < 8; ++b)
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int i = tx + 32 * a, j = ty + 16 * b;
      double x = (i == j) ? 1.0 : 0.0;          // identity padding beyond n: sweeping it is a no-op
      if (i < n && j < n) x = (i <= j) ? A[(size_t)j * ld + i] : A[(size_t)i * ld + j];
      v[a][b] = x;
    }
  const int npad = (n + kPB - 1) / kPB * kPB;
  for (int k0 = 0; k0 < npad; k0 += kPB) {
    const int buf = (k0 / kPB) & 1;
    const int kb = k0 >> 4, kt = k0 & 15;       // columns k0..k0+7 <=> b == kb, ty in [kt, kt + 8)
    // 1. publish the panel (and the pivot block)
    if (ty >= kt && ty < kt + kPB) {
      const int q = ty - kt;
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        double c = v[a][0];
#pragma unroll
        for (int b = 1; b < 8; ++b)
          if (b == kb) c = v[a][b];
        const int i = tx + 32 * a;
        P[buf][q][i] = c;
        if (i >= k0 && i < k0 + kPB) Dm[i - k0][q] = c;
      }
    }
    __syncthreads();
    // 2. warp 0: symmetric sweep of the 8 x 8 pivot block (lane l owns column l & 7, rows l >> 3 and + 4)
    if (ty == 0) {
      const int c = tx & 7, r0 = tx >> 3, r1 = r0 + 4;
      double e0 = Dm[r0][c], e1 = Dm[r1][c];
      for (int t = 0; t < kPB; ++t) {
        __syncwarp();
        Dm[r0][c] = e0; Dm[r1][c] = e1;
        __syncwarp();
        const double d = Dm[t][t];
        if (tx == 0 && !(d > 0.0)) bad = 1;
        const double rinv = fast_rcp(d);           // one short reciprocal chain per pivot
        const double ct = Dm[t][c] * rinv;         // a_tc / d
        const double rt0 = Dm[r0][t], rt1 = Dm[r1][t];
        double n0 = fma(-rt0, ct, e0), n1 = fma(-rt1, ct, e1);
        if (c == t) { n0 = rt0 * rinv; n1 = rt1 * rinv; }
        if (r0 == t) n0 = ct;
        if (r1 == t) n1 = ct;
        if (c == t && r0 == t) n0 = -rinv;
        if (c == t && r1 == t) n1 = -rinv;
        e0 = n0; e1 = n1;
      }
      __syncwarp();
      Dm[r0][c] = e0; Dm[r1][c] = e1;              // = -A_BB^-1
    }
    __syncthreads();
    // 3. S = P * A_BB^-1 : 128 x 8 entries, two per thread
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int e = tid + h * kLeafThreads;         // 0..1023
      const int i = e & (kLeaf - 1), q = e >> 7;
      double acc = 0.0;
#pragma unroll
      for (int r = 0; r < kPB; ++r) acc = fma(P[buf][r][i], -Dm[r][q], acc);
      S[buf][q][i] = acc;
    }
    __syncthreads();
    // 4. rank-8 update and the swept rows / columns
    bool inI[4], inJ[8];
#pragma unroll
    for (int a = 0; a < 4; ++a) { const int i = tx + 32 * a; inI[a] = (i >= k0 && i < k0 + kPB); }
#pragma unroll
    for (int b = 0; b < 8; ++b) inJ[b] = (b == kb) && (ty >= kt && ty < kt + kPB);
#pragma unroll
    for (int q = 0; q < kPB; ++q) {
      double si[4], pj[8];
#pragma unroll
      for (int a = 0; a < 4; ++a) si[a] = S[buf][q][tx + 32 * a];
#pragma unroll
      for (int b = 0; b < 8; ++b) pj[b] = P[buf][q][ty + 16 * b];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) v[a][b] = fma(-si[a], pj[b], v[a][b]);
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        if (inI[a] || inJ[b]) {
          const int i = tx + 32 * a, j = ty + 16 * b;
          double nv;
          if (inI[a] && inJ[b]) nv = Dm[i - k0][j - k0];            // -A_BB^-1
          else if (inJ[b]) nv = S[buf][j - k0][i];                  // (P A_BB^-1)(i, q)
          else nv = S[buf][i - k0][j];                              // symmetric counterpart
          v[a][b] = nv;
        }
      }
    }
    // P/S are double-buffered; Dm is rewritten only after the next group's first barrier... by
    // the owner warps *before* it, so fence the readers of Dm here.
    __syncthreads();
  }
#pragma unroll
  for (int b = 0; b < 8; ++b)
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int i = tx + 32 * a, j = ty + 16 * b;
      if (i < n && j < n) A[(size_t)j * ld + i] = -v[a][b];
    }
  if (tid == 0 && bad && *flag == 0) *flag = code;
}

struct Ctx {
  cublasHandle_t blas;
  cudaStream_t s;
  int* flag;
};

void leaf(const Ctx& c, double* A, int n, size_t ld) {
  leaf_inverse_kernel<<<1, dim3(32, 16), 0, c.s>>>(A, n, ld, c.flag, SPB_ERR_NOT_POSDEF);
  count_launch();
}

// Symmetric products  upper(C) += alpha * op(A) op(B)  on a UNIFORM grid of hb x hb blocks (hb ~ 528, a
// multiple of 16; the last block may be ragged): all diagonal blocks in ONE strided-batched GEMM (they are
// computed as full squares: hb / n of the product is wasted, 6 % at n = 8464), the blocks above the diagonal
// by bisection over block ranges -- one GEMM per node, square at the top (4232 x 4232 x k runs at the DGEMM
// rate), single blocks at the bottom.  The former element-wise recursion ended in sixteen separate
// 530-wide GEMMs of 41 us each at 22 TFLOP/s (profiles/r01_notes.md section 6).
struct SymGrid {
  int n, hb, nb, last;      // nb blocks: nb - 1 of hb rows and a last one of `last` rows (1 <= last <= hb)
  int off(int b) const { return b * hb; }
  int width(int b0, int b1) const { return (b1 == nb ? n : b1 * hb) - b0 * hb; }
};

SymGrid sym_grid(int n) {
  SymGrid g;
  g.n = n;
  int nb = (n + 264) / 528;
  if (nb < 2) nb = 2;
  g.hb = (int)round_up((size_t)ceil_div(n, nb), 16);
  g.nb = ceil_div(n, g.hb);
  g.last = n - (g.nb - 1) * g.hb;
  return g;
}

// blocks above the diagonal inside the block range [b0, b1): C[rows b0..bm, cols bm..b1] then both halves
template <bool TN>
void sym_offdiag(const Ctx& c, const SymGrid& g, int b0, int b1, int k, double alpha, const double* A, size_t lda,
                 const double* B, size_t ldb, double* C, size_t ldc) {
  if (b1 - b0 < 2) return;
  const int bm = (b0 + b1) / 2;
  const int r0 = g.off(b0), c0 = g.off(bm);
  const int m = g.width(b0, bm), w = g.width(bm, b1);
  const double one = 1.0;
  if (TN) {   // C[r0.., c0..] += alpha * A[:, r0..]' * B[:, c0..]
    SPB_CUBLAS(cublasDgemm(c.blas, CUBLAS_OP_T, CUBLAS_OP_N, m, w, k, &alpha, A + (size_t)r0 * lda, (int)lda,
                           B + (size_t)c0 * ldb, (int)ldb, &one, C + (size_t)c0 * ldc + r0, (int)ldc));
  } else {    // C[r0.., c0..] += alpha * A[r0.., :] * B[c0.., :]'
    SPB_CUBLAS(cublasDgemm(c.blas, CUBLAS_OP_N, CUBLAS_OP_T, m, w, k, &alpha, A + r0, (int)lda, B + c0, (int)ldb, &one,
                           C + (size_t)c0 * ldc + r0, (int)ldc));
  }
  sym_offdiag<TN>(c, g, b0, bm, k, alpha, A, lda, B, ldb, C, ldc);
  sym_offdiag<TN>(c, g, bm, b1, k, alpha, A, lda, B, ldb, C, ldc);
}

template <bool TN>
void gemm_upper(const Ctx& c, int n, int k, double alpha, const double* A, size_t lda, const double* B, size_t ldb,
                double* C, size_t ldc) {
  const double one = 1.0;
  const cublasOperation_t ta = TN ? CUBLAS_OP_T : CUBLAS_OP_N, tb = TN ? CUBLAS_OP_N : CUBLAS_OP_T;
  if (n <= kSymGemmMin) {
    SPB_CUBLAS(cublasDgemm(c.blas, ta, tb, n, n, k, &alpha, A, (int)lda, B, (int)ldb, &one, C, (int)ldc));
    return;
  }
  const SymGrid g = sym_grid(n);
  const int nfull = (g.last == g.hb) ? g.nb : g.nb - 1;
  // operand strides from one diagonal block to the next: TN walks columns of A and B, NT walks rows
  const long long sa = TN ? (long long)g.hb * (long long)lda : (long long)g.hb;
  const long long sb = TN ? (long long)g.hb * (long long)ldb : (long long)g.hb;
  const long long sc = (long long)g.hb * ((long long)ldc + 1);
  SPB_CUBLAS(cublasDgemmStridedBatched(c.blas, ta, tb, g.hb, g.hb, k, &alpha, A, (int)lda, sa, B, (int)ldb, sb, &one, C,
                                       (int)ldc, sc, nfull));
  if (nfull < g.nb) {
    const size_t o = (size_t)g.off(g.nb - 1);
    SPB_CUBLAS(cublasDgemm(c.blas, ta, tb, g.last, g.last, k, &alpha, A + (TN ? o * lda : o), (int)lda,
                           B + (TN ? o * ldb : o), (int)ldb, &one, C + o * ldc + o, (int)ldc));
  }
  sym_offdiag<TN>(c, g, 0, g.nb, k, alpha, A, lda, B, ldb, C, ldc);
}

// upper(C) += alpha * A' * B,   A: k x n, B: k x n, C: n x n symmetric result
void gemm_upper_TN(const Ctx& c, int n, int k, double alpha, const double* A, size_t lda, const double* B,
                   size_t ldb, double* C, size_t ldc) {
  gemm_upper<true>(c, n, k, alpha, A, lda, B, ldb, C, ldc);
}

// upper(C) += alpha * A * B',   A: n x k, B: n x k
void gemm_upper_NT(const Ctx& c, int n, int k, double alpha, const double* A, size_t lda, const double* B,
                   size_t ldb, double* C, size_t ldc) {
  gemm_upper<false>(c, n, k, alpha, A, lda, B, ldb, C, ldc);
}

// A (n x n, upper triangle valid) -> A^-1 in full storage.  W: workspace of >= n*n/2 doubles.
void invert(const Ctx& c, double* A, int n, size_t ld, double* W) {
  if (n <= kLeaf) {
    leaf(c, A, n, ld);
    return;
  }
  const int n1 = (n / 2 + 15) / 16 * 16, n2 = n - n1;
  double* A11 = A;
  double* A12 = A + (size_t)n1 * ld;
  double* A22 = A12 + n1;
  const double one = 1.0, zero = 0.0, mone = -1.0;
  invert(c, A11, n1, ld, W);                                             // A11^-1 (full)
  double* Wm = W;                                                        // n1 x n2
  SPB_CUBLAS(cublasDgemm(c.blas, CUBLAS_OP_N, CUBLAS_OP_N, n1, n2, n1, &one, A11, (int)ld, A12, (int)ld, &zero, Wm,
                         n1));                                           // W = A11^-1 A12
  gemm_upper_TN(c, n2, n1, -1.0, A12, ld, Wm, (size_t)n1, A22, ld);      // S = A22 - A12' W (upper)
  invert(c, A22, n2, ld, W + (size_t)n1 * n2);                           // S^-1 (full)
  SPB_CUBLAS(cublasDgemm(c.blas, CUBLAS_OP_N, CUBLAS_OP_N, n1, n2, n2, &mone, Wm, n1, A22, (int)ld, &zero, A12,
                         (int)ld));                                      // B12 = -W S^-1
  gemm_upper_NT(c, n1, n2, -1.0, A12, ld, Wm, (size_t)n1, A11, ld);      // B11 = A11^-1 - B12 W' (upper)
  launch_symmatu_inplace(A, n, ld, 0.0, c.s);                            // full storage for the parent
}

// dst (rows x cols, ldd) <- src (rows x cols, lds)
__global__ void copy_block_kernel(const double* __restrict__ src, size_t lds, double* __restrict__ dst, size_t ldd,
                                  int rows, int cols) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  for (int j = blockIdx.y; j < cols; j += gridDim.y) dst[(size_t)j * ldd + i] = src[(size_t)j * lds + i];
}

}  // namespace

size_t spd_inverse_workspace_doubles(int n) { return (size_t)n * n / 2 + 1024; }

void gemm_upper_tn(cublasHandle_t blas, int n, int k, double alpha, const double* A, size_t lda, const double* B,
                   size_t ldb, double* C, size_t ldc) {
  Ctx c{blas, nullptr, nullptr};
  gemm_upper_TN(c, n, k, alpha, A, lda, B, ldb, C, ldc);
}

Blocking spd_blocking_m(int p, int m_request) {
  Blocking b;
  constexpr int kAlign = 512;                      // the ADMM kernel's largest row tile
  int m = m_request < 1 ? 1 : (m_request > kMaxBlocks ? kMaxBlocks : m_request);
  int bs = (int)round_up((size_t)ceil_div(p, m), kAlign);
  m = ceil_div(p, bs);
  b.m = m;
  for (int j = 0; j < m; ++j) b.off[j] = j * bs;
  b.off[m] = p;
  return b;
}

Blocking spd_blocking(int p) {
  // SPB_FACTOR_BLOCKS = 1 forces explicit inverses.  Default: blocks of 1024 rows (larger ones once kMaxBlocks
  // = 12 would be exceeded, i.e. above p = 12288).  Measured on the C4 job (profiles/r01_notes.md sections 6, 9):
  // 5 / 7 / 10 / 20 blocks -> 1.43 / 1.22 / 1.17 / 1.16 s per job while an ADMM iteration on such a system takes
  // 258 / 282 / 344 / 609 us -- the factor's flop count falls as p^3/3 + p^3/m while every block adds two
  // dependent stream phases; 1024 rows keep the GEMMs deep enough (k = 1024) and the phase count moderate.
  // Below p = 2048 everything is latency-bound and the single-phase stream of an explicit inverse wins.
  static const int env_m = [] {
    const char* e = std::getenv("SPB_FACTOR_BLOCKS");
    return e != nullptr ? std::atoi(e) : 0;
  }();
  static const bool cusolver = [] {
    const char* e = std::getenv("SPB_INVERSE");
    return e != nullptr && std::string(e) == "cusolver";
  }();
  if (cusolver) return spd_blocking_m(p, 1);
  if (env_m > 0) return spd_blocking_m(p, env_m);
  if (p < 2048) return spd_blocking_m(p, 1);
  return spd_blocking_m(p, ceil_div(p, 1024));
}

Blocking spd_blocking_path(int p) {
  static const int env_m = [] {
    const char* e = std::getenv("SPB_PATH_BLOCKS");
    return e != nullptr ? std::atoi(e) : 0;
  }();
  if (env_m > 0) return spd_blocking_m(p, env_m);
  // Measured at C4 (p = 10 000, ~25 iterations per path system): against spd_blocking(p)'s many blocks, 2 blocks
  // cost +20 ms per job and explicit inverses +80 ms, while an iteration takes 344 us with 10 blocks, 170 us
  // with 2 and 144 us explicit -- 2 blocks win from ~46 iterations per system on and lose little before
  // (profiles/r01_notes.md sections 6, 8).
  const Blocking b = spd_blocking(p);
  return b.m <= 2 ? b : spd_blocking_m(p, 2);
}

// A (n x n, upper triangle valid) -> block factor F of A = U'DU over the partition `blk` (full storage):
// diagonal blocks D_j^-1 (D_j = the j-th Schur complement), blocks above the diagonal -U_jk =
// -D_j^-1 A_jk^(j), mirrored below.  Right-looking, every flop a cuBLAS GEMM or the leaf kernel:
// per block row  invert(D_j), Wm = -D_j^-1 R_j, trailing upper += R_j' Wm, R_j <- Wm.
void spd_factor_blocked(const Exec& ex, double* A, int n, size_t ld, double* workspace, const Blocking& blk) {
  Ctx c{ex.blas, ex.stream, ex.flag};
  if (blk.m <= 1) {
    invert(c, A, n, ld, workspace);
    SPB_CUDA(cudaGetLastError());
    return;
  }
  SPB_REQUIRE(blk.off[0] == 0 && blk.off[blk.m] == n, "invalid block partition");
  const double zero = 0.0, mone = -1.0;
  for (int j = 0; j < blk.m; ++j) {
    const int o = blk.off[j], bj = blk.off[j + 1] - o, rest = n - blk.off[j + 1];
    double* Ajj = A + (size_t)o * ld + o;
    invert(c, Ajj, bj, ld, workspace);                                   // D_j^-1 (full)
    if (rest <= 0) break;
    double* Rj = Ajj + (size_t)bj * ld;                                  // bj x rest
    double* A22 = Rj + bj;
    double* Wm = workspace;                                              // bj x rest, ld = bj
    SPB_CUBLAS(cublasDgemm(c.blas, CUBLAS_OP_N, CUBLAS_OP_N, bj, rest, bj, &mone, Ajj, (int)ld, Rj, (int)ld, &zero,
                           Wm, bj));                                     // Wm = -D_j^-1 R_j
    gemm_upper_TN(c, rest, bj, 1.0, Rj, ld, Wm, (size_t)bj, A22, ld);    // S = A22 - R_j' D_j^-1 R_j (upper)
    copy_block_kernel<<<dim3(ceil_div(bj, 256), rest < 1024 ? rest : 1024), 256, 0, c.s>>>(Wm, (size_t)bj, Rj, ld, bj,
                                                                                         rest);
    count_launch();
  }
  launch_symmatu_inplace(A, n, ld, 0.0, c.s);                            // full storage for the stream
  SPB_CUDA(cudaGetLastError());
}

void spd_inverse_recursive(const Exec& ex, double* A, int n, size_t ld, double* workspace) {
  Ctx c{ex.blas, ex.stream, ex.flag};
  invert(c, A, n, ld, workspace);
  SPB_CUDA(cudaGetLastError());
}

}  // namespace spb
