// Extended-precision products of the Omega build on the INT8 tensor cores (Ozaki-type slicing).
//
// The Newton-Schulz step of omega_refine.cu needs the residual  I - A X0  of a system with cond ~ 1e8 .. 1e10 to an
// absolute accuracy of ~1e-13: the terms a_ik x_kj reach 1e7 and cancel to delta_ij, i.e. ~70 significant bits.  In
// FP64 arithmetic that took three full DGEMMs (head / tail splitting, 6 q^3 flop, 168 ms at q = 10 003 on B200's
// 36 TFLOP/s).  B200's INT8 tensor cores run the same shape at 2.6 POP/s (measured, tools/int8_probe.py), so the
// product is rebuilt from exact integer pieces instead:
//
//   * every column of an operand is scaled by its own power of two and cut into S signed 7-bit digits
//         x = 2^(e - 6) * sum_s d_s 2^(-7 s),   |d_s| <= 64            (exact: scaling, rint and the remainder are exact)
//     (rows of the symmetric left operands A and X0 are their columns);
//   * digit planes multiply exactly: int8 x int8 -> int32, |sum| <= 64 * 64 * K < 2^31;
//   * all plane pairs with s + t = d share the weight 2^(-7 d), so they are ONE GEMM over a concatenated K axis:
//     the left planes are stored in order s = 0 .. S-1, the right planes in REVERSE order, and diagonal d multiplies
//     rows [0, (d+1) kq) of the left buffer with rows [(S-1-d) kq, S kq) of the right one -- S GEMMs of growing depth
//     instead of S (S + 1) / 2 launches, pointer arithmetic only;
//   * the S int32 results are accumulated in FP64 from the smallest weight up and the epilogue applies the scales
//     and the outer operation (I - ., base + .) in the same pass.
//   Pairs with s + t >= S are dropped: a relative error of ~S 2^(-7 S) of rowmax * colmax per term (S = 12: 2^-80,
//   below the 2^-74 of the FP64 head/tail scheme).  Cost: S (S + 1) / 2 = 78 plane products, 62 ms at q = 10 003.
//
// The GEMMs themselves are cuBLAS (cublasGemmEx, CUDA_R_8I x CUDA_R_8I -> CUDA_R_32I, TN): a plain library GEMM on
// planes that this file produces and consumes.  Everything else -- exponents, slicing, accumulation, epilogue -- is here.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>

#include "engine.cuh"

namespace spb {

namespace {

constexpr int kDigitBits = 7;             // planes after the first; the first holds 6 bits + sign
constexpr long long kMaxDepth = ((1LL << 31) - 1) / (64 * 64);   // int32 accumulators: depth * 64 * 64 < 2^31

// column maxima of |M| (rows x cols) -> power-of-two exponents e with max < 2^e (0 for an all-zero column);
// a non-finite entry raises `bad`
__global__ void __launch_bounds__(256) oz_expo_kernel(const double* __restrict__ M, size_t ld, int rows,
                                                      int* __restrict__ expo, int* __restrict__ bad) {
  __shared__ double red[8];
  __shared__ int nonfinite;
  if (threadIdx.x == 0) nonfinite = 0;
  __syncthreads();
  const double* col = M + (size_t)blockIdx.x * ld;
  double m = 0.0;
  int nf = 0;
  for (int i = threadIdx.x; i < rows; i += 256) {
    const double a = fabs(col[i]);
    if (!(a <= 1.7e308)) nf = 1;
    m = fmax(m, a);
  }
  if (nf) nonfinite = 1;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) m = fmax(m, red[w]);
    int e = 0;
    if (nonfinite) { *bad = 1; m = 0.0; }
    if (m > 0.0) frexp(m, &e);                  // m = f 2^e, f in [0.5, 1)
    expo[blockIdx.x] = e;
  }
}

// digit planes of M (rows x cols, column scaling): plane s of column j lives at out[j * S * kq + slot(s) * kq + k],
// slot(s) = s (forward) or S - 1 - s (reversed); rows >= `rows` and columns >= `cols` (padding) are zero planes.
// One thread: four consecutive k of one column.
__global__ void __launch_bounds__(256) oz_slice_kernel(const double* __restrict__ M, size_t ld, int rows, int cols,
                                                       const int* __restrict__ expo, int S, int reversed,
                                                       int8_t* __restrict__ out, size_t kq) {
  const size_t k4 = (size_t)blockIdx.x * 256 + threadIdx.x;
  const int j = blockIdx.y;
  if (k4 * 4 >= kq) return;
  double y[4] = {0.0, 0.0, 0.0, 0.0};
  if (j < cols) {
    const int e = expo[j];
    const double* col = M + (size_t)j * ld;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const size_t k = k4 * 4 + u;
      if (k < (size_t)rows) {
        const double v = col[k];
        y[u] = (fabs(v) <= 1.7e308) ? scalbn(v, 6 - e) : 0.0;      // |y| < 64
      }
    }
  }
  int8_t* base = out + (size_t)j * S * kq + k4 * 4;
  for (int s = 0; s < S; ++s) {
    char4 dg;
    double d;
    d = rint(y[0]); y[0] = (y[0] - d) * 128.0; dg.x = (signed char)(int)d;
    d = rint(y[1]); y[1] = (y[1] - d) * 128.0; dg.y = (signed char)(int)d;
    d = rint(y[2]); y[2] = (y[2] - d) * 128.0; dg.z = (signed char)(int)d;
    d = rint(y[3]); y[3] = (y[3] - d) * 128.0; dg.w = (signed char)(int)d;
    const int slot = reversed ? S - 1 - s : s;
    *reinterpret_cast<char4*>(base + (size_t)slot * kq) = dg;
  }
}

// acc (+)= w * C   (C: int32 planes product, mp x np, ld = mp)
__global__ void __launch_bounds__(256) oz_accumulate_kernel(const int* __restrict__ C, size_t ldc, int m, int n, double w,
                                                            int first, double* __restrict__ acc, size_t lda) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  const int j0 = blockIdx.y * 8;
  if (i >= m) return;
#pragma unroll
  for (int jj = 0; jj < 8; ++jj) {
    const int j = j0 + jj;
    if (j >= n) break;
    const double v = w * (double)C[(size_t)j * ldc + i];
    double* a = acc + (size_t)j * lda + i;
    *a = first ? v : *a + v;
  }
}

// last plane product + epilogue:  prod = 2^(ea_i + eb_j - 12) * (acc + w C);
//   mode 0: out = delta(i, j + col0) - prod        (residual columns)
//   mode 1: out = base + coef * prod               (correction, final product; out may alias base)
__global__ void __launch_bounds__(256) oz_finish_kernel(const int* __restrict__ C, size_t ldc, int m, int n, double w,
                                                        int first, const double* __restrict__ acc, size_t lda,
                                                        const int* __restrict__ ea, const int* __restrict__ eb, int mode,
                                                        int col0, double coef, const double* __restrict__ base,
                                                        size_t ldb, double* __restrict__ out, size_t ldo,
                                                        const int* __restrict__ bad) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  const int j0 = blockIdx.y * 8;
  if (i >= m) return;
  const int ei = ea[i] - 12;
  const bool poison = *bad != 0;
#pragma unroll
  for (int jj = 0; jj < 8; ++jj) {
    const int j = j0 + jj;
    if (j >= n) break;
    double v = w * (double)C[(size_t)j * ldc + i];
    if (!first) v += acc[(size_t)j * lda + i];
    v = scalbn(v, ei + eb[j]);
    if (poison) v = nan("");                       // a non-finite operand: behave like the FP64 product would
    if (mode == 0) out[(size_t)j * ldo + i] = ((i == j + col0) ? 1.0 : 0.0) - v;
    else out[(size_t)j * ldo + i] = base[(size_t)j * ldb + i] + coef * v;
  }
}

// dst = src' (n x n, tiled through shared memory)
__global__ void __launch_bounds__(256) oz_transpose_kernel(const double* __restrict__ src, size_t lds, int n,
                                                           double* __restrict__ dst, size_t ldd) {
  __shared__ double tile[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int i = bx + tx, j = by + r;
    tile[r][tx] = (i < n && j < n) ? src[(size_t)j * lds + i] : 0.0;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int i = by + tx, j = bx + r;                  // dst(i, j) = src(j, i) = tile[i - by][j - bx]
    if (i < n && j < n) dst[(size_t)j * ldd + i] = tile[tx][r];
  }
}

inline size_t oz_round_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

}  // namespace

int ozaki_env_slices(const char* name, int dflt, int lo, int hi) {
  const char* e = std::getenv(name);
  if (!e) return dflt;
  return std::min(hi, std::max(lo, std::atoi(e)));
}

OzakiGeom ozaki_geom(int q, int nc) {
  OzakiGeom g;
  g.kq = oz_round_up((size_t)q, 128);
  g.mp = (int)oz_round_up((size_t)q, 16);
  g.np = (int)oz_round_up((size_t)std::max(nc, 1), 16);
  return g;
}

size_t ozaki_workspace_bytes(int q, int nc, int S, int Sc) {
  const OzakiGeom g = ozaki_geom(q, nc);
  const int Sm = std::max(S, Sc);
  return (size_t)Sm * g.kq * g.mp + (size_t)Sm * g.kq * g.np + sizeof(int) * (size_t)g.mp * g.np + 64 * (size_t)g.mp;
}

void ozaki_transpose(cudaStream_t s, const double* src, size_t lds, int n, double* dst, size_t ldd) {
  const dim3 grid((unsigned)((n + 31) / 32), (unsigned)((n + 31) / 32));
  oz_transpose_kernel<<<grid, 256, 0, s>>>(src, lds, n, dst, ldd);
  count_launch();
  SPB_CUDA(cudaGetLastError());
}

// planes of `cols` columns of M into `out` (`colsp` columns are written, the surplus as zeros)
void ozaki_slice(cudaStream_t s, const double* M, size_t ld, int rows, int cols, int colsp, int S, bool reversed,
                 int8_t* out, size_t kq, int* expo, int* bad) {
  if (cols > 0) oz_expo_kernel<<<cols, 256, 0, s>>>(M, ld, rows, expo, bad);
  const dim3 grid((unsigned)((kq / 4 + 255) / 256), (unsigned)colsp);
  oz_slice_kernel<<<grid, 256, 0, s>>>(M, ld, rows, cols, expo, S, reversed ? 1 : 0, out, kq);
  count_launch(2);
  SPB_CUDA(cudaGetLastError());
}

// out (m x n) = epilogue( left' planes x right planes ):  left has Sa forward planes of >= m columns, right has Sb
// reversed planes of >= n columns; diagonals d = S-1 .. 0 with S <= min(Sa, Sb)
void ozaki_product(cublasHandle_t blas, cudaStream_t s, const OzakiGeom& g, const int8_t* left, int Sa,
                   const int8_t* right, int Sb, int S, int m, int n, int* C, double* acc, size_t lda, const int* ea,
                   const int* eb, int mode, int col0, double coef, const double* base, size_t ldb, double* out,
                   size_t ldo, const int* bad) {
  SPB_REQUIRE(S >= 1 && S <= Sa && S <= Sb, "ozaki_product: plane count out of range");
  const int one = 1, zero = 0;
  const size_t kq = g.kq;
  const int max_slots = (int)std::max<long long>(1, kMaxDepth / (long long)kq);
  const dim3 grid((unsigned)((m + 255) / 256), (unsigned)((n + 7) / 8));
  bool first = true;
  for (int d = S - 1; d >= 0; --d) {
    const double w = std::ldexp(1.0, -kDigitBits * d);
    for (int s0 = 0; s0 <= d; s0 += max_slots) {
      const int s1 = std::min(d + 1, s0 + max_slots);
      const int8_t* a = left + (size_t)s0 * kq;                       // planes s0 .. s1-1 of every left column
      const int8_t* b = right + (size_t)(Sb - 1 - d + s0) * kq;       // planes t = d-s0 .. d-s1+1 of every right column
      SPB_CUBLAS(cublasGemmEx(blas, CUBLAS_OP_T, CUBLAS_OP_N, g.mp, g.np, (int)((size_t)(s1 - s0) * kq), &one, a, CUDA_R_8I,
                              (int)((size_t)Sa * kq), b, CUDA_R_8I, (int)((size_t)Sb * kq), &zero, C, CUDA_R_32I, g.mp,
                              CUBLAS_COMPUTE_32I, CUBLAS_GEMM_DEFAULT));
      const bool last = (d == 0 && s1 == d + 1);
      if (!last) {
        oz_accumulate_kernel<<<grid, 256, 0, s>>>(C, (size_t)g.mp, m, n, w, first ? 1 : 0, acc, lda);
      } else {
        oz_finish_kernel<<<grid, 256, 0, s>>>(C, (size_t)g.mp, m, n, w, first ? 1 : 0, acc, lda, ea, eb, mode, col0, coef,
                                              base, ldb, out, ldo, bad);
      }
      count_launch();
      first = false;
    }
  }
  SPB_CUDA(cudaGetLastError());
}

// ozaki_product for a result of which only the part on / above the diagonal is read afterwards (a symmetric matrix whose
// upper triangle is mirrored later): column strips of `strip` columns, strip [j0, j0 + nb) is computed for the rows
// [0, j0 + nb) only -- 0.6 of the plane products with 5 strips (p = 10 000, strip = 2048).  Entries below the diagonal
// blocks of `out` are NOT written.  `strip` must be a multiple of 16.
void ozaki_product_upper(cublasHandle_t blas, cudaStream_t s, const OzakiGeom& g, const int8_t* left, int Sa,
                         const int8_t* right, int Sb, int S, int m, int n, int* C, double* acc, size_t lda, const int* ea,
                         const int* eb, int mode, int col0, double coef, const double* base, size_t ldb, double* out,
                         size_t ldo, const int* bad, int strip) {
  SPB_REQUIRE(strip >= 16 && strip % 16 == 0, "ozaki_product_upper: strip must be a multiple of 16");
  for (int j0 = 0; j0 < n; j0 += strip) {
    const int nb = std::min(strip, n - j0);
    const int mr = std::min(m, j0 + nb);
    OzakiGeom gs = g;
    gs.mp = (int)oz_round_up((size_t)mr, 16);
    gs.np = (int)oz_round_up((size_t)nb, 16);
    ozaki_product(blas, s, gs, left, Sa, right + (size_t)j0 * Sb * g.kq, Sb, S, mr, nb, C, acc + (size_t)j0 * lda, lda, ea,
                  eb + j0, mode, col0 + j0, coef, base ? base + (size_t)j0 * ldb : nullptr, ldb, out + (size_t)j0 * ldo, ldo,
                  bad);
  }
}

}  // namespace spb
