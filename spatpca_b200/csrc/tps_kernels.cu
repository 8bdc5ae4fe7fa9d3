// Thin-plate-spline kernels, matrix assembly and small utility kernels (sm_100a).
//
// Everything here is either write-bound (one coalesced 8-byte store per element,
// columns are contiguous) or a tiled transpose; grids are sized from the matrix,
// loads of the tiny p x d location table go through the read-only path.
#include "kernels.cuh"

namespace spb {

namespace {

constexpr double kEightPi = 8.0 * 3.14159265358979323846264338327950288;

// r^3 with a single final rounding (the reference calls libm pow(r, 3), which is
// correctly rounded in practice): r*r is kept as a double-double.
__device__ __forceinline__ double cube_accurate(double r) {
  double hi = __dmul_rn(r, r);
  double lo = fma(r, r, -hi);
  return fma(hi, r, __dmul_rn(lo, r));
}

// phi(r) of reference src/RcppSpatPCA.cpp:23-36, evaluated from the coordinate
// differences in the reference's operation order (no FMA contraction).
template <int D>
__device__ __forceinline__ double tps_phi(const double* __restrict__ xi, const double* __restrict__ xj) {
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < D; ++k) {
    double dx = __dsub_rn(xi[k], xj[k]);
    double sq = __dmul_rn(dx, dx);
    s = (k == 0) ? sq : __dadd_rn(s, sq);
  }
  double r = sqrt(s);
  if (D == 1) return cube_accurate(r) / 12.0;
  if (D == 2) return __dmul_rn(__dmul_rn(r, r), log(r)) / kEightPi;   // 0 * log 0 = NaN, as the reference
  return -r / kEightPi;
}

constexpr int kTpsRows = 128;   // rows per block (one per thread)
constexpr int kTpsCols = 8;     // columns per block

template <int D>
__global__ void __launch_bounds__(kTpsRows)
tps_fill_kernel(const double* __restrict__ x, int p, double* __restrict__ L, size_t ld,
                double ridge, int q) {
  const int i = blockIdx.x * kTpsRows + threadIdx.x;
  const int j0 = blockIdx.y * kTpsCols;
  if (i >= q) return;
  double xi[D];
  if (i < p) {
#pragma unroll
    for (int k = 0; k < D; ++k) xi[k] = __ldg(x + (size_t)k * p + i);
  }
#pragma unroll
  for (int jj = 0; jj < kTpsCols; ++jj) {
    const int j = j0 + jj;
    if (j >= q) break;
    double v;
    if (i == j) {
      v = ridge;
    } else if (i < p && j < p) {
      double xj[D];
#pragma unroll
      for (int k = 0; k < D; ++k) xj[k] = __ldg(x + (size_t)k * p + j);
      // the reference fills j > i and mirrors: evaluate with the smaller index first
      v = (i < j) ? tps_phi<D>(xi, xj) : tps_phi<D>(xj, xi);
    } else if (i < p) {           // border columns: [1, X]
      v = (j == p) ? 1.0 : xi[j - p - 1];
    } else if (j < p) {           // border rows: [1, X]'
      v = (i == p) ? 1.0 : __ldg(x + (size_t)(i - p - 1) * p + j);
    } else {
      v = 0.0;
    }
    L[(size_t)j * ld + i] = v;
  }
}

constexpr int kTile = 32;

// In-place symmatu: lower <- upper', one thread block per tile pair (bi <= bj).
__global__ void __launch_bounds__(kTile * 8)
symmatu_kernel(double* __restrict__ A, int n, size_t ld, double diag_add) {
  const int bi = blockIdx.x, bj = blockIdx.y;
  if (bi > bj) return;
  __shared__ double tile[kTile][kTile + 1];
  const int tx = threadIdx.x & (kTile - 1);
  const int ty = threadIdx.x / kTile;       // 0..7
  // load upper tile (rows bi*32.., cols bj*32..), coalesced along rows
  for (int c = ty; c < kTile; c += 8) {
    int i = bi * kTile + tx, j = bj * kTile + c;
    if (i < n && j < n) tile[c][tx] = A[(size_t)j * ld + i];
  }
  __syncthreads();
  if (bi == bj) {
    for (int c = ty; c < kTile; c += 8) {
      int i = bi * kTile + tx, j = bj * kTile + c;
      if (i < n && j < n) {
        if (i > j) A[(size_t)j * ld + i] = tile[tx][c];          // lower <- upper'
        else if (i == j && diag_add != 0.0) A[(size_t)j * ld + i] = __dadd_rn(tile[c][tx], diag_add);
      }
    }
  } else {
    // write transposed tile at (rows bj*32.., cols bi*32..)
    for (int c = ty; c < kTile; c += 8) {
      int i = bj * kTile + tx, j = bi * kTile + c;
      if (i < n && j < n) A[(size_t)j * ld + i] = tile[tx][c];
    }
  }
}

constexpr int kAsmRows = 128;
constexpr int kAsmCols = 8;

__global__ void __launch_bounds__(kAsmRows)
assemble_upper_kernel(const double* __restrict__ om, size_t ldo, const double* __restrict__ G,
                      size_t ldg, int p, double tau1, double scale, double rho,
                      double* __restrict__ out, size_t ldout) {
  const int i = blockIdx.x * kAsmRows + threadIdx.x;
  const int j0 = blockIdx.y * kAsmCols;
  if (blockIdx.x * kAsmRows > j0 + kAsmCols - 1) return;   // block entirely below the diagonal
  if (i >= p) return;
#pragma unroll
  for (int jj = 0; jj < kAsmCols; ++jj) {
    const int j = j0 + jj;
    if (j >= p) break;
    if (i <= j) {
      double t = __dmul_rn(tau1, om[(size_t)j * ldo + i]);
      double u = __dsub_rn(t, G[(size_t)j * ldg + i]);
      double v = __dmul_rn(scale, u);
      if (i == j) v = __dadd_rn(v, rho);
      out[(size_t)j * ldout + i] = v;
    }
  }
}

__global__ void gather_rows_kernel(const double* __restrict__ Y, int n, int ncol,
                                   const int* __restrict__ rows, int m, double* __restrict__ out) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = (size_t)m * ncol;
  if (idx >= total) return;
  int r = (int)(idx % m);
  size_t j = idx / m;
  out[idx] = Y[j * n + rows[r]];
}

__global__ void __launch_bounds__(kTile * 8)
transpose_kernel(const double* __restrict__ in, int rows, int cols, double* __restrict__ out) {
  __shared__ double tile[kTile][kTile + 1];
  const int tx = threadIdx.x & (kTile - 1);
  const int ty = threadIdx.x / kTile;
  const int r0 = blockIdx.x * kTile, c0 = blockIdx.y * kTile;
  for (int c = ty; c < kTile; c += 8) {
    int i = r0 + tx, j = c0 + c;
    if (i < rows && j < cols) tile[c][tx] = in[(size_t)j * rows + i];
  }
  __syncthreads();
  for (int c = ty; c < kTile; c += 8) {
    int i = c0 + tx, j = r0 + c;        // out is cols x rows
    if (i < cols && j < rows) out[(size_t)j * cols + i] = tile[tx][c];
  }
}

__global__ void copy_upper_kernel(const double* __restrict__ src, size_t lds, int n,
                                  double* __restrict__ dst, size_t ldd) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)n * n) return;
  int i = (int)(idx % n);
  size_t j = idx / n;
  dst[j * ldd + i] = (i <= (int)j) ? src[j * lds + i] : 0.0;
}

__global__ void set_identity_kernel(double* __restrict__ B, int q, int ncol, size_t ld) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = (size_t)q * ncol;
  if (idx >= total) return;
  int i = (int)(idx % q);
  size_t j = idx / q;
  B[j * ld + i] = (i == (int)j) ? 1.0 : 0.0;
}

constexpr int kRedBlocks = 256;
constexpr int kRedThreads = 256;

__device__ __forceinline__ double block_sum_256(double v, double* sh) {
  // fixed-shape tree: warp shuffles, then 8 warp partials summed in order
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sh[w] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0) {
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += sh[k];
  }
  __syncthreads();
  return t;   // valid on thread 0
}

__global__ void __launch_bounds__(kRedThreads)
sumsq_partial_kernel(const double* __restrict__ x, size_t n, double* __restrict__ part) {
  __shared__ double sh[8];
  double acc = 0.0;
  for (size_t i = (size_t)blockIdx.x * kRedThreads + threadIdx.x; i < n;
       i += (size_t)kRedBlocks * kRedThreads) {
    double v = x[i];
    acc = fma(v, v, acc);
  }
  double t = block_sum_256(acc, sh);
  if (threadIdx.x == 0) part[blockIdx.x] = t;
}

__global__ void __launch_bounds__(kRedThreads)
sum_final_kernel(const double* __restrict__ part, int n, double* __restrict__ out) {
  __shared__ double sh[8];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += kRedThreads) acc += part[i];
  double t = block_sum_256(acc, sh);
  if (threadIdx.x == 0) out[0] = t;
}

struct CoefPack { double c[SPB_MAX_K]; };

__global__ void lambda_init_kernel(const double* __restrict__ Phi, int p, int K, CoefPack coef,
                                   double* __restrict__ L2) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)p * K) return;
  int k = (int)(idx / p);
  L2[idx] = __dmul_rn(Phi[idx], coef.c[k]);
}

__global__ void scale_copy_kernel(const double* __restrict__ src, double alpha,
                                  double* __restrict__ dst, size_t n) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < n) dst[idx] = __dmul_rn(alpha, src[idx]);
}

__global__ void check_info_kernel(const int* info, int* flag, int code) {
  if (*info != 0 && *flag == 0) *flag = code;
}

__global__ void small_rmul_kernel(const double* __restrict__ Phi, int p, int K,
                                  const double* __restrict__ Q, double* __restrict__ out) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)p * K) return;
  int i = (int)(idx % p);
  int k = (int)(idx / p);
  double acc = 0.0;
  for (int a = 0; a < K; ++a) acc = fma(Phi[(size_t)a * p + i], Q[(size_t)k * K + a], acc);
  out[idx] = acc;
}

// eigenFunction evaluation, reference src/RcppSpatPCA.cpp:113-145: one thread per
// (new site, eigenfunction); the sum over original sites runs in index order like
// the reference's loop so that the result differs only through log().
template <int D>
__global__ void tps_eval_kernel(const double* __restrict__ xnew, int pnew,
                                const double* __restrict__ xorig, int p,
                                const double* __restrict__ para, size_t ldpara, int K,
                                double* __restrict__ out) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)pnew * K) return;
  const int ni = (int)(idx % pnew);
  const int k = (int)(idx / pnew);
  double xn[D];
#pragma unroll
  for (int c = 0; c < D; ++c) xn[c] = xnew[(size_t)c * pnew + ni];
  const double* pk = para + (size_t)k * ldpara;
  double psum = 0.0;
  for (int j = 0; j < p; ++j) {
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < D; ++c) {
      double dx = __dsub_rn(xn[c], __ldg(xorig + (size_t)c * p + j));
      double sq = __dmul_rn(dx, dx);
      s = (c == 0) ? sq : __dadd_rn(s, sq);
    }
    double r = (D == 1) ? fabs(__dsub_rn(xn[0], __ldg(xorig + j))) : sqrt(s);
    if (r > 0) {
      double pj = pk[j];
      if (D == 1) {
        psum = __dadd_rn(psum, __dmul_rn(pj, __dmul_rn(__dmul_rn(r, r), r)) / 12.0);
      } else if (D == 2) {
        psum = __dadd_rn(psum, __dmul_rn(__dmul_rn(__dmul_rn(pj, r), r), log(r)) / kEightPi);
      } else {
        psum = __dsub_rn(psum, __dmul_rn(pj, r) / kEightPi);
      }
    }
  }
  double v = psum;
#pragma unroll
  for (int c = 0; c < D; ++c) v = __dadd_rn(v, __dmul_rn(pk[p + 1 + c], xn[c]));
  v = __dadd_rn(v, pk[p]);
  out[idx] = v;
}

// ---- power iteration for lambda_max(Omega_u): start vector and normalisation (one CTA) ----------------
__global__ void power_init_kernel(double* x, int p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p) return;
  unsigned int h = (unsigned int)i * 2654435761u + 0x9e3779b9u;      // fixed pseudo-random signs and magnitudes
  h ^= h >> 15; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
  x[i] = ((double)(h & 0xffffffu) / 16777216.0 - 0.5) * rsqrt((double)p);
}
// x <- y / |y|, norm_out[it] <- |y|   (deterministic tree; p is small against one CTA's throughput)
__global__ void __launch_bounds__(1024) power_normalize_kernel(const double* __restrict__ y, int p, double* __restrict__ x,
                                                               double* __restrict__ norm_out) {
  __shared__ double red[32];
  __shared__ double nrm;
  double t = 0.0;
  for (int i = threadIdx.x; i < p; i += 1024) t = fma(y[i], y[i], t);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = red[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) { nrm = sqrt(v); *norm_out = nrm; }
  }
  __syncthreads();
  const double inv = nrm > 0.0 ? 1.0 / nrm : 0.0;
  for (int i = threadIdx.x; i < p; i += 1024) x[i] = y[i] * inv;
}

}  // namespace

void launch_tps_fill(const double* x_dev, int p, int d, double* L, size_t ld, double ridge,
                     bool border, cudaStream_t s) {
  const int q = border ? p + d + 1 : p;
  dim3 grid(ceil_div(q, kTpsRows), ceil_div(q, kTpsCols));
  switch (d) {
    case 1: tps_fill_kernel<1><<<grid, kTpsRows, 0, s>>>(x_dev, p, L, ld, ridge, q); break;
    case 2: tps_fill_kernel<2><<<grid, kTpsRows, 0, s>>>(x_dev, p, L, ld, ridge, q); break;
    case 3: tps_fill_kernel<3><<<grid, kTpsRows, 0, s>>>(x_dev, p, L, ld, ridge, q); break;
    default: throw Error(SPB_ERR_INVALID_ARG, "location dimension must be 1, 2 or 3");
  }
  count_launch();
  SPB_CUDA(cudaGetLastError());
}

void launch_symmatu_inplace(double* A, int n, size_t ld, double diag_add, cudaStream_t s) {
  int nb = ceil_div(n, kTile);
  dim3 grid(nb, nb);
  symmatu_kernel<<<grid, kTile * 8, 0, s>>>(A, n, ld, diag_add);
  count_launch();
  SPB_CUDA(cudaGetLastError());
}

void launch_assemble_upper(const double* omega_u, size_t ldo, const double* G, size_t ldg, int p,
                           double tau1, double scale, double rho, double* out, size_t ldout,
                           cudaStream_t s) {
  dim3 grid(ceil_div(p, kAsmRows), ceil_div(p, kAsmCols));
  assemble_upper_kernel<<<grid, kAsmRows, 0, s>>>(omega_u, ldo, G, ldg, p, tau1, scale, rho, out,
                                                  ldout);
  count_launch();
  SPB_CUDA(cudaGetLastError());
}

void launch_gather_rows(const double* Y, int n, int ncol, const int* rows, int m, double* out,
                        cudaStream_t s) {
  size_t total = (size_t)m * ncol;
  if (total == 0) return;
  gather_rows_kernel<<<ceil_div(total, 256), 256, 0, s>>>(Y, n, ncol, rows, m, out);
  count_launch();
  SPB_CUDA(cudaGetLastError());
}

void launch_transpose(const double* in, int rows, int cols, double* out, cudaStream_t s) {
  dim3 grid(ceil_div(rows, kTile), ceil_div(cols, kTile));
  transpose_kernel<<<grid, kTile * 8, 0, s>>>(in, rows, cols, out);
  count_launch();
  SPB_CUDA(cudaGetLastError());
}

void launch_copy_upper(const double* src, size_t lds, int n, double* dst, size_t ldd, cudaStream_t s) {
  size_t total = (size_t)n * n;
  copy_upper_kernel<<<ceil_div(total, 256), 256, 0, s>>>(src, lds, n, dst, ldd);
  count_launch();
  SPB_CUDA(cudaGetLastError());
}

void launch_set_identity(double* B, int q, int ncol, size_t ld, cudaStream_t s) {
  size_t total = (size_t)q * ncol;
  set_identity_kernel<<<ceil_div(total, 256), 256, 0, s>>>(B, q, ncol, ld);
  count_launch();
  SPB_CUDA(cudaGetLastError());
}

void launch_sumsq(const double* x, size_t n, double* scratch, double* out, cudaStream_t s) {
  sumsq_partial_kernel<<<kRedBlocks, kRedThreads, 0, s>>>(x, n, scratch);
  sum_final_kernel<<<1, kRedThreads, 0, s>>>(scratch, kRedBlocks, out);
  count_launch(2);
  SPB_CUDA(cudaGetLastError());
}

void launch_lambda_init_host(const double* Phi, int p, int K, const double* coef_host, double* L2,
                             cudaStream_t s) {
  CoefPack c;
  for (int k = 0; k < SPB_MAX_K; ++k) c.c[k] = (k < K) ? coef_host[k] : 0.0;
  size_t total = (size_t)p * K;
  lambda_init_kernel<<<ceil_div(total, 256), 256, 0, s>>>(Phi, p, K, c, L2);
  count_launch();
  SPB_CUDA(cudaGetLastError());
}

void launch_scale_copy(const double* src, double alpha, double* dst, size_t n, cudaStream_t s) {
  if (n == 0) return;
  scale_copy_kernel<<<ceil_div(n, 256), 256, 0, s>>>(src, alpha, dst, n);
  count_launch();
  SPB_CUDA(cudaGetLastError());
}

void launch_check_info(const int* info, int* flag, int code, cudaStream_t s) {
  check_info_kernel<<<1, 1, 0, s>>>(info, flag, code);
  count_launch();
  SPB_CUDA(cudaGetLastError());
}

void launch_small_rmul(const double* Phi, int p, int K, const double* Q_dev, double* out,
                       cudaStream_t s) {
  size_t total = (size_t)p * K;
  small_rmul_kernel<<<ceil_div(total, 256), 256, 0, s>>>(Phi, p, K, Q_dev, out);
  count_launch();
  SPB_CUDA(cudaGetLastError());
}

void launch_power_init(double* x, int p, cudaStream_t s) {
  power_init_kernel<<<ceil_div(p, 256), 256, 0, s>>>(x, p);
  count_launch();
  SPB_CUDA(cudaGetLastError());
}

void launch_power_normalize(const double* y, int p, double* x, double* norm_out, cudaStream_t s) {
  power_normalize_kernel<<<1, 1024, 0, s>>>(y, p, x, norm_out);
  count_launch();
  SPB_CUDA(cudaGetLastError());
}

void launch_tps_eval(const double* xnew, int pnew, const double* xorig, int p, int d,
                     const double* para, size_t ldpara, int K, double* out, cudaStream_t s) {
  size_t total = (size_t)pnew * K;
  if (total == 0) return;
  int grid = ceil_div(total, 128);
  switch (d) {
    case 1: tps_eval_kernel<1><<<grid, 128, 0, s>>>(xnew, pnew, xorig, p, para, ldpara, K, out); break;
    case 2: tps_eval_kernel<2><<<grid, 128, 0, s>>>(xnew, pnew, xorig, p, para, ldpara, K, out); break;
    case 3: tps_eval_kernel<3><<<grid, 128, 0, s>>>(xnew, pnew, xorig, p, para, ldpara, K, out); break;
    default: throw Error(SPB_ERR_INVALID_ARG, "location dimension must be 1, 2 or 3");
  }
  count_launch();
  SPB_CUDA(cudaGetLastError());
}

}  // namespace spb
