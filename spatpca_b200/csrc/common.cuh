// Shared host/device helpers for the spatpca_b200 engine.
#pragma once

#include <cuda_runtime.h>
#include <cublas_v2.h>
#include <cusolverDn.h>

#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

#include <nvtx3/nvToolsExt.h>

#include "../../include/spatpca_b200.h"

namespace spb {

// Internal exception carrying an ABI status; translated at the extern "C" boundary.
struct Error : public std::runtime_error {
  int code;
  Error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};

void set_last_error(const std::string& s);

#define SPB_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess)                                                               \
      throw ::spb::Error(SPB_ERR_CUDA, std::string("CUDA error: ") + cudaGetErrorString(_e) + \
                                           " at " + __FILE__ + ":" + std::to_string(__LINE__)); \
  } while (0)

#define SPB_CUBLAS(expr)                                                                 \
  do {                                                                                   \
    cublasStatus_t _s = (expr);                                                          \
    if (_s != CUBLAS_STATUS_SUCCESS)                                                     \
      throw ::spb::Error(SPB_ERR_CUDA, std::string("cuBLAS error ") + std::to_string((int)_s) + \
                                           " at " + __FILE__ + ":" + std::to_string(__LINE__)); \
  } while (0)

#define SPB_CUSOLVER(expr)                                                               \
  do {                                                                                   \
    cusolverStatus_t _s = (expr);                                                        \
    if (_s != CUSOLVER_STATUS_SUCCESS)                                                   \
      throw ::spb::Error(SPB_ERR_CUDA, std::string("cuSOLVER error ") + std::to_string((int)_s) + \
                                           " at " + __FILE__ + ":" + std::to_string(__LINE__)); \
  } while (0)

#define SPB_REQUIRE(cond, msg)                                                           \
  do {                                                                                   \
    if (!(cond)) throw ::spb::Error(SPB_ERR_INVALID_ARG, std::string(msg));              \
  } while (0)

inline size_t round_up(size_t x, size_t m) { return (x + m - 1) / m * m; }
inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// Leading dimension used for every p x p device matrix: columns start 128-byte
// aligned so that 16-byte vector loads and bulk copies never straddle a column.
inline size_t padded_ld(size_t rows) { return round_up(rows, 16); }

// Size-bucketed caching allocator.  cudaMalloc / cudaFree synchronise the device and were seen to
// stall for 10-500 ms inside a multi-second job (profiles/r01_notes.md), so released buffers are
// kept per device and handed back on the next request of the same size; spb_trim_memory() (or a
// failed cudaMalloc) returns the cache to the driver.
void* cached_alloc(size_t bytes);
void cached_free(void* ptr, size_t bytes);
void cached_trim();

// RAII device buffer.
template <typename T>
struct DevBuf {
  T* ptr = nullptr;
  size_t count = 0;
  DevBuf() = default;
  explicit DevBuf(size_t n) { alloc(n); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : ptr(o.ptr), count(o.count) { o.ptr = nullptr; o.count = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { release(); ptr = o.ptr; count = o.count; o.ptr = nullptr; o.count = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  void alloc(size_t n) {
    release();
    if (n == 0) return;
    ptr = static_cast<T*>(cached_alloc(n * sizeof(T)));
    count = n;
  }
  void ensure(size_t n) { if (n > count) alloc(n); }
  void release() { if (ptr) { cached_free(ptr, count * sizeof(T)); ptr = nullptr; count = 0; } }
  T* get() const { return ptr; }
  operator T*() const { return ptr; }
};

// NVTX range for the host-side span that ENQUEUES a stage (tracing); shows up in nsys / ncu
// timelines next to the kernels it launched.  Header-only NVTX3: no link dependency, a no-op without a tool attached.
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

// Kernel-launch counter (claimed in bench.py as `gpu_launches`): only launches of
// this library's own kernels are counted, never cuBLAS / cuSOLVER internals.
extern thread_local long long g_launch_count;
inline void count_launch(int n = 1) { g_launch_count += n; }

}  // namespace spb
