// Shared host-side plumbing: status codes, thread-local error text, launch counting,
// and the strided tensor view the generic kernels operate on.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <string>

#include "../../include/bayesic_b200.h"

// Length of the fp32 TMEM accumulation chains of the BF16x3 kernels, as a divisor of the default
// 2048 rows (the tensor core truncates its fp32 accumulate, so chain length trades accuracy against
// drain traffic; profiles/r02_parity_report.txt has the measured trade).  Build-time knob.
#ifndef BB_CHAIN_DIV
#define BB_CHAIN_DIV 1
#endif

namespace bb {

void set_error(const char* fmt, ...);
const char* get_error();

// Every kernel launch of the library goes through this counter so the host can
// report `gpu_launches` (bench.py) and tests can assert the CUDA path really ran.
extern thread_local int64_t g_launch_count;
inline void note_launch(int n = 1) { g_launch_count += n; }

#define BB_CUDA_OK(expr)                                                                   \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      ::bb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,    \
                      __LINE__);                                                           \
      return BB_ERR_CUDA;                                                                  \
    }                                                                                      \
  } while (0)

#define BB_CHECK_LAUNCH(name)                                                              \
  do {                                                                                     \
    cudaError_t _e = cudaGetLastError();                                                   \
    if (_e != cudaSuccess) {                                                               \
      ::bb::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));            \
      return BB_ERR_CUDA;                                                                  \
    }                                                                                      \
    ::bb::note_launch();                                                                   \
  } while (0)

#define BB_TRY(expr)              \
  do {                            \
    int _s = (expr);              \
    if (_s != BB_OK) return _s;   \
  } while (0)

constexpr int kMaxDims = BB_MAX_DIMS;

// A (possibly strided, possibly broadcast) float32 tensor, or a host-known scalar.
struct View {
  float* ptr = nullptr;
  int ndim = 0;
  int64_t shape[kMaxDims] = {0};
  int64_t stride[kMaxDims] = {0};  // in elements; 0 on broadcast axes
  bool is_host = false;            // value known on the host (literal / shape arithmetic)
  double host_value = 0.0;

  int64_t numel() const {
    int64_t n = 1;
    for (int i = 0; i < ndim; ++i) n *= shape[i];
    return n;
  }
  bool is_contiguous() const {
    int64_t expect = 1;
    for (int i = ndim - 1; i >= 0; --i) {
      if (shape[i] != 1 && stride[i] != expect) return false;
      expect *= shape[i];
    }
    return true;
  }
  void set_contiguous_strides() {
    int64_t s = 1;
    for (int i = ndim - 1; i >= 0; --i) {
      stride[i] = s;
      s *= shape[i];
    }
  }
};

int device_sm_count();

// Function attributes are per device: remember, per kernel, on which devices the opt-in to large
// dynamic shared memory has been made (one process may drive several GPUs).
struct SmemOptIn {
  bool done[64] = {false};
  template <typename Kernel>
  cudaError_t ensure(Kernel kernel, int bytes) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64 && done[dev]) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess && dev >= 0 && dev < 64) done[dev] = true;
    return e;
  }
};

inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

}  // namespace bb
