// Batched log-determinant of symmetric positive-definite matrices: out[b] = log|X[b]|, X[b] float32
// [d, d] row-major, as 2 sum_j log L_jj of the Cholesky factor, factored in float64.
//
// What it replaces: the reference's MultivariateNormal log-normaliser calls `T.logdet(precision)`
// (bayesic/distribution/core.py:49-52, :51) -- an op Theano never had, so the reference cannot evaluate
// it; it is the one piece of vocabulary the Wishart / Gaussian-Wishart families and
// an MVN log-normaliser written as an expression need beyond log/exp/pow (BB_NODE_LOGDET).
//
// Parameter-space arithmetic (K matrices of D x D, K = 256, D = 64 at cfg3): latency-bound, one CTA per
// matrix.  d <= 128: the matrix lives in shared memory (d^2 float64 <= 128 KB); larger d: the same
// right-looking column Cholesky on a float64 copy in the caller's scratch (one CTA, ~d^3/3 FMAs --
// milliseconds at d = 1024; a convenience path, not a hot one).  A non-positive pivot gives NaN, like
// log of a non-positive number.
#include <math_constants.h>

#include "common.cuh"
#include "kernels.h"

namespace bb {
namespace {

constexpr int kLogdetThreads = 256;
constexpr int kLogdetSmemDim = 128;

__global__ void __launch_bounds__(kLogdetThreads)
logdet_spd_kernel(const float* __restrict__ x, int d, float* __restrict__ out, double* __restrict__ scratch) {
  extern __shared__ double smem_a[];
  __shared__ double piv_s;
  const int b = blockIdx.x, t = threadIdx.x;
  const float* src = x + static_cast<int64_t>(b) * d * d;
  double* a = scratch != nullptr ? scratch + static_cast<int64_t>(b) * d * d : smem_a;
  for (int i = t; i < d * d; i += kLogdetThreads) a[i] = static_cast<double>(src[i]);
  __syncthreads();
  double logdet = 0.0;                       // thread 0 only
  bool bad = false;
  for (int j = 0; j < d; ++j) {
    if (t == 0) {
      const double p = a[j * d + j];
      if (!(p > 0.0)) bad = true;
      const double l = sqrt(p);
      logdet += 2.0 * log(l);
      piv_s = l;
    }
    __syncthreads();
    const double inv = 1.0 / piv_s;
    for (int i = j + 1 + t; i < d; i += kLogdetThreads) a[i * d + j] *= inv;
    __syncthreads();
    // trailing lower triangle: a[i][k] -= a[i][j] a[k][j],  j < k <= i < d
    const int m = d - j - 1;
    for (int e = t; e < m * m; e += kLogdetThreads) {
      const int i = j + 1 + e / m, k = j + 1 + e % m;
      if (k <= i) a[i * d + k] -= a[i * d + j] * a[k * d + j];
    }
    __syncthreads();
  }
  if (t == 0) out[b] = bad ? CUDART_NAN_F : static_cast<float>(logdet);
}

}  // namespace

int64_t logdet_scratch_bytes(int64_t batch, int64_t d) {
  return d > kLogdetSmemDim ? batch * d * d * static_cast<int64_t>(sizeof(double)) : 0;
}

int launch_logdet_spd(const float* x, int64_t batch, int d, float* out, void* scratch, cudaStream_t stream) {
  if (batch < 0 || d < 0 || batch > 2147483647LL) { set_error("logdet: bad extents"); return BB_ERR_INVALID; }
  if (batch == 0) return BB_OK;
  if (d == 0) return launch_fill(out, batch, 0.f, stream);        // determinant of the empty matrix is 1
  if (x == nullptr || out == nullptr || (d > kLogdetSmemDim && scratch == nullptr)) {
    set_error("logdet: null argument");
    return BB_ERR_INVALID;
  }
  const int smem = d > kLogdetSmemDim ? 0 : d * d * static_cast<int>(sizeof(double));
  static SmemOptIn opt_in;
  BB_CUDA_OK(opt_in.ensure(logdet_spd_kernel, kLogdetSmemDim * kLogdetSmemDim * static_cast<int>(sizeof(double))));
  logdet_spd_kernel<<<static_cast<unsigned int>(batch), kLogdetThreads, smem, stream>>>(
      x, d, out, d > kLogdetSmemDim ? static_cast<double*>(scratch) : nullptr);
  BB_CHECK_LAUNCH("logdet_spd_kernel");
  return BB_OK;
}

}  // namespace bb
