// Small float64 kernels that turn sufficient statistics into ELBO terms on the device, so the
// update loop never has to synchronise with the host.
//
// gaussian_expected_loglik: E_q[ log N(X | mu, Lambda^-1) ] summed over the batch, i.e.
//   log_likelihood = data_term + interaction_term - log_normalizer
//                                                   (bayesic/distribution/base.py:25-100)
// with the multivariate-normal parametrisation  s = (x, x x^T),  eta = (Lambda mu, -1/2 Lambda)
// (distribution/core.py:41-47) and log-normaliser -1/2 D log 2pi ... (core.py:49-52), the
// interaction term being sum_i <flatten s_i, flatten eta_i> (base.py:279-291) and the
// normaliser multiplied by the number of draws (base.py:235-242):
//   out = -n D/2 log(2 pi) + n/2 E[log|Lambda|] - 1/2 tr(E[Lambda] S2) + S1 . E[Lambda mu]
//         - n/2 E[mu^T Lambda mu]
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace bb {

namespace {

__global__ void f32_to_f64_kernel(const float* in, double* out, int64_t n) {
  const int64_t step = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += step)
    out[i] = static_cast<double>(in[i]);
}

__global__ void __launch_bounds__(256)
gaussian_expected_loglik_kernel(const double* __restrict__ s1, const double* __restrict__ s2,
                                double n, const double* __restrict__ e_lambda,
                                const double* __restrict__ e_lambda_mu, double e_mu_l_mu,
                                double e_logdet, int d, double* __restrict__ out) {
  __shared__ double part[256];
  double acc = 0.0;
  for (int i = threadIdx.x; i < d * d; i += blockDim.x) acc -= 0.5 * e_lambda[i] * s2[i];
  for (int i = threadIdx.x; i < d; i += blockDim.x) acc += s1[i] * e_lambda_mu[i];
  part[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) part[threadIdx.x] += part[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double log_2pi = 1.8378770664093454835606594728112;
    out[0] = part[0] - 0.5 * n * d * log_2pi + 0.5 * n * e_logdet - 0.5 * n * e_mu_l_mu;
  }
}

}  // namespace

int launch_f32_to_f64(const float* in, double* out, int64_t n, cudaStream_t stream) {
  if (n == 0) return BB_OK;
  const int64_t blocks = std::min<int64_t>((n + 255) / 256, 1184);
  f32_to_f64_kernel<<<static_cast<int>(std::max<int64_t>(1, blocks)), 256, 0, stream>>>(in, out, n);
  BB_CHECK_LAUNCH("f32_to_f64_kernel");
  return BB_OK;
}

int launch_gaussian_expected_loglik(const double* s1, const double* s2, double n,
                                    const double* e_lambda, const double* e_lambda_mu,
                                    double e_mu_l_mu, double e_logdet, int d, double* out,
                                    cudaStream_t stream) {
  gaussian_expected_loglik_kernel<<<1, 256, 0, stream>>>(s1, s2, n, e_lambda, e_lambda_mu, e_mu_l_mu,
                                                         e_logdet, d, out);
  BB_CHECK_LAUNCH("gaussian_expected_loglik_kernel");
  return BB_OK;
}

}  // namespace bb
