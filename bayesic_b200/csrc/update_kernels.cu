// Parameter-space update steps either side of the data pass, as device kernels
// so that a VMP / SVI / reparameterised-gradient loop never synchronises with the host:
//
//   gmm_global_update   all-reduced {N_k, sum r x, sum r x x^T}  ->  Dirichlet / Gaussian-Wishart
//                       posterior, the whitened logit parameters (U_k, t_k, c_k) the next local
//                       step consumes, and the KL terms of the ELBO (Bishop PRML 10.58-10.77)
//   svi_blend           eta <- (1 - rho) eta + rho (eta_prior + scale * statistic)
//   reparam_draws       W[s, :] = mu + exp(log_sigma) * eps[s, :]            (float32 out)
//   reparam_gradient    {G[D, S], loglik[S]} -> ELBO, grad mu, grad log sigma (prior N(0, I))
//   adam_step           the optimiser step on those gradients
//
// The reference names these algorithms in prose only (README.md:30-37 VMP, :47-51 reparameterised
// gradients, :69-80 SVI); everything here is float64 arithmetic on K * D^2 or D * S numbers --
// latency-bound, one CTA per mixture component.
#include <math_constants.h>

#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace bb {
namespace {

constexpr int kUpdThreads = 256;
constexpr int kMaxUpdateDim = 96;      // 3 D^2 float64 in shared memory: 216 KB at D = 96

__device__ __forceinline__ double digamma_pos(double x) {
  // psi(x) for x > 0: recurrence up to x >= 10, then the asymptotic series (error < 1e-15 there)
  double acc = 0.0;
  while (x < 10.0) {
    acc -= 1.0 / x;
    x += 1.0;
  }
  const double inv = 1.0 / x, inv2 = inv * inv;
  const double series = inv2 * (1.0 / 12.0 - inv2 * (1.0 / 120.0 - inv2 * (1.0 / 252.0 - inv2 * (1.0 / 240.0 -
                        inv2 * (1.0 / 132.0 - inv2 * (691.0 / 32760.0 - inv2 * (1.0 / 12.0)))))));
  return acc + log(x) - 0.5 * inv - series;
}

__device__ __forceinline__ double block_sum(double v, double* scratch) {
  // scratch: kUpdThreads / 32 doubles; every thread gets the total
  for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  double total = 0.0;
  for (int w = 0; w < kUpdThreads / 32; ++w) total += scratch[w];
  return total;
}

// log B(W, nu) of the Wishart normaliser given log|W| and sum_i lgamma((nu - i) / 2)  (Bishop B.79)
__device__ __forceinline__ double log_wishart_b(double logdet_w, double nu, int d, double lgamma_sum) {
  return -0.5 * nu * logdet_w -
         (0.5 * nu * d * 0.6931471805599453 + 0.25 * d * (d - 1) * 1.1447298858494002 + lgamma_sum);
}

struct GmmUpdateParams {
  const double* nk;
  const double* sum_rx;
  const double* sum_rxx;
  int k, d;
  double alpha0, beta0, nu0;
  const double* m0;
  const double* w0_inv;
  const double* prior_consts;   // [1]: log|W0^-1|, written by gmm_prior_logdet_kernel
  double* alpha;
  double* beta;
  double* nu;
  double* m;
  double* w_inv;
  float* u;
  float* t;
  float* c;
  double* kl;
  int* status;
};

// In-place factorisation A = R R^T with R UPPER triangular (Cholesky taken from the last
// column backwards), so that W = A^-1 = R^-T R^-1 and U = sqrt(nu) R^-1 is upper triangular --
// the form whose zero blocks the logits kernel skips.  Returns false if A is not SPD.
__device__ bool factor_upper(double* a, int d) {
  for (int j = d - 1; j >= 0; --j) {
    __syncthreads();
    const double piv = a[j * d + j];
    if (!(piv > 0.0)) return false;          // same value for every thread
    const double rjj = sqrt(piv);
    __syncthreads();
    for (int i = threadIdx.x; i <= j; i += kUpdThreads) a[i * d + j] = (i == j) ? rjj : a[i * d + j] / rjj;
    __syncthreads();
    for (int idx = threadIdx.x; idx < j * j; idx += kUpdThreads) {
      const int i = idx / j, l = idx - i * j;
      if (l >= i) a[i * d + l] -= a[i * d + j] * a[l * d + j];
    }
  }
  __syncthreads();
  return true;
}

// rinv = R^-1 (upper triangular) by back substitution: four lanes share one column.
__device__ void invert_upper(const double* r, double* rinv, int d) {
  const int sub = threadIdx.x & 3, grp = threadIdx.x >> 2;
  for (int c0 = 0; c0 < d; c0 += kUpdThreads / 4) {
    const int c = c0 + grp;
    const bool live = c < d;
    for (int i = d - 1; i >= 0; --i) {          // rows below the diagonal are written as zeros
      double part = 0.0;
      if (live && i < c)
        for (int l = i + 1 + sub; l <= c; l += 4) part += r[i * d + l] * rinv[l * d + c];
      part += __shfl_xor_sync(0xffffffffu, part, 1);
      part += __shfl_xor_sync(0xffffffffu, part, 2);
      if (live && sub == 0) {
        if (i <= c) rinv[i * d + c] = ((i == c ? 1.0 : 0.0) - part) / r[i * d + i];
        else rinv[i * d + c] = 0.0;
      }
      __syncwarp();
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kUpdThreads) gmm_prior_logdet_kernel(const double* w0_inv, int d, double* out,
                                                                      int* status) {
  extern __shared__ double smem_prior[];
  double* a = smem_prior;
  for (int i = threadIdx.x; i < d * d; i += kUpdThreads) {
    const int r = i / d, c = i - r * d;
    a[i] = 0.5 * (w0_inv[i] + w0_inv[c * d + r]);
  }
  const bool ok = factor_upper(a, d);
  if (threadIdx.x == 0) {
    double ld = 0.0;
    if (ok) for (int j = 0; j < d; ++j) ld += 2.0 * log(a[j * d + j]);
    else { ld = CUDART_NAN; atomicMax(status, 1 << 30); }
    out[0] = ld;                                  // log|W0^-1|
  }
}

__global__ void __launch_bounds__(kUpdThreads) gmm_global_update_kernel(const GmmUpdateParams p) {
  extern __shared__ double smem_upd[];
  const int d = p.d, comp = blockIdx.x, tid = threadIdx.x;
  double* a = smem_upd;                 // W_k^-1, then its factor R (upper), then W0^-1 R^-T
  double* rinv = a + d * d;             // R^-1
  double* w0 = rinv + d * d;            // symmetrised W0^-1
  __shared__ double red[kUpdThreads / 32];
  __shared__ double mk[kMaxUpdateDim], dm[kMaxUpdateDim];

  const double n_k = p.nk[comp];
  const double beta_k = p.beta0 + n_k, nu_k = p.nu0 + n_k, alpha_k = p.alpha0 + n_k;
  double alpha_sum_part = 0.0;
  for (int j = tid; j < p.k; j += kUpdThreads) alpha_sum_part += p.alpha0 + p.nk[j];
  const double alpha_sum = block_sum(alpha_sum_part, red);
  for (int i = tid; i < d; i += kUpdThreads) {
    const double v = (p.beta0 * p.m0[i] + p.sum_rx[static_cast<int64_t>(comp) * d + i]) / beta_k;
    mk[i] = v;
    dm[i] = v - p.m0[i];
    p.m[static_cast<int64_t>(comp) * d + i] = v;
  }
  __syncthreads();
  // W_k^-1 = W0^-1 + sum r x x^T + beta0 m0 m0^T - beta_k m_k m_k^T   (Bishop 10.62 regrouped so that
  // an empty component needs no division by N_k); symmetrised
  const double* sxx = p.sum_rxx + static_cast<int64_t>(comp) * d * d;
  double* w_inv_out = p.w_inv + static_cast<int64_t>(comp) * d * d;
  for (int i = tid; i < d * d; i += kUpdThreads) {
    const int r = i / d, c = i - r * d;
    const double w0_rc = 0.5 * (p.w0_inv[i] + p.w0_inv[c * d + r]);
    w0[i] = w0_rc;
    const double v = w0_rc + 0.5 * (sxx[i] + sxx[c * d + r]) +
                     p.beta0 * p.m0[r] * p.m0[c] - beta_k * mk[r] * mk[c];
    a[i] = v;
    w_inv_out[i] = v;
  }
  if (tid == 0) {
    p.alpha[comp] = alpha_k;
    p.beta[comp] = beta_k;
    p.nu[comp] = nu_k;
  }
  const bool ok = factor_upper(a, d);
  if (!ok) {
    if (tid == 0) {
      atomicMax(p.status, comp + 1);
      p.c[comp] = CUDART_NAN_F;
      p.kl[comp] = CUDART_NAN;
    }
    for (int i = tid; i < d * d; i += kUpdThreads) p.u[static_cast<int64_t>(comp) * d * d + i] = CUDART_NAN_F;
    for (int i = tid; i < d; i += kUpdThreads) p.t[static_cast<int64_t>(comp) * d + i] = CUDART_NAN_F;
    return;
  }
  double ld_part = 0.0;
  for (int j = tid; j < d; j += kUpdThreads) ld_part += 2.0 * log(a[j * d + j]);
  const double logdet_winv = block_sum(ld_part, red);            // log|W_k^-1|
  invert_upper(a, rinv, d);
  // U = sqrt(nu) R^-1, t = U m
  const double snu = sqrt(nu_k);
  float* u_out = p.u + static_cast<int64_t>(comp) * d * d;
  for (int i = tid; i < d * d; i += kUpdThreads) u_out[i] = static_cast<float>(snu * rinv[i]);
  for (int j = tid; j < d; j += kUpdThreads) {
    double acc = 0.0;
    for (int i = j; i < d; ++i) acc += rinv[j * d + i] * mk[i];
    p.t[static_cast<int64_t>(comp) * d + j] = static_cast<float>(snu * acc);
  }
  // expectations
  double psi_part = 0.0;
  for (int i = tid; i < d; i += kUpdThreads) psi_part += digamma_pos(0.5 * (nu_k - i));
  const double e_logdet = block_sum(psi_part, red) + d * 0.6931471805599453 - logdet_winv;   // E log|Lambda_k|
  const double e_log_pi = digamma_pos(alpha_k) - digamma_pos(alpha_sum);
  if (tid == 0)
    p.c[comp] = static_cast<float>(e_log_pi + 0.5 * e_logdet - 0.5 * d * 1.8378770664093453 - 0.5 * d / beta_k);
  // KL(q(mu_k, Lambda_k) || p(mu, Lambda)): needs tr(W0^-1 W_k) and (m_k - m0)^T W_k (m_k - m0), W_k = R^-T R^-1
  double quad_part = 0.0;
  for (int j = tid; j < d; j += kUpdThreads) {
    double acc = 0.0;
    for (int i = j; i < d; ++i) acc += rinv[j * d + i] * dm[i];
    quad_part += acc * acc;
  }
  const double quad = block_sum(quad_part, red);
  // tr(W0^-1 R^-T R^-1) = sum_j rinv[j,:] W0^-1 rinv[j,:]^T ; a (R) is no longer needed: a <- W0^-1 rinv^T
  __syncthreads();
  for (int idx = tid; idx < d * d; idx += kUpdThreads) {
    const int r = idx / d, j = idx - r * d;        // a[r][j] = sum_b W0inv[r][b] rinv[j][b]
    double acc = 0.0;
    for (int b = j; b < d; ++b) acc += w0[r * d + b] * rinv[j * d + b];
    a[idx] = acc;
  }
  __syncthreads();
  double tr_part = 0.0;
  for (int idx = tid; idx < d * d; idx += kUpdThreads) {
    const int r = idx / d, j = idx - r * d;
    if (r >= j) tr_part += rinv[j * d + r] * a[idx];
  }
  const double tr = block_sum(tr_part, red);
  double lg_k_part = 0.0, lg_0_part = 0.0;
  for (int i = tid; i < d; i += kUpdThreads) {
    lg_k_part += lgamma(0.5 * (nu_k - i));
    lg_0_part += lgamma(0.5 * (p.nu0 - i));
  }
  const double lg_k = block_sum(lg_k_part, red), lg_0 = block_sum(lg_0_part, red);
  if (tid == 0) {
    const double logdet_w0 = -p.prior_consts[0];
    const double logdet_wk = -logdet_winv;
    const double kl_wishart = log_wishart_b(logdet_wk, nu_k, d, lg_k) - log_wishart_b(logdet_w0, p.nu0, d, lg_0) +
                              0.5 * (nu_k - p.nu0) * e_logdet - 0.5 * nu_k * d + 0.5 * nu_k * tr;
    const double kl_gauss = 0.5 * (d * p.beta0 / beta_k + p.beta0 * nu_k * quad - d + d * log(beta_k / p.beta0));
    p.kl[comp] = kl_wishart + kl_gauss;
  }
  // KL(q(pi) || p(pi)) of the Dirichlet factor, by the first CTA
  if (comp == 0) {
    double part = 0.0;
    for (int j = tid; j < p.k; j += kUpdThreads) {
      const double aj = p.alpha0 + p.nk[j];
      part += -lgamma(aj) + lgamma(p.alpha0) + (aj - p.alpha0) * (digamma_pos(aj) - digamma_pos(alpha_sum));
    }
    const double s = block_sum(part, red);
    if (tid == 0) p.kl[p.k] = lgamma(alpha_sum) - lgamma(p.k * p.alpha0) + s;
  }
}

__global__ void svi_blend_kernel(double* eta, const double* eta_prior, const double* stat, double scale, double rho,
                                 int64_t count) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < count;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    eta[i] = (1.0 - rho) * eta[i] + rho * (eta_prior[i] + scale * stat[i]);
}

__global__ void reparam_draws_kernel(const double* mu, const double* log_sigma, const double* eps, int d, int s,
                                     float* w) {
  const int64_t total = static_cast<int64_t>(d) * s;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int j = static_cast<int>(i % d);
    w[i] = static_cast<float>(mu[j] + exp(log_sigma[j]) * eps[i]);
  }
}

// grid of ceil(d / 256) CTAs over features; CTA 0 also forms the ELBO.  KL per feature is
// accumulated with one atomicAdd per CTA into elbo[0] (zeroed by the launcher).
__global__ void __launch_bounds__(kUpdThreads) reparam_gradient_kernel(const double* g, const double* loglik,
                                                                      const double* eps, const double* mu,
                                                                      const double* log_sigma, int d, int s,
                                                                      double* grad_mu, double* grad_ls, double* elbo) {
  __shared__ double red[kUpdThreads / 32];
  const int j = blockIdx.x * kUpdThreads + threadIdx.x;
  double kl = 0.0;
  if (j < d) {
    const double sg = exp(log_sigma[j]);
    double gm = 0.0, gs = 0.0;
    for (int i = 0; i < s; ++i) {
      const double v = g[static_cast<int64_t>(j) * s + i];
      gm += v;
      gs += v * eps[static_cast<int64_t>(i) * d + j];
    }
    grad_mu[j] = gm / s - mu[j];
    grad_ls[j] = gs / s * sg - sg * sg + 1.0;
    kl = 0.5 * (sg * sg + mu[j] * mu[j] - 1.0 - 2.0 * log_sigma[j]);
  }
  double total = -block_sum(kl, red);
  if (blockIdx.x == 0) {
    double ll = 0.0;
    for (int i = threadIdx.x; i < s; i += kUpdThreads) ll += loglik[i];
    total += block_sum(ll, red) / s;
  }
  if (threadIdx.x == 0) atomicAdd(elbo, total);
}

__global__ void adam_step_kernel(double* param, const double* grad, double* m, double* v, int64_t count, double lr,
                                 double b1, double b2, double eps, double corr1, double corr2, double sign) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < count;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const double gi = grad[i];
    const double mi = b1 * m[i] + (1.0 - b1) * gi;
    const double vi = b2 * v[i] + (1.0 - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    param[i] += sign * lr * (mi / corr1) / (sqrt(vi / corr2) + eps);
  }
}

// out[j, :] = x[idx[j], :]; rows with an index outside [0, n) are NaN-filled and counted in *bad
__global__ void gather_rows_kernel(const float* __restrict__ x, int64_t n, int d, const int64_t* __restrict__ idx,
                                   int64_t m, float* __restrict__ out, int* __restrict__ bad) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int64_t n_warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  const bool vec = (d & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  for (int64_t j = warp; j < m; j += n_warps) {
    const int64_t src = idx[j];
    const bool ok = src >= 0 && src < n;
    if (!ok && lane == 0) atomicAdd(bad, 1);
    float* dst = out + j * d;
    if (vec) {
      const float4* s4 = reinterpret_cast<const float4*>(x + (ok ? src : 0) * d);
      float4* d4 = reinterpret_cast<float4*>(dst);
      for (int c = lane; c < d / 4; c += 32)
        d4[c] = ok ? __ldg(s4 + c) : make_float4(CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F);
    } else {
      for (int c = lane; c < d; c += 32) dst[c] = ok ? __ldg(x + src * d + c) : CUDART_NAN_F;
    }
  }
}

SmemOptIn g_optin_update, g_optin_prior;

inline int elementwise_grid(int64_t count) {
  return static_cast<int>(std::min<int64_t>((count + 255) / 256, 148 * 8));
}

}  // namespace

int gmm_update_max_dim() { return kMaxUpdateDim; }

int launch_gmm_global_update(const double* nk, const double* sum_rx, const double* sum_rxx, int k, int d,
                             double alpha0, double beta0, double nu0, const double* m0, const double* w0_inv,
                             double* alpha, double* beta, double* nu, double* m, double* w_inv, float* u, float* t,
                             float* c, double* kl, int* status, cudaStream_t stream) {
  if (k < 1 || d < 1 || d > kMaxUpdateDim) {
    set_error("gmm_global_update: need k >= 1 and 1 <= d <= %d (got k=%d d=%d)", kMaxUpdateDim, k, d);
    return BB_ERR_UNSUPPORTED;
  }
  if (!(alpha0 > 0.0) || !(beta0 > 0.0) || !(nu0 > d - 1)) {
    set_error("gmm_global_update: need alpha0 > 0, beta0 > 0, nu0 > d - 1");
    return BB_ERR_INVALID;
  }
  BB_CUDA_OK(cudaMemsetAsync(status, 0, sizeof(int), stream));
  const int prior_smem = d * d * static_cast<int>(sizeof(double));
  const int smem = 3 * prior_smem;
  // the opt-in is made once per device, so it must cover the largest d
  constexpr int kMaxMatrixBytes = kMaxUpdateDim * kMaxUpdateDim * static_cast<int>(sizeof(double));
  BB_CUDA_OK(g_optin_prior.ensure(gmm_prior_logdet_kernel, kMaxMatrixBytes));
  BB_CUDA_OK(g_optin_update.ensure(gmm_global_update_kernel, 3 * kMaxMatrixBytes));
  // kl has k + 2 slots: per-component KL, the Dirichlet KL, and log|W0^-1| (the prior constant the
  // per-component CTAs read)
  gmm_prior_logdet_kernel<<<1, kUpdThreads, prior_smem, stream>>>(w0_inv, d, kl + k + 1, status);
  BB_CHECK_LAUNCH("gmm_prior_logdet_kernel");
  GmmUpdateParams p;
  p.nk = nk; p.sum_rx = sum_rx; p.sum_rxx = sum_rxx; p.k = k; p.d = d;
  p.alpha0 = alpha0; p.beta0 = beta0; p.nu0 = nu0; p.m0 = m0; p.w0_inv = w0_inv;
  p.prior_consts = kl + k + 1;
  p.alpha = alpha; p.beta = beta; p.nu = nu; p.m = m; p.w_inv = w_inv; p.u = u; p.t = t; p.c = c; p.kl = kl;
  p.status = status;
  gmm_global_update_kernel<<<k, kUpdThreads, smem, stream>>>(p);
  BB_CHECK_LAUNCH("gmm_global_update_kernel");
  return BB_OK;
}

int launch_gather_rows(const float* x, int64_t n, int d, const int64_t* idx, int64_t m, float* out, int* bad,
                       cudaStream_t stream) {
  BB_CUDA_OK(cudaMemsetAsync(bad, 0, sizeof(int), stream));
  if (m == 0 || d == 0) return BB_OK;
  const int64_t warps = std::min<int64_t>(m, 148 * 32);
  gather_rows_kernel<<<static_cast<int>((warps * 32 + 255) / 256), 256, 0, stream>>>(x, n, d, idx, m, out, bad);
  BB_CHECK_LAUNCH("gather_rows_kernel");
  return BB_OK;
}

int launch_svi_blend(double* eta, const double* eta_prior, const double* stat, double scale, double rho,
                     int64_t count, cudaStream_t stream) {
  if (count == 0) return BB_OK;
  svi_blend_kernel<<<elementwise_grid(count), 256, 0, stream>>>(eta, eta_prior, stat, scale, rho, count);
  BB_CHECK_LAUNCH("svi_blend_kernel");
  return BB_OK;
}

int launch_reparam_draws(const double* mu, const double* log_sigma, const double* eps, int d, int s, float* w,
                         cudaStream_t stream) {
  const int64_t total = static_cast<int64_t>(d) * s;
  if (total == 0) return BB_OK;
  reparam_draws_kernel<<<elementwise_grid(total), 256, 0, stream>>>(mu, log_sigma, eps, d, s, w);
  BB_CHECK_LAUNCH("reparam_draws_kernel");
  return BB_OK;
}

int launch_reparam_gradient(const double* g, const double* loglik, const double* eps, const double* mu,
                            const double* log_sigma, int d, int s, double* grad_mu, double* grad_ls, double* elbo,
                            cudaStream_t stream) {
  BB_CUDA_OK(cudaMemsetAsync(elbo, 0, sizeof(double), stream));
  reparam_gradient_kernel<<<(d + kUpdThreads - 1) / kUpdThreads, kUpdThreads, 0, stream>>>(
      g, loglik, eps, mu, log_sigma, d, s, grad_mu, grad_ls, elbo);
  BB_CHECK_LAUNCH("reparam_gradient_kernel");
  return BB_OK;
}

int launch_adam_step(double* param, const double* grad, double* m, double* v, int64_t count, double lr, double b1,
                     double b2, double eps, int64_t step, int maximize, cudaStream_t stream) {
  if (count == 0) return BB_OK;
  const double corr1 = 1.0 - pow(b1, static_cast<double>(step)), corr2 = 1.0 - pow(b2, static_cast<double>(step));
  adam_step_kernel<<<elementwise_grid(count), 256, 0, stream>>>(param, grad, m, v, count, lr, b1, b2, eps, corr1,
                                                               corr2, maximize ? 1.0 : -1.0);
  BB_CHECK_LAUNCH("adam_step_kernel");
  return BB_OK;
}

}  // namespace bb
