// Internal launcher declarations (one per kernel family).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace bb {

// generic_kernels.cu
int launch_elementwise(int op, const View& out, const View* operands, int n_operands,
                       cudaStream_t stream);
int launch_strided_copy(const View& out, const View& in, cudaStream_t stream);
int launch_fill(float* out, int64_t n, float value, cudaStream_t stream);
int launch_eye(float* out, int64_t n, cudaStream_t stream);
int launch_f64_to_f32(const double* in, float* out, int64_t n, cudaStream_t stream);
int64_t reduce_sum_scratch_bytes(int64_t kept_total);
int launch_reduce_sum(const View& in, const bool* reduce_axis, const View& out, void* scratch,
                      cudaStream_t stream);
int64_t gemm_workspace_bytes(int64_t M, int64_t N, int64_t K, int64_t batch);
// tall-skinny matrix-vector contractions: out[d] = sum_n X[n,d] y[n]; out[n] = sum_d X[n,d] w[d]
bool gemv_cols_supported(int64_t n, int64_t d, const void* x);
int64_t gemv_cols_workspace(int64_t n, int64_t d);
int launch_gemv_cols(const float* x, const float* y, int64_t n, int d, float* out, void* workspace,
                     cudaStream_t stream);
int launch_gemv_rows(const float* x, const float* w, int64_t n, int d, float* out, cudaStream_t stream);
int launch_gemm(const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K,
                int64_t batch, int64_t sAb, int64_t sAm, int64_t sAk, int64_t sBb, int64_t sBk,
                int64_t sBn, void* workspace, cudaStream_t stream);

// suffstats_sm100.cu
#define BB_GAUSSIAN_PASS_SLICES 128
// What the statistics kernel does after its main loop, in the same launch: where the reduced
// statistics go, the optional expected-log-likelihood consumer, and (world > 1) the peer-memory
// exchange.  d, scratch and ticket are filled in by the launcher.
struct SuffstatsTail {
  double* s2;                  // [d, d] out (added to when accumulate)
  double* s1;                  // [d] out, may be null
  double* count_out;           // reduced row count out, may be null
  int d, accumulate;
  const double* e_lambda;      // consumer: E_q[sum_n log N(x_n | mu, Lambda^-1)] -> loglik (null: none)
  const double* e_lambda_mu;
  double e_mu_l_mu, e_logdet, n_total;   // n_total is used when world == 1; the reduced count otherwise
  double* loglik;
  double* accum;               // [64 * 64 + 64] float64 accumulator block, zero between launches (null: carved from
                               // the workspace and zeroed per launch)
  unsigned int* ticket;        // completion ticket, zero between launches
  unsigned int* tile_counter;  // dynamic tile scheduler: next unclaimed tile, zero at launch (null: static partition);
                               // launches that may overlap (pdl) alternate between two counters
  int pdl;                     // launch with programmatic stream serialization (handle-owned persistent state only)
  double local_count;          // this rank's row count (payload element d*d + d)
  int rank, world;
  double* const* peer_recv;    // device array [world]: receive buffers, [2][world][stride] float64 each
  uint32_t* const* peer_flags; // device array [world]: flag arrays, [world][BB_GAUSSIAN_PASS_SLICES] uint32 each (word 0 of a row is used)
  int64_t stride;
  uint32_t* epoch_dev;         // device word: epochs completed so far (this launch is stored + 1)
  long long spin_limit;        // clock64 ticks
  int* status;                 // 1 + rank of a lost peer
};
bool suffstats_tc_supported(int64_t n, int d, const void* x);
int64_t suffstats_tc_workspace(int64_t n);
int launch_suffstats_tc_fused(const float* x, int64_t n, int d, void* workspace, int64_t workspace_bytes,
                              SuffstatsTail tail, cudaStream_t stream);
int launch_suffstats_tc_loglik(const float* x, int64_t n, int d, double* s1, double* s2, double n_total,
                               const double* e_lambda, const double* e_lambda_mu, double e_mu_l_mu,
                               double e_logdet, double* loglik, void* workspace, int64_t workspace_bytes,
                               cudaStream_t stream);
int launch_suffstats_tc(const float* x, int64_t n, int d, double* s1, double* s2, void* workspace,
                        int64_t workspace_bytes, cudaStream_t stream);

// gram_sm100.cu: X^T X (+ X^T y, y^T y) for d % 4 == 0, 64 < d <= 4096 on tcgen05 CTA pairs (BF16x3)
bool gram_tc_supported(int64_t n, int d, const void* x);
int64_t gram_tc_workspace(int64_t n, int d);
int launch_gram_tc(const float* x, const float* y, int64_t n, int d, double* xtx, double* xty,
                   double* yty, void* workspace, int64_t workspace_bytes, cudaStream_t stream);

// rowproj_sm100.cu: Z = X W^T per data row on tcgen05 (W resident in shared memory), optional
// fused logistic epilogue (out = y - sigmoid(z), colsum[q] = sum_n y z - log(1 + exp z))
bool rowproj_tc_supported(int64_t n, int d, int q, const void* x);
int64_t rowproj_tc_workspace(int64_t n, int d, int q);
int launch_rowproj_tc(const float* x, const float* w, const float* y, int64_t n, int d, int q, float* out,
                      double* colsum, void* workspace, int64_t workspace_bytes, cudaStream_t stream);

// colproj_sm100.cu: G = X^T R over the data axis on tcgen05 (float64 out)
bool colproj_tc_supported(int64_t n, int d, int q, const void* x, const void* r);
int64_t colproj_tc_workspace(int64_t n, int d, int q);
int launch_colproj_tc(const float* x, const float* r, int64_t n, int d, int q, double* out, void* workspace,
                      int64_t workspace_bytes, cudaStream_t stream);

// logistic_fused2_sm100.cu: the whole reparameterised logistic pass in one kernel -- X read once, the X tile
// resident (converted once), W streamed
bool logistic_fused2_supported(int64_t n, int d, int s, const void* x);
int64_t logistic_fused2_workspace(int64_t n, int d, int s);
int launch_logistic_fused2(const float* x, const float* y, const float* w, int64_t n, int d, int s, double* loglik,
                           double* g, void* workspace, int64_t workspace_bytes, cudaStream_t stream);

// mixture_logits_sm100.cu: logit[n,k] = c_k - 1/2 |U_k x_n - t_k|^2 (+ row log-sum-exp) on tcgen05
bool mixture_logits_supported(int64_t n, int d, int k, const void* x);
int64_t mixture_logits_workspace(int64_t n, int d, int k);
int launch_mixture_logits(const float* x, const float* u, const float* t, const float* c, int64_t n, int d, int k,
                          int upper_triangular, float* logits, float* lse, double* sum_lse, void* workspace,
                          int64_t workspace_bytes, cudaStream_t stream);

// mixture_kernels.cu
int launch_logsoftmax_rows(const float* logits, int64_t n, int k, float* log_resp, float* lse,
                           double* sum_lse, bool responsibilities, cudaStream_t stream);
int64_t weighted_stats_workspace(int64_t n, int d, int k);
int launch_weighted_stats(const float* x, const float* r, int64_t n, int d, int k, double* nk,
                          double* sum_rx, double* sum_rxx, void* workspace,
                          int64_t workspace_bytes, cudaStream_t stream);

// weighted_sm100.cu
bool weighted_tc_supported(int64_t n, int d, int k, const void* x, const void* r);
int64_t weighted_tc_workspace(int64_t n, int k);
int launch_weighted_stats_tc(const float* x, const float* r, int64_t n, int d, int k, double* nk,
                             double* sum_rx, double* sum_rxx, void* workspace,
                             int64_t workspace_bytes, cudaStream_t stream);
// weighted_pairs_sm100.cu: the same statistics as R^T . (X (x) X) on BF16 tcgen05 (d % 8 == 0, k <= 256)
bool weighted_pairs_supported(int64_t n, int d, int k, const void* x, const void* r);
int64_t weighted_pairs_workspace(int64_t n, int d, int k);
int launch_weighted_pairs(const float* x, const float* r, const float* lse, int64_t n, int d, int k, double* nk,
                          double* sum_rx, double* sum_rxx, void* workspace, int64_t workspace_bytes,
                          cudaStream_t stream);
// pre-split responsibilities (BF16 operand tiles, see weighted_pairs_sm100.cu): layout producer in mixture_kernels.cu
bool weighted_pairs_split_supported(int64_t n, int d, int k, const void* x, const void* rsplit);
int64_t weighted_pairs_split_stages(int64_t n);
int64_t weighted_pairs_split_bytes(int64_t n, int k);
int launch_weighted_pairs_split(const float* x, const void* rsplit, int64_t n, int d, int k, double* nk,
                                double* sum_rx, double* sum_rxx, void* workspace, int64_t workspace_bytes,
                                cudaStream_t stream);
int launch_softmax_rows_split(const float* logits, int64_t n, int k, void* rsplit, float* lse, double* sum_lse,
                              cudaStream_t stream);
// tcgen05 kernel when the shape allows it (and BB_WEIGHTED_SIMT is unset), SIMT kernel otherwise
int64_t weighted_stats_auto_workspace(int64_t n, int d, int k);
int launch_weighted_stats_auto(const float* x, const float* r, int64_t n, int d, int k, double* nk,
                               double* sum_rx, double* sum_rxx, void* workspace,
                               int64_t workspace_bytes, cudaStream_t stream);

// update_kernels.cu: parameter-space steps (VMP global update, SVI blend, reparameterised gradient, Adam)
int launch_gmm_global_update(const double* nk, const double* sum_rx, const double* sum_rxx, int k, int d,
                             double alpha0, double beta0, double nu0, const double* m0, const double* w0_inv,
                             double* alpha, double* beta, double* nu, double* m, double* w_inv, float* u, float* t,
                             float* c, double* kl, int* status, cudaStream_t stream);
int launch_gather_rows(const float* x, int64_t n, int d, const int64_t* idx, int64_t m, float* out, int* bad,
                       cudaStream_t stream);
int launch_svi_blend(double* eta, const double* eta_prior, const double* stat, double scale, double rho,
                     int64_t count, cudaStream_t stream);
int launch_reparam_draws(const double* mu, const double* log_sigma, const double* eps, int d, int s, float* w,
                         cudaStream_t stream);
int launch_reparam_gradient(const double* g, const double* loglik, const double* eps, const double* mu,
                            const double* log_sigma, int d, int s, double* grad_mu, double* grad_ls, double* elbo,
                            cudaStream_t stream);
int launch_adam_step(double* param, const double* grad, double* m, double* v, int64_t count, double lr, double b1,
                     double b2, double eps, int64_t step, int maximize, cudaStream_t stream);

// p2p_reduce.cu: two-shot all-reduce over peer memory, counterpart-CTA handshakes (no grid barrier)
#define BB_COMM_MAX_CTAS 64
int comm_grid_for(int64_t count, int world);
int launch_p2p_allreduce(const double* const* in, double* const* out, uint32_t* const* flags, int rank, int world,
                         int64_t count, uint32_t* epoch_dev, unsigned int* ticket, double spin_limit_ms, int* status,
                         cudaStream_t stream);

// linalg_kernels.cu: batched log|X| of SPD float32 matrices via float64 Cholesky
int64_t logdet_scratch_bytes(int64_t batch, int64_t d);
int launch_logdet_spd(const float* x, int64_t batch, int d, float* out, void* scratch, cudaStream_t stream);

// stats_kernels.cu
int launch_f32_to_f64(const float* in, double* out, int64_t n, cudaStream_t stream);
int launch_gaussian_expected_loglik(const double* s1, const double* s2, double n,
                                    const double* e_lambda, const double* e_lambda_mu,
                                    double e_mu_l_mu, double e_logdet, int d, double* out,
                                    cudaStream_t stream);

}  // namespace bb
