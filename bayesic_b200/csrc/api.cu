// extern "C" surface declared in include/bayesic_b200.h.  Plain pointers and sizes only.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <cstdlib>
#include <new>
#include <vector>

#include "common.cuh"
#include "kernels.h"
#include "plan.h"


namespace bb {
int plan_infer(bb_plan*, const bb_tensor_arg*, int32_t, bb_result_info*, int64_t*);
int plan_execute(bb_plan*, const bb_tensor_arg*, int32_t, void* const*, void*, int64_t, cudaStream_t);
int plan_validate(const bb_node_desc*, int32_t, const int32_t*, int32_t, int32_t);
// suffstats_sm100.cu (accumulating variant used by the host-streaming path)
int launch_suffstats_tc_acc(const float* x, int64_t n, int d, double* s1, double* s2, void* workspace,
                            int64_t workspace_bytes, bool accumulate, cudaStream_t stream);
}  // namespace bb

using namespace bb;

extern "C" {

BB_API int bb_abi_version(void) { return BB_ABI_VERSION; }

BB_API const char* bb_last_error(void) { return get_error(); }

BB_API int64_t bb_launch_count(void) { return g_launch_count; }

BB_API int bb_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor) {
  int dev = 0;
  BB_CUDA_OK(cudaGetDevice(&dev));
  int sms = 0, major = 0, minor = 0;
  BB_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  BB_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  BB_CUDA_OK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (sm_count) *sm_count = sms;
  if (cc_major) *cc_major = major;
  if (cc_minor) *cc_minor = minor;
  return BB_OK;
}

BB_API int bb_plan_create(const bb_node_desc* nodes, int32_t n_nodes, const int32_t* outputs,
                   int32_t n_outputs, int32_t n_inputs, bb_plan** plan) {
  if (plan == nullptr) { set_error("null plan out-pointer"); return BB_ERR_INVALID; }
  *plan = nullptr;
  BB_TRY(plan_validate(nodes, n_nodes, outputs, n_outputs, n_inputs));
  bb_plan* p = new (std::nothrow) bb_plan();
  if (p == nullptr) { set_error("out of host memory"); return BB_ERR_INVALID; }
  p->nodes.assign(nodes, nodes + n_nodes);
  p->outputs.assign(outputs, outputs + n_outputs);
  p->n_inputs = n_inputs;
  *plan = p;
  return BB_OK;
}

BB_API int bb_plan_destroy(bb_plan* plan) {
  delete plan;
  return BB_OK;
}

BB_API int bb_plan_infer(bb_plan* plan, const bb_tensor_arg* inputs, int32_t n_inputs,
                  bb_result_info* results, int64_t* workspace_bytes) {
  return plan_infer(plan, inputs, n_inputs, results, workspace_bytes);
}

BB_API int bb_plan_execute(bb_plan* plan, const bb_tensor_arg* inputs, int32_t n_inputs,
                    void* const* out_ptrs, void* workspace, int64_t workspace_bytes, void* stream) {
  return plan_execute(plan, inputs, n_inputs, out_ptrs, workspace, workspace_bytes,
                      static_cast<cudaStream_t>(stream));
}

BB_API int bb_plan_last_launch_count(bb_plan* plan, int32_t* count) {
  if (plan == nullptr || count == nullptr) { set_error("null argument"); return BB_ERR_INVALID; }
  *count = plan->last_launches;
  return BB_OK;
}

// ---- fused statistics ------------------------------------------------------------

static int64_t generic_suffstats_workspace(int64_t n, int32_t d) {
  // float32 S2 + split-K partials + float32 S1 + float64 reduce scratch
  return align_up(static_cast<int64_t>(d) * d * 4, 256) + gemm_workspace_bytes(d, d, n, 1) +
         align_up(static_cast<int64_t>(d) * 4, 256) + reduce_sum_scratch_bytes(d) + 1024;
}

BB_API int64_t bb_suffstats_gaussian_workspace(int64_t n, int32_t d) {
  int64_t need = generic_suffstats_workspace(n, d);
  if (d >= 4 && d <= 64 && d % 4 == 0) need = std::max(need, suffstats_tc_workspace(n));
  return need;
}

static int suffstats_device(const float* X, int64_t n, int32_t d, double* sum_x, double* sum_xxT,
                            void* workspace, int64_t workspace_bytes, bool accumulate,
                            cudaStream_t st) {
  if ((X == nullptr && n > 0) || sum_xxT == nullptr || n < 0 || d < 1) {
    set_error("suffstats_gaussian: bad arguments (n=%lld d=%d)", static_cast<long long>(n), d);
    return BB_ERR_INVALID;
  }
  if (workspace_bytes < bb_suffstats_gaussian_workspace(n, d)) {
    set_error("suffstats_gaussian: workspace %lld < %lld bytes", static_cast<long long>(workspace_bytes),
              static_cast<long long>(bb_suffstats_gaussian_workspace(n, d)));
    return BB_ERR_WORKSPACE;
  }
  if (n == 0) {
    if (!accumulate) {
      BB_CUDA_OK(cudaMemsetAsync(sum_xxT, 0, sizeof(double) * d * d, st));
      if (sum_x) BB_CUDA_OK(cudaMemsetAsync(sum_x, 0, sizeof(double) * d, st));
    }
    return BB_OK;
  }
  if (suffstats_tc_supported(n, d, X))
    return launch_suffstats_tc_acc(X, n, d, sum_x, sum_xxT, workspace, workspace_bytes, accumulate, st);
  if (accumulate) {
    set_error("suffstats_gaussian: accumulating mode needs d <= 64, d %% 4 == 0");
    return BB_ERR_UNSUPPORTED;
  }
  // generic path: split-K GEMM for X^T X and a column reduction for sum x
  char* ws = static_cast<char*>(workspace);
  float* s2f = reinterpret_cast<float*>(ws);
  ws += align_up(static_cast<int64_t>(d) * d * 4, 256);
  void* gemm_ws = ws;
  ws += gemm_workspace_bytes(d, d, n, 1);
  float* s1f = reinterpret_cast<float*>(ws);
  ws += align_up(static_cast<int64_t>(d) * 4, 256);
  void* red_ws = ws;
  BB_TRY(launch_gemm(X, X, s2f, d, d, n, 1, 0, 1, d, 0, d, 1, gemm_ws, st));
  BB_TRY(launch_f32_to_f64(s2f, sum_xxT, static_cast<int64_t>(d) * d, st));
  if (sum_x != nullptr) {
    View in, out;
    in.ptr = const_cast<float*>(X);
    in.ndim = 2; in.shape[0] = n; in.shape[1] = d; in.set_contiguous_strides();
    out.ptr = s1f; out.ndim = 1; out.shape[0] = d; out.set_contiguous_strides();
    const bool reduce[kMaxDims] = {true, false};
    BB_TRY(launch_reduce_sum(in, reduce, out, red_ws, st));
    BB_TRY(launch_f32_to_f64(s1f, sum_x, d, st));
  }
  return BB_OK;
}

BB_API int bb_suffstats_gaussian(const float* X, int64_t n, int32_t d, double* sum_x, double* sum_xxT,
                          void* workspace, int64_t workspace_bytes, void* stream) {
  return suffstats_device(X, n, d, sum_x, sum_xxT, workspace, workspace_bytes, false,
                          static_cast<cudaStream_t>(stream));
}

BB_API int bb_suffstats_gaussian_loglik(const float* X, int64_t n, int32_t d, double* sum_x, double* sum_xxT,
                                 double n_total, const double* E_Lambda, const double* E_Lambda_mu,
                                 double E_mu_L_mu, double E_logdet, double* out, void* workspace,
                                 int64_t workspace_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!sum_x || !sum_xxT || !E_Lambda || !E_Lambda_mu || !out) {
    set_error("suffstats_gaussian_loglik: bad arguments");
    return BB_ERR_INVALID;
  }
  if (n > 0 && X != nullptr && d >= 1 && suffstats_tc_supported(n, d, X)) {
    if (workspace_bytes < bb_suffstats_gaussian_workspace(n, d)) {
      set_error("suffstats_gaussian_loglik: workspace %lld < %lld bytes", static_cast<long long>(workspace_bytes),
                static_cast<long long>(bb_suffstats_gaussian_workspace(n, d)));
      return BB_ERR_WORKSPACE;
    }
    // one launch: statistics, cross-CTA reduction, log-likelihood by the last CTA
    return launch_suffstats_tc_loglik(X, n, d, sum_x, sum_xxT, n_total, E_Lambda, E_Lambda_mu, E_mu_L_mu, E_logdet,
                                      out, workspace, workspace_bytes, st);
  }
  BB_TRY(suffstats_device(X, n, d, sum_x, sum_xxT, workspace, workspace_bytes, false, st));
  return launch_gaussian_expected_loglik(sum_x, sum_xxT, n_total, E_Lambda, E_Lambda_mu, E_mu_L_mu, E_logdet, d, out,
                                         st);
}

// Staging pool of the host-streaming entry point: two device chunk buffers, statistics and
// kernel workspace; grown on demand, kept for the life of the process.
namespace {
struct Staging {
  float* chunk[2] = {nullptr, nullptr};
  int64_t chunk_bytes = 0;
  double* stats = nullptr;      // [d*d + d]
  int64_t stats_doubles = 0;
  void* ws = nullptr;
  int64_t ws_bytes = 0;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t copied[2] = {nullptr, nullptr};
  cudaEvent_t consumed[2] = {nullptr, nullptr};
  int device = -1;
};
Staging g_staging;
}  // namespace

BB_API int bb_suffstats_gaussian_host(const float* X_host, int64_t n, int32_t d, double* sum_x_host,
                               double* sum_xxT_host, int64_t chunk_rows, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (X_host == nullptr || sum_xxT_host == nullptr || n < 0 || d < 1) {
    set_error("suffstats_gaussian_host: bad arguments");
    return BB_ERR_INVALID;
  }
  if (!(d >= 4 && d <= 64 && d % 4 == 0)) {
    set_error("suffstats_gaussian_host: needs d <= 64 and d %% 4 == 0 (got %d)", d);
    return BB_ERR_UNSUPPORTED;
  }
  if (chunk_rows <= 0) chunk_rows = int64_t(1) << 20;
  chunk_rows = std::min<int64_t>(std::max<int64_t>(chunk_rows, 128), std::max<int64_t>(n, 128));
  chunk_rows = (chunk_rows + 127) / 128 * 128;
  Staging& s = g_staging;
  int dev = 0;
  BB_CUDA_OK(cudaGetDevice(&dev));
  if (s.device != dev) {
    if (s.device >= 0) { set_error("suffstats_gaussian_host: staging pool bound to device %d", s.device); return BB_ERR_UNSUPPORTED; }
    s.device = dev;
    BB_CUDA_OK(cudaStreamCreateWithFlags(&s.copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      BB_CUDA_OK(cudaEventCreateWithFlags(&s.copied[i], cudaEventDisableTiming));
      BB_CUDA_OK(cudaEventCreateWithFlags(&s.consumed[i], cudaEventDisableTiming));
    }
  }
  const int64_t need_chunk = chunk_rows * d * static_cast<int64_t>(sizeof(float));
  if (s.chunk_bytes < need_chunk) {
    for (int i = 0; i < 2; ++i) {
      if (s.chunk[i]) BB_CUDA_OK(cudaFree(s.chunk[i]));
      s.chunk[i] = nullptr;
      BB_CUDA_OK(cudaMalloc(&s.chunk[i], need_chunk));
    }
    s.chunk_bytes = need_chunk;
  }
  const int64_t need_stats = static_cast<int64_t>(d) * d + d;
  if (s.stats_doubles < need_stats) {
    if (s.stats) BB_CUDA_OK(cudaFree(s.stats));
    s.stats = nullptr;
    BB_CUDA_OK(cudaMalloc(&s.stats, need_stats * sizeof(double)));
    s.stats_doubles = need_stats;
  }
  const int64_t need_ws = bb_suffstats_gaussian_workspace(chunk_rows, d);
  if (s.ws_bytes < need_ws) {
    if (s.ws) BB_CUDA_OK(cudaFree(s.ws));
    s.ws = nullptr;
    BB_CUDA_OK(cudaMalloc(&s.ws, need_ws));
    s.ws_bytes = need_ws;
  }
  double* s2 = s.stats;
  double* s1 = s.stats + static_cast<int64_t>(d) * d;
  BB_CUDA_OK(cudaMemsetAsync(s.stats, 0, need_stats * sizeof(double), st));
  int64_t done = 0;
  for (int64_t c = 0; done < n; ++c) {
    const int b = static_cast<int>(c & 1);
    const int64_t rows = std::min(chunk_rows, n - done);
    if (c >= 2) BB_CUDA_OK(cudaStreamWaitEvent(s.copy_stream, s.consumed[b], 0));
    BB_CUDA_OK(cudaMemcpyAsync(s.chunk[b], X_host + done * d, rows * d * sizeof(float),
                               cudaMemcpyHostToDevice, s.copy_stream));
    BB_CUDA_OK(cudaEventRecord(s.copied[b], s.copy_stream));
    BB_CUDA_OK(cudaStreamWaitEvent(st, s.copied[b], 0));
    BB_TRY(suffstats_device(s.chunk[b], rows, d, s1, s2, s.ws, s.ws_bytes, true, st));
    BB_CUDA_OK(cudaEventRecord(s.consumed[b], st));
    done += rows;
  }
  BB_CUDA_OK(cudaMemcpyAsync(sum_xxT_host, s2, sizeof(double) * d * d, cudaMemcpyDeviceToHost, st));
  if (sum_x_host)
    BB_CUDA_OK(cudaMemcpyAsync(sum_x_host, s1, sizeof(double) * d, cudaMemcpyDeviceToHost, st));
  BB_CUDA_OK(cudaStreamSynchronize(st));
  return BB_OK;
}

BB_API int bb_release_staging(void) {
  Staging& s = g_staging;
  if (s.device < 0) return BB_OK;
  for (int i = 0; i < 2; ++i) {
    if (s.chunk[i]) cudaFree(s.chunk[i]);
    if (s.copied[i]) cudaEventDestroy(s.copied[i]);
    if (s.consumed[i]) cudaEventDestroy(s.consumed[i]);
  }
  if (s.stats) cudaFree(s.stats);
  if (s.ws) cudaFree(s.ws);
  if (s.copy_stream) cudaStreamDestroy(s.copy_stream);
  s = Staging();
  return BB_OK;
}

BB_API int64_t bb_suffstats_regression_workspace(int64_t n, int32_t d) {
  // generic path: float32 results (d*d + d + 1) + the largest split-K scratch of the three GEMMs
  int64_t need = align_up((static_cast<int64_t>(d) * d + d + 1) * 4, 256) +
                 std::max(gemm_workspace_bytes(d, d, n, 1),
                          std::max(gemm_workspace_bytes(d, 1, n, 1), gemm_workspace_bytes(1, 1, n, 1))) + 1024;
  if (d > 64 && d % 4 == 0 && d <= 4096 && n > 0) need = std::max(need, gram_tc_workspace(n, d) + 256);
  return need;
}

BB_API int bb_suffstats_regression(const float* X, const float* y, int64_t n, int32_t d, double* xtx,
                            double* xty, double* yty, void* workspace, int64_t workspace_bytes,
                            void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if ((X == nullptr && n > 0) || xtx == nullptr || n < 0 || d < 1 || (xty == nullptr) != (yty == nullptr) ||
      (n > 0 && (y == nullptr) != (xty == nullptr))) {
    set_error("suffstats_regression: bad arguments (n=%lld d=%d; y, xty, yty go together)",
              static_cast<long long>(n), d);
    return BB_ERR_INVALID;
  }
  if (workspace == nullptr || workspace_bytes < bb_suffstats_regression_workspace(n, d)) {
    set_error("suffstats_regression: workspace %lld < %lld bytes", static_cast<long long>(workspace_bytes),
              static_cast<long long>(bb_suffstats_regression_workspace(n, d)));
    return BB_ERR_WORKSPACE;
  }
  if (n == 0) {
    BB_CUDA_OK(cudaMemsetAsync(xtx, 0, sizeof(double) * d * d, st));
    if (xty) BB_CUDA_OK(cudaMemsetAsync(xty, 0, sizeof(double) * d, st));
    if (yty) BB_CUDA_OK(cudaMemsetAsync(yty, 0, sizeof(double), st));
    return BB_OK;
  }
  if (gram_tc_supported(n, d, X) && (y == nullptr || reinterpret_cast<uintptr_t>(y) % 4 == 0))
    return launch_gram_tc(X, y, n, d, xtx, xty, yty, workspace, workspace_bytes, st);
  // generic path: three split-K contractions over the data axis
  char* ws = static_cast<char*>(workspace);
  float* out32 = reinterpret_cast<float*>(ws);
  ws += align_up((static_cast<int64_t>(d) * d + d + 1) * 4, 256);
  BB_TRY(launch_gemm(X, X, out32, d, d, n, 1, 0, 1, d, 0, d, 1, ws, st));
  BB_TRY(launch_f32_to_f64(out32, xtx, static_cast<int64_t>(d) * d, st));
  if (y != nullptr) {
    float* xty32 = out32 + static_cast<int64_t>(d) * d;
    BB_TRY(launch_gemm(X, y, xty32, d, 1, n, 1, 0, 1, d, 0, 1, 1, ws, st));
    BB_TRY(launch_f32_to_f64(xty32, xty, d, st));
    BB_TRY(launch_gemm(y, y, xty32 + d, 1, 1, n, 1, 0, 1, 1, 0, 1, 1, ws, st));
    BB_TRY(launch_f32_to_f64(xty32 + d, yty, 1, st));
  }
  return BB_OK;
}

BB_API int64_t bb_rowproj_workspace(int64_t n, int32_t d, int32_t q) { return rowproj_tc_workspace(n, d, q) + 256; }

BB_API int bb_rowproj(const float* X, const float* W, int64_t n, int32_t d, int32_t q, float* Z, void* workspace,
               int64_t workspace_bytes, void* stream) {
  if (n < 0 || !W || !Z || (n > 0 && !X)) { set_error("rowproj: bad arguments"); return BB_ERR_INVALID; }
  if (n == 0) return BB_OK;
  return launch_rowproj_tc(X, W, nullptr, n, d, q, Z, nullptr, workspace, workspace_bytes,
                           static_cast<cudaStream_t>(stream));
}

BB_API int64_t bb_colproj_workspace(int64_t n, int32_t d, int32_t q) { return colproj_tc_workspace(n, d, q) + 256; }

BB_API int bb_colproj(const float* X, const float* R, int64_t n, int32_t d, int32_t q, double* G, void* workspace,
               int64_t workspace_bytes, void* stream) {
  if (n < 0 || !G || (n > 0 && (!X || !R))) { set_error("colproj: bad arguments"); return BB_ERR_INVALID; }
  if (n == 0) {
    BB_CUDA_OK(cudaMemsetAsync(G, 0, sizeof(double) * d * q, static_cast<cudaStream_t>(stream)));
    return BB_OK;
  }
  return launch_colproj_tc(X, R, n, d, q, G, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

BB_API int64_t bb_logistic_reparam_workspace(int64_t n, int32_t d, int32_t s) {
  int64_t need = align_up(n * static_cast<int64_t>(s) * 4, 256) +
                 std::max(rowproj_tc_workspace(n, d, s), colproj_tc_workspace(n, d, s)) + 512;
  if (s == 64 && d >= 128 && d % 128 == 0 && d <= 512 && n > 0)
    need = std::max(need, logistic_fused2_workspace(n, d, s) + 512);
  return need;
}

BB_API int bb_logistic_reparam_pass(const float* X, const float* y, const float* W, int64_t n, int32_t d, int32_t s,
                             double* loglik, double* G, void* workspace, int64_t workspace_bytes,
                             void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n < 0 || !W || !loglik || !G || (n > 0 && (!X || !y))) { set_error("logistic_reparam_pass: bad arguments"); return BB_ERR_INVALID; }
  if (n == 0) {
    BB_CUDA_OK(cudaMemsetAsync(loglik, 0, sizeof(double) * s, st));
    BB_CUDA_OK(cudaMemsetAsync(G, 0, sizeof(double) * d * s, st));
    return BB_OK;
  }
  if (!rowproj_tc_supported(n, d, s, X) || !colproj_tc_supported(n, d, s, X, X)) {
    set_error("logistic_reparam_pass: unsupported shape n=%lld d=%d s=%d (needs d %% 128 == 0, s %% 64 == 0, "
              "s d <= 32768, (d/128)(s/64) <= 4); use the compiled plan instead",
              static_cast<long long>(n), d, s);
    return BB_ERR_UNSUPPORTED;
  }
  if (workspace == nullptr || workspace_bytes < bb_logistic_reparam_workspace(n, d, s)) {
    set_error("logistic_reparam_pass: workspace %lld < %lld bytes", static_cast<long long>(workspace_bytes),
              static_cast<long long>(bb_logistic_reparam_workspace(n, d, s)));
    return BB_ERR_WORKSPACE;
  }
  // single-kernel path (X read from HBM once); BB_LOGISTIC_UNFUSED=1 keeps the two-kernel path
  static const bool unfused = getenv("BB_LOGISTIC_UNFUSED") != nullptr;
  // (the resident-tile design, logistic_fused2_sm100.cu; the first, W-resident design served the same shapes
  // 1.2x slower and was removed in round 2)
  if (!unfused && logistic_fused2_supported(n, d, s, X))
    return launch_logistic_fused2(X, y, W, n, d, s, loglik, G, workspace, workspace_bytes, st);
  char* ws = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~static_cast<uintptr_t>(255));
  float* resid = reinterpret_cast<float*>(ws);
  ws += align_up(n * static_cast<int64_t>(s) * 4, 256);
  const int64_t rest = workspace_bytes - (ws - static_cast<char*>(workspace));
  BB_TRY(launch_rowproj_tc(X, W, y, n, d, s, resid, loglik, ws, rest, st));
  BB_TRY(launch_colproj_tc(X, resid, n, d, s, G, ws, rest, st));
  return BB_OK;
}

BB_API int bb_gaussian_expected_loglik(const double* sum_x, const double* sum_xxT, double n,
                                const double* E_Lambda, const double* E_Lambda_mu, double E_muLmu,
                                double E_logdet, int32_t d, double* out, void* stream) {
  if (!sum_x || !sum_xxT || !E_Lambda || !E_Lambda_mu || !out || d < 1) {
    set_error("gaussian_expected_loglik: null argument");
    return BB_ERR_INVALID;
  }
  return launch_gaussian_expected_loglik(sum_x, sum_xxT, n, E_Lambda, E_Lambda_mu, E_muLmu, E_logdet,
                                         d, out, static_cast<cudaStream_t>(stream));
}

BB_API int bb_logsoftmax_rows(const float* logits, int64_t n, int32_t k, float* log_resp, float* lse,
                       double* sum_lse, void* stream) {
  if (n < 0 || (n > 0 && (!logits || !log_resp))) { set_error("logsoftmax_rows: bad arguments"); return BB_ERR_INVALID; }
  return launch_logsoftmax_rows(logits, n, k, log_resp, lse, sum_lse, false, static_cast<cudaStream_t>(stream));
}

BB_API int bb_softmax_rows(const float* logits, int64_t n, int32_t k, float* resp, float* lse, double* sum_lse,
                    void* stream) {
  if (n < 0 || (n > 0 && (!logits || !resp))) { set_error("softmax_rows: bad arguments"); return BB_ERR_INVALID; }
  return launch_logsoftmax_rows(logits, n, k, resp, lse, sum_lse, true, static_cast<cudaStream_t>(stream));
}

BB_API int64_t bb_mixture_logits_workspace(int64_t n, int32_t d, int32_t k) {
  return mixture_logits_workspace(n, d, k) + 256;
}

BB_API int bb_mixture_logits(const float* X, const float* U, const float* t, const float* c, int64_t n, int32_t d,
                      int32_t k, int32_t upper_triangular, float* logits, float* lse, double* sum_lse,
                      void* workspace, int64_t workspace_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n < 0 || !U || !t || !c || (n > 0 && (!X || !logits))) { set_error("mixture_logits: bad arguments"); return BB_ERR_INVALID; }
  if (n == 0) {
    if (sum_lse) BB_CUDA_OK(cudaMemsetAsync(sum_lse, 0, sizeof(double), st));
    return BB_OK;
  }
  return launch_mixture_logits(X, U, t, c, n, d, k, upper_triangular, logits, lse, sum_lse, workspace, workspace_bytes, st);
}

BB_API int64_t bb_suffstats_weighted_workspace(int64_t n, int32_t d, int32_t k) {
  return weighted_stats_auto_workspace(n, d, k) + 256;
}

BB_API int bb_suffstats_weighted(const float* X, const float* R, int64_t n, int32_t d, int32_t k, double* Nk,
                          double* sum_rx, double* sum_rxx, void* workspace, int64_t workspace_bytes,
                          void* stream) {
  if (n < 0 || !sum_rxx || (n > 0 && (!X || !R))) { set_error("suffstats_weighted: bad arguments"); return BB_ERR_INVALID; }
  return launch_weighted_stats_auto(X, R, n, d, k, Nk, sum_rx, sum_rxx, workspace, workspace_bytes,
                                    static_cast<cudaStream_t>(stream));
}

BB_API int64_t bb_softmax_rows_split_bytes(int64_t n, int32_t k) {
  if (n < 0 || k < 256 || k % 256 != 0) return 0;
  return weighted_pairs_split_bytes(n, k);
}

BB_API int bb_softmax_rows_split(const float* logits, int64_t n, int32_t k, void* rsplit, float* lse, double* sum_lse,
                          void* stream) {
  if (n < 0 || (n > 0 && (!logits || !rsplit))) { set_error("softmax_rows_split: bad arguments"); return BB_ERR_INVALID; }
  return launch_softmax_rows_split(logits, n, k, rsplit, lse, sum_lse, static_cast<cudaStream_t>(stream));
}

BB_API int bb_suffstats_weighted_split(const float* X, const void* rsplit, int64_t n, int32_t d, int32_t k, double* Nk,
                                double* sum_rx, double* sum_rxx, void* workspace, int64_t workspace_bytes,
                                void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n < 0 || !sum_rxx || (n > 0 && (!X || !rsplit))) { set_error("suffstats_weighted_split: bad arguments"); return BB_ERR_INVALID; }
  if (n == 0) {
    BB_CUDA_OK(cudaMemsetAsync(sum_rxx, 0, sizeof(double) * k * d * d, st));
    if (sum_rx) BB_CUDA_OK(cudaMemsetAsync(sum_rx, 0, sizeof(double) * k * d, st));
    if (Nk) BB_CUDA_OK(cudaMemsetAsync(Nk, 0, sizeof(double) * k, st));
    return BB_OK;
  }
  return launch_weighted_pairs_split(X, rsplit, n, d, k, Nk, sum_rx, sum_rxx, workspace, workspace_bytes, st);
}

BB_API int bb_suffstats_weighted_from_logits(const float* X, const float* logits, const float* lse, int64_t n,
                                      int32_t d, int32_t k, double* Nk, double* sum_rx, double* sum_rxx,
                                      void* workspace, int64_t workspace_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n < 0 || !sum_rxx || (n > 0 && (!X || !logits || !lse))) { set_error("suffstats_weighted_from_logits: bad arguments"); return BB_ERR_INVALID; }
  if (n == 0) {
    BB_CUDA_OK(cudaMemsetAsync(sum_rxx, 0, sizeof(double) * k * d * d, st));
    if (sum_rx) BB_CUDA_OK(cudaMemsetAsync(sum_rx, 0, sizeof(double) * k * d, st));
    if (Nk) BB_CUDA_OK(cudaMemsetAsync(Nk, 0, sizeof(double) * k, st));
    return BB_OK;
  }
  return launch_weighted_pairs(X, logits, lse, n, d, k, Nk, sum_rx, sum_rxx, workspace, workspace_bytes, st);
}

BB_API int bb_gmm_global_update(const double* Nk, const double* sum_rx, const double* sum_rxx, int32_t k, int32_t d,
                         double alpha0, double beta0, double nu0, const double* m0, const double* W0_inv,
                         double* alpha, double* beta, double* nu, double* m, double* W_inv, float* U, float* t,
                         float* c, double* kl, int32_t* status, void* stream) {
  if (!Nk || !sum_rx || !sum_rxx || !m0 || !W0_inv || !alpha || !beta || !nu || !m || !W_inv || !U || !t || !c ||
      !kl || !status) {
    set_error("gmm_global_update: bad arguments");
    return BB_ERR_INVALID;
  }
  return launch_gmm_global_update(Nk, sum_rx, sum_rxx, k, d, alpha0, beta0, nu0, m0, W0_inv, alpha, beta, nu, m,
                                  W_inv, U, t, c, kl, status, static_cast<cudaStream_t>(stream));
}

// ---- multi-GPU: peer-memory communicator and the one-launch Gaussian pass ---------------------------

struct bb_comm {
  int rank = 0, world = 1, device = 0;
  int64_t capacity = 0;
  const double** in = nullptr;       // device arrays of `world` pointers
  double** out = nullptr;
  uint32_t** flags = nullptr;
  int* status = nullptr;             // device: [0] status word, [1] epochs completed, [2] CTA ticket
  double spin_limit_ms = 2000.0;
};

static int upload_ptrs(const void* const* host, int world, void** dev_out) {
  void* dev = nullptr;
  BB_CUDA_OK(cudaMalloc(&dev, sizeof(void*) * world));
  cudaError_t e = cudaMemcpy(dev, host, sizeof(void*) * world, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    cudaFree(dev);
    set_error("cudaMemcpy of the peer pointer table failed: %s", cudaGetErrorString(e));
    return BB_ERR_CUDA;
  }
  *dev_out = dev;
  return BB_OK;
}

BB_API int64_t bb_comm_flag_bytes(int32_t world) {
  return static_cast<int64_t>(2) * world * BB_COMM_MAX_CTAS * sizeof(uint32_t);
}

BB_API int bb_comm_create(int32_t rank, int32_t world, const void* const* peer_in, const void* const* peer_out,
                   const void* const* peer_flags, int64_t capacity, double spin_limit_ms, bb_comm** comm) {
  if (comm == nullptr) { set_error("comm_create: null out-pointer"); return BB_ERR_INVALID; }
  *comm = nullptr;
  if (world < 1 || rank < 0 || rank >= world || capacity < 1 || !peer_in || !peer_out || !peer_flags) {
    set_error("comm_create: bad arguments (rank %d world %d capacity %lld)", rank, world,
              static_cast<long long>(capacity));
    return BB_ERR_INVALID;
  }
  for (int r = 0; r < world; ++r)
    if (!peer_in[r] || !peer_out[r] || !peer_flags[r] || reinterpret_cast<uintptr_t>(peer_in[r]) % 16 ||
        reinterpret_cast<uintptr_t>(peer_out[r]) % 16) {
      set_error("comm_create: peer buffer %d is null or not 16-byte aligned", r);
      return BB_ERR_INVALID;
    }
  bb_comm* c = new (std::nothrow) bb_comm();
  if (c == nullptr) { set_error("out of host memory"); return BB_ERR_INVALID; }
  c->rank = rank; c->world = world; c->capacity = capacity;
  if (spin_limit_ms > 0) c->spin_limit_ms = spin_limit_ms;
  int st = BB_OK;
  if (cudaGetDevice(&c->device) != cudaSuccess) st = BB_ERR_CUDA;
  if (st == BB_OK) st = upload_ptrs(peer_in, world, reinterpret_cast<void**>(&c->in));
  if (st == BB_OK) st = upload_ptrs(peer_out, world, reinterpret_cast<void**>(&c->out));
  if (st == BB_OK) st = upload_ptrs(peer_flags, world, reinterpret_cast<void**>(&c->flags));
  if (st == BB_OK && cudaMalloc(&c->status, 4 * sizeof(int)) != cudaSuccess) st = BB_ERR_CUDA;
  if (st == BB_OK && cudaMemset(c->status, 0, 4 * sizeof(int)) != cudaSuccess) st = BB_ERR_CUDA;
  if (st != BB_OK) {
    if (st == BB_ERR_CUDA) set_error("comm_create: CUDA allocation failed");
    bb_comm_destroy(c);
    return st;
  }
  *comm = c;
  return BB_OK;
}

BB_API int bb_comm_destroy(bb_comm* c) {
  if (c == nullptr) return BB_OK;
  if (c->in) cudaFree(c->in);
  if (c->out) cudaFree(c->out);
  if (c->flags) cudaFree(c->flags);
  if (c->status) cudaFree(c->status);
  delete c;
  return BB_OK;
}

BB_API int bb_comm_allreduce_sum(bb_comm* c, int64_t count, void* stream) {
  if (c == nullptr || count < 0 || count > c->capacity) {
    set_error("comm_allreduce_sum: bad arguments (count %lld, capacity %lld)", static_cast<long long>(count),
              c ? static_cast<long long>(c->capacity) : 0LL);
    return BB_ERR_INVALID;
  }
  return launch_p2p_allreduce(c->in, c->out, c->flags, c->rank, c->world, count,
                              reinterpret_cast<uint32_t*>(c->status + 1), reinterpret_cast<unsigned int*>(c->status + 2),
                              c->spin_limit_ms, c->status, static_cast<cudaStream_t>(stream));
}

BB_API int bb_comm_status(bb_comm* c, int32_t* status, void* stream) {
  if (c == nullptr || status == nullptr) { set_error("comm_status: null argument"); return BB_ERR_INVALID; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int host = 0;
  BB_CUDA_OK(cudaMemcpyAsync(&host, c->status, sizeof(int), cudaMemcpyDeviceToHost, st));
  BB_CUDA_OK(cudaStreamSynchronize(st));
  *status = host;
  return BB_OK;
}

struct bb_gaussian_pass {
  int d = 0, device = 0;
  void* ws = nullptr;
  int64_t ws_bytes = 0;
  int* status = nullptr;           // device: [0] status word, [1] epochs completed, [2] completion ticket,
                                   // [4], [5] tile counters of even / odd launches
  uint64_t launches = 0;
  int rank = 0, world = 1;
  double** peer_recv = nullptr;
  uint32_t** peer_flags = nullptr;
  int64_t stride = 0;
  double spin_limit_ms = 2000.0;
};

static int64_t pass_stride(int d) { return align_up(static_cast<int64_t>(d) * d + d + 1, 32); }

BB_API int bb_gaussian_pass_peer_bytes(int32_t d, int32_t world, int64_t* recv_bytes, int64_t* flag_bytes) {
  if (d < 1 || world < 1 || !recv_bytes || !flag_bytes) { set_error("gaussian_pass_peer_bytes: bad arguments"); return BB_ERR_INVALID; }
  *recv_bytes = 2 * static_cast<int64_t>(world) * pass_stride(d) * sizeof(double);
  *flag_bytes = static_cast<int64_t>(world) * BB_GAUSSIAN_PASS_SLICES * sizeof(uint32_t);
  return BB_OK;
}

BB_API int bb_gaussian_pass_create(int32_t d, bb_gaussian_pass** pass) {
  if (pass == nullptr) { set_error("gaussian_pass_create: null out-pointer"); return BB_ERR_INVALID; }
  *pass = nullptr;
  if (!(d >= 4 && d <= 64 && d % 4 == 0)) {
    set_error("gaussian_pass_create: needs d <= 64 and d %% 4 == 0 (got %d)", d);
    return BB_ERR_UNSUPPORTED;
  }
  bb_gaussian_pass* p = new (std::nothrow) bb_gaussian_pass();
  if (p == nullptr) { set_error("out of host memory"); return BB_ERR_INVALID; }
  p->d = d;
  p->ws_bytes = suffstats_tc_workspace(int64_t(1) << 30) + 256;
  if (cudaGetDevice(&p->device) != cudaSuccess || cudaMalloc(&p->ws, p->ws_bytes) != cudaSuccess ||
      cudaMemset(p->ws, 0, p->ws_bytes) != cudaSuccess ||
      cudaMalloc(&p->status, 8 * sizeof(int)) != cudaSuccess || cudaMemset(p->status, 0, 8 * sizeof(int)) != cudaSuccess) {
    set_error("gaussian_pass_create: CUDA allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
    bb_gaussian_pass_destroy(p);
    return BB_ERR_CUDA;
  }
  *pass = p;
  return BB_OK;
}

BB_API int bb_gaussian_pass_attach_peers(bb_gaussian_pass* p, int32_t rank, int32_t world, const void* const* peer_recv,
                                  const void* const* peer_flags, double spin_limit_ms) {
  if (p == nullptr || world < 1 || rank < 0 || rank >= world || !peer_recv || !peer_flags) {
    set_error("gaussian_pass_attach_peers: bad arguments");
    return BB_ERR_INVALID;
  }
  for (int r = 0; r < world; ++r)
    if (!peer_recv[r] || !peer_flags[r]) { set_error("gaussian_pass_attach_peers: peer %d has a null buffer", r); return BB_ERR_INVALID; }
  if (p->peer_recv) cudaFree(p->peer_recv);
  if (p->peer_flags) cudaFree(p->peer_flags);
  p->peer_recv = nullptr; p->peer_flags = nullptr;
  BB_TRY(upload_ptrs(peer_recv, world, reinterpret_cast<void**>(&p->peer_recv)));
  BB_TRY(upload_ptrs(peer_flags, world, reinterpret_cast<void**>(&p->peer_flags)));
  p->rank = rank; p->world = world; p->stride = pass_stride(p->d);
  BB_CUDA_OK(cudaMemset(p->status, 0, 2 * sizeof(int)));
  if (spin_limit_ms > 0) p->spin_limit_ms = spin_limit_ms;
  return BB_OK;
}

BB_API int bb_gaussian_pass_run(bb_gaussian_pass* p, const float* X, int64_t n, const double* E_Lambda,
                         const double* E_Lambda_mu, double E_mu_L_mu, double E_logdet, double n_total,
                         double* sum_x, double* sum_xxT, double* count_out, double* loglik, void* stream) {
  if (p == nullptr || sum_xxT == nullptr || n < 0 || (n > 0 && X == nullptr)) {
    set_error("gaussian_pass_run: bad arguments");
    return BB_ERR_INVALID;
  }
  SuffstatsTail t;
  memset(&t, 0, sizeof(t));
  t.s2 = sum_xxT; t.s1 = sum_x; t.count_out = count_out;
  t.e_lambda = E_Lambda; t.e_lambda_mu = E_Lambda_mu; t.e_mu_l_mu = E_mu_L_mu; t.e_logdet = E_logdet;
  t.n_total = n_total; t.loglik = loglik;
  t.local_count = static_cast<double>(n);
  t.rank = p->rank; t.world = p->world;
  t.status = p->status;
  t.accum = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(p->ws) + 255) & ~static_cast<uintptr_t>(255));
  t.ticket = reinterpret_cast<unsigned int*>(p->status + 2);
  // consecutive launches may overlap (the next one streams while this one's last CTA finishes), so they
  // alternate between two tile counters; each launch's last CTA re-zeroes its own
  t.tile_counter = reinterpret_cast<unsigned int*>(p->status + 4 + (p->launches & 1));
  t.pdl = 1;
  ++p->launches;
  if (p->world > 1) {
    t.peer_recv = p->peer_recv; t.peer_flags = p->peer_flags; t.stride = p->stride;
    t.epoch_dev = reinterpret_cast<uint32_t*>(p->status + 1);
    t.spin_limit = static_cast<long long>(p->spin_limit_ms * 2.0e6);
  }
  return launch_suffstats_tc_fused(X, n, p->d, p->ws, p->ws_bytes, t, static_cast<cudaStream_t>(stream));
}

BB_API int bb_gaussian_pass_status(bb_gaussian_pass* p, int32_t* status, void* stream) {
  if (p == nullptr || status == nullptr) { set_error("gaussian_pass_status: null argument"); return BB_ERR_INVALID; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int host = 0;
  BB_CUDA_OK(cudaMemcpyAsync(&host, p->status, sizeof(int), cudaMemcpyDeviceToHost, st));
  BB_CUDA_OK(cudaStreamSynchronize(st));
  *status = host;
  return BB_OK;
}

BB_API int bb_gaussian_pass_destroy(bb_gaussian_pass* p) {
  if (p == nullptr) return BB_OK;
  if (p->ws) cudaFree(p->ws);
  if (p->status) cudaFree(p->status);
  if (p->peer_recv) cudaFree(p->peer_recv);
  if (p->peer_flags) cudaFree(p->peer_flags);
  delete p;
  return BB_OK;
}

BB_API int bb_gather_rows(const float* X, int64_t n, int32_t d, const int64_t* index, int64_t m, float* out,
                   int32_t* n_out_of_range, void* stream) {
  if (n < 0 || d < 0 || m < 0 || !n_out_of_range || (m > 0 && d > 0 && (!X || !index || !out))) {
    set_error("gather_rows: bad arguments");
    return BB_ERR_INVALID;
  }
  return launch_gather_rows(X, n, d, index, m, out, n_out_of_range, static_cast<cudaStream_t>(stream));
}

BB_API int bb_svi_natural_blend(double* eta, const double* eta_prior, const double* stat, double scale, double rho,
                         int64_t count, void* stream) {
  if (count < 0 || (count > 0 && (!eta || !eta_prior || !stat)) || !(rho >= 0.0 && rho <= 1.0)) {
    set_error("svi_natural_blend: bad arguments (need 0 <= rho <= 1)");
    return BB_ERR_INVALID;
  }
  return launch_svi_blend(eta, eta_prior, stat, scale, rho, count, static_cast<cudaStream_t>(stream));
}

BB_API int bb_reparam_draws(const double* mu, const double* log_sigma, const double* eps, int32_t d, int32_t s,
                     float* W, void* stream) {
  if (d < 0 || s < 0 || (d > 0 && s > 0 && (!mu || !log_sigma || !eps || !W))) {
    set_error("reparam_draws: bad arguments");
    return BB_ERR_INVALID;
  }
  return launch_reparam_draws(mu, log_sigma, eps, d, s, W, static_cast<cudaStream_t>(stream));
}

BB_API int bb_reparam_gradient(const double* G, const double* loglik, const double* eps, const double* mu,
                        const double* log_sigma, int32_t d, int32_t s, double* grad_mu, double* grad_log_sigma,
                        double* elbo, void* stream) {
  if (d < 1 || s < 1 || !G || !loglik || !eps || !mu || !log_sigma || !grad_mu || !grad_log_sigma || !elbo) {
    set_error("reparam_gradient: bad arguments");
    return BB_ERR_INVALID;
  }
  return launch_reparam_gradient(G, loglik, eps, mu, log_sigma, d, s, grad_mu, grad_log_sigma, elbo,
                                 static_cast<cudaStream_t>(stream));
}

BB_API int bb_adam_step(double* param, const double* grad, double* m, double* v, int64_t count, double lr,
                 double beta1, double beta2, double eps, int64_t step, int32_t maximize, void* stream) {
  if (count < 0 || step < 1 || (count > 0 && (!param || !grad || !m || !v))) {
    set_error("adam_step: bad arguments (step counts from 1)");
    return BB_ERR_INVALID;
  }
  return launch_adam_step(param, grad, m, v, count, lr, beta1, beta2, eps, step, maximize,
                          static_cast<cudaStream_t>(stream));
}

}  // extern "C"
