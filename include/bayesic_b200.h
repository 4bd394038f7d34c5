/*
 * bayesic_b200 -- C-ABI of the B200 (sm_100a) executor for Bayesic's einsum plans.
 *
 * This is the drop-in boundary for the one hot path of mjwillson/Bayesic: the
 * per-minibatch evaluation of a compiled expression.  The reference has no FFI;
 * its seam is the Python protocol
 *
 *     fn = theano.function(inputs, expression)        bayesic/algebra.py:54
 *     f(**inputs) -> fn(*arrays)                      bayesic/algebra.py:55-56
 *     node._apply_to_parents(*parent_vars)            bayesic/algebra.py:34-40,
 *                                                     1290, 1302, 1318, 1347, 1405
 *
 * Every entry point below replaces one of those call sites (cited per function).
 * Conventions:
 *   - plain C types only; all tensors are raw pointers + int64 extents;
 *   - device tensors are dense row-major float32 unless stated; statistics that
 *     are sums over the data axis are returned as float64;
 *   - the library BORROWS every pointer for the duration of the call; device
 *     entry points never allocate: scratch comes from a caller-provided workspace
 *     whose size is queried first.  The one exception is the host-buffer entry
 *     point bb_suffstats_gaussian_host, which owns a small staging pool
 *     (released by bb_release_staging);
 *   - calls are ordered on the caller's `stream` (a cudaStream_t passed as
 *     void*); nothing synchronises unless documented;
 *   - every function returns a bb_status (0 = ok); on failure
 *     bb_last_error() gives a thread-local message.  No exceptions cross the ABI.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point
 *     fails with BB_ERR_CUDA.
 */
#ifndef BAYESIC_B200_H
#define BAYESIC_B200_H

#include <stdint.h>

#if defined(__GNUC__)
#define BB_API __attribute__((visibility("default")))
#else
#define BB_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define BB_ABI_VERSION 2
#define BB_MAX_DIMS 8
#define BB_MAX_PARENTS 8
#define BB_MAX_IPARAMS 40

typedef enum {
  BB_OK = 0,
  BB_ERR_INVALID = 1,      /* malformed descriptor / argument                      */
  BB_ERR_CUDA = 2,         /* CUDA runtime or driver error (message has details)   */
  BB_ERR_UNSUPPORTED = 3,  /* valid request this build cannot serve                */
  BB_ERR_SHAPE = 4,        /* runtime extents inconsistent with the plan           */
  BB_ERR_WORKSPACE = 5     /* workspace too small                                  */
} bb_status;

/* ---- library ------------------------------------------------------------ */

BB_API int bb_abi_version(void);
BB_API const char* bb_last_error(void);
/* Kernels launched by this library on the calling thread since load (monotonic). */
BB_API int64_t bb_launch_count(void);
/* SM count and compute capability of the current device. */
BB_API int bb_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor);

/* ---- plan executor -------------------------------------------------------
 * A plan is the flat, topologically ordered form of the plan-IR tree the
 * planner emits (bayesic/algebra.py:527-765).  Node kinds map one-to-one onto
 * the reference's executor vocabulary; iparams carry the integer attributes. */

typedef enum {
  BB_NODE_INPUT = 0,      /* var            algebra.py:108-126   iparams: [input slot]            */
  BB_NODE_SCALAR = 1,     /* constant       algebra.py:129-144   fparam: value (host scalar)      */
  BB_NODE_SHAPE = 2,      /* shape          algebra.py:147-161   parents: [x]  iparams: [axis]    */
  BB_NODE_EYE = 3,        /* eye            algebra.py:236-258   parents: [n]                     */
  BB_NODE_SUM = 4,        /* _sum           algebra.py:1284-1294 parents: [x]  iparams: axes      */
  BB_NODE_MUL = 5,        /* _mul           algebra.py:1297-1309 parents: factors (<= 8)          */
  BB_NODE_DIMSHUFFLE = 6, /* _dimshuffle    algebra.py:1312-1326 parents: [x]  iparams: axes, -1 = 'x' */
  BB_NODE_TENSORDOT = 7,  /* _tensordot     algebra.py:1329-1396 parents: [x, y]
                             iparams: [n_dot, n_batch, x_dot.., y_dot.., x_batch.., y_batch..]
                             result axes = batch + x_other + y_other (algebra.py:1161-1171)    */
  BB_NODE_DIAGONAL = 8,   /* _diagonal      algebra.py:1398-1414 parents: [x]  iparams: [a1, a2];
                             the diagonal axis is appended last                                  */
  BB_NODE_ELEMWISE = 9,   /* elemwise/add   algebra.py:195-233, 1435-1448
                             parents: args (<= 8)  iparams: [bb_elemwise_op]                     */
  /* fused nodes recognised by the host-side lowering */
  BB_NODE_LOGSOFTMAX = 20,/* x - log(sum(exp(x), axis=-1)) over the last axis, max-subtracted;
                             the pattern add(Lg, einsum(-1 * log(einsum(sum exp(Lg)))))          */
  BB_NODE_SYRK = 21,      /* _tensordot(_dimshuffle(X,1,0), X, [1],[0]): X^T X over the data axis */
  BB_NODE_WEIGHTED_SCATTER = 22,/* sum_n R[n,k] X[n,d] X[n,e] -> [k,d,e] without materialising
                             the K x D x N intermediate of algebra.py's plan; parents: [R, X]    */
  /* extension of the vocabulary (like BB_OP_LGAMMA) */
  BB_NODE_LOGDET = 23     /* log|X| over the last two axes of a stack of symmetric positive-definite
                             matrices X[..., d, d] -> [...]: the `T.logdet(precision)` that the
                             reference's MultivariateNormal log-normaliser calls
                             (distribution/core.py:49-52) and Theano never had; float64 Cholesky;
                             NaN for a matrix that is not positive definite.  parents: [x]        */
} bb_node_kind;

typedef enum {
  BB_OP_ADD = 0, BB_OP_MUL = 1, BB_OP_LOG = 2, BB_OP_EXP = 3, BB_OP_POW = 4, BB_OP_ABS = 5,
  BB_OP_LGAMMA = 6   /* extension: log Gamma, for exponential-family log-normalisers */
} bb_elemwise_op;

typedef struct {
  int32_t kind;                       /* bb_node_kind */
  int32_t n_parents;
  int32_t parents[BB_MAX_PARENTS];    /* indices of earlier nodes */
  int32_t n_iparams;
  int32_t iparams[BB_MAX_IPARAMS];
  double fparam;
} bb_node_desc;

/* A tensor argument: a dense row-major float32 device array, or a host scalar
 * (0-dim inputs, e.g. the int32 scalar `a` of eye(a), test_algebra.py:40). */
typedef struct {
  const void* data;                   /* device pointer (ignored for host scalars) */
  int32_t ndim;
  int32_t is_host_scalar;
  int64_t shape[BB_MAX_DIMS];
  double host_value;
} bb_tensor_arg;

typedef struct {
  int32_t ndim;
  int32_t is_host_scalar;             /* value known on the host (shape arithmetic, literals) */
  int64_t shape[BB_MAX_DIMS];
  double host_value;
} bb_result_info;

typedef struct bb_plan bb_plan;

/* Replaces theano.function(inputs, expression), algebra.py:54. */
BB_API int bb_plan_create(const bb_node_desc* nodes, int32_t n_nodes,
                   const int32_t* outputs, int32_t n_outputs,
                   int32_t n_inputs, bb_plan** plan);
BB_API int bb_plan_destroy(bb_plan* plan);

/* Shape pass: result extents and scratch size for the given input extents
 * (plans are shape-polymorphic, algebra.py:539-546). */
BB_API int bb_plan_infer(bb_plan* plan, const bb_tensor_arg* inputs, int32_t n_inputs,
                  bb_result_info* results, int64_t* workspace_bytes);

/* Replaces fn(*arrays), algebra.py:55-56.  out_ptrs[i] receives result i as a
 * dense float32 array (untouched for host-scalar results). */
BB_API int bb_plan_execute(bb_plan* plan, const bb_tensor_arg* inputs, int32_t n_inputs,
                    void* const* out_ptrs, void* workspace, int64_t workspace_bytes,
                    void* stream);

/* Number of kernels the last bb_plan_execute on this plan launched. */
BB_API int bb_plan_last_launch_count(bb_plan* plan, int32_t* count);

/* ---- fused sufficient-statistic passes --------------------------------------
 * The iid-summed exponential-family statistics of
 * ExpFamIndependentObservations.sufficient_statistics
 * (bayesic/distribution/base.py:328-332) for MultivariateNormal's (x, x x^T)
 * (bayesic/distribution/core.py:41-44), in ONE pass over X. */

/* sum_x[d] = sum_n X[n,d];  sum_xxT[d,e] = sum_n X[n,d] X[n,e]   (float64 out): the plans of
 * sum(X, 0) and dot(X.T, X) (algebra.py:1284-1294, :1151-1158), i.e. base.py:328-332 for core.py:41-44.
 * X is device float32 [n, d] row-major.  d <= 64 and d % 4 == 0 runs the
 * tcgen05 kernel; other d use the generic contraction kernels. */
BB_API int64_t bb_suffstats_gaussian_workspace(int64_t n, int32_t d);
BB_API int bb_suffstats_gaussian(const float* X, int64_t n, int32_t d,
                          double* sum_x, double* sum_xxT,
                          void* workspace, int64_t workspace_bytes, void* stream);

/* Same pass from a HOST buffer (the reference-facing call: numpy in, numpy out,
 * algebra.py:55-56).  Streams X through pinned staging in chunks so the copy
 * overlaps the kernel; synchronises before returning.  sum_x / sum_xxT are host
 * pointers. */
BB_API int bb_suffstats_gaussian_host(const float* X_host, int64_t n, int32_t d,
                               double* sum_x_host, double* sum_xxT_host,
                               int64_t chunk_rows, void* stream);

/* Frees the staging pool of bb_suffstats_gaussian_host (device chunk buffers, stream, events). */
BB_API int bb_release_staging(void);

/* Expected Gaussian log-likelihood of the batch under q(mu, Lambda), from the
 * statistics (log_likelihood = data_term + interaction_term - log_normalizer,
 * bayesic/distribution/base.py:25-100; eta = (Lambda mu, -1/2 Lambda),
 * core.py:46-47; A = -1/2 D log 2pi ... core.py:49-52):
 *   out = -n d/2 log(2 pi) + n/2 E_logdet - 1/2 tr(E_Lambda S2)
 *         + S1 . E_Lambda_mu - n/2 E_muLmu
 * All pointers are device float64. */
BB_API int bb_gaussian_expected_loglik(const double* sum_x, const double* sum_xxT, double n,
                                const double* E_Lambda, const double* E_Lambda_mu,
                                double E_muLmu, double E_logdet, int32_t d,
                                double* out, void* stream);

/* bb_suffstats_gaussian followed by bb_gaussian_expected_loglik in one call (the cfg2 step on one
 * GPU; statistics of bayesic/distribution/base.py:328-332, log-likelihood of base.py:25-100): on the
 * tcgen05 path statistics, cross-CTA reduction and log-likelihood are ONE launch (the last CTA to
 * finish evaluates the log-likelihood; see bb_gaussian_pass for N GPUs).  sum_x is required; n_total is the
 * row count the log-likelihood refers to (= n on one GPU). */
BB_API int bb_suffstats_gaussian_loglik(const float* X, int64_t n, int32_t d, double* sum_x,
                                 double* sum_xxT, double n_total, const double* E_Lambda,
                                 const double* E_Lambda_mu, double E_mu_L_mu, double E_logdet,
                                 double* out, void* workspace, int64_t workspace_bytes,
                                 void* stream);

/* Minibatch statistics of conjugate Bayesian linear regression / factor analysis
 * (natural-gradient SVI, README.md:69-80), i.e. the plans
 *   _tensordot(_dimshuffle(X,1,0), X, [1],[0]), _tensordot(_dimshuffle(X,1,0), y, [1],[0]),
 *   _tensordot(y, y, [0],[0])                                  (algebra.py:527-551, 1347-1351)
 * in ONE pass over X:  xtx[d,e] = sum_n X[n,d] X[n,e];  xty[d] = sum_n X[n,d] y[n];
 * yty = sum_n y[n]^2   (float64 out, device).  y, xty, yty may all be NULL (Gram matrix only).
 * d % 4 == 0, 64 < d <= 4096 runs the tcgen05 CTA-pair kernel (error-compensated BF16; the feature axis is
 * zero-padded to a multiple of 256 on the fly); other d use the
 * generic contraction kernels. */
BB_API int64_t bb_suffstats_regression_workspace(int64_t n, int32_t d);
BB_API int bb_suffstats_regression(const float* X, const float* y, int64_t n, int32_t d,
                            double* xtx, double* xty, double* yty,
                            void* workspace, int64_t workspace_bytes, void* stream);

/* ---- tensor-core contractions of the plan executor ---------------------------
 * Z[n,q] = sum_d X[n,d] W[q,d]: the plan _tensordot(X, _dimshuffle(W,1,0), [1],[0])
 * ("dot(X, W.T)", algebra.py:1151-1158 -> 1347-1351) on tcgen05, W resident in shared memory.
 * Needs d % 64 == 0, q % 16 == 0, q <= 256, q * d <= 32768 (BB_ERR_UNSUPPORTED otherwise). */
BB_API int64_t bb_rowproj_workspace(int64_t n, int32_t d, int32_t q);
BB_API int bb_rowproj(const float* X, const float* W, int64_t n, int32_t d, int32_t q, float* Z,
               void* workspace, int64_t workspace_bytes, void* stream);

/* G[d,q] = sum_n X[n,d] R[n,q] (float64 out): the plan _tensordot(_dimshuffle(X,1,0), R, [1],[0])
 * ("dot(X.T, R)", algebra.py:1151-1158 -> 1347-1351) on tcgen05.  Needs d % 128 == 0, q % 64 == 0, (d/128)(q/64) <= 4. */
BB_API int64_t bb_colproj_workspace(int64_t n, int32_t d, int32_t q);
BB_API int bb_colproj(const float* X, const float* R, int64_t n, int32_t d, int32_t q, double* G,
               void* workspace, int64_t workspace_bytes, void* stream);

/* Reparameterised-gradient pass of Bayesian logistic regression (README.md:47-51) for S parameter
 * draws W[s,d] at once:  Z = X W^T;  loglik[s] = sum_n (y_n z_ns - log(1 + exp z_ns));
 * G[d,s] = sum_n X[n,d] (y_n - (1 + exp(-z_ns))^-1)     (float64 out, device).
 * These are the plans of  sum(ycol * Z - log(1 + exp(Z)), axis=0)  and
 * dot(X.T, ycol - (1 + exp(-1 * Z)) ** -1)  (algebra.py:1435-1448 vocabulary) with the
 * elementwise chain fused into the TMEM epilogue of the first contraction; the workspace holds
 * the n x s residual.  Same shape limits as bb_rowproj (q = s) and bb_colproj. */
BB_API int64_t bb_logistic_reparam_workspace(int64_t n, int32_t d, int32_t s);
BB_API int bb_logistic_reparam_pass(const float* X, const float* y, const float* W, int64_t n, int32_t d,
                             int32_t s, double* loglik, double* G,
                             void* workspace, int64_t workspace_bytes, void* stream);

/* ---- mixture responsibilities ------------------------------------------------
 * log r[n,k] = logits[n,k] - logsumexp_k logits[n,:]  (max-subtracted; the
 * reference can only spell the unstabilised form, algebra.py:1435-1448).
 * lse[n] and sum_lse (float64, device) are optional (may be NULL). */
BB_API int bb_logsoftmax_rows(const float* logits, int64_t n, int32_t k,
                       float* log_resp, float* lse, double* sum_lse, void* stream);
/* The responsibilities themselves, r[n,k] = exp(logits[n,k] - logsumexp_k logits[n,:]) -- the value of
 * exp(Lg + (-1 * log(sum(exp(Lg), 1))).dimshuffle(0, 'x')) in the reference's vocabulary (algebra.py:1435-1448)
 * -- in the same single pass (resp may alias logits); lse / sum_lse as above. */
BB_API int bb_softmax_rows(const float* logits, int64_t n, int32_t k,
                    float* resp, float* lse, double* sum_lse, void* stream);

/* Gaussian-mixture expected log-densities in whitened form (VMP local step, README.md:30-37):
 *   logits[n,k] = c[k] - 1/2 || U[k] x_n - t[k] ||^2      (= c' + x.b_k - 1/2 x^T A_k x with
 *   A_k = U_k^T U_k, t_k = U_k m_k), optionally lse[n] = logsumexp_k logits[n,:] and
 *   sum_lse = sum_n lse[n] (float64).  U [k,d,d], t [k,d], c [k], logits [n,k], lse [n]: device
 *   float32.  This is the value of the user expression dot(X, bk.T) - 0.5 einsum(X, Ak, X) + ck
 *   (the reference's plan for it is a batched _tensordot whose evaluation is broken,
 *   algebra.py:1370-1373, :1380).  Needs d % 8 == 0, 8 <= d <= 64 (the feature axis is zero-padded to a
 *   multiple of 16 on chip), k % 4 == 0 (BB_ERR_UNSUPPORTED otherwise).
 *   upper_triangular != 0 promises U[k,j,i] == 0 for i < j (Cholesky factors): 10 of the 16 MMA
 *   K-steps at d = 64 are then skipped; with general factors pass 0. */
BB_API int64_t bb_mixture_logits_workspace(int64_t n, int32_t d, int32_t k);
BB_API int bb_mixture_logits(const float* X, const float* U, const float* t, const float* c,
                      int64_t n, int32_t d, int32_t k, int32_t upper_triangular,
                      float* logits, float* lse, double* sum_lse,
                      void* workspace, int64_t workspace_bytes, void* stream);

/* Responsibility-weighted statistics in one pass over (R, X) -- the plans of sum(R, 0), dot(R.T, X)
 * and einsum(R_nk X_nd X_ne -> kde), whose reference plan materialises K x D x N
 * (algebra.py:527-765):
 *   Nk[k] = sum_n R[n,k];  sum_rx[k,d] = sum_n R[n,k] X[n,d];
 *   sum_rxx[k,d,e] = sum_n R[n,k] X[n,d] X[n,e]           (float64 out, device) */
BB_API int64_t bb_suffstats_weighted_workspace(int64_t n, int32_t d, int32_t k);
BB_API int bb_suffstats_weighted(const float* X, const float* R, int64_t n, int32_t d, int32_t k,
                          double* Nk, double* sum_rx, double* sum_rxx,
                          void* workspace, int64_t workspace_bytes, void* stream);

/* Same statistics (same plans, algebra.py:527-765) with the responsibilities formed on the fly from the logits and their row
 * log-sum-exp (as bb_mixture_logits returns them): r[n,k] = exp(logits[n,k] - lse[n]) -- the
 * N x K responsibility matrix is never written.  tcgen05 path only: d % 8 == 0, d <= 64,
 * k % 4 == 0, k <= 4096 -- more than 256 components run as slices of 256 -- (BB_ERR_UNSUPPORTED otherwise;
 * normalise with bb_logsoftmax_rows and use
 * bb_suffstats_weighted instead). */
BB_API int bb_suffstats_weighted_from_logits(const float* X, const float* logits, const float* lse,
                                      int64_t n, int32_t d, int32_t k,
                                      double* Nk, double* sum_rx, double* sum_rxx,
                                      void* workspace, int64_t workspace_bytes, void* stream);

/* The same statistics with the responsibilities handed over ALREADY in the tensor core's operand format: the
 * statistics kernel is bound by its operand conversion, and every one of its ten column-tile CTAs per row range
 * converts the same rows of R -- so the row pass that normalises the logits writes r = exp(logits - lse) split
 * into error-compensated BF16 (b1 | b2) operand tiles instead of float32 (same bytes), and the statistics kernel
 * bulk-copies them (cp.async.bulk) and only forms the X (x) X operand.  Same value as
 * bb_logsoftmax_rows + exp + bb_suffstats_weighted (plans of algebra.py:527-765 over the elementwise chain of
 * :1435-1448).  Needs k in {256, 512, 768, 1024} (slices of 256 components), d % 8 == 0, d <= 64.
 * rsplit: device buffer of bb_softmax_rows_split_bytes(n, k) bytes = 4 k * 32 ceil(n / 32). */
BB_API int64_t bb_softmax_rows_split_bytes(int64_t n, int32_t k);
BB_API int bb_softmax_rows_split(const float* logits, int64_t n, int32_t k, void* rsplit, float* lse,
                          double* sum_lse, void* stream);
BB_API int bb_suffstats_weighted_split(const float* X, const void* rsplit, int64_t n, int32_t d, int32_t k,
                                double* Nk, double* sum_rx, double* sum_rxx,
                                void* workspace, int64_t workspace_bytes, void* stream);

/* ---- parameter-space update steps either side of the data pass ----
 * The reference names these algorithms in prose only (README.md:30-37 VMP, :47-51
 * reparameterised gradients, :69-80 SVI).  All pointers are device pointers; nothing here
 * synchronises with the host, so a whole iteration can be enqueued on one stream. */

/* VMP global step (README.md:30-37) of a K-component Gaussian mixture with a Dirichlet(alpha0) prior on the
 * weights and Gaussian-Wishart(m0, beta0, W0, nu0) priors on the components (Bishop PRML
 * 10.58-10.63), from the (all-reduced) statistics of the local step:
 *   alpha_k = alpha0 + N_k, beta_k = beta0 + N_k, nu_k = nu0 + N_k,
 *   m_k = (beta0 m0 + sum r x) / beta_k,
 *   W_k^-1 = W0^-1 + sum r x x^T + beta0 m0 m0^T - beta_k m_k m_k^T.
 * Also emits what the next local step consumes (bb_mixture_logits, upper_triangular = 1):
 *   U_k upper triangular with nu_k W_k = U_k^T U_k, t_k = U_k m_k,
 *   c_k = E[log pi_k] + 1/2 E[log|Lambda_k|] - D/2 log 2 pi - D / (2 beta_k),
 * and kl[0..k-1] = KL(q(mu_k, Lambda_k) || p), kl[k] = KL(q(pi) || p(pi)); kl has k + 2 slots
 * (the last one is scratch: log|W0^-1|).  status (device int32) is 0, or 1 + the index of a
 * component whose W_k^-1 is not positive definite (its outputs are NaN).  1 <= d <= 96. */
BB_API int bb_gmm_global_update(const double* Nk, const double* sum_rx, const double* sum_rxx,
                         int32_t k, int32_t d, double alpha0, double beta0, double nu0,
                         const double* m0, const double* W0_inv,
                         double* alpha, double* beta, double* nu, double* m, double* W_inv,
                         float* U, float* t, float* c, double* kl, int32_t* status, void* stream);

/* ---- multi-GPU combine over NVLink peer memory ---------------------------------------------
 * What it replaces: nothing the reference has -- ExpFamIndependentObservations.sufficient_statistics
 * (bayesic/distribution/base.py:328-332) sums the per-point statistics over the iid axis; with the
 * data axis sharded over one process per GPU that sum needs one combine of the per-rank partial
 * statistics per minibatch.  The library does not bootstrap inter-process memory:
 * the host binding allocates the buffers below on every rank, maps its peers' buffers into its own
 * address space (CUDA IPC / cuMem fabric handles; bayesic_b200/parallel.py uses torch's symmetric
 * memory for exactly that) and passes HOST arrays of `world` DEVICE pointers, index = rank.
 *
 * bb_comm: a two-shot all-reduce(sum) of float64[count] as ONE kernel per rank: counterpart CTAs
 * handshake through epoch flags (no grid barrier), each rank pulls and reduces its 1/world chunk from
 * all ranks' input buffers in rank order (bit-identical everywhere) and pushes the sums into all
 * ranks' output buffers (csrc/p2p_reduce.cu).
 *   peer_in / peer_out   float64[capacity] per rank, 16-byte aligned; this rank writes its partial
 *                        statistics into its own input buffer (stream order) before the call and finds
 *                        the reduced payload in its own output buffer after it;
 *   peer_flags           bb_comm_flag_bytes(world) bytes per rank, ZEROED on every rank (and a
 *                        host-side barrier passed) before the first call;
 *   every rank must call bb_comm_allreduce_sum the same number of times with the same count;
 *   consumers of the output must be ordered on `stream` before the next call;
 *   a peer that does not answer within spin_limit_ms (<= 0: 2000) sets the status word to 1 + its
 *   rank and the output is filled with NaN -- never a partial sum; bb_comm_status reads the word
 *   (synchronises `stream`). */
typedef struct bb_comm bb_comm;
BB_API int64_t bb_comm_flag_bytes(int32_t world);
BB_API int bb_comm_create(int32_t rank, int32_t world, const void* const* peer_in,
                   const void* const* peer_out, const void* const* peer_flags, int64_t capacity,
                   double spin_limit_ms, bb_comm** comm);
BB_API int bb_comm_allreduce_sum(bb_comm* comm, int64_t count, void* stream);
BB_API int bb_comm_status(bb_comm* comm, int32_t* status, void* stream);
BB_API int bb_comm_destroy(bb_comm* comm);

/* bb_gaussian_pass: the whole cfg2 step -- statistics of base.py:328-332 for core.py:41-44, their
 * combine over the ranks, and the expected log-likelihood of base.py:25-100 -- as ONE kernel launch
 * per step on any number of GPUs (csrc/suffstats_sm100.cu: tcgen05 pass with dynamically claimed tiles,
 * L2 float64 reductions of the CTAs' partial statistics, then the LAST CTA pushes the rank's payload into
 * the peers' receive buffers + one flag per peer, sums the world's slots in rank order and evaluates the
 * ELBO term).  Back-to-back runs of one handle are launched with programmatic stream serialization: the
 * next run streams its rows while this run's last CTA finishes (it touches the shared state only after
 * griddepcontrol.wait), so consecutive runs must not depend on each other's OUTPUTS through device memory
 * other than by stream order of other kernels in between -- which is the ordinary stream contract: a
 * consumer kernel enqueued between two runs sees the first run complete.  Assumes one rank per GPU.
 * The handle owns its workspace (the only device allocation; one handle = one stream at a time).
 *   d                    4 <= d <= 64, d % 4 == 0;
 *   attach_peers         optional (world > 1): peer_recv = bb_gaussian_pass_peer_bytes' recv_bytes per
 *                        rank, peer_flags = flag_bytes per rank, zeroed + barrier before the first run;
 *   run                  X device float32 [n, d] (this rank's rows; n may be 0 on a rank).  Outputs
 *                        (device float64): sum_x[d] (may be NULL), sum_xxT[d, d], count_out (may be
 *                        NULL), loglik (may be NULL; needs E_Lambda[d, d], E_Lambda_mu[d]) -- all
 *                        REDUCED over the ranks when peers are attached, bit-identical on every rank.
 *                        n_total is the row count the log-likelihood refers to on one GPU; with
 *                        peers the reduced count is used;
 *   status               1 + rank of a peer that did not answer (outputs are NaN then). */
typedef struct bb_gaussian_pass bb_gaussian_pass;
BB_API int bb_gaussian_pass_create(int32_t d, bb_gaussian_pass** pass);
BB_API int bb_gaussian_pass_peer_bytes(int32_t d, int32_t world, int64_t* recv_bytes, int64_t* flag_bytes);
BB_API int bb_gaussian_pass_attach_peers(bb_gaussian_pass* pass, int32_t rank, int32_t world,
                                  const void* const* peer_recv, const void* const* peer_flags,
                                  double spin_limit_ms);
BB_API int bb_gaussian_pass_run(bb_gaussian_pass* pass, const float* X, int64_t n,
                         const double* E_Lambda, const double* E_Lambda_mu, double E_mu_L_mu,
                         double E_logdet, double n_total, double* sum_x, double* sum_xxT,
                         double* count_out, double* loglik, void* stream);
BB_API int bb_gaussian_pass_status(bb_gaussian_pass* pass, int32_t* status, void* stream);
BB_API int bb_gaussian_pass_destroy(bb_gaussian_pass* pass);

/* Minibatch selection on the device (README.md:71-73 "subsample the data"): out[j, :] = X[index[j], :]
 * for j < m.  Indices outside [0, n) are counted in *n_out_of_range (device int32) and their rows
 * NaN-filled; nothing is read back by the library. */
BB_API int bb_gather_rows(const float* X, int64_t n, int32_t d, const int64_t* index, int64_t m,
                   float* out, int32_t* n_out_of_range, void* stream);

/* SVI natural-parameter blend (Hoffman et al. 2013; README.md:69-80), in place:
 *   eta[i] <- (1 - rho) eta[i] + rho (eta_prior[i] + scale * stat[i]),  i < count. */
BB_API int bb_svi_natural_blend(double* eta, const double* eta_prior, const double* stat,
                         double scale, double rho, int64_t count, void* stream);

/* Reparameterised draws W[s, j] = mu[j] + exp(log_sigma[j]) * eps[s, j]  (float32 out, [S, D];
 * README.md:47-51). */
BB_API int bb_reparam_draws(const double* mu, const double* log_sigma, const double* eps,
                     int32_t d, int32_t s, float* W, void* stream);

/* ELBO and its reparameterised gradient (README.md:47-51) for q(w) = N(mu, diag sigma^2), prior N(0, I), from the
 * outputs of bb_logistic_reparam_pass (G[D, S], loglik[S]):
 *   elbo = mean_s loglik_s - KL,  grad_mu = mean_s G_s - mu,
 *   grad_log_sigma = mean_s (G_s * eps_s) * sigma - sigma^2 + 1. */
BB_API int bb_reparam_gradient(const double* G, const double* loglik, const double* eps,
                        const double* mu, const double* log_sigma, int32_t d, int32_t s,
                        double* grad_mu, double* grad_log_sigma, double* elbo, void* stream);

/* One Adam step on count float64 parameters (step counts from 1); maximize != 0 ascends -- the
 * optimiser step of the reparameterised-gradient loop (README.md:47-51). */
BB_API int bb_adam_step(double* param, const double* grad, double* m, double* v, int64_t count,
                 double lr, double beta1, double beta2, double eps, int64_t step,
                 int32_t maximize, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BAYESIC_B200_H */
