// Plan executor: walks the flat plan descriptor (include/bayesic_b200.h) in topological order,
// infers extents, carves results out of the caller's workspace and launches kernels.
//
// It replaces the reference's per-call path  f(**inputs) -> theano_fn(*arrays)
// (bayesic/algebra.py:55-56), i.e. the Theano VM running one thunk per plan-IR node
// (node._apply_to_parents, algebra.py:34-40).  Views (_dimshuffle, _diagonal, transposes) never
// move data; literals and shape arithmetic stay on the host and are folded into kernel
// immediates.  The same walk runs "dry" (no pointers, no launches) for bb_plan_infer.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "kernels.h"
#include "plan.h"


namespace bb {

namespace {

struct Arena {
  char* base;
  int64_t size;
  int64_t used;
  bool dry;
  // Returns nullptr in dry mode; sets *ok = false on overflow.
  void* alloc(int64_t bytes, bool* ok) {
    used = align_up(used, 256);
    void* p = dry ? nullptr : static_cast<void*>(base + used);
    used += std::max<int64_t>(bytes, 1);
    if (!dry && used > size) *ok = false;
    return p;
  }
};

struct Exec {
  bb_plan* plan;
  const bb_tensor_arg* inputs;
  void* const* out_ptrs;
  Arena arena;
  cudaStream_t stream;
  std::vector<View> vals;
  std::vector<int> out_slot;   // first output slot of a node, or -1
  std::vector<char> needed;    // value (not just extents) required

  bool dry() const { return arena.dry; }

  int alloc_floats(int node, int64_t count, float** out) {
    if (out_slot[node] >= 0 && !dry()) {
      *out = static_cast<float*>(out_ptrs[out_slot[node]]);
      return BB_OK;
    }
    bool ok = true;
    *out = static_cast<float*>(arena.alloc(count * 4, &ok));
    if (!ok) {
      set_error("workspace too small (need > %lld bytes)", static_cast<long long>(arena.used));
      return BB_ERR_WORKSPACE;
    }
    return BB_OK;
  }
  int alloc_scratch(int64_t bytes, void** out) {
    bool ok = true;
    *out = arena.alloc(bytes, &ok);
    if (!ok) {
      set_error("workspace too small (need > %lld bytes)", static_cast<long long>(arena.used));
      return BB_ERR_WORKSPACE;
    }
    return BB_OK;
  }

  // Dense row-major copy of v (device), or a 1-element buffer for a host scalar.
  int materialize(const View& v, View* out) {
    if (v.is_host) {
      View r = v;
      r.is_host = false;
      void* p = nullptr;
      BB_TRY(alloc_scratch(4, &p));
      r.ptr = static_cast<float*>(p);
      for (int d = 0; d < r.ndim; ++d) r.stride[d] = 0;
      if (!dry()) BB_TRY(launch_fill(r.ptr, 1, static_cast<float>(v.host_value), stream));
      *out = r;
      return BB_OK;
    }
    if (v.is_contiguous()) {
      *out = v;
      return BB_OK;
    }
    View r = v;
    r.set_contiguous_strides();
    void* p = nullptr;
    BB_TRY(alloc_scratch(r.numel() * 4, &p));
    r.ptr = static_cast<float*>(p);
    if (!dry()) BB_TRY(launch_strided_copy(r, v, stream));
    *out = r;
    return BB_OK;
  }
};

double host_elemwise(int op, const double* v, int n) {
  switch (op) {
    case BB_OP_ADD: { double r = 0; for (int i = 0; i < n; ++i) r += v[i]; return r; }
    case BB_OP_MUL: { double r = 1; for (int i = 0; i < n; ++i) r *= v[i]; return r; }
    case BB_OP_LOG: return log(v[0]);
    case BB_OP_EXP: return exp(v[0]);
    case BB_OP_POW: return pow(v[0], v[1]);
    case BB_OP_ABS: return fabs(v[0]);
    case BB_OP_LGAMMA: return lgamma(v[0]);
  }
  return v[0];
}

// n-ary broadcasting pointwise node (MUL and ELEMWISE share this).
int run_pointwise(Exec& ex, int idx, const bb_node_desc& nd, int op) {
  const int n = nd.n_parents;
  if (n < 1) { set_error("node %d: pointwise op without operands", idx); return BB_ERR_INVALID; }
  const int arity = (op == BB_OP_POW) ? 2 : (op == BB_OP_ADD || op == BB_OP_MUL) ? -1 : 1;
  if (arity > 0 && n != arity) {
    set_error("node %d: elementwise op %d expects %d operands, got %d", idx, op, arity, n);
    return BB_ERR_INVALID;
  }
  View ops[BB_MAX_PARENTS];
  bool all_host = true;
  int ndim = 0;
  for (int i = 0; i < n; ++i) {
    ops[i] = ex.vals[nd.parents[i]];
    all_host = all_host && ops[i].is_host;
    ndim = std::max(ndim, ops[i].ndim);
  }
  View out;
  out.ndim = ndim;
  for (int d = 0; d < ndim; ++d) out.shape[d] = 1;
  for (int i = 0; i < n; ++i) {
    if (ops[i].ndim != ndim) {
      if (ops[i].ndim == 0) {            // scalars broadcast (algebra.py:179-192)
        ops[i].ndim = ndim;
        for (int d = 0; d < ndim; ++d) { ops[i].shape[d] = 1; ops[i].stride[d] = 0; }
      } else {
        set_error("node %d: rank mismatch %d vs %d", idx, ops[i].ndim, ndim);
        return BB_ERR_SHAPE;
      }
    }
    for (int d = 0; d < ndim; ++d) {
      const int64_t e = ops[i].shape[d];
      if (e == 1) continue;
      if (out.shape[d] == 1) out.shape[d] = e;
      else if (out.shape[d] != e) {
        set_error("node %d: cannot broadcast extents %lld and %lld on axis %d", idx,
                  static_cast<long long>(out.shape[d]), static_cast<long long>(e), d);
        return BB_ERR_SHAPE;
      }
    }
  }
  if (all_host) {
    double v[BB_MAX_PARENTS];
    for (int i = 0; i < n; ++i) v[i] = ops[i].host_value;
    out.is_host = true;
    out.host_value = host_elemwise(op, v, n);
    ex.vals[idx] = out;
    return BB_OK;
  }
  out.set_contiguous_strides();
  if (ex.needed[idx]) {
    BB_TRY(ex.alloc_floats(idx, out.numel(), &out.ptr));
    if (!ex.dry()) BB_TRY(launch_elementwise(op, out, ops, n, ex.stream));
  }
  ex.vals[idx] = out;
  return BB_OK;
}

// Merge the listed axes of v (in the given order) into one (extent, stride); false if the
// axes are not laid out so that a single stride walks them.
bool merge_axes(const View& v, const int* axes, int n, int64_t* extent, int64_t* stride) {
  int64_t e = 1, s = 0;
  bool have = false;
  for (int i = n - 1; i >= 0; --i) {
    const int64_t ae = v.shape[axes[i]], as = v.stride[axes[i]];
    if (ae == 1) continue;
    if (!have) { e = ae; s = as; have = true; }
    else {
      if (as != s * e) return false;
      e *= ae;
    }
  }
  *extent = e;
  *stride = have ? s : 0;
  return true;
}

int run_tensordot(Exec& ex, int idx, const bb_node_desc& nd) {
  if (nd.n_parents != 2 || nd.n_iparams < 2) { set_error("node %d: bad tensordot", idx); return BB_ERR_INVALID; }
  const int n_dot = nd.iparams[0], n_batch = nd.iparams[1];
  if (nd.n_iparams != 2 + 2 * n_dot + 2 * n_batch) { set_error("node %d: bad tensordot params", idx); return BB_ERR_INVALID; }
  const int* x_dot = nd.iparams + 2;
  const int* y_dot = x_dot + n_dot;
  const int* x_batch = y_dot + n_dot;
  const int* y_batch = x_batch + n_batch;
  View X = ex.vals[nd.parents[0]], Y = ex.vals[nd.parents[1]];
  int x_other[kMaxDims], y_other[kMaxDims], nxo = 0, nyo = 0;
  for (int a = 0; a < X.ndim; ++a) {
    bool used = false;
    for (int i = 0; i < n_dot; ++i) used |= (x_dot[i] == a);
    for (int i = 0; i < n_batch; ++i) used |= (x_batch[i] == a);
    if (!used) x_other[nxo++] = a;
  }
  for (int a = 0; a < Y.ndim; ++a) {
    bool used = false;
    for (int i = 0; i < n_dot; ++i) used |= (y_dot[i] == a);
    for (int i = 0; i < n_batch; ++i) used |= (y_batch[i] == a);
    if (!used) y_other[nyo++] = a;
  }
  for (int i = 0; i < n_dot; ++i) {
    if (x_dot[i] < 0 || x_dot[i] >= X.ndim || y_dot[i] < 0 || y_dot[i] >= Y.ndim) {
      set_error("node %d: tensordot axis out of range", idx); return BB_ERR_INVALID;
    }
    if (X.shape[x_dot[i]] != Y.shape[y_dot[i]]) {
      set_error("node %d: contracted extents differ (%lld vs %lld)", idx,
                static_cast<long long>(X.shape[x_dot[i]]), static_cast<long long>(Y.shape[y_dot[i]]));
      return BB_ERR_SHAPE;
    }
  }
  for (int i = 0; i < n_batch; ++i)
    if (X.shape[x_batch[i]] != Y.shape[y_batch[i]]) {
      set_error("node %d: batch extents differ", idx); return BB_ERR_SHAPE;
    }
  View out;
  out.ndim = n_batch + nxo + nyo;
  if (out.ndim > kMaxDims) { set_error("node %d: result rank %d too large", idx, out.ndim); return BB_ERR_UNSUPPORTED; }
  int o = 0;
  for (int i = 0; i < n_batch; ++i) out.shape[o++] = X.shape[x_batch[i]];
  for (int i = 0; i < nxo; ++i) out.shape[o++] = X.shape[x_other[i]];
  for (int i = 0; i < nyo; ++i) out.shape[o++] = Y.shape[y_other[i]];
  out.set_contiguous_strides();
  if (!ex.needed[idx]) { ex.vals[idx] = out; return BB_OK; }

  // Bring each operand to (batch, other, dot) single-stride form; copy if its layout
  // cannot be walked that way.
  auto canon = [&](View& V, const int* batch_axes, const int* other_axes, int n_other,
                   const int* dot_axes, int64_t ext[3], int64_t str[3]) -> int {
    if (V.is_host) BB_TRY(ex.materialize(V, &V));
    bool ok = merge_axes(V, batch_axes, n_batch, &ext[0], &str[0]) &&
              merge_axes(V, other_axes, n_other, &ext[1], &str[1]) &&
              merge_axes(V, dot_axes, n_dot, &ext[2], &str[2]);
    if (ok) return BB_OK;
    // permuted dense copy in (batch, other, dot) order
    View perm;
    perm.ndim = V.ndim;
    int order[kMaxDims], n = 0;
    for (int i = 0; i < n_batch; ++i) order[n++] = batch_axes[i];
    for (int i = 0; i < n_other; ++i) order[n++] = other_axes[i];
    for (int i = 0; i < n_dot; ++i) order[n++] = dot_axes[i];
    for (int i = 0; i < n; ++i) { perm.shape[i] = V.shape[order[i]]; perm.stride[i] = V.stride[order[i]]; }
    perm.ptr = V.ptr;
    View dense = perm;
    dense.set_contiguous_strides();
    void* p = nullptr;
    BB_TRY(ex.alloc_scratch(dense.numel() * 4, &p));
    dense.ptr = static_cast<float*>(p);
    if (!ex.dry()) BB_TRY(launch_strided_copy(dense, perm, ex.stream));
    ext[0] = ext[1] = ext[2] = 1;
    for (int i = 0; i < n_batch; ++i) ext[0] *= dense.shape[i];
    for (int i = 0; i < n_other; ++i) ext[1] *= dense.shape[n_batch + i];
    for (int i = 0; i < n_dot; ++i) ext[2] *= dense.shape[n_batch + n_other + i];
    str[2] = 1; str[1] = ext[2]; str[0] = ext[1] * ext[2];
    V = dense;
    return BB_OK;
  };
  int64_t xe[3], xs[3], ye[3], ys[3];
  BB_TRY(canon(X, x_batch, x_other, nxo, x_dot, xe, xs));
  BB_TRY(canon(Y, y_batch, y_other, nyo, y_dot, ye, ys));
  const int64_t batch = xe[0], M = xe[1], K = xe[2], N = ye[1];
  BB_TRY(ex.alloc_floats(idx, out.numel(), &out.ptr));
  void* ws = nullptr;
  const int64_t ws_bytes = gemm_workspace_bytes(M, N, K, batch);
  if (ws_bytes > 0) BB_TRY(ex.alloc_scratch(ws_bytes, &ws));

  // Large contractions of the two shapes the hot path is made of go to the tcgen05 kernels
  // (decided from extents and strides only, so the dry pass reserves the same scratch):
  //   rows x features . (q x features)^T  -> rowproj   (dot(X, W.T): per-row projection)
  //   (rows x d)^T . rows x q             -> colproj   (dot(X.T, R): contraction over the data axis)
  const int64_t kMinRows = 4096;
  const bool small_dims = M <= 2147483647LL && N <= 4096 && K <= 2147483647LL;
  const bool x_k_major = xs[2] == 1 && xs[1] == K, y_k_major = ys[2] == 1 && ys[1] == K;
  const bool x_mn_major = xs[1] == 1 && xs[2] == M, y_mn_major = ys[1] == 1 && ys[2] == N;
  const bool row_shape = batch == 1 && small_dims && M >= kMinRows && x_k_major && (y_k_major || y_mn_major) &&
                         K >= 64 && K % 64 == 0 && K <= 32768 && N >= 16 && N % 16 == 0 && N <= 256 && N * K <= 32768;
  const bool col_shape = batch == 1 && small_dims && K >= kMinRows && x_mn_major && y_mn_major && M >= 128 &&
                         M % 128 == 0 && N >= 64 && N % 64 == 0 && (M / 128) * (N / 64) <= 4;
  // Matrix-vector contractions over a tall matrix (dot(X.T, y), dot(X, w) and their mirror images with
  // the vector on the left): one pass over the matrix at HBM rate instead of GEMM tiles that are 1 wide.
  //   mat [rows, feats] dense row-major; vec dense.  cols: contract rows; rows: contract feats.
  const bool vec_right = batch == 1 && N == 1 && ys[2] == 1, vec_left = batch == 1 && M == 1 && xs[2] == 1;
  const bool gemv_cols_r = vec_right && K >= kMinRows && x_mn_major && M <= 4096;          // X^T y
  const bool gemv_rows_r = vec_right && M >= kMinRows && x_k_major && K <= 65536;           // X w
  const bool gemv_cols_l = !vec_right && vec_left && K >= kMinRows && ys[1] == 1 && ys[2] == N && N <= 4096;   // y^T X
  const bool gemv_rows_l = !vec_right && vec_left && N >= kMinRows && ys[2] == 1 && ys[1] == K && K <= 65536;  // (X w)^T
  void* gemv_ws = nullptr;
  if (gemv_cols_r || gemv_cols_l) {
    const int64_t feats = gemv_cols_r ? M : N;
    BB_TRY(ex.alloc_scratch(gemv_cols_workspace(K, feats), &gemv_ws));
  }
  void* tc_ws = nullptr;
  void* tc_aux = nullptr;      // rowproj: W transposed to (q, features); colproj: float64 result
  int64_t tc_ws_bytes = 0;
  if (row_shape) {
    tc_ws_bytes = rowproj_tc_workspace(M, static_cast<int>(K), static_cast<int>(N));
    BB_TRY(ex.alloc_scratch(tc_ws_bytes, &tc_ws));
    if (!y_k_major) BB_TRY(ex.alloc_scratch(N * K * 4, &tc_aux));
  } else if (col_shape) {
    tc_ws_bytes = colproj_tc_workspace(K, static_cast<int>(M), static_cast<int>(N));
    BB_TRY(ex.alloc_scratch(tc_ws_bytes, &tc_ws));
    BB_TRY(ex.alloc_scratch(M * N * 8, &tc_aux));
  }
  if (!ex.dry()) {
    bool done = false;
    if (row_shape && rowproj_tc_supported(M, static_cast<int>(K), static_cast<int>(N), X.ptr) &&
        reinterpret_cast<uintptr_t>(out.ptr) % 16 == 0) {
      const float* w = Y.ptr;
      if (!y_k_major) {      // W given as (features, q): dense transposed copy (tiny)
        View src, dst;
        src.ptr = Y.ptr; src.ndim = 2; src.shape[0] = N; src.shape[1] = K; src.stride[0] = ys[1]; src.stride[1] = ys[2];
        dst = src; dst.ptr = static_cast<float*>(tc_aux); dst.set_contiguous_strides();
        BB_TRY(launch_strided_copy(dst, src, ex.stream));
        w = dst.ptr;
      }
      BB_TRY(launch_rowproj_tc(X.ptr, w, nullptr, M, static_cast<int>(K), static_cast<int>(N), out.ptr, nullptr,
                               tc_ws, tc_ws_bytes, ex.stream));
      done = true;
    } else if (col_shape && colproj_tc_supported(K, static_cast<int>(M), static_cast<int>(N), X.ptr, Y.ptr)) {
      BB_TRY(launch_colproj_tc(X.ptr, Y.ptr, K, static_cast<int>(M), static_cast<int>(N),
                               static_cast<double*>(tc_aux), tc_ws, tc_ws_bytes, ex.stream));
      BB_TRY(launch_f64_to_f32(static_cast<const double*>(tc_aux), out.ptr, M * N, ex.stream));
      done = true;
    }
    if (!done && (gemv_cols_r || gemv_cols_l)) {
      const float* mat = gemv_cols_r ? X.ptr : Y.ptr;
      const float* vec = gemv_cols_r ? Y.ptr : X.ptr;
      const int64_t feats = gemv_cols_r ? M : N;
      if (gemv_cols_supported(K, feats, mat)) {
        BB_TRY(launch_gemv_cols(mat, vec, K, static_cast<int>(feats), out.ptr, gemv_ws, ex.stream));
        done = true;
      }
    }
    if (!done && (gemv_rows_r || gemv_rows_l)) {
      const float* mat = gemv_rows_r ? X.ptr : Y.ptr;
      const float* vec = gemv_rows_r ? Y.ptr : X.ptr;
      BB_TRY(launch_gemv_rows(mat, vec, gemv_rows_r ? M : N, static_cast<int>(K), out.ptr, ex.stream));
      done = true;
    }
    if (!done)
      BB_TRY(launch_gemm(X.ptr, Y.ptr, out.ptr, M, N, K, batch, xs[0], xs[1], xs[2], ys[0], ys[2],
                         ys[1], ws, ex.stream));
  }
  ex.vals[idx] = out;
  return BB_OK;
}

int run_sum(Exec& ex, int idx, const bb_node_desc& nd) {
  const View& X = ex.vals[nd.parents[0]];
  bool reduce[kMaxDims] = {false};
  for (int i = 0; i < nd.n_iparams; ++i) {
    const int a = nd.iparams[i];
    if (a < 0 || a >= X.ndim || reduce[a]) { set_error("node %d: bad sum axis %d", idx, a); return BB_ERR_INVALID; }
    reduce[a] = true;
  }
  View out;
  out.ndim = 0;
  int64_t reduced = 1;
  for (int d = 0; d < X.ndim; ++d) {
    if (reduce[d]) reduced *= X.shape[d];
    else out.shape[out.ndim++] = X.shape[d];
  }
  if (X.is_host) {
    out.is_host = true;
    out.host_value = X.host_value * static_cast<double>(reduced);
    ex.vals[idx] = out;
    return BB_OK;
  }
  out.set_contiguous_strides();
  if (ex.needed[idx]) {
    BB_TRY(ex.alloc_floats(idx, out.numel(), &out.ptr));
    void* scratch = nullptr;
    BB_TRY(ex.alloc_scratch(reduce_sum_scratch_bytes(out.numel()), &scratch));
    if (!ex.dry()) BB_TRY(launch_reduce_sum(X, reduce, out, scratch, ex.stream));
  }
  ex.vals[idx] = out;
  return BB_OK;
}

int run_node(Exec& ex, int idx) {
  const bb_node_desc& nd = ex.plan->nodes[idx];
  View out;
  switch (nd.kind) {
    case BB_NODE_INPUT: {
      const bb_tensor_arg& a = ex.inputs[nd.iparams[0]];
      if (a.ndim < 0 || a.ndim > kMaxDims) { set_error("input %d: bad rank", nd.iparams[0]); return BB_ERR_INVALID; }
      out.ndim = a.ndim;
      for (int d = 0; d < a.ndim; ++d) out.shape[d] = a.shape[d];
      if (a.is_host_scalar) {
        out.is_host = true;
        out.host_value = a.host_value;
        for (int d = 0; d < a.ndim; ++d) out.shape[d] = 1;
      } else {
        out.ptr = const_cast<float*>(static_cast<const float*>(a.data));
        out.set_contiguous_strides();
        if (!ex.dry() && out.ptr == nullptr && out.numel() > 0) {
          set_error("input %d: null device pointer", nd.iparams[0]);
          return BB_ERR_INVALID;
        }
      }
      break;
    }
    case BB_NODE_SCALAR:
      out.ndim = 0;
      out.is_host = true;
      out.host_value = nd.fparam;
      break;
    case BB_NODE_SHAPE: {
      const View& X = ex.vals[nd.parents[0]];
      const int a = nd.iparams[0];
      if (a < 0 || a >= X.ndim) { set_error("node %d: shape axis %d out of range", idx, a); return BB_ERR_INVALID; }
      out.ndim = 0;
      out.is_host = true;
      out.host_value = static_cast<double>(X.shape[a]);
      break;
    }
    case BB_NODE_EYE: {
      const View& nv = ex.vals[nd.parents[0]];
      if (!nv.is_host) { set_error("node %d: eye extent must be a host scalar", idx); return BB_ERR_UNSUPPORTED; }
      const int64_t n = static_cast<int64_t>(llround(nv.host_value));
      if (n < 0) { set_error("node %d: negative eye extent", idx); return BB_ERR_SHAPE; }
      out.ndim = 2;
      out.shape[0] = out.shape[1] = n;
      out.set_contiguous_strides();
      if (ex.needed[idx]) {
        BB_TRY(ex.alloc_floats(idx, n * n, &out.ptr));
        if (!ex.dry()) BB_TRY(launch_eye(out.ptr, n, ex.stream));
      }
      break;
    }
    case BB_NODE_DIMSHUFFLE: {
      const View& X = ex.vals[nd.parents[0]];
      out = X;
      out.ndim = nd.n_iparams;
      if (out.ndim > kMaxDims) { set_error("node %d: rank too large", idx); return BB_ERR_UNSUPPORTED; }
      for (int i = 0; i < nd.n_iparams; ++i) {
        const int a = nd.iparams[i];
        if (a < 0) { out.shape[i] = 1; out.stride[i] = 0; }
        else if (a < X.ndim) { out.shape[i] = X.shape[a]; out.stride[i] = X.stride[a]; }
        else { set_error("node %d: dimshuffle axis %d out of range", idx, a); return BB_ERR_INVALID; }
      }
      break;
    }
    case BB_NODE_DIAGONAL: {
      const View& X = ex.vals[nd.parents[0]];
      const int a1 = nd.iparams[0], a2 = nd.iparams[1];
      if (a1 < 0 || a2 < 0 || a1 >= X.ndim || a2 >= X.ndim || a1 == a2) { set_error("node %d: bad diagonal axes", idx); return BB_ERR_INVALID; }
      if (X.shape[a1] != X.shape[a2]) {
        set_error("node %d: diagonal of unequal extents %lld, %lld", idx,
                  static_cast<long long>(X.shape[a1]), static_cast<long long>(X.shape[a2]));
        return BB_ERR_SHAPE;
      }
      out = X;
      out.ndim = 0;
      for (int d = 0; d < X.ndim; ++d)
        if (d != a1 && d != a2) { out.shape[out.ndim] = X.shape[d]; out.stride[out.ndim] = X.stride[d]; ++out.ndim; }
      out.shape[out.ndim] = X.shape[a1];
      out.stride[out.ndim] = X.stride[a1] + X.stride[a2];
      ++out.ndim;
      break;
    }
    case BB_NODE_SUM:
      return run_sum(ex, idx, nd);
    case BB_NODE_MUL:
      return run_pointwise(ex, idx, nd, BB_OP_MUL);
    case BB_NODE_ELEMWISE:
      if (nd.n_iparams != 1) { set_error("node %d: elemwise needs an opcode", idx); return BB_ERR_INVALID; }
      return run_pointwise(ex, idx, nd, nd.iparams[0]);
    case BB_NODE_TENSORDOT:
      return run_tensordot(ex, idx, nd);
    case BB_NODE_LOGSOFTMAX: {
      View X = ex.vals[nd.parents[0]];
      if (X.ndim < 1 || X.is_host) { set_error("node %d: logsoftmax needs a device tensor of rank >= 1", idx); return BB_ERR_UNSUPPORTED; }
      out = X;
      out.set_contiguous_strides();
      if (ex.needed[idx]) {
        BB_TRY(ex.materialize(X, &X));
        BB_TRY(ex.alloc_floats(idx, out.numel(), &out.ptr));
        const int64_t k = X.shape[X.ndim - 1];
        if (k > 2147483647LL || k < 1) { set_error("node %d: bad last extent", idx); return BB_ERR_SHAPE; }
        if (!ex.dry())
          BB_TRY(launch_logsoftmax_rows(X.ptr, out.numel() / k, static_cast<int>(k), out.ptr, nullptr,
                                        nullptr, false, ex.stream));
      }
      break;
    }
    case BB_NODE_SYRK: {
      View X = ex.vals[nd.parents[0]];
      if (X.ndim != 2 || X.is_host) { set_error("node %d: syrk needs a device matrix", idx); return BB_ERR_UNSUPPORTED; }
      const int64_t n = X.shape[0], d = X.shape[1];
      out.ndim = 2;
      out.shape[0] = out.shape[1] = d;
      out.set_contiguous_strides();
      if (ex.needed[idx]) {
        BB_TRY(ex.materialize(X, &X));
        BB_TRY(ex.alloc_floats(idx, d * d, &out.ptr));
        // Both variants reserve scratch so the dry pass (which sees no pointers) is an upper bound.
        const bool tc_shape = d >= 4 && d <= 64 && d % 4 == 0 && n > 0 && n < (int64_t(1) << 31) - 128;
        void* s2 = nullptr;
        void* ws = nullptr;
        void* gws = nullptr;
        int64_t ws_bytes = 0;
        if (tc_shape) {
          ws_bytes = suffstats_tc_workspace(n);
          BB_TRY(ex.alloc_scratch(d * d * 8, &s2));
          BB_TRY(ex.alloc_scratch(ws_bytes, &ws));
        }
        const bool gram_shape = d > 64 && d % 4 == 0 && d <= 4096 && n > 0;   // tcgen05 CTA-pair kernel (zero-padded to 256 k)
        if (gram_shape) {
          ws_bytes = gram_tc_workspace(n, static_cast<int>(d));
          BB_TRY(ex.alloc_scratch(d * d * 8, &s2));
          BB_TRY(ex.alloc_scratch(ws_bytes, &ws));
        }
        const int64_t gemm_bytes = gemm_workspace_bytes(d, d, n, 1);
        if (gemm_bytes > 0) BB_TRY(ex.alloc_scratch(gemm_bytes, &gws));
        if (!ex.dry()) {
          if (gram_shape && gram_tc_supported(n, static_cast<int>(d), X.ptr)) {
            BB_TRY(launch_gram_tc(X.ptr, nullptr, n, static_cast<int>(d), static_cast<double*>(s2), nullptr,
                                  nullptr, ws, ws_bytes, ex.stream));
            BB_TRY(launch_f64_to_f32(static_cast<const double*>(s2), out.ptr, d * d, ex.stream));
          } else if (tc_shape && suffstats_tc_supported(n, static_cast<int>(d), X.ptr)) {
            BB_TRY(launch_suffstats_tc(X.ptr, n, static_cast<int>(d), nullptr, static_cast<double*>(s2),
                                       ws, ws_bytes, ex.stream));
            BB_TRY(launch_f64_to_f32(static_cast<const double*>(s2), out.ptr, d * d, ex.stream));
          } else {
            BB_TRY(launch_gemm(X.ptr, X.ptr, out.ptr, d, d, n, 1, 0, 1, d, 0, d, 1, gws, ex.stream));
          }
        }
      }
      break;
    }
    case BB_NODE_WEIGHTED_SCATTER: {
      View R = ex.vals[nd.parents[0]], X = ex.vals[nd.parents[1]];
      if (R.ndim != 2 || X.ndim != 2 || R.is_host || X.is_host) { set_error("node %d: weighted scatter needs two device matrices", idx); return BB_ERR_UNSUPPORTED; }
      if (R.shape[0] != X.shape[0]) { set_error("node %d: R and X disagree on the data axis", idx); return BB_ERR_SHAPE; }
      const int64_t n = X.shape[0], d = X.shape[1], k = R.shape[1];
      out.ndim = 3;
      out.shape[0] = k; out.shape[1] = d; out.shape[2] = d;
      out.set_contiguous_strides();
      if (ex.needed[idx]) {
        BB_TRY(ex.materialize(R, &R));
        BB_TRY(ex.materialize(X, &X));
        BB_TRY(ex.alloc_floats(idx, out.numel(), &out.ptr));
        void* acc = nullptr;
        void* wws = nullptr;
        BB_TRY(ex.alloc_scratch(out.numel() * 8, &acc));
        const int64_t wws_bytes = weighted_stats_auto_workspace(n, static_cast<int>(d), static_cast<int>(k));
        BB_TRY(ex.alloc_scratch(wws_bytes, &wws));
        if (!ex.dry()) {
          BB_TRY(launch_weighted_stats_auto(X.ptr, R.ptr, n, static_cast<int>(d), static_cast<int>(k), nullptr,
                                            nullptr, static_cast<double*>(acc), wws, wws_bytes, ex.stream));
          BB_TRY(launch_f64_to_f32(static_cast<const double*>(acc), out.ptr, out.numel(), ex.stream));
        }
      }
      break;
    }
    case BB_NODE_LOGDET: {
      View X = ex.vals[nd.parents[0]];
      if (X.ndim < 2 || X.is_host) { set_error("node %d: logdet needs a device tensor of rank >= 2", idx); return BB_ERR_UNSUPPORTED; }
      const int64_t d = X.shape[X.ndim - 1];
      if (d != X.shape[X.ndim - 2]) {
        set_error("node %d: logdet of non-square matrices (%lld x %lld)", idx,
                  static_cast<long long>(X.shape[X.ndim - 2]), static_cast<long long>(d));
        return BB_ERR_SHAPE;
      }
      if (d > 4096) { set_error("node %d: logdet supports d <= 4096 (got %lld)", idx, static_cast<long long>(d)); return BB_ERR_UNSUPPORTED; }
      out.ndim = X.ndim - 2;
      for (int a = 0; a < out.ndim; ++a) out.shape[a] = X.shape[a];
      out.set_contiguous_strides();
      if (ex.needed[idx]) {
        const int64_t batch = out.numel();
        BB_TRY(ex.materialize(X, &X));
        BB_TRY(ex.alloc_floats(idx, batch, &out.ptr));
        void* scratch = nullptr;
        const int64_t bytes = logdet_scratch_bytes(batch, d);
        if (bytes > 0) BB_TRY(ex.alloc_scratch(bytes, &scratch));
        if (!ex.dry()) BB_TRY(launch_logdet_spd(X.ptr, batch, static_cast<int>(d), out.ptr, scratch, ex.stream));
      }
      break;
    }
    default:
      set_error("node %d: unknown kind %d", idx, nd.kind);
      return BB_ERR_INVALID;
  }
  ex.vals[idx] = out;
  return BB_OK;
}

int run_plan(bb_plan* plan, const bb_tensor_arg* inputs, int32_t n_inputs, void* const* out_ptrs,
             void* workspace, int64_t workspace_bytes, cudaStream_t stream, bool dry,
             bb_result_info* results, int64_t* used_bytes) {
  if (plan == nullptr) { set_error("null plan"); return BB_ERR_INVALID; }
  if (n_inputs != plan->n_inputs) {
    set_error("plan expects %d inputs, got %d", plan->n_inputs, n_inputs);
    return BB_ERR_INVALID;
  }
  const int n = static_cast<int>(plan->nodes.size());
  Exec ex;
  ex.plan = plan;
  ex.inputs = inputs;
  ex.out_ptrs = out_ptrs;
  ex.arena = Arena{static_cast<char*>(workspace), workspace_bytes, 0, dry};
  ex.stream = stream;
  ex.vals.resize(n);
  ex.out_slot.assign(n, -1);
  ex.needed.assign(n, 0);
  for (size_t j = 0; j < plan->outputs.size(); ++j) {
    const int node = plan->outputs[j];
    if (ex.out_slot[node] < 0) ex.out_slot[node] = static_cast<int>(j);
    ex.needed[node] = 1;
  }
  for (int i = n - 1; i >= 0; --i) {
    if (!ex.needed[i]) continue;
    const bb_node_desc& nd = plan->nodes[i];
    if (nd.kind == BB_NODE_SHAPE) continue;    // needs extents only
    for (int p = 0; p < nd.n_parents; ++p) ex.needed[nd.parents[p]] = 1;
  }
  // Views must not alias an output buffer they do not own: only compute nodes write in place.
  const int64_t launches_before = g_launch_count;
  for (int i = 0; i < n; ++i) BB_TRY(run_node(ex, i));
  for (size_t j = 0; j < plan->outputs.size(); ++j) {
    const View& v = ex.vals[plan->outputs[j]];
    if (results != nullptr) {
      bb_result_info& r = results[j];
      r.ndim = v.ndim;
      r.is_host_scalar = v.is_host ? 1 : 0;
      r.host_value = v.host_value;
      for (int d = 0; d < kMaxDims; ++d) r.shape[d] = d < v.ndim ? v.shape[d] : 0;
    }
    if (dry || v.is_host) continue;
    float* dst = static_cast<float*>(out_ptrs[j]);
    if (v.ptr == dst && v.is_contiguous()) continue;
    View o = v;
    o.ptr = dst;
    o.set_contiguous_strides();
    BB_TRY(launch_strided_copy(o, v, stream));
  }
  plan->last_launches = static_cast<int32_t>(g_launch_count - launches_before);
  if (used_bytes != nullptr) *used_bytes = align_up(ex.arena.used, 256) + 256;
  return BB_OK;
}

}  // namespace

int plan_infer(bb_plan* plan, const bb_tensor_arg* inputs, int32_t n_inputs,
               bb_result_info* results, int64_t* workspace_bytes) {
  return run_plan(plan, inputs, n_inputs, nullptr, nullptr, 0, nullptr, true, results,
                  workspace_bytes);
}

int plan_execute(bb_plan* plan, const bb_tensor_arg* inputs, int32_t n_inputs,
                 void* const* out_ptrs, void* workspace, int64_t workspace_bytes,
                 cudaStream_t stream) {
  return run_plan(plan, inputs, n_inputs, out_ptrs, workspace, workspace_bytes, stream, false,
                  nullptr, nullptr);
}

int plan_validate(const bb_node_desc* nodes, int32_t n_nodes, const int32_t* outputs,
                  int32_t n_outputs, int32_t n_inputs) {
  if (nodes == nullptr || n_nodes <= 0 || outputs == nullptr || n_outputs <= 0) {
    set_error("plan needs at least one node and one output");
    return BB_ERR_INVALID;
  }
  for (int i = 0; i < n_nodes; ++i) {
    const bb_node_desc& nd = nodes[i];
    if (nd.n_parents < 0 || nd.n_parents > BB_MAX_PARENTS || nd.n_iparams < 0 ||
        nd.n_iparams > BB_MAX_IPARAMS) {
      set_error("node %d: parent/param count out of range", i);
      return BB_ERR_INVALID;
    }
    for (int p = 0; p < nd.n_parents; ++p)
      if (nd.parents[p] < 0 || nd.parents[p] >= i) {
        set_error("node %d: parent %d is not an earlier node", i, nd.parents[p]);
        return BB_ERR_INVALID;
      }
    int want_parents = -1;
    switch (nd.kind) {
      case BB_NODE_INPUT:
        want_parents = 0;
        if (nd.n_iparams != 1 || nd.iparams[0] < 0 || nd.iparams[0] >= n_inputs) {
          set_error("node %d: input slot out of range", i);
          return BB_ERR_INVALID;
        }
        break;
      case BB_NODE_SCALAR: want_parents = 0; break;
      case BB_NODE_SHAPE:
        want_parents = 1;
        if (nd.n_iparams != 1) { set_error("node %d: shape needs one axis", i); return BB_ERR_INVALID; }
        break;
      case BB_NODE_EYE: case BB_NODE_SUM: case BB_NODE_DIMSHUFFLE: case BB_NODE_LOGSOFTMAX:
      case BB_NODE_SYRK: case BB_NODE_LOGDET:
        want_parents = 1; break;
      case BB_NODE_DIAGONAL:
        want_parents = 1;
        if (nd.n_iparams != 2) { set_error("node %d: diagonal needs two axes", i); return BB_ERR_INVALID; }
        break;
      case BB_NODE_TENSORDOT: case BB_NODE_WEIGHTED_SCATTER: want_parents = 2; break;
      case BB_NODE_MUL: case BB_NODE_ELEMWISE:
        if (nd.n_parents < 1) { set_error("node %d: needs operands", i); return BB_ERR_INVALID; }
        break;
      default:
        set_error("node %d: unknown kind %d", i, nd.kind);
        return BB_ERR_INVALID;
    }
    if (want_parents >= 0 && nd.n_parents != want_parents) {
      set_error("node %d (kind %d): expected %d parents, got %d", i, nd.kind, want_parents, nd.n_parents);
      return BB_ERR_INVALID;
    }
  }
  for (int j = 0; j < n_outputs; ++j)
    if (outputs[j] < 0 || outputs[j] >= n_nodes) {
      set_error("output %d refers to node %d", j, outputs[j]);
      return BB_ERR_INVALID;
    }
  return BB_OK;
}

}  // namespace bb
