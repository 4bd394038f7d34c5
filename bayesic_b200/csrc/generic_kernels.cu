// Generic float32 kernels behind the plan-IR vocabulary (bayesic/algebra.py:1280-1414 and
// the elementwise ops :195-233, :1435-1448) for arbitrary ranks/strides.  These make every
// expression the reference can build executable on the device; the hot shapes are served by
// the specialised kernels (suffstats_sm100.cu, mixture kernels), not by these.
//
//   elementwise_kernel  n-ary add / mul, log, exp, pow, abs, copy over broadcast strides
//   reduce_sum_kernel   _sum over any axis set, fp32 chunks folded into float64
//   gemm_splitk_kernel  _tensordot as a batched strided GEMM, split along the contracted
//                       axis when the output is small (contractions over the data axis N)
//   eye / fill / f64->f32 helpers
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "kernels.h"

namespace bb {

// ============================ elementwise ====================================

constexpr int kMaxOperands = BB_MAX_PARENTS;

struct ElemParams {
  int op;
  int n_operands;
  int ndim;                                  // after collapsing
  int64_t shape[kMaxDims];
  int64_t stride[kMaxOperands][kMaxDims];    // element strides, 0 = broadcast
  const float* ptr[kMaxOperands];
  float host_value[kMaxOperands];
  int is_host[kMaxOperands];
  float scale;                               // folded host scalars (mul) / offset (add)
  int64_t total;
  float* out;
};

constexpr int kOpCopy = 100;

__device__ __forceinline__ float apply_op(int op, const float* v, int n, float scale) {
  switch (op) {
    case BB_OP_ADD: {
      float r = scale;
      for (int i = 0; i < n; ++i) r += v[i];
      return r;
    }
    case BB_OP_MUL: {
      float r = scale;
      for (int i = 0; i < n; ++i) r *= v[i];
      return r;
    }
    case BB_OP_LOG: return logf(v[0]);
    case BB_OP_EXP: return expf(v[0]);
    case BB_OP_POW: return powf(v[0], v[1]);
    case BB_OP_ABS: return fabsf(v[0]);
    case BB_OP_LGAMMA: return lgammaf(v[0]);
    default: return v[0];
  }
}

__global__ void elementwise_kernel(const ElemParams p) {
  const int64_t step = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < p.total;
       i += step) {
    int64_t rem = i;
    int64_t off[kMaxOperands];
#pragma unroll
    for (int o = 0; o < kMaxOperands; ++o) off[o] = 0;
    for (int d = p.ndim - 1; d >= 0; --d) {
      const int64_t c = rem % p.shape[d];
      rem /= p.shape[d];
#pragma unroll
      for (int o = 0; o < kMaxOperands; ++o)
        if (o < p.n_operands) off[o] += c * p.stride[o][d];
    }
    float v[kMaxOperands];
#pragma unroll
    for (int o = 0; o < kMaxOperands; ++o)
      if (o < p.n_operands) v[o] = p.is_host[o] ? p.host_value[o] : __ldg(p.ptr[o] + off[o]);
    p.out[i] = apply_op(p.op, v, p.n_operands, p.scale);
  }
}

static int grid_for(int64_t total, int threads) {
  int64_t blocks = (total + threads - 1) / threads;
  const int64_t cap = static_cast<int64_t>(std::max(1, device_sm_count())) * 16;
  return static_cast<int>(std::max<int64_t>(1, std::min(blocks, cap)));
}

// out must be contiguous with the broadcast result shape; operands have the same ndim.
int launch_elementwise(int op, const View& out, const View* operands, int n_operands,
                       cudaStream_t stream) {
  if (n_operands < 1 || n_operands > kMaxOperands) {
    set_error("elementwise: %d operands (max %d)", n_operands, kMaxOperands);
    return BB_ERR_INVALID;
  }
  ElemParams p;
  p.op = op;
  p.out = out.ptr;
  p.total = out.numel();
  p.scale = (op == BB_OP_MUL) ? 1.f : 0.f;
  if (p.total == 0) return BB_OK;
  // Fold host scalars of n-ary add/mul into `scale`; keep them as operands otherwise.
  std::vector<const View*> dev;
  int n = 0;
  for (int o = 0; o < n_operands; ++o) {
    const View& v = operands[o];
    if (v.is_host && (op == BB_OP_MUL || op == BB_OP_ADD)) {
      if (op == BB_OP_MUL) p.scale *= static_cast<float>(v.host_value);
      else p.scale += static_cast<float>(v.host_value);
      continue;
    }
    p.is_host[n] = v.is_host ? 1 : 0;
    p.host_value[n] = static_cast<float>(v.host_value);
    p.ptr[n] = v.ptr;
    dev.push_back(&v);
    ++n;
  }
  p.n_operands = n;
  // Collapse: drop extent-1 axes, merge neighbours that are contiguous for every operand.
  int nd = 0;
  for (int d = 0; d < out.ndim; ++d) {
    if (out.shape[d] == 1) continue;
    int64_t st[kMaxOperands];
    for (int o = 0; o < n; ++o) {
      const View& v = *dev[o];
      st[o] = (v.is_host || v.shape[d] == 1) ? 0 : v.stride[d];
      if (!v.is_host && v.shape[d] != 1 && v.shape[d] != out.shape[d]) {
        set_error("elementwise: extent mismatch on axis %d (%lld vs %lld)", d,
                  static_cast<long long>(v.shape[d]), static_cast<long long>(out.shape[d]));
        return BB_ERR_SHAPE;
      }
    }
    bool merged = false;
    if (nd > 0) {
      merged = true;
      for (int o = 0; o < n; ++o)
        if (p.stride[o][nd - 1] != st[o] * out.shape[d]) merged = false;
      if (merged) {
        p.shape[nd - 1] *= out.shape[d];
        for (int o = 0; o < n; ++o) p.stride[o][nd - 1] = st[o];
      }
    }
    if (!merged) {
      p.shape[nd] = out.shape[d];
      for (int o = 0; o < n; ++o) p.stride[o][nd] = st[o];
      ++nd;
    }
  }
  p.ndim = nd;
  const int threads = 256;
  elementwise_kernel<<<grid_for(p.total, threads), threads, 0, stream>>>(p);
  BB_CHECK_LAUNCH("elementwise_kernel");
  return BB_OK;
}

int launch_strided_copy(const View& out, const View& in, cudaStream_t stream) {
  return launch_elementwise(kOpCopy, out, &in, 1, stream);
}

__global__ void fill_kernel(float* out, int64_t n, float value) {
  const int64_t step = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += step)
    out[i] = value;
}

int launch_fill(float* out, int64_t n, float value, cudaStream_t stream) {
  if (n == 0) return BB_OK;
  fill_kernel<<<grid_for(n, 256), 256, 0, stream>>>(out, n, value);
  BB_CHECK_LAUNCH("fill_kernel");
  return BB_OK;
}

__global__ void eye_kernel(float* out, int64_t n) {
  const int64_t total = n * n;
  const int64_t step = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += step)
    out[i] = (i / n == i % n) ? 1.f : 0.f;
}

int launch_eye(float* out, int64_t n, cudaStream_t stream) {
  if (n == 0) return BB_OK;
  eye_kernel<<<grid_for(n * n, 256), 256, 0, stream>>>(out, n);
  BB_CHECK_LAUNCH("eye_kernel");
  return BB_OK;
}

__global__ void f64_to_f32_kernel(const double* in, float* out, int64_t n) {
  const int64_t step = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += step)
    out[i] = static_cast<float>(in[i]);
}

int launch_f64_to_f32(const double* in, float* out, int64_t n, cudaStream_t stream) {
  if (n == 0) return BB_OK;
  f64_to_f32_kernel<<<grid_for(n, 256), 256, 0, stream>>>(in, out, n);
  BB_CHECK_LAUNCH("f64_to_f32_kernel");
  return BB_OK;
}

// ============================ reduce-sum =====================================

struct ReduceParams {
  const float* in;
  double* scratch;          // [kept] float64, zero-initialised
  int n_kept, n_red;        // collapsed dim counts
  int64_t kept_shape[kMaxDims], kept_stride[kMaxDims];
  int64_t red_shape[kMaxDims], red_stride[kMaxDims];
  int64_t kept_total, red_total;
  int kw, rw;               // block = kw kept elements x rw reduce lanes (kw * rw = 256)
  int64_t red_per_split;
};

__device__ __forceinline__ int64_t offset_of(int64_t idx, int nd, const int64_t* shape,
                                             const int64_t* stride) {
  int64_t off = 0;
  for (int d = nd - 1; d >= 0; --d) {
    off += (idx % shape[d]) * stride[d];
    idx /= shape[d];
  }
  return off;
}

__global__ void reduce_sum_kernel(const ReduceParams p) {
  __shared__ double partial[256];
  const int kx = threadIdx.x % p.kw;
  const int rx = threadIdx.x / p.kw;
  const int64_t k = static_cast<int64_t>(blockIdx.x) * p.kw + kx;
  const int64_t r_begin = static_cast<int64_t>(blockIdx.y) * p.red_per_split;
  const int64_t r_end = min(r_begin + p.red_per_split, p.red_total);
  double acc = 0.0;
  if (k < p.kept_total) {
    const float* base = p.in + offset_of(k, p.n_kept, p.kept_shape, p.kept_stride);
    float chunk = 0.f;
    int in_chunk = 0;
    if (p.n_red == 1) {
      const int64_t st = p.red_stride[0];
      for (int64_t r = r_begin + rx; r < r_end; r += p.rw) {
        chunk += __ldg(base + r * st);
        if (++in_chunk == 64) { acc += chunk; chunk = 0.f; in_chunk = 0; }
      }
    } else {
      for (int64_t r = r_begin + rx; r < r_end; r += p.rw) {
        chunk += __ldg(base + offset_of(r, p.n_red, p.red_shape, p.red_stride));
        if (++in_chunk == 64) { acc += chunk; chunk = 0.f; in_chunk = 0; }
      }
    }
    acc += chunk;
  }
  partial[threadIdx.x] = acc;
  __syncthreads();
  if (rx == 0 && k < p.kept_total) {
    double total = 0.0;
    for (int j = 0; j < p.rw; ++j) total += partial[j * p.kw + kx];
    atomicAdd(p.scratch + k, total);
  }
}

static int pow2_ceil(int64_t x) {
  int p = 1;
  while (p < x) p <<= 1;
  return p;
}

// Collapse a list of (extent, stride) pairs given in logical order.
static int collapse_dims(int n, int64_t* shape, int64_t* stride) {
  int m = 0;
  for (int i = 0; i < n; ++i) {
    if (shape[i] == 1) continue;
    if (m > 0 && stride[m - 1] == stride[i] * shape[i]) {
      shape[m - 1] *= shape[i];
      stride[m - 1] = stride[i];
    } else {
      shape[m] = shape[i];
      stride[m] = stride[i];
      ++m;
    }
  }
  return m;
}

int64_t reduce_sum_scratch_bytes(int64_t kept_total) { return align_up(kept_total * 8, 256); }

// out: contiguous float32 [kept dims in order]; scratch: reduce_sum_scratch_bytes(out.numel()).
int launch_reduce_sum(const View& in, const bool* reduce_axis, const View& out, void* scratch,
                      cudaStream_t stream) {
  ReduceParams p;
  p.in = in.ptr;
  p.scratch = static_cast<double*>(scratch);
  int nk = 0, nr = 0;
  for (int d = 0; d < in.ndim; ++d) {
    if (reduce_axis[d]) {
      p.red_shape[nr] = in.shape[d];
      p.red_stride[nr] = in.stride[d];
      ++nr;
    } else {
      p.kept_shape[nk] = in.shape[d];
      p.kept_stride[nk] = in.stride[d];
      ++nk;
    }
  }
  p.n_kept = collapse_dims(nk, p.kept_shape, p.kept_stride);
  p.n_red = collapse_dims(nr, p.red_shape, p.red_stride);
  p.kept_total = 1;
  for (int i = 0; i < p.n_kept; ++i) p.kept_total *= p.kept_shape[i];
  p.red_total = 1;
  for (int i = 0; i < p.n_red; ++i) p.red_total *= p.red_shape[i];
  if (p.kept_total == 0) return BB_OK;
  BB_CUDA_OK(cudaMemsetAsync(scratch, 0, p.kept_total * sizeof(double), stream));
  if (p.red_total > 0) {
    // Which side owns the unit stride decides the thread mapping (coalescing).
    int64_t min_kept = INT64_MAX, min_red = INT64_MAX;
    for (int i = 0; i < p.n_kept; ++i) min_kept = std::min(min_kept, std::abs(p.kept_stride[i]));
    for (int i = 0; i < p.n_red; ++i) min_red = std::min(min_red, std::abs(p.red_stride[i]));
    if (min_kept <= min_red) {
      p.kw = std::min(256, pow2_ceil(p.kept_total));
      p.rw = 256 / p.kw;
    } else {
      p.rw = std::min(256, pow2_ceil(p.red_total));
      p.kw = 256 / p.rw;
    }
    const int64_t kept_blocks = (p.kept_total + p.kw - 1) / p.kw;
    const int64_t target = static_cast<int64_t>(std::max(1, device_sm_count())) * 8;
    int64_t splits = std::max<int64_t>(1, target / kept_blocks);
    const int64_t min_per_split = static_cast<int64_t>(p.rw) * 32;
    splits = std::min(splits, std::max<int64_t>(1, p.red_total / min_per_split));
    splits = std::min<int64_t>(splits, 65535);
    p.red_per_split = (p.red_total + splits - 1) / splits;
    splits = (p.red_total + p.red_per_split - 1) / p.red_per_split;
    if (kept_blocks > 2147483647LL) {
      set_error("reduce_sum: too many kept elements");
      return BB_ERR_UNSUPPORTED;
    }
    dim3 grid(static_cast<unsigned>(kept_blocks), static_cast<unsigned>(splits));
    reduce_sum_kernel<<<grid, 256, 0, stream>>>(p);
    BB_CHECK_LAUNCH("reduce_sum_kernel");
  }
  return launch_f64_to_f32(p.scratch, out.ptr, p.kept_total, stream);
}

// ============================ batched strided GEMM ==============================

struct GemmParams {
  const float* A;
  const float* B;
  float* C;               // [batch][M][N] (splits == 1) or partial [split][batch][M][N]
  int64_t M, N, K, batch;
  int64_t sAb, sAm, sAk, sBb, sBk, sBn;
  int64_t k_per_split;
  int splits;
  int m_on_x;             // M tiles on grid.x (else N tiles)
};

constexpr int kGemmTile = 64;
constexpr int kGemmK = 16;

template <bool A_K_CONTIG, bool B_N_CONTIG>
__global__ void __launch_bounds__(256) gemm_splitk_kernel(const GemmParams p) {
  __shared__ __align__(16) float As[kGemmK][kGemmTile + 4];
  __shared__ __align__(16) float Bs[kGemmK][kGemmTile + 4];
  const int t = threadIdx.x;
  const int tx = t % 16, ty = t / 16;
  // the longer tile axis rides on grid.x (2^31 - 1 blocks), the other on grid.y (65535)
  const int64_t m0 = static_cast<int64_t>(p.m_on_x ? blockIdx.x : blockIdx.y) * kGemmTile;
  const int64_t n0 = static_cast<int64_t>(p.m_on_x ? blockIdx.y : blockIdx.x) * kGemmTile;
  const int64_t bz = blockIdx.z;
  const int64_t split = bz / p.batch, batch = bz % p.batch;
  const int64_t k_begin = split * p.k_per_split;
  const int64_t k_end = min(k_begin + p.k_per_split, p.K);
  const float* A = p.A + batch * p.sAb;
  const float* B = p.B + batch * p.sBb;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int64_t k0 = k_begin; k0 < k_end; k0 += kGemmK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int kk, mm;
      if (A_K_CONTIG) { kk = t % 16; mm = t / 16 + 16 * i; }
      else { mm = t % 64; kk = t / 64 + 4 * i; }
      const int64_t m = m0 + mm, k = k0 + kk;
      As[kk][mm] = (m < p.M && k < k_end) ? __ldg(A + m * p.sAm + k * p.sAk) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int kk, nn;
      if (B_N_CONTIG) { nn = t % 64; kk = t / 64 + 4 * i; }
      else { kk = t % 16; nn = t / 16 + 16 * i; }
      const int64_t n = n0 + nn, k = k0 + kk;
      Bs[kk][nn] = (n < p.N && k < k_end) ? __ldg(B + k * p.sBk + n * p.sBn) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kGemmK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* C = p.C + (split * p.batch + batch) * p.M * p.N;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t n = n0 + tx * 4 + j;
      if (n < p.N) C[m * p.N + n] = acc[i][j];
    }
  }
}

__global__ void splitk_reduce_kernel(const float* partial, float* out, int64_t n, int splits) {
  const int64_t step = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += step) {
    double acc = 0.0;
    for (int s = 0; s < splits; ++s) acc += static_cast<double>(partial[s * n + i]);
    out[i] = static_cast<float>(acc);
  }
}

static void gemm_plan(int64_t M, int64_t N, int64_t K, int64_t batch, int* splits,
                      int64_t* k_per_split) {
  if (K <= 0 || M <= 0 || N <= 0 || batch <= 0) {
    *splits = 1;
    *k_per_split = kGemmK;
    return;
  }
  const int64_t tiles = ((M + kGemmTile - 1) / kGemmTile) * ((N + kGemmTile - 1) / kGemmTile) * batch;
  const int64_t target = static_cast<int64_t>(std::max(1, device_sm_count())) * 4;
  int64_t s = std::max<int64_t>(1, target / std::max<int64_t>(1, tiles));
  s = std::min(s, std::max<int64_t>(1, K / 256));
  s = std::min<int64_t>(s, 65535 / std::max<int64_t>(1, batch));
  s = std::max<int64_t>(1, s);
  int64_t kps = (K + s - 1) / s;
  kps = (kps + kGemmK - 1) / kGemmK * kGemmK;
  s = std::max<int64_t>(1, (K + kps - 1) / kps);
  *splits = static_cast<int>(s);
  *k_per_split = kps;
}

int64_t gemm_workspace_bytes(int64_t M, int64_t N, int64_t K, int64_t batch) {
  int splits;
  int64_t kps;
  gemm_plan(M, N, K, batch, &splits, &kps);
  if (splits <= 1) return 0;
  return align_up(static_cast<int64_t>(splits) * batch * M * N * 4, 256);
}

// C[b,m,n] = sum_k A[b,m,k] B[b,k,n]; A/B given by element strides, C dense.
int launch_gemm(const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K,
                int64_t batch, int64_t sAb, int64_t sAm, int64_t sAk, int64_t sBb, int64_t sBk,
                int64_t sBn, void* workspace, cudaStream_t stream) {
  if (M == 0 || N == 0 || batch == 0) return BB_OK;
  if (K == 0) return launch_fill(C, batch * M * N, 0.f, stream);
  GemmParams p;
  p.A = A; p.B = B; p.M = M; p.N = N; p.K = K; p.batch = batch;
  p.sAb = sAb; p.sAm = sAm; p.sAk = sAk; p.sBb = sBb; p.sBk = sBk; p.sBn = sBn;
  gemm_plan(M, N, K, batch, &p.splits, &p.k_per_split);
  p.C = (p.splits > 1) ? static_cast<float*>(workspace) : C;
  const int64_t gx = (N + kGemmTile - 1) / kGemmTile, gy = (M + kGemmTile - 1) / kGemmTile;
  const int64_t gz = batch * p.splits;
  p.m_on_x = gy > gx ? 1 : 0;
  if (std::min(gx, gy) > 65535 || std::max(gx, gy) > 2147483647LL || gz > 65535) {
    set_error("gemm: grid too large (M tiles %lld, N tiles %lld, batch*splits %lld)",
              static_cast<long long>(gy), static_cast<long long>(gx), static_cast<long long>(gz));
    return BB_ERR_UNSUPPORTED;
  }
  dim3 grid(static_cast<unsigned>(p.m_on_x ? gy : gx), static_cast<unsigned>(p.m_on_x ? gx : gy),
            static_cast<unsigned>(gz));
  const bool a_k = (sAk == 1), b_n = (sBn == 1);
  if (a_k && b_n) gemm_splitk_kernel<true, true><<<grid, 256, 0, stream>>>(p);
  else if (a_k) gemm_splitk_kernel<true, false><<<grid, 256, 0, stream>>>(p);
  else if (b_n) gemm_splitk_kernel<false, true><<<grid, 256, 0, stream>>>(p);
  else gemm_splitk_kernel<false, false><<<grid, 256, 0, stream>>>(p);
  BB_CHECK_LAUNCH("gemm_splitk_kernel");
  if (p.splits > 1) {
    const int64_t n = batch * M * N;
    splitk_reduce_kernel<<<grid_for(n, 256), 256, 0, stream>>>(p.C, C, n, p.splits);
    BB_CHECK_LAUNCH("splitk_reduce_kernel");
  }
  return BB_OK;
}

// ---- tall-skinny matrix-vector contractions (HBM-bound: X is read once) ---------------------------
// gemv_cols: out[d] = sum_n X[n, d] y[n]   (plan of dot(X.T, y): contraction over the data axis)
// gemv_rows: out[n] = sum_d X[n, d] w[d]   (plan of dot(X, w): the linear predictor)
// The split-K GEMM tiles waste 63/64 of their work on these shapes.
namespace {

constexpr int kGemvThreads = 256;
constexpr int kGemvMaxColBlocks = 4;          // gemv_cols: d <= 4 * 256 * 4 = 4096 in the vector path

// X row-major [n, d], d % 4 == 0, 16-byte aligned rows.  Thread = (row lane, float4 column): fp32
// partial sums over this CTA's row range, combined in float64 across row lanes, CTAs in the finalize.
__global__ void __launch_bounds__(kGemvThreads)
gemv_cols_kernel(const float* __restrict__ x, const float* __restrict__ y, int64_t n, int d,
                 double* __restrict__ partial) {
  __shared__ double red[kGemvThreads * 4];
  const int vec_cols = d >> 2;
  const int cols_here = min(vec_cols, kGemvThreads);
  const int rows_per_pass = kGemvThreads / cols_here;
  const int tid = threadIdx.x;
  const int r_lane = tid / cols_here, c_lane = tid - r_lane * cols_here;
  const bool active = r_lane < rows_per_pass;
  const int64_t row_begin = n * blockIdx.x / gridDim.x, row_end = n * (blockIdx.x + 1) / gridDim.x;
  const int col_blocks = (vec_cols + kGemvThreads - 1) / kGemvThreads;
  for (int cb = 0; cb < col_blocks; ++cb) {
    const int c4 = cb * kGemvThreads + c_lane;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active && c4 < vec_cols) {
      const float4* col = reinterpret_cast<const float4*>(x) + c4;
      for (int64_t r = row_begin + r_lane; r < row_end; r += rows_per_pass) {
        const float4 v = __ldg(col + r * vec_cols);
        const float w = __ldg(y + r);
        acc.x = fmaf(v.x, w, acc.x); acc.y = fmaf(v.y, w, acc.y);
        acc.z = fmaf(v.z, w, acc.z); acc.w = fmaf(v.w, w, acc.w);
      }
    }
    red[tid * 4 + 0] = acc.x; red[tid * 4 + 1] = acc.y; red[tid * 4 + 2] = acc.z; red[tid * 4 + 3] = acc.w;
    __syncthreads();
    if (r_lane == 0 && c4 < vec_cols) {
      double t0 = 0, t1 = 0, t2 = 0, t3 = 0;
      for (int rl = 0; rl < rows_per_pass; ++rl) {
        const double* src = red + (rl * cols_here + c_lane) * 4;
        t0 += src[0]; t1 += src[1]; t2 += src[2]; t3 += src[3];
      }
      double* dst = partial + static_cast<int64_t>(blockIdx.x) * d + c4 * 4;
      dst[0] = t0; dst[1] = t1; dst[2] = t2; dst[3] = t3;
    }
    __syncthreads();
  }
}

__global__ void gemv_cols_finalize_kernel(const double* __restrict__ partial, int n_ctas, int d,
                                          float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d) return;
  double acc = 0.0;
  for (int b = 0; b < n_ctas; ++b) acc += partial[static_cast<int64_t>(b) * d + c];
  out[c] = static_cast<float>(acc);
}

// one warp per row; float4 loads when d % 4 == 0 and the rows are 16-byte aligned
__global__ void __launch_bounds__(kGemvThreads)
gemv_rows_kernel(const float* __restrict__ x, const float* __restrict__ w, int64_t n, int d, int vec,
                 float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int64_t n_warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t r = warp; r < n; r += n_warps) {
    const float* row = x + r * d;
    float acc = 0.f;
    if (vec) {
      const float4* r4 = reinterpret_cast<const float4*>(row);
      const float4* w4 = reinterpret_cast<const float4*>(w);
      for (int c = lane; c < (d >> 2); c += 32) {
        const float4 a = __ldg(r4 + c), b = __ldg(w4 + c);
        acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
      }
    } else {
      for (int c = lane; c < d; c += 32) acc = fmaf(__ldg(row + c), __ldg(w + c), acc);
    }
    for (int off = 16; off >= 1; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) out[r] = acc;
  }
}

int gemv_cols_grid(int64_t n) {
  int sms = device_sm_count();
  if (sms <= 0) sms = 148;
  return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(static_cast<int64_t>(sms) * 4, (n + 63) / 64)));
}

}  // namespace

bool gemv_cols_supported(int64_t n, int64_t d, const void* x) {
  return n > 0 && d >= 4 && d % 4 == 0 && d <= kGemvMaxColBlocks * kGemvThreads * 4 &&
         reinterpret_cast<uintptr_t>(x) % 16 == 0;
}

int64_t gemv_cols_workspace(int64_t n, int64_t d) { return static_cast<int64_t>(gemv_cols_grid(n)) * d * 8 + 256; }

int launch_gemv_cols(const float* x, const float* y, int64_t n, int d, float* out, void* workspace,
                     cudaStream_t stream) {
  double* partial = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~static_cast<uintptr_t>(255));
  const int grid = gemv_cols_grid(n);
  gemv_cols_kernel<<<grid, kGemvThreads, 0, stream>>>(x, y, n, d, partial);
  BB_CHECK_LAUNCH("gemv_cols_kernel");
  gemv_cols_finalize_kernel<<<(d + 255) / 256, 256, 0, stream>>>(partial, grid, d, out);
  BB_CHECK_LAUNCH("gemv_cols_finalize_kernel");
  return BB_OK;
}

int launch_gemv_rows(const float* x, const float* w, int64_t n, int d, float* out, cudaStream_t stream) {
  if (n == 0) return BB_OK;
  int sms = device_sm_count();
  if (sms <= 0) sms = 148;
  const int vec = (d % 4 == 0 && reinterpret_cast<uintptr_t>(x) % 16 == 0 && reinterpret_cast<uintptr_t>(w) % 16 == 0) ? 1 : 0;
  const int64_t blocks = std::min<int64_t>((n + 7) / 8, static_cast<int64_t>(sms) * 16);
  gemv_rows_kernel<<<static_cast<int>(blocks), kGemvThreads, 0, stream>>>(x, w, n, d, vec, out);
  BB_CHECK_LAUNCH("gemv_rows_kernel");
  return BB_OK;
}

}  // namespace bb
