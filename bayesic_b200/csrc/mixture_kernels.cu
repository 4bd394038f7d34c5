// Mixture-responsibility kernels.
//
// logsoftmax_rows: log r[n,k] = logits[n,k] - logsumexp_k logits[n,:], the K-way
// responsibility pass of a mixture VMP step.  The reference can only spell it as
//   Lg + (-1 * log(sum(exp(Lg), axis=1)))            (bayesic/algebra.py:1435-1448)
// = three elementwise passes plus a reduction, unstabilised (overflows at logits ~ +89 in
// float32).  Here: ONE pass, one warp per row, 128-bit loads, the row kept in registers,
// max-subtracted, warp-shuffle reductions; the dominant component is excluded from the sum and
// re-added through log1p so log r of the winning component keeps full relative precision.
// Algorithmic traffic: 4K bytes in + 4K bytes out + 4 bytes (lse) per row.
//
// weighted_stats (SIMT): Nk, sum_n r x, sum_n r x x^T in one pass over (R, X) without the
// K x D x N intermediate the reference plan materialises
//   _tensordot(_mul(_dimshuffle(R,1,'x',0), _dimshuffle(X,'x',1,0)), X, [2],[0])
// (algebra.py:741-765 + 1297-1306).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace bb {

namespace {

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// K = 128 * ITERS floats per row, row held in registers (ITERS float4 per lane).
// kResp: write the responsibilities exp(logit - lse) instead of their logarithm.
template <int ITERS, bool kResp>
__global__ void __launch_bounds__(256)
logsoftmax_rows_vec_kernel(const float* __restrict__ logits, int64_t n, float* __restrict__ log_resp,
                           float* __restrict__ lse_out, double* __restrict__ sum_lse) {
  constexpr int K = 128 * ITERS;
  const int lane = threadIdx.x & 31;
  const int warp_in_block = threadIdx.x >> 5;
  const int64_t warps_total = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
  double lse_acc = 0.0;
  for (int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + warp_in_block; row < n;
       row += warps_total) {
    const float4* src = reinterpret_cast<const float4*>(logits + row * K);
    float4 v[ITERS];
#pragma unroll
    for (int i = 0; i < ITERS; ++i) v[i] = __ldcs(src + i * 32 + lane);
    // row maximum and the position (lane-local slot) that owns it
    float m = -INFINITY;
    int slot = 0;
#pragma unroll
    for (int i = 0; i < ITERS; ++i) {
      const float e[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (e[j] > m) { m = e[j]; slot = i * 4 + j; }
    }
    const float row_max = warp_max(m);
    const unsigned owners = __ballot_sync(0xffffffffu, m == row_max);
    const bool i_own = (lane == (__ffs(owners) - 1));
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ITERS; ++i) {
      const float e[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float t = __expf(e[j] - row_max);
        s += (i_own && slot == i * 4 + j) ? 0.f : t;
      }
    }
    s = warp_sum(s);                       // sum over all but the winning component
    const float log_sum = log1pf(s);       // log(1 + rest), accurate when rest is tiny
    float4* dst = reinterpret_cast<float4*>(log_resp + row * K);
#pragma unroll
    for (int i = 0; i < ITERS; ++i) {
      float4 o;
      o.x = (v[i].x - row_max) - log_sum;
      o.y = (v[i].y - row_max) - log_sum;
      o.z = (v[i].z - row_max) - log_sum;
      o.w = (v[i].w - row_max) - log_sum;
      if (kResp) {
        o.x = __expf(o.x); o.y = __expf(o.y); o.z = __expf(o.z); o.w = __expf(o.w);
      }
      __stcs(dst + i * 32 + lane, o);
    }
    const float lse = row_max + log_sum;
    if (lane == 0) {
      if (lse_out != nullptr) lse_out[row] = lse;
      lse_acc += static_cast<double>(lse);
    }
  }
  if (sum_lse != nullptr) {
    __shared__ double block_acc[8];
    if (lane == 0) block_acc[warp_in_block] = lse_acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < (blockDim.x >> 5); ++w) t += block_acc[w];
      atomicAdd(sum_lse, t);
    }
  }
}

// Responsibilities r = exp(logit - lse) written ALREADY SPLIT into the BF16 (b1 | b2) operand tiles the
// weighted-statistics kernel feeds to the tensor core (weighted_pairs_sm100.cu, kMode 2): per slice of 256
// components and per 16-row stage one 16 KB image [b1 | b2], each [64-component block (4)][8-row group (2)]
// [8 rows x 128 B, 16-byte chunks XOR-swizzled with the row].  Same single pass as the log-softmax: K = 256 ITERS2
// floats per row in registers, lane l holds components 128 i + 4 l .. + 3.  Rows n .. 32 ceil(n / 32) - 1 are written as
// zeros (they are operand rows of the last MMAs).
template <int ITERS>      // K = 128 * ITERS, ITERS even
__global__ void __launch_bounds__(256)
softmax_rows_split_kernel(const float* __restrict__ logits, int64_t n, uint8_t* __restrict__ rsplit,
                          float* __restrict__ lse_out, double* __restrict__ sum_lse) {
  constexpr int K = 128 * ITERS;
  const int lane = threadIdx.x & 31;
  const int warp_in_block = threadIdx.x >> 5;
  const int64_t warps_total = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
  const int64_t stages = 2 * ((n + 31) / 32);       // = weighted_pairs_split_stages(n): an even number of 16-row images
  const int64_t n_pad = stages * 16;
  double lse_acc = 0.0;
  for (int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + warp_in_block; row < n_pad;
       row += warps_total) {
    float4 v[ITERS];
    float lse = 0.f;
    if (row < n) {
      const float4* src = reinterpret_cast<const float4*>(logits + row * K);
#pragma unroll
      for (int i = 0; i < ITERS; ++i) v[i] = __ldcs(src + i * 32 + lane);
      float m = -INFINITY;
#pragma unroll
      for (int i = 0; i < ITERS; ++i) m = fmaxf(m, fmaxf(fmaxf(v[i].x, v[i].y), fmaxf(v[i].z, v[i].w)));
      const float row_max = warp_max(m);
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < ITERS; ++i) {
        v[i].x = __expf(v[i].x - row_max); v[i].y = __expf(v[i].y - row_max);
        v[i].z = __expf(v[i].z - row_max); v[i].w = __expf(v[i].w - row_max);
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      }
      s = warp_sum(s);
      const float inv = 1.f / s;
#pragma unroll
      for (int i = 0; i < ITERS; ++i) {
        v[i].x *= inv; v[i].y *= inv; v[i].z *= inv; v[i].w *= inv;
      }
      lse = row_max + logf(s);
    } else {
#pragma unroll
      for (int i = 0; i < ITERS; ++i) v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const int64_t stage = row >> 4;
    const int kk = static_cast<int>(row & 15);
    const uint32_t in_tile = (lane >> 4) * 2048 + (kk >> 3) * 1024 + (kk & 7) * 128 +
                             ((((lane & 15) >> 1) ^ (kk & 7)) << 4) + (lane & 1) * 8;
#pragma unroll
    for (int i = 0; i < ITERS; ++i) {
      __nv_bfloat162 p0 = __floats2bfloat162_rn(v[i].x, v[i].y);
      __nv_bfloat162 p1 = __floats2bfloat162_rn(v[i].z, v[i].w);
      const uint32_t h0 = *reinterpret_cast<uint32_t*>(&p0), h1 = *reinterpret_cast<uint32_t*>(&p1);
      __nv_bfloat162 q0 = __floats2bfloat162_rn(v[i].x - __uint_as_float(h0 << 16), v[i].y - __uint_as_float(h0 & 0xFFFF0000u));
      __nv_bfloat162 q1 = __floats2bfloat162_rn(v[i].z - __uint_as_float(h1 << 16), v[i].w - __uint_as_float(h1 & 0xFFFF0000u));
      uint8_t* tile = rsplit + ((static_cast<int64_t>(i >> 1) * stages + stage) << 14) + (i & 1) * 4096 + in_tile;
      *reinterpret_cast<uint2*>(tile) = make_uint2(h0, h1);
      *reinterpret_cast<uint2*>(tile + 8192) = make_uint2(*reinterpret_cast<uint32_t*>(&q0), *reinterpret_cast<uint32_t*>(&q1));
    }
    if (lane == 0 && row < n) {
      if (lse_out != nullptr) lse_out[row] = lse;
      lse_acc += static_cast<double>(lse);
    }
  }
  if (sum_lse != nullptr) {
    __shared__ double block_acc[8];
    if (lane == 0) block_acc[warp_in_block] = lse_acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < (blockDim.x >> 5); ++w) t += block_acc[w];
      atomicAdd(sum_lse, t);
    }
  }
}

// Any K: three sweeps over the row (it stays in L1/L2 after the first).
template <bool kResp>
__global__ void __launch_bounds__(256)
logsoftmax_rows_any_kernel(const float* __restrict__ logits, int64_t n, int k,
                           float* __restrict__ log_resp, float* __restrict__ lse_out,
                           double* __restrict__ sum_lse) {
  const int lane = threadIdx.x & 31;
  const int warp_in_block = threadIdx.x >> 5;
  const int64_t warps_total = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
  double lse_acc = 0.0;
  for (int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + warp_in_block; row < n;
       row += warps_total) {
    const float* src = logits + row * k;
    float m = -INFINITY;
    int slot = -1;
    for (int j = lane; j < k; j += 32) {
      const float e = src[j];
      if (e > m) { m = e; slot = j; }
    }
    const float row_max = warp_max(m);
    const unsigned owners = __ballot_sync(0xffffffffu, m == row_max && slot >= 0);
    const bool i_own = (lane == (__ffs(owners) - 1));
    float s = 0.f;
    for (int j = lane; j < k; j += 32) {
      const float t = __expf(src[j] - row_max);
      s += (i_own && slot == j) ? 0.f : t;
    }
    s = warp_sum(s);
    const float log_sum = log1pf(s);
    float* dst = log_resp + row * k;
    for (int j = lane; j < k; j += 32) {
      const float o = (src[j] - row_max) - log_sum;
      dst[j] = kResp ? __expf(o) : o;
    }
    const float lse = row_max + log_sum;
    if (lane == 0) {
      if (lse_out != nullptr) lse_out[row] = lse;
      lse_acc += static_cast<double>(lse);
    }
  }
  if (sum_lse != nullptr) {
    __shared__ double block_acc[8];
    if (lane == 0) block_acc[warp_in_block] = lse_acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < (blockDim.x >> 5); ++w) t += block_acc[w];
      atomicAdd(sum_lse, t);
    }
  }
}

}  // namespace

int launch_logsoftmax_rows(const float* logits, int64_t n, int k, float* log_resp, float* lse,
                           double* sum_lse, bool responsibilities, cudaStream_t stream) {
  if (k <= 0) {
    set_error("logsoftmax_rows: k must be positive");
    return BB_ERR_INVALID;
  }
  if (sum_lse != nullptr) BB_CUDA_OK(cudaMemsetAsync(sum_lse, 0, sizeof(double), stream));
  if (n == 0) return BB_OK;
  const int threads = 256, warps = threads / 32;
  const int64_t want = (n + warps - 1) / warps;
  const int64_t cap = static_cast<int64_t>(std::max(1, device_sm_count())) * 8;
  const int grid = static_cast<int>(std::max<int64_t>(1, std::min(want, cap)));
  const bool aligned = (reinterpret_cast<uintptr_t>(logits) % 16 == 0) &&
                       (reinterpret_cast<uintptr_t>(log_resp) % 16 == 0);
  const int iters = (k % 128 == 0) ? k / 128 : 0;
  if (aligned && iters >= 1 && iters <= 8) {
    switch (iters) {
#define BB_LSE_CASE(I)                                                                                    \
  case I:                                                                                                 \
    if (responsibilities)                                                                                 \
      logsoftmax_rows_vec_kernel<I, true><<<grid, threads, 0, stream>>>(logits, n, log_resp, lse, sum_lse);  \
    else                                                                                                  \
      logsoftmax_rows_vec_kernel<I, false><<<grid, threads, 0, stream>>>(logits, n, log_resp, lse, sum_lse); \
    break;
      BB_LSE_CASE(1) BB_LSE_CASE(2) BB_LSE_CASE(3) BB_LSE_CASE(4)
      BB_LSE_CASE(5) BB_LSE_CASE(6) BB_LSE_CASE(7) BB_LSE_CASE(8)
#undef BB_LSE_CASE
    }
    BB_CHECK_LAUNCH("logsoftmax_rows_vec_kernel");
  } else {
    if (responsibilities)
      logsoftmax_rows_any_kernel<true><<<grid, threads, 0, stream>>>(logits, n, k, log_resp, lse, sum_lse);
    else
      logsoftmax_rows_any_kernel<false><<<grid, threads, 0, stream>>>(logits, n, k, log_resp, lse, sum_lse);
    BB_CHECK_LAUNCH("logsoftmax_rows_any_kernel");
  }
  return BB_OK;
}

int launch_softmax_rows_split(const float* logits, int64_t n, int k, void* rsplit, float* lse, double* sum_lse,
                              cudaStream_t stream) {
  if (k < 256 || k % 256 != 0 || k > 1024 || reinterpret_cast<uintptr_t>(logits) % 16 != 0 ||
      reinterpret_cast<uintptr_t>(rsplit) % 16 != 0) {
    set_error("softmax_rows_split: needs k in {256, 512, 768, 1024} and 16-byte aligned buffers (got k=%d)", k);
    return BB_ERR_UNSUPPORTED;
  }
  if (sum_lse != nullptr) BB_CUDA_OK(cudaMemsetAsync(sum_lse, 0, sizeof(double), stream));
  if (n == 0) return BB_OK;
  const int threads = 256, warps = threads / 32;
  const int64_t want = ((n + 31) / 32 * 32 + warps - 1) / warps;
  const int64_t cap = static_cast<int64_t>(std::max(1, device_sm_count())) * 8;
  const int grid = static_cast<int>(std::max<int64_t>(1, std::min(want, cap)));
  uint8_t* out = static_cast<uint8_t*>(rsplit);
  switch (k / 128) {
    case 2: softmax_rows_split_kernel<2><<<grid, threads, 0, stream>>>(logits, n, out, lse, sum_lse); break;
    case 4: softmax_rows_split_kernel<4><<<grid, threads, 0, stream>>>(logits, n, out, lse, sum_lse); break;
    case 6: softmax_rows_split_kernel<6><<<grid, threads, 0, stream>>>(logits, n, out, lse, sum_lse); break;
    default: softmax_rows_split_kernel<8><<<grid, threads, 0, stream>>>(logits, n, out, lse, sum_lse); break;
  }
  BB_CHECK_LAUNCH("softmax_rows_split_kernel");
  return BB_OK;
}

// ============================ weighted statistics (SIMT) =========================

namespace {

constexpr int kWsFeat = 64;     // padded feature extent
constexpr int kWsComp = 4;      // components per CTA
constexpr int kWsRows = 32;     // rows per smem tile
constexpr int kWsFlushTiles = 32;  // fp32 -> fp64 every 1024 rows

__global__ void __launch_bounds__(256, 1)
weighted_stats_simt_kernel(const float* __restrict__ x, const float* __restrict__ r, int64_t n,
                           int d, int k, int64_t rows_per_block, double* __restrict__ nk,
                           double* __restrict__ sum_rx, double* __restrict__ sum_rxx) {
  __shared__ __align__(16) float xs[kWsRows][kWsFeat];
  __shared__ float rs[kWsRows][kWsComp];
  const int t = threadIdx.x;
  const int e0 = (t % 16) * 4, d0 = (t / 16) * 4;
  const int k0 = blockIdx.x * kWsComp;
  const int64_t row_begin = static_cast<int64_t>(blockIdx.y) * rows_per_block;
  const int64_t row_end = min(row_begin + rows_per_block, n);

  float acc[kWsComp][4][4];
  double dacc[kWsComp][4][4];
  float rx_acc[kWsComp] = {0.f, 0.f, 0.f, 0.f};
  double rx_dacc[kWsComp] = {0.0, 0.0, 0.0, 0.0};
  float nk_acc = 0.f;
  double nk_dacc = 0.0;
#pragma unroll
  for (int c = 0; c < kWsComp; ++c)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) { acc[c][i][j] = 0.f; dacc[c][i][j] = 0.0; }

  int tiles_since_flush = 0;
  for (int64_t row0 = row_begin; row0 < row_end; row0 += kWsRows) {
    for (int idx = t; idx < kWsRows * kWsFeat; idx += 256) {
      const int rr = idx / kWsFeat, cc = idx % kWsFeat;
      const int64_t row = row0 + rr;
      xs[rr][cc] = (row < row_end && cc < d) ? __ldg(x + row * d + cc) : 0.f;
    }
    if (t < kWsRows * kWsComp) {
      const int rr = t / kWsComp, cc = t % kWsComp;
      const int64_t row = row0 + rr;
      rs[rr][cc] = (row < row_end && k0 + cc < k) ? __ldg(r + row * k + k0 + cc) : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int rr = 0; rr < kWsRows; ++rr) {
      const float4 xd = *reinterpret_cast<const float4*>(&xs[rr][d0]);
      const float4 xe = *reinterpret_cast<const float4*>(&xs[rr][e0]);
      const float xdv[4] = {xd.x, xd.y, xd.z, xd.w};
      const float xev[4] = {xe.x, xe.y, xe.z, xe.w};
#pragma unroll
      for (int c = 0; c < kWsComp; ++c) {
        const float w = rs[rr][c];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float a = w * xdv[i];
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[c][i][j] = fmaf(a, xev[j], acc[c][i][j]);
        }
      }
      if (t < kWsFeat) {
        const float xv = xs[rr][t];
#pragma unroll
        for (int c = 0; c < kWsComp; ++c) rx_acc[c] = fmaf(rs[rr][c], xv, rx_acc[c]);
      } else if (t < kWsFeat + kWsComp) {
        nk_acc += rs[rr][t - kWsFeat];
      }
    }
    __syncthreads();
    if (++tiles_since_flush == kWsFlushTiles) {
      tiles_since_flush = 0;
#pragma unroll
      for (int c = 0; c < kWsComp; ++c) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) { dacc[c][i][j] += acc[c][i][j]; acc[c][i][j] = 0.f; }
        rx_dacc[c] += rx_acc[c];
        rx_acc[c] = 0.f;
      }
      nk_dacc += nk_acc;
      nk_acc = 0.f;
    }
  }
#pragma unroll
  for (int c = 0; c < kWsComp; ++c) {
    const int kk = k0 + c;
    if (kk >= k) continue;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int dd = d0 + i, ee = e0 + j;
        if (dd < d && ee < d)
          atomicAdd(sum_rxx + (static_cast<int64_t>(kk) * d + dd) * d + ee,
                    dacc[c][i][j] + static_cast<double>(acc[c][i][j]));
      }
    if (t < d && sum_rx != nullptr)
      atomicAdd(sum_rx + static_cast<int64_t>(kk) * d + t, rx_dacc[c] + static_cast<double>(rx_acc[c]));
  }
  if (nk != nullptr && t >= kWsFeat && t < kWsFeat + kWsComp && k0 + t - kWsFeat < k)
    atomicAdd(nk + k0 + t - kWsFeat, nk_dacc + static_cast<double>(nk_acc));
}

}  // namespace

int64_t weighted_stats_workspace(int64_t, int, int) { return 0; }

int launch_weighted_stats(const float* x, const float* r, int64_t n, int d, int k, double* nk,
                          double* sum_rx, double* sum_rxx, void*, int64_t, cudaStream_t stream) {
  if (d < 1 || d > kWsFeat || k < 1) {
    set_error("weighted_stats: needs 1 <= d <= %d and k >= 1 (got d=%d k=%d)", kWsFeat, d, k);
    return BB_ERR_UNSUPPORTED;
  }
  if (nk != nullptr) BB_CUDA_OK(cudaMemsetAsync(nk, 0, sizeof(double) * k, stream));
  if (sum_rx != nullptr)
    BB_CUDA_OK(cudaMemsetAsync(sum_rx, 0, sizeof(double) * static_cast<int64_t>(k) * d, stream));
  BB_CUDA_OK(cudaMemsetAsync(sum_rxx, 0, sizeof(double) * static_cast<int64_t>(k) * d * d, stream));
  if (n == 0) return BB_OK;
  const int kblocks = (k + kWsComp - 1) / kWsComp;
  const int64_t target = static_cast<int64_t>(std::max(1, device_sm_count())) * 2;
  int64_t splits = std::max<int64_t>(1, target / kblocks);
  splits = std::min<int64_t>(splits, std::max<int64_t>(1, n / (kWsRows * 8)));
  splits = std::min<int64_t>(splits, 65535);
  int64_t rows_per_block = (n + splits - 1) / splits;
  rows_per_block = (rows_per_block + kWsRows - 1) / kWsRows * kWsRows;
  splits = (n + rows_per_block - 1) / rows_per_block;
  dim3 grid(kblocks, static_cast<unsigned>(splits));
  weighted_stats_simt_kernel<<<grid, 256, 0, stream>>>(x, r, n, d, k, rows_per_block, nk, sum_rx,
                                                       sum_rxx);
  BB_CHECK_LAUNCH("weighted_stats_simt_kernel");
  return BB_OK;
}


int64_t weighted_stats_auto_workspace(int64_t n, int d, int k) {
  int64_t need = 256;
  if (d >= 4 && d <= 64 && d % 4 == 0 && k % 4 == 0 && n > 0) need = std::max(need, weighted_tc_workspace(n, k));
  if (d >= 8 && d <= 64 && d % 8 == 0 && k >= 4 && k <= 4096 && k % 4 == 0 && n > 0)
    need = std::max(need, weighted_pairs_workspace(n, d, k));
  return need;
}

int launch_weighted_stats_auto(const float* x, const float* r, int64_t n, int d, int k, double* nk,
                               double* sum_rx, double* sum_rxx, void* workspace,
                               int64_t workspace_bytes, cudaStream_t stream) {
  static const bool force_simt = getenv("BB_WEIGHTED_SIMT") != nullptr;
  static const bool no_pairs = getenv("BB_WEIGHTED_NO_PAIRS") != nullptr;
  // R^T . (X (x) X) grouping on BF16 tcgen05 (weighted_pairs_sm100.cu): tensor-pipe bound
  if (!force_simt && !no_pairs && n >= 1024 && weighted_pairs_supported(n, d, k, x, r) && workspace != nullptr &&
      workspace_bytes >= weighted_pairs_workspace(n, d, k))
    return launch_weighted_pairs(x, r, nullptr, n, d, k, nk, sum_rx, sum_rxx, workspace, workspace_bytes, stream);
  // the tensor-core kernel pays off once there are enough rows to amortise its per-CTA setup
  if (!force_simt && n >= 1024 && weighted_tc_supported(n, d, k, x, r) && workspace != nullptr &&
      workspace_bytes >= weighted_tc_workspace(n, k))
    return launch_weighted_stats_tc(x, r, n, d, k, nk, sum_rx, sum_rxx, workspace, workspace_bytes, stream);
  return launch_weighted_stats(x, r, n, d, k, nk, sum_rx, sum_rxx, workspace, workspace_bytes, stream);
}

}  // namespace bb
