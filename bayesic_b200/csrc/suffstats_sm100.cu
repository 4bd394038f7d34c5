// One-pass Gaussian sufficient statistics  S1 = sum_n x_n,  S2 = sum_n x_n x_n^T
// over X[n, d] (float32, row-major, d <= 64) on sm_100a.
//
// What it replaces: the reference evaluates Sigma x x^T as the plan
//   _tensordot(_dimshuffle(X,1,0), X, [1],[0])        (bayesic/algebra.py:527-551, 1347-1351)
// i.e. one Theano/BLAS sgemm over the data axis, and Sigma x as _sum(X, 0) -- a second full
// pass (algebra.py:1290-1291).  These are the iid-summed statistics of
// ExpFamIndependentObservations (bayesic/distribution/base.py:328-332) for
// MultivariateNormal's (x, x x^T) (distribution/core.py:41-44).
//
// Design (HBM-bound: 256 B/row must stream once at ~6.5 TB/s, but 8 256 flop/row would need
// ~210 TFLOP/s -- more than the FP32 SIMT pipes have -- so the outer products go to tcgen05):
//
//   * persistent kernel, one CTA per SM, each owning a contiguous range of 128-row tiles;
//   * a TMA producer thread streams tiles (two 128x32 boxes, SWIZZLE_128B_ATOM_32B -- the
//     only shared-memory layout the tensor core accepts for an MN-major 32-bit operand;
//     measured with tests/cuda/tc_probe.cu: plain SWIZZLE_128B / no-swizzle MN-major TF32
//     operands silently read as zero) through a 6-stage mbarrier ring: 192 KB in flight per SM;
//   * error-compensated TF32 ("3xTF32" for the price of one MMA): x = hi + lo with
//     hi = x truncated to TF32 (exactly what the tensor core reads from a raw FP32 word) and
//     lo = x - hi.  Eight "split" warps write A = [hi ; lo] (M = 128 rows: 64 features of hi
//     stacked on 64 features of lo, K = data rows) into TMEM with tcgen05.st; the B operand
//     is the raw FP32 tile itself in shared memory (MN-major, read as hi by the hardware).
//     One M=128, N=64, K=8 tcgen05.mma per 8 data rows therefore yields
//         D[0:64]   += hi^T hi         D[64:128] += lo^T hi
//     and  S2 = hi^T hi + lo^T hi + (lo^T hi)^T  (+ O(2^-20) lo^T lo, dropped);
//   * FP32 accumulation in TMEM is drained every 512 rows into float64 registers by eight
//     epilogue warps (double-buffered accumulators, so the MMA never waits);
//   * Sigma x rides along for free in the split warps (they already touch every element);
//   * per-CTA float64 partials go to a workspace; a tiny finalize kernel adds them up in a
//     fixed order (deterministic) and applies the symmetrisation above.
//
// Algorithmic traffic: 4*d bytes per row, read once.  Nothing else touches HBM except
// 148 x 65.5 KB of partials.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "sm100_ptx.cuh"

namespace bb {

namespace {

// epilogue warps wait ~5 000 cycles per 512-row chunk; a backing-off wait (mbar_wait_sleep) was
// measured here and made no difference (0.772 ms either way), so they spin
#ifdef BB_SUFFSTATS_EPI_SLEEP
#define BB_SUFFSTATS_EPI_WAIT(bar, parity) ptx::mbar_wait_sleep(bar, parity, 64)
#else
#define BB_SUFFSTATS_EPI_WAIT(bar, parity) ptx::mbar_wait(bar, parity)
#endif

constexpr int kFeat = 64;          // padded feature extent (= MMA N, = half of MMA M)
constexpr int kTileRows = 128;     // data rows per pipeline stage
constexpr int kStages = 6;
// TMEM accumulators are drained to float64 every kFlushTiles tiles.  The tensor core truncates
// its fp32 accumulate, so the chunk length sets the (systematic) error: measured on B200 at
// N = 16 Mi, D = 64:  4 tiles (512 rows) 2.5e-6 relative, 0.79 ms/pass;  1 tile 7e-7, 0.93 ms.
// Both are far inside the 1e-4 parity bar; the default favours throughput.
#ifndef BB_SUFFSTATS_FLUSH_TILES
#define BB_SUFFSTATS_FLUSH_TILES 4
#endif
constexpr int kFlushTiles = BB_SUFFSTATS_FLUSH_TILES;
constexpr int kBoxCols = 32;       // one TMA box = 128 rows x 32 floats (one 128-byte swizzle span per row)
constexpr int kHalfBytes = kTileRows * kBoxCols * 4;   // 16 KB
constexpr int kStageBytes = 2 * kHalfBytes;            // 32 KB
constexpr int kKBlocks = kTileRows / 8;                // 16 MMAs (K = 8) per tile
constexpr int kSplitWarps = 8;     // warps 0-7: quadrant = w & 3, k-half = w >> 2
constexpr int kEpiWarps = 8;       // warps 8-15: quadrant = w & 3, column half = (w - 8) >> 2
constexpr int kTmaWarp = kSplitWarps + kEpiWarps;
constexpr int kMmaWarp = kTmaWarp + 1;
constexpr int kThreads = (kMmaWarp + 1) * 32;   // 576
constexpr int kTmemCols = 512;
constexpr int kTmemAcc = 0;        // 2 accumulators x 64 columns
constexpr int kTmemA = 128;        // 2 A buffers x 128 columns

struct __align__(1024) SmemLayout {
  uint8_t stage[kStages][kStageBytes];
  uint64_t full[kStages];
  uint64_t empty[kStages];
  uint64_t a_ready[2];
  uint64_t a_free[2];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint32_t tmem_base;
};

constexpr uint32_t kIdesc = ptx::make_idesc(128, kFeat, /*tf32*/ 2, /*A K-major*/ 0, /*B MN-major*/ 1);

__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}

// One split warp: for every tile, read its 64 rows x 32 features from the swizzled stage
// (lane = feature, conflict-free: the 32 lanes of a load cover one 128-byte row), split, and
// store 8 data rows at a time as 8 TMEM columns.
template <bool kIsLo>
__device__ __forceinline__ void split_warp_loop(SmemLayout& sm, uint32_t tmem, int q, int khalf,
                                                int lane, int my_tiles,
                                                double* __restrict__ partial_s1) {
  const int half = q & 1;             // which 32-feature column block
  // byte offset of (row j of an 8-row group, feature = lane) inside a stage: rows are 128 B,
  // 32-byte chunks XOR-swizzled with (row & 3)  (TMA SWIZZLE_128B_ATOM_32B)
  uint32_t off[8];
#pragma unroll
  for (int j = 0; j < 8; ++j)
    off[j] = half * kHalfBytes + khalf * (kKBlocks / 2) * 1024 + j * 128 +
             ((((lane >> 3) ^ j) & 3) << 5) + (lane & 7) * 4;
  const uint32_t lane_addr = tmem + (static_cast<uint32_t>(q * 32) << 16) + kTmemA +
                             khalf * (kKBlocks / 2) * 8;
  double s1 = 0.0;
  for (int i = 0; i < my_tiles; ++i) {
    const int s = i % kStages;
    const int b = i & 1;
    ptx::mbar_wait(&sm.full[s], (i / kStages) & 1);
    ptx::mbar_wait(&sm.a_free[b], ((i >> 1) & 1) ^ 1);
    ptx::tc_fence_after_sync();
    const uint32_t stage_addr = ptx::smem_u32(sm.stage[s]);
    const uint32_t a_addr = lane_addr + b * kTileRows;
    float s1_tile = 0.f;
#pragma unroll
    for (int kb = 0; kb < kKBlocks / 2; ++kb) {
      float x[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = lds_f32(stage_addr + off[j] + kb * 1024);
      uint32_t v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t hi_bits = __float_as_uint(x[j]) & 0xFFFFE000u;
        if (kIsLo) {
          // lo = x - hi exactly; round it to TF32 (nearest, on the magnitude bits) so the
          // tensor core's own truncation of A changes nothing
          const float lo = x[j] - __uint_as_float(hi_bits);
          v[j] = (__float_as_uint(lo) + 0x1000u) & 0xFFFFE000u;
        } else {
          v[j] = hi_bits;
          s1_tile += x[j];
        }
      }
      ptx::tmem_st_32x32b_x8(a_addr + kb * 8, v);
    }
    ptx::tmem_wait_st();
    ptx::tc_fence_before_sync();
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(&sm.a_ready[b]);
    if (!kIsLo) s1 += static_cast<double>(s1_tile);
  }
  if (!kIsLo)
    partial_s1[(static_cast<int64_t>(blockIdx.x) * 2 + khalf) * kFeat + half * 32 + lane] = s1;
}

__global__ void __launch_bounds__(kThreads, 1)
suffstats_tc_kernel(const __grid_constant__ CUtensorMap x_map, int64_t n_tiles,
                    double* __restrict__ partial_s2,   // [grid][128][64]
                    double* __restrict__ partial_s1,   // [grid][2][64]
                    unsigned int* __restrict__ fin_counter) {   // block counter of the finalize kernel's tail
  extern __shared__ uint8_t smem_raw[];
  if (blockIdx.x == 0 && threadIdx.x == 0) *fin_counter = 0u;
  SmemLayout& sm = *reinterpret_cast<SmemLayout*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t tile_begin = n_tiles * blockIdx.x / gridDim.x;
  const int64_t tile_end = n_tiles * (blockIdx.x + 1) / gridDim.x;
  const int my_tiles = static_cast<int>(tile_end - tile_begin);

  if (warp == kMmaWarp) {
    if (lane == 0) {
      for (int s = 0; s < kStages; ++s) {
        ptx::mbar_init(&sm.full[s], 1);
        ptx::mbar_init(&sm.empty[s], 1);
      }
      for (int b = 0; b < 2; ++b) {
        ptx::mbar_init(&sm.a_ready[b], kSplitWarps);     // one elected arrival per split warp
        ptx::mbar_init(&sm.a_free[b], 1);
        ptx::mbar_init(&sm.acc_full[b], 1);
        ptx::mbar_init(&sm.acc_empty[b], kEpiWarps);
      }
      ptx::fence_mbar_init();
    }
    __syncwarp();
    ptx::tmem_alloc(&sm.tmem_base, kTmemCols);
  } else if (warp == kTmaWarp && lane == 0) {
    ptx::prefetch_tensormap(&x_map);
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem = sm.tmem_base;

  if (warp == kTmaWarp) {
    // ---------------- TMA producer ----------------
    if (ptx::elect_one()) {
      for (int i = 0; i < my_tiles; ++i) {
        const int s = i % kStages;
        const uint32_t ph = (i / kStages) & 1;
        ptx::mbar_wait(&sm.empty[s], ph ^ 1);
        ptx::mbar_arrive_expect_tx(&sm.full[s], kStageBytes);
        const int32_t row0 = static_cast<int32_t>((tile_begin + i) * kTileRows);
        ptx::tma_load_2d(sm.stage[s], &x_map, &sm.full[s], 0, row0);
        ptx::tma_load_2d(sm.stage[s] + kHalfBytes, &x_map, &sm.full[s], kBoxCols, row0);
      }
    }
  } else if (warp == kMmaWarp) {
    // ---------------- MMA issuer ----------------
    if (ptx::elect_one()) {
      for (int i = 0; i < my_tiles; ++i) {
        const int s = i % kStages;
        const int b = i & 1;
        const int chunk = i / kFlushTiles;
        const int ab = chunk & 1;
        const bool first_in_chunk = (i % kFlushTiles) == 0;
        if (first_in_chunk) ptx::mbar_wait(&sm.acc_empty[ab], ((chunk >> 1) & 1) ^ 1);
        ptx::mbar_wait(&sm.full[s], (i / kStages) & 1);
        ptx::mbar_wait(&sm.a_ready[b], (i >> 1) & 1);
        ptx::tc_fence_after_sync();
        const uint32_t stage_addr = ptx::smem_u32(sm.stage[s]);
        const uint32_t d_tmem = tmem + kTmemAcc + ab * kFeat;
        const uint32_t a_tmem = tmem + kTmemA + b * kTileRows;
#pragma unroll
        for (int kb = 0; kb < kKBlocks; ++kb) {
          // B = rows [8kb, 8kb+8) of the tile: K = 8 rows of 128 B, N = 64 features in two
          // 32-wide blocks kHalfBytes apart (LBO); the 32-byte-atom swizzle repeats every 4 rows,
          // so the two 4-row groups of one MMA are 512 B apart (SBO).
          const uint64_t b_desc = ptx::make_smem_desc(stage_addr + kb * 1024, kHalfBytes, 512,
                                                      ptx::kLayoutSwizzle128B32BAtom);
          ptx::mma_tf32_ts(d_tmem, a_tmem + kb * 8, b_desc, kIdesc,
                           (first_in_chunk && kb == 0) ? 0u : 1u);
        }
        ptx::mma_commit(&sm.empty[s]);
        ptx::mma_commit(&sm.a_free[b]);
        if ((i % kFlushTiles) == kFlushTiles - 1 || i == my_tiles - 1)
          ptx::mma_commit(&sm.acc_full[ab]);
      }
    }
  } else if (warp < kSplitWarps) {
    // ---------------- split warps: A = [hi ; lo] into TMEM, and Sigma x ----------------
    // Warp w owns TMEM lanes [32 (w&3), +32): quadrants 0,1 hold hi of features 0-31 / 32-63,
    // quadrants 2,3 hold lo.  The two warps sharing a quadrant take rows 0-63 / 64-127 of a tile.
    const int q = warp & 3;
    if (q < 2) split_warp_loop<false>(sm, tmem, q, warp >> 2, lane, my_tiles, partial_s1);
    else split_warp_loop<true>(sm, tmem, q, warp >> 2, lane, my_tiles, partial_s1);
  } else {
    // ---------------- epilogue warps: TMEM fp32 chunks -> float64 registers ----------------
    const int q = warp & 3;
    const int chalf = (warp - kSplitWarps) >> 2;       // columns [32 chalf, +32)
    constexpr int kCols = kFeat / 2;
    double acc[kCols];
#pragma unroll
    for (int c = 0; c < kCols; ++c) acc[c] = 0.0;
    const int n_chunks = (my_tiles + kFlushTiles - 1) / kFlushTiles;
    for (int chunk = 0; chunk < n_chunks; ++chunk) {
      const int ab = chunk & 1;
      BB_SUFFSTATS_EPI_WAIT(&sm.acc_full[ab], (chunk >> 1) & 1);
      ptx::tc_fence_after_sync();
      const uint32_t d_addr = tmem + (static_cast<uint32_t>(q * 32) << 16) + kTmemAcc + ab * kFeat +
                              chalf * kCols;
      uint32_t v0[16], v1[16];
      ptx::tmem_ld_32x32b_x16(d_addr, v0);
      ptx::tmem_ld_32x32b_x16(d_addr + 16, v1);
      ptx::tmem_wait_ld();
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&sm.acc_empty[ab]);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        acc[j] += static_cast<double>(__uint_as_float(v0[j]));
        acc[16 + j] += static_cast<double>(__uint_as_float(v1[j]));
      }
    }
    double* out = partial_s2 + (static_cast<int64_t>(blockIdx.x) * 128 + q * 32 + lane) * kFeat +
                  chalf * kCols;
#pragma unroll
    for (int c = 0; c < kCols; ++c) out[c] = acc[c];
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == kMmaWarp) ptx::tmem_dealloc(tmem, kTmemCols);
}

// S2[r,c] = sum_cta ( P[r][c] + P[64+r][c] + P[64+c][r] ),  S1[d] = sum_cta p1[d].
// Blocks 0..nb-1: 32 consecutive outputs x 8 groups of partials each (fixed summation order:
// deterministic); the last block reduces S1 (64 features x 4 groups).
constexpr int kFinThreads = 256;

// Optional consumer fused into the finalize kernel: the last block to finish evaluates the Gaussian
// expected log-likelihood from the finished statistics (stats_kernels.cu has the stand-alone
// kernel), which saves a launch on the single-GPU step.
struct LoglikTail {
  const double* e_lambda;
  const double* e_lambda_mu;
  double e_mu_l_mu, e_logdet, n;
  double* out;                    // nullptr: no tail
  unsigned int* counter;          // zeroed by suffstats_tc_kernel
};

__global__ void __launch_bounds__(kFinThreads)
suffstats_finalize_kernel(const double* __restrict__ partial_s2,
                          const double* __restrict__ partial_s1, int n_partials, int d,
                          int accumulate, double* __restrict__ s2, double* __restrict__ s1,
                          const LoglikTail tail) {
  __shared__ double red[kFinThreads];
  __shared__ bool is_last;
  const int t = threadIdx.x;
  if (blockIdx.x + 1 < gridDim.x) {
    const int lane_c = t & 31, g = t >> 5;
    const int idx = blockIdx.x * 32 + lane_c;
    double acc = 0.0;
    if (idx < d * d) {
      const int r = idx / d, c = idx % d;
#pragma unroll 4
      for (int p = g; p < n_partials; p += 8) {
        const double* P = partial_s2 + static_cast<int64_t>(p) * 128 * kFeat;
        acc += P[r * kFeat + c] + P[(kFeat + r) * kFeat + c] + P[(kFeat + c) * kFeat + r];
      }
    }
    red[t] = acc;
    __syncthreads();
    if (g == 0 && idx < d * d) {
      double total = 0.0;
#pragma unroll
      for (int j = 0; j < 8; ++j) total += red[j * 32 + lane_c];
      s2[idx] = accumulate ? s2[idx] + total : total;
    }
  } else if (s1 != nullptr) {
    const int f = t & 63, g = t >> 6;
    double acc = 0.0;
#pragma unroll 4
    for (int p = g; p < 2 * n_partials; p += 4) acc += partial_s1[static_cast<int64_t>(p) * kFeat + f];
    red[t] = acc;
    __syncthreads();
    if (g == 0 && f < d) {
      const double total = red[f] + red[64 + f] + red[128 + f] + red[192 + f];
      s1[f] = accumulate ? s1[f] + total : total;
    }
  }
  if (tail.out == nullptr) return;
  __threadfence();                                   // this block's statistics are visible device-wide
  __syncthreads();
  if (t == 0) is_last = atomicAdd(tail.counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  double acc = 0.0;
  for (int i = t; i < d * d; i += kFinThreads) acc -= 0.5 * tail.e_lambda[i] * __ldcg(s2 + i);
  for (int i = t; i < d; i += kFinThreads) acc += __ldcg(s1 + i) * tail.e_lambda_mu[i];
  red[t] = acc;
  __syncthreads();
  for (int w = kFinThreads / 2; w > 0; w >>= 1) {
    if (t < w) red[t] += red[t + w];
    __syncthreads();
  }
  if (t == 0) {
    const double log_2pi = 1.8378770664093454835606594728112;
    tail.out[0] = red[0] - 0.5 * tail.n * d * log_2pi + 0.5 * tail.n * tail.e_logdet - 0.5 * tail.n * tail.e_mu_l_mu;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

}  // namespace

bool suffstats_tc_supported(int64_t n, int d, const void* x) {
  return n > 0 && d >= 4 && d <= kFeat && (d % 4) == 0 &&
         (reinterpret_cast<uintptr_t>(x) % 16) == 0 && n < (int64_t(1) << 31) - kTileRows;
}

int64_t suffstats_tc_workspace(int64_t n) {
  const int64_t tiles = (n + kTileRows - 1) / kTileRows;
  int64_t grid = device_sm_count();
  if (grid <= 0) grid = 148;
  if (tiles < grid) grid = tiles > 0 ? tiles : 1;
  return grid * (128 * kFeat + 2 * kFeat) * static_cast<int64_t>(sizeof(double)) + 512;   // + finalize counter
}

namespace {
// s1 may be nullptr.  s1/s2 are device float64; with `accumulate` the results are added to them.
int launch_suffstats_tc_impl(const float* x, int64_t n, int d, double* s1, double* s2,
                             void* workspace, int64_t workspace_bytes, bool accumulate, LoglikTail tail,
                             cudaStream_t stream) {
  if (!suffstats_tc_supported(n, d, x)) {
    set_error("suffstats_tc: unsupported shape n=%lld d=%d", static_cast<long long>(n), d);
    return BB_ERR_UNSUPPORTED;
  }
  const int64_t need = suffstats_tc_workspace(n);
  if (workspace == nullptr || workspace_bytes < need) {
    set_error("suffstats_tc: workspace %lld < %lld bytes", static_cast<long long>(workspace_bytes),
              static_cast<long long>(need));
    return BB_ERR_WORKSPACE;
  }
  EncodeTiledFn encode = get_encode_tiled();
  if (encode == nullptr) {
    set_error("cuTensorMapEncodeTiled unavailable from the driver");
    return BB_ERR_CUDA;
  }
  CUtensorMap map;
  const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(d), static_cast<cuuint64_t>(n)};
  const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(d) * sizeof(float)};
  const cuuint32_t box[2] = {kBoxCols, kTileRows};
  const cuuint32_t estride[2] = {1, 1};
  CUresult cr = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(x), gdim,
                       gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (n=%lld d=%d)", static_cast<int>(cr),
              static_cast<long long>(n), d);
    return BB_ERR_CUDA;
  }
  const int64_t tiles = (n + kTileRows - 1) / kTileRows;
  int grid = device_sm_count();
  if (tiles < grid) grid = static_cast<int>(tiles);
  double* partial_s2 = reinterpret_cast<double*>(
      (reinterpret_cast<uintptr_t>(workspace) + 255) & ~static_cast<uintptr_t>(255));
  double* partial_s1 = partial_s2 + static_cast<int64_t>(grid) * 128 * kFeat;
  unsigned int* fin_counter = reinterpret_cast<unsigned int*>(partial_s1 + static_cast<int64_t>(grid) * 2 * kFeat);
  tail.counter = fin_counter;

  const int smem_bytes = static_cast<int>(sizeof(SmemLayout)) + 1024;
  static SmemOptIn smem_opt_in_0;
  BB_CUDA_OK(smem_opt_in_0.ensure(suffstats_tc_kernel, smem_bytes));
  suffstats_tc_kernel<<<grid, kThreads, smem_bytes, stream>>>(map, tiles, partial_s2, partial_s1, fin_counter);
  BB_CHECK_LAUNCH("suffstats_tc_kernel");
  const int fin_blocks = (d * d + 31) / 32 + 1;
  suffstats_finalize_kernel<<<fin_blocks, kFinThreads, 0, stream>>>(partial_s2, partial_s1, grid, d,
                                                                    accumulate ? 1 : 0, s2, s1, tail);
  BB_CHECK_LAUNCH("suffstats_finalize_kernel");
  return BB_OK;
}
}  // namespace

int launch_suffstats_tc_acc(const float* x, int64_t n, int d, double* s1, double* s2,
                            void* workspace, int64_t workspace_bytes, bool accumulate,
                            cudaStream_t stream) {
  LoglikTail none = {nullptr, nullptr, 0.0, 0.0, 0.0, nullptr, nullptr};
  return launch_suffstats_tc_impl(x, n, d, s1, s2, workspace, workspace_bytes, accumulate, none, stream);
}

// statistics and, in the finalize kernel's last block, the expected log-likelihood from them (s1 required)
int launch_suffstats_tc_loglik(const float* x, int64_t n, int d, double* s1, double* s2, double n_total,
                               const double* e_lambda, const double* e_lambda_mu, double e_mu_l_mu,
                               double e_logdet, double* loglik, void* workspace, int64_t workspace_bytes,
                               cudaStream_t stream) {
  LoglikTail tail = {e_lambda, e_lambda_mu, e_mu_l_mu, e_logdet, n_total, loglik, nullptr};
  return launch_suffstats_tc_impl(x, n, d, s1, s2, workspace, workspace_bytes, false, tail, stream);
}

int launch_suffstats_tc(const float* x, int64_t n, int d, double* s1, double* s2, void* workspace,
                        int64_t workspace_bytes, cudaStream_t stream) {
  return launch_suffstats_tc_acc(x, n, d, s1, s2, workspace, workspace_bytes, false, stream);
}

}  // namespace bb
