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
//   * per-CTA float64 partials go to a workspace (L2-resident, 148 x 65.5 KB) and the SAME launch
//     finishes the job (round 2; round 1 needed a finalize launch + an all-reduce launch):
//     one grid-wide barrier (cooperative launch), then the d*d + d + 1 outputs are cut into 128
//     slices and CTA j adds up slice j over all partials in a fixed order (deterministic), applying
//     the symmetrisation above;
//   * multi-GPU (world > 1): CTA j then PUSHES its slice into every peer's receive buffer over
//     NVLink (plain stores into peer memory), publishes a per-slice flag (st.release.sys), waits for
//     the peers' flags of the same slice only (so the exchange pipelines slice by slice, no second
//     grid barrier), and sums the world's contributions from LOCAL memory in rank order -- the result
//     is bit-identical on every rank.  Receive buffers are double-buffered by epoch parity;
//   * the expected log-likelihood (ELBO term) of the reduced statistics is a per-slice dot product
//     with E[Lambda], E[Lambda mu]; the last CTA to finish (atomic ticket) adds the 128 slice terms
//     in order.  One launch per step on any number of GPUs.
//
// Algorithmic traffic: 4*d bytes per row, read once.  Nothing else touches HBM except
// 148 x 65.5 KB of partials.
#include <cooperative_groups.h>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "common.cuh"
#include "kernels.h"
#include "sm100_ptx.cuh"

namespace cg = cooperative_groups;

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
  double s1_part[2][kFeat];      // Sigma x of the two row halves (split warps)
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
                                                int lane, int my_tiles) {
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
  if (!kIsLo) sm.s1_part[khalf][half * 32 + lane] = s1;
}

// Developer timeline (variant build -DBB_SUFFSTATS_TIMELINE): CTA 0 / thread 0 stamps %globaltimer at the
// phase boundaries and prints the deltas (ns) at the end of every launch.
#ifdef BB_SUFFSTATS_TIMELINE
__device__ unsigned long long g_tl[16];
__device__ unsigned long long g_cta[2][256];      // per-CTA time of kernel entry / end of the main loop
__device__ __forceinline__ void tl_stamp(int i) {
  if (threadIdx.x == 0 && (blockIdx.x == 0 || i == 0 || i == 2)) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (blockIdx.x == 0) g_tl[i] = t;
    if (i == 0) g_cta[0][blockIdx.x] = t;
    if (i == 2) g_cta[1][blockIdx.x] = t;
  }
}
#define BB_TL(i) tl_stamp(i)
#else
#define BB_TL(i)
#endif

// ---- fused tail: cross-CTA reduction, cross-GPU exchange, expected log-likelihood ----------------
constexpr int kPartialDoubles = kFeat * kFeat + kFeat;      // one CTA's partial: [S2 (64 x 64) | S1 (64)] float64
constexpr long long kGridBarrierSpinLimit = 4000000000LL;  // ~2 s of clock64 ticks
constexpr int kSlices = BB_GAUSSIAN_PASS_SLICES;   // output slices (and per-peer flags); grid-size independent so
                                                   // that ranks whose grids differ agree on the slicing

__device__ __forceinline__ void st_release_sys_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_gpu_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Runs in every CTA after its partials are written and the grid barrier has passed.
// `red` is kThreads doubles of shared memory (the drained pipeline stages are reused).
__device__ __forceinline__ void fused_tail(const SuffstatsTail& tp, const double* __restrict__ partial,
                                           unsigned int bar_base, double* red, volatile int* sflag) {
  const int t = threadIdx.x;
  const int d = tp.d;
  const int elements = d * d + d + 1;                     // [S2 | S1 | row count]
  const int per = (elements + kSlices - 1) / kSlices;     // outputs per slice (33 at d = 64)
  const int groups = kThreads / per;                      // partial groups summed in parallel
  const int e = t % per, g = t / per;
  const int grid = gridDim.x;
  const bool want_ll = tp.loglik != nullptr;
  // the epoch lives in device memory (this launch uses stored + 1; the last CTA stores it back), so a
  // captured CUDA graph can be replayed: no per-launch host argument changes
  const uint32_t epoch = tp.world > 1 ? __ldcg(tp.epoch_dev) + 1u : 0u;
  const int64_t parity_off = static_cast<int64_t>(epoch & 1u) * tp.world * tp.stride;
  for (int slice = blockIdx.x; slice < kSlices && slice * per < elements; slice += grid) {
    const int idx = slice * per + e;
    double acc = 0.0;
    if (g < groups && idx < elements) {
      if (idx < d * d + d) {
        // element of the packed output -> element of a CTA's [64 x 64 | 64] partial
        const int pidx = idx < d * d ? (idx / d) * kFeat + idx % d : kFeat * kFeat + (idx - d * d);
        const double* src = partial + pidx;
        // fixed assignment of partials to groups and a fixed combination order: deterministic; four
        // independent accumulators keep the L2 loads in flight together
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int p = g;
        for (; p + 3 * groups < grid; p += 4 * groups) {
          a0 += __ldcg(src + static_cast<int64_t>(p) * kPartialDoubles);
          a1 += __ldcg(src + static_cast<int64_t>(p + groups) * kPartialDoubles);
          a2 += __ldcg(src + static_cast<int64_t>(p + 2 * groups) * kPartialDoubles);
          a3 += __ldcg(src + static_cast<int64_t>(p + 3 * groups) * kPartialDoubles);
        }
        for (; p < grid; p += groups) a0 += __ldcg(src + static_cast<int64_t>(p) * kPartialDoubles);
        acc = (a0 + a1) + (a2 + a3);
      } else if (g == 0) {
        acc = tp.local_count;
      }
    }
    if (g < groups) red[g * per + e] = acc;
    if (t == 0) *sflag = 0;
    __syncthreads();
    const bool owner = g == 0 && idx < elements;          // threads 0 .. per-1
    double v = 0.0;
    if (owner)
      for (int j = 0; j < groups; ++j) v += red[j * per + e];     // fixed order: deterministic
    if (tp.world > 1) {
      if (owner) {
        // push this rank's slice into every rank's receive buffer (own included), slot [rank]
        for (int r = 0; r < tp.world; ++r) tp.peer_recv[r][parity_off + tp.rank * tp.stride + idx] = v;
        __threadfence_system();
      }
      __syncthreads();
      if (t < tp.world) {
        st_release_sys_u32(tp.peer_flags[t] + tp.rank * kSlices + slice, epoch);
        const uint32_t* mine = tp.peer_flags[tp.rank] + t * kSlices + slice;
        const long long t0 = clock64();
        // epochs are compared as signed distances so that the counter may wrap
        while (static_cast<int32_t>(ld_acquire_sys_u32(mine) - epoch) < 0) {
          if (clock64() - t0 > tp.spin_limit) {
            atomicMax(tp.status, t + 1);
            *sflag = 1;
            break;
          }
        }
      }
      __syncthreads();
      if (owner) {
        if (*sflag) {
          v = __longlong_as_double(0x7ff8000000000000LL);  // lost peer: poison, never a partial sum
        } else {
          const double* mine = tp.peer_recv[tp.rank] + parity_off + idx;
          v = 0.0;
          for (int r = 0; r < tp.world; ++r) v += __ldcv(mine + r * tp.stride);   // rank order: same bits everywhere
        }
      }
    }
    double term = 0.0;
    if (owner) {
      if (idx < d * d) {
        if (tp.accumulate) v += tp.s2[idx];
        tp.s2[idx] = v;
        if (want_ll) term = -0.5 * tp.e_lambda[idx] * v;
      } else if (idx < d * d + d) {
        const int f = idx - d * d;
        if (tp.s1 != nullptr) {
          if (tp.accumulate) v += tp.s1[f];
          tp.s1[f] = v;
        }
        if (want_ll) term = v * tp.e_lambda_mu[f];
      } else {
        if (tp.count_out != nullptr) *tp.count_out = v;
        tp.scratch[kSlices] = v;
      }
    }
    __syncthreads();                                       // red[] is reused below
    if (want_ll) {
      if (t < per) red[t] = term;
      __syncthreads();
      if (t == 0) {
        double total = 0.0;
        for (int j = 0; j < per; ++j) total += red[j];
        tp.scratch[slice] = total;
      }
      __syncthreads();
    }
  }
  BB_TL(4);
  // completion ticket = second round of arrivals on the barrier counter
  __threadfence();
  __syncthreads();
  if (t == 0) *sflag = atomicAdd(tp.bar_arrive, 1u) == bar_base + 2u * static_cast<unsigned int>(grid) - 1u;
  __syncthreads();
  if (!*sflag || t != 0) return;
  // last CTA of the grid: every CTA has read the stored base / epoch and finished
  __threadfence();
  *tp.bar_base = bar_base + 2u * static_cast<unsigned int>(grid);
  if (tp.world > 1) *tp.epoch_dev = epoch;
  if (!want_ll) return;
  double total = 0.0;
  const int used = (elements + per - 1) / per;
  for (int j = 0; j < used; ++j) total += __ldcg(tp.scratch + j);
  const double n = tp.world > 1 ? __ldcg(tp.scratch + kSlices) : tp.n_total;
  const double log_2pi = 1.8378770664093454835606594728112;
  tp.loglik[0] = total - 0.5 * n * d * log_2pi + 0.5 * n * tp.e_logdet - 0.5 * n * tp.e_mu_l_mu;
}

__global__ void __launch_bounds__(kThreads, 1)
suffstats_tc_kernel(const __grid_constant__ CUtensorMap x_map, int64_t n_tiles,
                    double* __restrict__ partial,      // [grid][kPartialDoubles]
                    const SuffstatsTail tail) {
  extern __shared__ uint8_t smem_raw[];
  BB_TL(0);
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
  } else if (warp == kTmaWarp && lane == 0 && n_tiles > 0) {
    ptx::prefetch_tensormap(&x_map);
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem = sm.tmem_base;
  BB_TL(1);

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
    if (q < 2) split_warp_loop<false>(sm, tmem, q, warp >> 2, lane, my_tiles);
    else split_warp_loop<true>(sm, tmem, q, warp >> 2, lane, my_tiles);
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
    // every MMA of this CTA has completed (last acc_full), so every pipeline stage has been consumed: the
    // first two stages become the CTA's [128][64] float64 scratch (hi^T hi on rows 0-63, lo^T hi on 64-127)
    double* out = reinterpret_cast<double*>(sm.stage[0]) + (q * 32 + lane) * kFeat + chalf * kCols;
#pragma unroll
    for (int c = 0; c < kCols; ++c) out[c] = acc[c];
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  BB_TL(2);
  if (warp == kMmaWarp) ptx::tmem_dealloc(tmem, kTmemCols);
  // this CTA's partial statistics in the packed output layout [S2 (64 x 64) | S1 (64)]:
  // S2[r][c] = hi^T hi [r][c] + lo^T hi [r][c] + lo^T hi [c][r]
  {
    const double* P = reinterpret_cast<const double*>(sm.stage[0]);
    double* mine = partial + static_cast<int64_t>(blockIdx.x) * kPartialDoubles;
    for (int e = threadIdx.x; e < kPartialDoubles; e += kThreads) {
      double v;
      if (e < kFeat * kFeat) {
        const int r = e >> 6, c = e & 63;
        v = P[r * kFeat + c] + P[(kFeat + r) * kFeat + c] + P[(kFeat + c) * kFeat + r];
      } else {
        v = sm.s1_part[0][e - kFeat * kFeat] + sm.s1_part[1][e - kFeat * kFeat];
      }
      mine[e] = v;
    }
  }
  // every CTA's partials are complete and visible before any CTA reads them
  __threadfence();
  __syncthreads();
#ifdef BB_SUFFSTATS_COOP
  cg::this_grid().sync();
  const unsigned int bar_base = __ldcg(tail.bar_base);
#else
  // Grid barrier on the monotonic arrival counter (all CTAs are resident: grid <= SM count, one CTA per
  // SM).  The counter is never reset: this launch's arrivals run from bar_base (stored by the previous
  // launch's last CTA) to bar_base + 2 grid -- first the barrier, then the completion ticket.
  __shared__ unsigned int bar_base_s;
  __shared__ int bar_lost_s;
  if (threadIdx.x == 0) {
    const unsigned int base = __ldcg(tail.bar_base);
    bar_base_s = base;
    bar_lost_s = 0;
    atomicAdd(tail.bar_arrive, 1u);
    const unsigned int target = base + gridDim.x;
    const long long t0 = clock64();
    while (static_cast<int>(ld_acquire_gpu_u32(tail.bar_arrive) - target) < 0) {
      if (clock64() - t0 > kGridBarrierSpinLimit) {         // a CTA of this grid is not resident: never hang
        if (tail.status != nullptr) atomicMax(tail.status, 0x40000000);
        bar_lost_s = 1;
        break;
      }
    }
  }
  __syncthreads();
  const unsigned int bar_base = bar_base_s;
  if (bar_lost_s) return;
#endif
  BB_TL(3);
  // the pipeline is drained (every TMA load was consumed): its third stage is scratch now
  fused_tail(tail, partial, bar_base, reinterpret_cast<double*>(sm.stage[2]),
             reinterpret_cast<volatile int*>(&sm.tmem_base));
#ifdef BB_SUFFSTATS_TIMELINE
  BB_TL(5);
  if (blockIdx.x == 0 && threadIdx.x == 0)
    printf("timeline ns: init %llu  main+partials %llu  gridsync %llu  slice-reduce %llu  rest-of-tail(cta0) %llu  | since previous launch's end %lld\n",
           g_tl[1] - g_tl[0], g_tl[2] - g_tl[1], g_tl[3] - g_tl[2], g_tl[4] - g_tl[3], g_tl[5] - g_tl[4],
           static_cast<long long>(g_tl[0] - g_tl[6]));
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long s0 = ~0ull, s1 = 0, m0 = ~0ull, m1 = 0;
    for (unsigned int b = 0; b < gridDim.x; ++b) {
      s0 = min(s0, g_cta[0][b]); s1 = max(s1, g_cta[0][b]);
      m0 = min(m0, g_cta[1][b]); m1 = max(m1, g_cta[1][b]);
    }
    printf("   CTA entry skew %llu ns, main-loop-end skew %llu ns (CTA 0 entry at +%llu, main end at +%llu of the earliest)\n",
           s1 - s0, m1 - m0, g_tl[0] - s0, g_tl[2] - m0);
    g_tl[6] = g_tl[5];
  }
#endif
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

namespace {
int grid_for(int64_t n) {
  const int64_t tiles = (n + kTileRows - 1) / kTileRows;
  int64_t grid = device_sm_count();
  if (grid <= 0) grid = 148;
  if (tiles < grid) grid = tiles > 0 ? tiles : 1;
  return static_cast<int>(grid);
}
}  // namespace

// partials + [kSlices + 2] float64 tail scratch (the last one holds the two barrier words of the
// caller-workspace entry points)
int64_t suffstats_tc_workspace(int64_t n) {
  return grid_for(n) * kPartialDoubles * static_cast<int64_t>(sizeof(double)) +
         (kSlices + 2) * static_cast<int64_t>(sizeof(double)) + 512;
}

// ONE cooperative launch: statistics, cross-CTA reduction, (world > 1) cross-GPU exchange, expected
// log-likelihood.  `tail` carries outputs / consumers / peers; its scratch and ticket are carved from the
// workspace here.  n == 0 is allowed (a rank with no rows still takes part in the exchange).
int launch_suffstats_tc_fused(const float* x, int64_t n, int d, void* workspace, int64_t workspace_bytes,
                              SuffstatsTail tail, cudaStream_t stream) {
  if (n < 0 || d < 4 || d > kFeat || (d % 4) != 0 || (n > 0 && !suffstats_tc_supported(n, d, x))) {
    set_error("suffstats_tc: unsupported shape n=%lld d=%d", static_cast<long long>(n), d);
    return BB_ERR_UNSUPPORTED;
  }
  const int64_t need = suffstats_tc_workspace(n);
  if (workspace == nullptr || workspace_bytes < need) {
    set_error("suffstats_tc: workspace %lld < %lld bytes", static_cast<long long>(workspace_bytes),
              static_cast<long long>(need));
    return BB_ERR_WORKSPACE;
  }
  if (tail.s2 == nullptr || (tail.loglik != nullptr && (tail.e_lambda == nullptr || tail.e_lambda_mu == nullptr)) ||
      (tail.world > 1 && (tail.peer_recv == nullptr || tail.peer_flags == nullptr || tail.status == nullptr ||
                          tail.accumulate || tail.rank < 0 || tail.rank >= tail.world ||
                          tail.stride < static_cast<int64_t>(d) * d + d + 1 || tail.world > kThreads ||
                          tail.epoch_dev == nullptr))) {
    set_error("suffstats_tc: bad tail arguments");
    return BB_ERR_INVALID;
  }
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  if (n > 0) {
    EncodeTiledFn encode = get_encode_tiled();
    if (encode == nullptr) {
      set_error("cuTensorMapEncodeTiled unavailable from the driver");
      return BB_ERR_CUDA;
    }
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
  }
  int64_t tiles = (n + kTileRows - 1) / kTileRows;
  const int grid = grid_for(n);
  double* partial = reinterpret_cast<double*>(
      (reinterpret_cast<uintptr_t>(workspace) + 255) & ~static_cast<uintptr_t>(255));
  tail.scratch = partial + static_cast<int64_t>(grid) * kPartialDoubles;
  if (tail.bar_arrive == nullptr) {
    // caller-provided (uninitialised) workspace: the barrier words live behind the scratch and are zeroed
    // per launch; a bb_gaussian_pass handle owns persistent ones instead (no memset in its step)
    tail.bar_arrive = reinterpret_cast<unsigned int*>(tail.scratch + kSlices + 1);
    tail.bar_base = tail.bar_arrive + 1;
    BB_CUDA_OK(cudaMemsetAsync(tail.bar_arrive, 0, 2 * sizeof(unsigned int), stream));
  }
  tail.d = d;
  if (tail.world < 1) tail.world = 1;

  const int smem_bytes = static_cast<int>(sizeof(SmemLayout)) + 1024;
  static SmemOptIn smem_opt_in_0;
  BB_CUDA_OK(smem_opt_in_0.ensure(suffstats_tc_kernel, smem_bytes));
#ifdef BB_SUFFSTATS_COOP
  void* args[] = {&map, &tiles, &partial, &tail};
  BB_CUDA_OK(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(suffstats_tc_kernel), dim3(grid), dim3(kThreads),
                                         args, smem_bytes, stream));
#else
  // a plain launch: the grid barrier inside needs every CTA resident, which grid <= SM count with one CTA
  // per SM gives as long as no other kernel occupies SMs indefinitely (one pass at a time per device; a CTA
  // that cannot become resident turns into a status flag after ~2 s, not a hang).  A cooperative launch
  // would guarantee residency but was measured to add ~10 us of launch latency per step.
  suffstats_tc_kernel<<<grid, kThreads, smem_bytes, stream>>>(map, tiles, partial, tail);
#endif
  BB_CHECK_LAUNCH("suffstats_tc_kernel");
  return BB_OK;
}

namespace {
SuffstatsTail local_tail(double* s1, double* s2, bool accumulate, double local_count) {
  SuffstatsTail t;
  memset(&t, 0, sizeof(t));
  t.s1 = s1;
  t.s2 = s2;
  t.accumulate = accumulate ? 1 : 0;
  t.local_count = local_count;
  t.world = 1;
  return t;
}
}  // namespace

int launch_suffstats_tc_acc(const float* x, int64_t n, int d, double* s1, double* s2,
                            void* workspace, int64_t workspace_bytes, bool accumulate,
                            cudaStream_t stream) {
  return launch_suffstats_tc_fused(x, n, d, workspace, workspace_bytes,
                                   local_tail(s1, s2, accumulate, static_cast<double>(n)), stream);
}

// statistics and the expected log-likelihood from them in the same launch
int launch_suffstats_tc_loglik(const float* x, int64_t n, int d, double* s1, double* s2, double n_total,
                               const double* e_lambda, const double* e_lambda_mu, double e_mu_l_mu,
                               double e_logdet, double* loglik, void* workspace, int64_t workspace_bytes,
                               cudaStream_t stream) {
  SuffstatsTail t = local_tail(s1, s2, false, static_cast<double>(n));
  t.e_lambda = e_lambda;
  t.e_lambda_mu = e_lambda_mu;
  t.e_mu_l_mu = e_mu_l_mu;
  t.e_logdet = e_logdet;
  t.n_total = n_total;
  t.loglik = loglik;
  return launch_suffstats_tc_fused(x, n, d, workspace, workspace_bytes, t, stream);
}

int launch_suffstats_tc(const float* x, int64_t n, int d, double* s1, double* s2, void* workspace,
                        int64_t workspace_bytes, cudaStream_t stream) {
  return launch_suffstats_tc_acc(x, n, d, s1, s2, workspace, workspace_bytes, false, stream);
}

}  // namespace bb
