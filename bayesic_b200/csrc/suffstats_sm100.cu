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
//   * the SAME launch finishes the job (round 1 needed a finalize launch + an all-reduce launch): every CTA
//     adds its float64 partial statistics (upper triangle of S2, S1: 2144 values) into one 33 KB accumulator
//     block with L2 reductions (red.global.add.f64) and takes a ticket; the LAST CTA to finish reads the
//     block, re-zeroes it for the next launch and does the rest alone -- no grid barrier, no partial
//     workspace.  (The sum over CTAs is in arrival order: float64, so run-to-run differences are ~1e-16
//     relative; everything after it is in a fixed order.)
//   * multi-GPU (world > 1): the last CTA PUSHES the rank's 33 KB payload into every peer's receive buffer
//     over NVLink (plain stores into peer memory), publishes one flag per peer (st.release.sys), waits for the
//     peers' flags, and sums the world's contributions from LOCAL memory in rank order -- the result is
//     bit-identical on every rank.  Receive buffers are double-buffered by epoch parity;
//   * the expected log-likelihood (ELBO term) of the reduced statistics is a dot product with E[Lambda],
//     E[Lambda mu] in the same CTA (fixed reduction tree).  One launch per step on any number of GPUs.
//
// Algorithmic traffic: 4*d bytes per row, read once.  Nothing else touches HBM.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <cstdlib>

#include "common.cuh"
#include "kernels.h"
#include "sm100_ptx.cuh"

namespace bb {

namespace {

constexpr int kFeat = 64;          // padded feature extent (= MMA N, = half of MMA M)
constexpr int kTileRows = 128;     // data rows per pipeline stage
constexpr int kStages = 6;
constexpr int kClaimTiles = 4;     // tiles per claim of the dynamic scheduler (512 rows = 128 KB, ~2 us of one SM's share)
// TMEM accumulators are drained to float64 every kFlushTiles tiles.  The tensor core truncates
// its fp32 accumulate, so the chunk length sets the (systematic) error: measured on B200 at
// N = 16 Mi, D = 64:  4 tiles (512 rows) 2.5e-6 relative, 0.79 ms/pass;  1 tile 7e-7, 0.93 ms.
// Both are far inside the 1e-4 parity bar; the default favours throughput.
#ifndef BB_SUFFSTATS_FLUSH_TILES
#define BB_SUFFSTATS_FLUSH_TILES 4
#endif
constexpr int kFlushTiles = BB_SUFFSTATS_FLUSH_TILES;
// a claim of the dynamic scheduler is exactly one fp32 accumulation chunk, aligned: whichever CTA processes it,
// the chunk's fp32 sums are the same, so the statistics do not depend on the tile assignment beyond the order
// of the float64 additions
static_assert(kClaimTiles == kFlushTiles, "claims must coincide with the fp32 accumulation chunks");
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
  uint64_t done;                 // MMA issuer: the tile stream has ended, total_tiles is valid
  double s1_part[2][kFeat];      // Sigma x of the two row halves (split warps)
  volatile int stage_valid[kStages];   // 1: the stage holds a tile; 0: end of this CTA's tile stream
  volatile int total_tiles;
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
__device__ __forceinline__ void split_warp_loop(SmemLayout& sm, uint32_t tmem, int q, int khalf, int lane) {
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
  for (int i = 0;; ++i) {
    const int s = i % kStages;
    const int b = i & 1;
    ptx::mbar_wait(&sm.full[s], (i / kStages) & 1);
    if (!sm.stage_valid[s]) break;                      // end of the tile stream
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

// Developer timeline (variant build -DBB_SUFFSTATS_TIMELINE): thread 0 of CTA 0 (stamps 0-3) and of the last
// CTA to finish (stamps 4-7) record %globaltimer at the phase boundaries; the last CTA prints the deltas (ns).
#ifdef BB_SUFFSTATS_TIMELINE
__device__ unsigned long long g_tl[16];
__device__ unsigned long long g_cta[2][256];      // per-CTA time of kernel entry / end of the main loop
__device__ __forceinline__ void tl_stamp(int i) {
  if (threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (blockIdx.x == 0 || i >= 3) g_tl[i] = t;
    if (i == 0) g_cta[0][blockIdx.x] = t;
    if (i == 2) g_cta[1][blockIdx.x] = t;
  }
}
#define BB_TL(i) tl_stamp(i)
#else
#define BB_TL(i)
#endif

// ---- fused tail: cross-CTA reduction, cross-GPU exchange, expected log-likelihood ----------------
// Round 2, second design.  The first one-launch version wrote per-CTA partials, crossed a grid barrier and cut
// the outputs into slices summed by all CTAs: 16 us of tail on one GPU (timeline build) and ~60 us more with
// peers (128 CTAs x system-scope fences).  Now every CTA adds its partial statistics into ONE float64
// accumulator block with L2 reductions (red.global.add.f64; upper triangle only: 2144 per CTA), takes a ticket,
// and the LAST CTA alone reads the block (33 KB), re-zeroes it for the next launch, exchanges it with the peers
// and evaluates the ELBO term.  No grid barrier, no co-residency requirement, no partial workspace.
constexpr int kAccumDoubles = kFeat * kFeat + kFeat;        // [S2 (64 x 64, upper triangle used) | S1 (64)] float64
constexpr int kScratchStride = kFeat + 1;                   // padded row stride of the CTA's [128][64] float64 scratch
constexpr int kSlices = BB_GAUSSIAN_PASS_SLICES;            // stride of a rank's flag words in a peer's flag array

__device__ __forceinline__ void st_release_sys_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Runs in the last CTA to finish only.  `pay` is d*d + d + 1 doubles of shared memory (drained pipeline
// stages), `coef` as many, `red` one double per warp.  Every phase issues all of a thread's loads before the first dependent
// instruction or store (one memory latency per phase, not one per element).
constexpr int kPerThread = (kFeat * kFeat + kFeat + 1 + kThreads - 1) / kThreads;     // 8 payload elements per thread
constexpr int kAccPerThread = (kAccumDoubles + kThreads - 1) / kThreads;              // 8 accumulator slots per thread

__device__ __forceinline__ void last_cta_tail(const SuffstatsTail& tp, double* pay, double* coef, double* red,
                                              volatile int* sflag) {
  const int t = threadIdx.x;
  const int d = tp.d;
  const int elements = d * d + d + 1;                     // packed payload [S2 | S1 | row count]
  const bool want_ll = tp.loglik != nullptr;
  // 0. the consumer's coefficients for this thread's payload elements: issued first, needed last
  //    (parked in shared memory: registers are short in the exchange)
#pragma unroll
  for (int k = 0; k < kPerThread; ++k) {
    const int idx = t + k * kThreads;
    double c = 0.0;
    if (want_ll && idx < d * d) c = -0.5 * __ldg(tp.e_lambda + idx);
    else if (want_ll && idx < d * d + d) c = __ldg(tp.e_lambda_mu + (idx - d * d));
    if (idx < elements) coef[idx] = c;
  }
  // 1. this rank's statistics out of the accumulator block (re-zeroed for the next launch), mirrored into the
  //    packed payload
  {
    double v[kAccPerThread];
#pragma unroll
    for (int k = 0; k < kAccPerThread; ++k) {
      const int e = t + k * kThreads;
      const int r = e >> 6, c = e & 63;
      const bool used = e < kFeat * kFeat ? (r <= c && c < d) : (e < kAccumDoubles && e - kFeat * kFeat < d);
      v[k] = used ? __ldcg(tp.accum + e) : 0.0;
    }
#pragma unroll
    for (int k = 0; k < kAccPerThread; ++k) {
      const int e = t + k * kThreads;
      const int r = e >> 6, c = e & 63;
      if (e < kFeat * kFeat) {
        if (r <= c && c < d) {
          tp.accum[e] = 0.0;
          pay[r * d + c] = v[k];
          pay[c * d + r] = v[k];
        }
      } else if (e < kAccumDoubles && e - kFeat * kFeat < d) {
        tp.accum[e] = 0.0;
        pay[d * d + (e - kFeat * kFeat)] = v[k];
      }
    }
  }
  if (t == 0) {
    pay[d * d + d] = tp.local_count;
    *sflag = 0;
  }
  __syncthreads();
  BB_TL(4);
  double val[kPerThread];
#pragma unroll
  for (int k = 0; k < kPerThread; ++k) {
    const int idx = t + k * kThreads;
    val[k] = idx < elements ? pay[idx] : 0.0;
  }
  // 2. exchange: push the payload into every PEER's receive buffer, slot [rank], over NVLink; one flag per
  //    (sender, receiver); sum the world's slots (own payload from registers, the peers' from LOCAL memory) in
  //    rank order, so the result is bit-identical on every rank.  Receive buffers are double-buffered by epoch
  //    parity; the epoch lives in device memory (this launch uses stored + 1), so a captured CUDA graph can be
  //    replayed.
  uint32_t epoch = 0u;
  if (tp.world > 1) {
    epoch = __ldcg(tp.epoch_dev) + 1u;
    const int64_t parity_off = static_cast<int64_t>(epoch & 1u) * tp.world * tp.stride;
    for (int i = 1; i < tp.world; ++i) {
      double* dst = tp.peer_recv[(tp.rank + i) % tp.world] + parity_off + tp.rank * tp.stride;
#pragma unroll
      for (int k = 0; k < kPerThread; ++k)
        if (t + k * kThreads < elements) dst[t + k * kThreads] = val[k];
    }
    // the CTA barrier orders every thread's pushes before the flag threads' release (cumulativity): one
    // system-scope fence per flag thread, not one per thread
    __syncthreads();
    BB_TL(5);
    if (t < tp.world && t != tp.rank) {
      st_release_sys_u32(tp.peer_flags[t] + tp.rank * kSlices, epoch);
      const uint32_t* mine = tp.peer_flags[tp.rank] + t * kSlices;
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
    BB_TL(6);
    const bool lost = *sflag != 0;
    const double* mine = tp.peer_recv[tp.rank] + parity_off;
    double sum[kPerThread];
#pragma unroll
    for (int k = 0; k < kPerThread; ++k) sum[k] = 0.0;
    for (int r0 = 0; r0 < tp.world; r0 += 2) {               // two ranks' slots in flight together
      double in[2][kPerThread];
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int k = 0; k < kPerThread; ++k) {
          const int r = r0 + j, idx = t + k * kThreads;
          in[j][k] = (r < tp.world && r != tp.rank && idx < elements) ? __ldcv(mine + r * tp.stride + idx) : 0.0;
        }
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int k = 0; k < kPerThread; ++k)
          if (r0 + j < tp.world) sum[k] += (r0 + j == tp.rank) ? val[k] : in[j][k];
    }
#pragma unroll
    for (int k = 0; k < kPerThread; ++k)
      val[k] = lost ? __longlong_as_double(0x7ff8000000000000LL) : sum[k];      // lost peer: poison, never a partial sum
  }
  // 3. outputs and the expected log-likelihood of the reduced statistics
  double term = 0.0;
#pragma unroll
  for (int k = 0; k < kPerThread; ++k) {
    const int idx = t + k * kThreads;
    double v = val[k];
    if (idx < d * d) {
      if (tp.accumulate) v += tp.s2[idx];
      tp.s2[idx] = v;
    } else if (idx < d * d + d) {
      if (tp.s1 != nullptr) {
        if (tp.accumulate) v += tp.s1[idx - d * d];
        tp.s1[idx - d * d] = v;
      }
    } else if (idx == d * d + d) {
      if (tp.count_out != nullptr) *tp.count_out = v;
      pay[idx] = v;                                        // the reduced row count, read by thread 0 below
    }
    if (idx < elements) term += coef[idx] * v;
  }
  if (want_ll) {
    // fixed reduction tree: the value depends only on the reduced statistics
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) term += __shfl_xor_sync(0xffffffffu, term, off);
    if ((t & 31) == 0) red[t >> 5] = term;
  }
  __syncthreads();
  if (t == 0) {
    if (want_ll) {
      double total = 0.0;
      for (int w = 0; w < kThreads / 32; ++w) total += red[w];
      const double n = tp.world > 1 ? pay[d * d + d] : tp.n_total;
      const double log_2pi = 1.8378770664093454835606594728112;
      tp.loglik[0] = total - 0.5 * n * d * log_2pi + 0.5 * n * tp.e_logdet - 0.5 * n * tp.e_mu_l_mu;
    }
    *tp.ticket = 0u;                                       // the next launch touches these after its griddepcontrol.wait
    if (tp.tile_counter != nullptr) *tp.tile_counter = 0u;  // every CTA of this launch has finished claiming
    if (tp.world > 1) *tp.epoch_dev = epoch;
  }
  BB_TL(7);
}

__global__ void __launch_bounds__(kThreads, 1)
suffstats_tc_kernel(const __grid_constant__ CUtensorMap x_map, int64_t n_tiles,
                    const SuffstatsTail tail) {
  extern __shared__ uint8_t smem_raw[];
  BB_TL(0);
  SmemLayout& sm = *reinterpret_cast<SmemLayout*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

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
      ptx::mbar_init(&sm.done, 1);
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
    // Tiles are handed out dynamically: claims of kClaimTiles consecutive tiles from a device-wide counter
    // (the next claim is requested before the current one is issued, so its latency is hidden).  A CTA that
    // starts late -- the one whose SM ran the previous launch's tail, see the PDL note at the launch -- simply
    // takes fewer claims, and the CTAs end within one claim of each other.  tile_counter == nullptr: the
    // static partition (contiguous range per CTA).
    if (ptx::elect_one()) {
      int i = 0;
      auto issue = [&](int64_t tile) {
        const int s = i % kStages;
        ptx::mbar_wait(&sm.empty[s], ((i / kStages) & 1) ^ 1);
        sm.stage_valid[s] = 1;
        ptx::mbar_arrive_expect_tx(&sm.full[s], kStageBytes);
        const int32_t row0 = static_cast<int32_t>(tile * kTileRows);
        ptx::tma_load_2d(sm.stage[s], &x_map, &sm.full[s], 0, row0);
        ptx::tma_load_2d(sm.stage[s] + kHalfBytes, &x_map, &sm.full[s], kBoxCols, row0);
        ++i;
      };
      if (tail.tile_counter != nullptr) {
        int64_t next = atomicAdd(tail.tile_counter, static_cast<unsigned int>(kClaimTiles));
        while (next < n_tiles) {
          const int64_t first = next;
          next = atomicAdd(tail.tile_counter, static_cast<unsigned int>(kClaimTiles));
          const int64_t last = first + kClaimTiles < n_tiles ? first + kClaimTiles : n_tiles;
          for (int64_t tile = first; tile < last; ++tile) issue(tile);
        }
      } else {
        const int64_t tile_end = n_tiles * (blockIdx.x + 1) / gridDim.x;
        for (int64_t tile = n_tiles * blockIdx.x / gridDim.x; tile < tile_end; ++tile) issue(tile);
      }
      // end marker: a stage without a tile
      const int s = i % kStages;
      ptx::mbar_wait(&sm.empty[s], ((i / kStages) & 1) ^ 1);
      sm.stage_valid[s] = 0;
      ptx::mbar_arrive(&sm.full[s]);
    }
  } else if (warp == kMmaWarp) {
    // ---------------- MMA issuer ----------------
    if (ptx::elect_one()) {
      int i = 0;
      for (;; ++i) {
        const int s = i % kStages;
        const int b = i & 1;
        const int chunk = i / kFlushTiles;
        const int ab = chunk & 1;
        const bool first_in_chunk = (i % kFlushTiles) == 0;
        ptx::mbar_wait(&sm.full[s], (i / kStages) & 1);
        if (!sm.stage_valid[s]) break;                    // end of the tile stream: i tiles in total
        if (first_in_chunk) ptx::mbar_wait(&sm.acc_empty[ab], ((chunk >> 1) & 1) ^ 1);
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
        if ((i % kFlushTiles) == kFlushTiles - 1) ptx::mma_commit(&sm.acc_full[ab]);
      }
      if ((i % kFlushTiles) != 0) ptx::mma_commit(&sm.acc_full[(i / kFlushTiles) & 1]);     // the partial last chunk
      sm.total_tiles = i;
      ptx::mbar_arrive(&sm.done);                         // release: the epilogue warps read total_tiles after it
    }
  } else if (warp < kSplitWarps) {
    // ---------------- split warps: A = [hi ; lo] into TMEM, and Sigma x ----------------
    // Warp w owns TMEM lanes [32 (w&3), +32): quadrants 0,1 hold hi of features 0-31 / 32-63,
    // quadrants 2,3 hold lo.  The two warps sharing a quadrant take rows 0-63 / 64-127 of a tile.
    const int q = warp & 3;
    if (q < 2) split_warp_loop<false>(sm, tmem, q, warp >> 2, lane);
    else split_warp_loop<true>(sm, tmem, q, warp >> 2, lane);
  } else {
    // ---------------- epilogue warps: TMEM fp32 chunks -> float64 registers ----------------
    const int q = warp & 3;
    const int chalf = (warp - kSplitWarps) >> 2;       // columns [32 chalf, +32)
    constexpr int kCols = kFeat / 2;
    double acc[kCols];
#pragma unroll
    for (int c = 0; c < kCols; ++c) acc[c] = 0.0;
    // the number of chunks is known only when the tile stream ends (`done`): wait for the next chunk OR the end
    int n_chunks = 0x7fffffff;
    for (int chunk = 0;; ++chunk) {
      const int ab = chunk & 1;
      bool ready = false;
      while (!(ready = ptx::mbar_try_wait(&sm.acc_full[ab], (chunk >> 1) & 1))) {
        if (n_chunks == 0x7fffffff && ptx::mbar_try_wait(&sm.done, 0))
          n_chunks = (sm.total_tiles + kFlushTiles - 1) / kFlushTiles;
        if (chunk >= n_chunks) break;
      }
      if (!ready) break;
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
    // first three stages become the CTA's [128][64] float64 scratch (hi^T hi on rows 0-63, lo^T hi on 64-127);
    // rows are padded to 65 doubles so that neither these row-per-lane stores nor the transposed reads below
    // meet in one shared-memory bank
    double* out = reinterpret_cast<double*>(sm.stage[0]) + (q * 32 + lane) * kScratchStride + chalf * kCols;
#pragma unroll
    for (int c = 0; c < kCols; ++c) out[c] = acc[c];
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  BB_TL(2);
  if (warp == kMmaWarp) ptx::tmem_dealloc(tmem, kTmemCols);
  // Programmatic dependent launch: everything above reads only X and this launch's own tile counter, so it may
  // overlap the previous launch's tail.  The accumulator block, the ticket and the outputs are shared with the
  // previous launch: wait for it to complete (and its memory to be visible) here, then let the NEXT launch start
  // as CTAs of this one exit.  (Without the launch attribute both instructions are no-ops.)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  // this CTA's partial statistics, S2[r][c] = hi^T hi [r][c] + lo^T hi [r][c] + lo^T hi [c][r] (upper triangle)
  // and S1, added into the accumulator block at L2.  The start is rotated per CTA so that the 148 CTAs, which
  // finish together, do not all hit the same address at the same moment.
  {
    const double* P = reinterpret_cast<const double*>(sm.stage[0]);
    const int d = tail.d;
    const int rot = static_cast<int>((blockIdx.x * 416u) % kAccumDoubles);
    for (int e0 = threadIdx.x; e0 < kAccumDoubles; e0 += kThreads) {
      int e = e0 + rot;
      if (e >= kAccumDoubles) e -= kAccumDoubles;
      if (e < kFeat * kFeat) {
        const int r = e >> 6, c = e & 63;
        if (r <= c && c < d)
          atomicAdd(tail.accum + e, P[r * kScratchStride + c] + P[(kFeat + r) * kScratchStride + c] +
                                        P[(kFeat + c) * kScratchStride + r]);
      } else if (e - kFeat * kFeat < d) {
        atomicAdd(tail.accum + e, sm.s1_part[0][e - kFeat * kFeat] + sm.s1_part[1][e - kFeat * kFeat]);
      }
    }
  }
  // completion ticket: the last CTA to arrive sees every CTA's reductions
  __threadfence();
  __syncthreads();
  volatile int* sflag = reinterpret_cast<volatile int*>(&sm.tmem_base);
  if (threadIdx.x == 0) *sflag = atomicAdd(tail.ticket, 1u) == gridDim.x - 1u;
  __syncthreads();
  BB_TL(3);
  if (!*sflag) return;
  __threadfence();
  // scratch of the tail: the payload in stages 3-4 (33 KB), the consumer's coefficients in stages 0-1 (the
  // CTA's statistics scratch is dead: its reductions were issued before the barrier above), per-warp terms in 5
  last_cta_tail(tail, reinterpret_cast<double*>(sm.stage[3]), reinterpret_cast<double*>(sm.stage[0]),
                reinterpret_cast<double*>(sm.stage[5]), sflag);
#ifdef BB_SUFFSTATS_TIMELINE
  if (threadIdx.x == 0) {
    printf("timeline ns (CTA 0): init %llu  main %llu  reductions+ticket %llu | last CTA %d: gather %llu  push+fence %llu  "
           "flags %llu  sum+outputs %llu | launch total %llu  since previous launch's end %lld\n",
           g_tl[1] - g_tl[0], g_tl[2] - g_tl[1], g_cta[1][0] ? g_tl[3] - g_cta[1][blockIdx.x] : 0ull, static_cast<int>(blockIdx.x),
           g_tl[4] - g_tl[3], g_tl[5] - g_tl[4], g_tl[6] - g_tl[5], g_tl[7] - (tail.world > 1 ? g_tl[6] : g_tl[4]),
           g_tl[7] - g_tl[0], static_cast<long long>(g_tl[0] - g_tl[8]));
    unsigned long long s0 = ~0ull, s1 = 0, m0 = ~0ull, m1 = 0;
    for (unsigned int b = 0; b < gridDim.x; ++b) {
      s0 = min(s0, g_cta[0][b]); s1 = max(s1, g_cta[0][b]);
      m0 = min(m0, g_cta[1][b]); m1 = max(m1, g_cta[1][b]);
    }
    printf("   CTA entry skew %llu ns, main-loop-end skew %llu ns; last main-loop end -> kernel end %llu ns\n",
           s1 - s0, m1 - m0, g_tl[7] - m1);
    g_tl[8] = g_tl[7];
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
// developer switches (read once): BB_SUFFSTATS_DYNAMIC=0 static tile partition, BB_SUFFSTATS_PDL=0 plain launches
bool dynamic_tiles() {
  static const bool on = !(getenv("BB_SUFFSTATS_DYNAMIC") && atoi(getenv("BB_SUFFSTATS_DYNAMIC")) == 0);
  return on;
}
bool pdl_enabled() {
  static const bool on = !(getenv("BB_SUFFSTATS_PDL") && atoi(getenv("BB_SUFFSTATS_PDL")) == 0);
  return on;
}
int grid_for(int64_t n) {
  const int64_t tiles = (n + kTileRows - 1) / kTileRows;
  int64_t grid = device_sm_count();
  if (grid <= 0) grid = 148;
  if (tiles < grid) grid = tiles > 0 ? tiles : 1;
  return static_cast<int>(grid);
}
}  // namespace

// the accumulator block + the ticket word (caller-workspace entry points; a bb_gaussian_pass handle owns
// persistent ones)
int64_t suffstats_tc_workspace(int64_t) {
  return kAccumDoubles * static_cast<int64_t>(sizeof(double)) + 64 + 512;     // block, ticket, tile counter
}

// ONE launch: statistics, cross-CTA reduction, (world > 1) cross-GPU exchange, expected
// log-likelihood.  `tail` carries outputs / consumers / peers; without its own accumulator block and ticket
// (zero between launches) they are carved from the workspace here and zeroed by a memset node.  n == 0 is allowed (a rank with no rows still takes part in the exchange).
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
  if (tail.accum == nullptr || tail.ticket == nullptr) {
    double* block = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~static_cast<uintptr_t>(255));
    tail.accum = block;
    tail.ticket = reinterpret_cast<unsigned int*>(block + kAccumDoubles);
    tail.tile_counter = dynamic_tiles() ? tail.ticket + 1 : nullptr;
    tail.pdl = 0;
    BB_CUDA_OK(cudaMemsetAsync(block, 0, kAccumDoubles * sizeof(double) + 2 * sizeof(unsigned int), stream));
  }
  if (!dynamic_tiles()) tail.tile_counter = nullptr;
  tail.d = d;
  if (tail.world < 1) tail.world = 1;

  const int smem_bytes = static_cast<int>(sizeof(SmemLayout)) + 1024;
  static SmemOptIn smem_opt_in_0;
  BB_CUDA_OK(smem_opt_in_0.ensure(suffstats_tc_kernel, smem_bytes));
  if (tail.pdl && pdl_enabled()) {
    // back-to-back passes of one handle: the next launch's streaming overlaps this launch's single-CTA tail
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    BB_CUDA_OK(cudaLaunchKernelEx(&cfg, suffstats_tc_kernel, map, tiles, tail));
  } else {
    suffstats_tc_kernel<<<grid, kThreads, smem_bytes, stream>>>(map, tiles, tail);
  }
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
