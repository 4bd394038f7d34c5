// Responsibility-weighted sufficient statistics on tcgen05 (D <= 64, K % 4 == 0):
//     Nk[k] = sum_n r_nk      Srx[k,d] = sum_n r_nk x_nd      Srxx[k,d,e] = sum_n r_nk x_nd x_ne
// in one pass over (R, X), never forming the K x D x N tensor that the reference's plan
//     _tensordot(_mul(_dimshuffle(R,1,'x',0), _dimshuffle(X,'x',1,0)), X, [2],[0])
// materialises (bayesic/algebra.py:741-765 -> 1297-1306 -> 1347-1351).
//
// Same machinery as suffstats_sm100.cu (TMA ring -> split warps -> TMEM A operand -> one
// M128 x N64 x K8 TF32 MMA per 8 rows against the raw FP32 tile as B), with the weight folded
// into the A operand.  Error compensation: with u = r*x (fp32), u = uh + ul, x = xh + xl
// (h = truncated to TF32, exactly what the tensor core reads from the raw B tile),
//     A = [ uh ; ul + r*xl ]            (rows 0-63 / 64-127, one row per feature)
//     G = A^T-stack . xh :   G[d,e] = sum u_d xh_e + sum r xl_d xh_e
// and because Srxx is symmetric,  Srxx = (G + G^T) / 2  to second order (the dropped terms are
// O(2^-22)).  So the compensated product again costs ONE M=128 MMA per 8 rows per component.
//
// Work split: a CTA owns a group of 4 components (4 x 64 TMEM accumulator columns) and a
// contiguous range of 128-row tiles; the CTAs of all groups walk the same rows in step so the
// X tiles they share come from L2.  FP32 TMEM accumulation is drained every 16 tiles
// (2048 rows) into the CTA's private float64 partials (read-modify-write in L2, no atomics).
//
// Measured on B200 (round 1): N = 1 Mi, D = 64, K = 256: 20.6 ms = 50.8 M rows/s = 107 TFLOP/s of
// useful 2 K D^2 flop/row (the FP32 SIMT kernel: 113 ms); relative error 1.5e-5 (fp32 TMEM
// accumulate truncation over 2048-row intervals).  Ablations (skip tcgen05.st / MMA / R load)
// move the time by < 12 %: the limiter is the split warps' ALU work (the lo half lives on TMEM
// lanes 64-127 = warps 2,3 mod 4 = two of the four schedulers), not the tensor pipe -- one
// M128 x N64 x K8 TF32 MMA issues every 45 cycles (tests/cuda/tc_rate.cu).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

#include "common.cuh"
#include "kernels.h"
#include "sm100_ptx.cuh"

namespace bb {

namespace {

constexpr int kFeat = 64;
constexpr int kTileRows = 128;
constexpr int kGroup = 4;                    // components per CTA
constexpr int kStages = 5;
constexpr int kBoxCols = 32;
constexpr int kHalfBytes = kTileRows * kBoxCols * 4;     // 16 KB
constexpr int kXBytes = 2 * kHalfBytes;                   // 32 KB
constexpr int kRBytes = kTileRows * kGroup * 4;           // 2 KB
constexpr int kStageBytes = kXBytes + kRBytes;
constexpr int kKBlocks = kTileRows / 8;
constexpr int kFlushTiles = 16;
constexpr int kSplitWarps = 16;   // quadrant = w & 3, row quarter = w >> 2 (32 rows of each tile)
constexpr int kRowParts = kSplitWarps / 4;
constexpr int kKbPerWarp = kKBlocks / kRowParts;
constexpr int kEpiWarps = 4;     // one per TMEM lane quadrant, all 256 accumulator columns
constexpr int kTmaWarp = kSplitWarps + kEpiWarps;
constexpr int kMmaWarp = kTmaWarp + 1;
constexpr int kThreads = (kMmaWarp + 1) * 32;
constexpr int kTmemCols = 512;
constexpr int kTmemAcc = 0;                  // 4 components x 64 columns
constexpr int kTmemA = 256;                  // 2 A buffers x 128 columns

struct __align__(1024) SmemLayout {
  uint8_t stage[kStages][kXBytes];           // X tiles (1024-aligned each)
  float rtile[kStages][kTileRows][kGroup];   // R[rows, 4 components]
  uint64_t full[kStages];
  uint64_t empty[kStages];
  uint64_t a_ready[2];
  uint64_t a_free[2];
  uint64_t acc_full;
  uint64_t acc_empty;
  uint32_t tmem_base;
};

constexpr uint32_t kIdesc = ptx::make_idesc(128, kFeat, 2, 0, 1);

// One split warp (TMEM lane quadrant q, rows [32 khalf, +32) of every tile; khalf = row part).  Lane = feature.
//   hi warps (q < 2):  A rows 0-63   <- u = r*x        (raw fp32; the tensor core reads its top
//                                                        19 bits = uh), and Sigma r x, Sigma r
//   lo warps (q >= 2): A rows 64-127 <- r*(x + xl) - uh = (r*x - uh) + r*xl   (two FMAs)
template <bool kIsLo>
__device__ __forceinline__ void split_loop(SmemLayout& sm, uint32_t tmem, int q, int khalf, int lane,
                                           int my_tiles, double* __restrict__ partial_rx,
                                           double* __restrict__ partial_nk) {
  const int half = q & 1;
  int off[8];
#pragma unroll
  for (int j = 0; j < 8; ++j)
    off[j] = half * kHalfBytes + khalf * (kKbPerWarp) * 1024 + j * 128 +
             ((((lane >> 3) ^ j) & 3) << 5) + (lane & 7) * 4;
  const uint32_t lane_addr = tmem + (static_cast<uint32_t>(q * 32) << 16) + kTmemA +
                             khalf * (kKbPerWarp) * 8;
  double rx[kGroup] = {0.0, 0.0, 0.0, 0.0};
  double nk[kGroup] = {0.0, 0.0, 0.0, 0.0};
  int slot = 0;                               // running (tile, component) counter
  for (int i = 0; i < my_tiles; ++i) {
    const int s = i % kStages;
    ptx::mbar_wait(&sm.full[s], (i / kStages) & 1);
    const uint8_t* stage = sm.stage[s];
    const float* rrow = &sm.rtile[s][khalf * (kTileRows / kRowParts)][0];
    // this warp's 64 rows x 1 feature stay in registers for all 4 components
    float x[kKbPerWarp][8];
#pragma unroll
    for (int kb = 0; kb < kKbPerWarp; ++kb)
#pragma unroll
      for (int j = 0; j < 8; ++j) x[kb][j] = *reinterpret_cast<const float*>(stage + off[j] + kb * 1024);
#pragma unroll
    for (int c = 0; c < kGroup; ++c, ++slot) {
      const int b = slot & 1;
      ptx::mbar_wait(&sm.a_free[b], ((slot >> 1) & 1) ^ 1);
      ptx::tc_fence_after_sync();
      const uint32_t a_addr = lane_addr + b * kTileRows;
      float rx_t = 0.f, nk_t = 0.f;
#pragma unroll
      for (int kb = 0; kb < kKbPerWarp; ++kb) {
        uint32_t v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float r = rrow[(kb * 8 + j) * kGroup + c];          // warp-wide broadcast
          const float xv = x[kb][j];
          const float u = r * xv;
          if (kIsLo) {
            const float uh = __uint_as_float(__float_as_uint(u) & 0xFFFFE000u);
            const float xh = __uint_as_float(__float_as_uint(xv) & 0xFFFFE000u);
            const float y = fmaf(2.f, xv, -xh);                     // x + xl, exact
            v[j] = __float_as_uint(fmaf(r, y, -uh));                // (r x - uh) + r xl
          } else {
            v[j] = __float_as_uint(u);
            rx_t += u;
            nk_t += r;
          }
        }
        ptx::tmem_st_32x32b_x8(a_addr + kb * 8, v);
      }
      ptx::tmem_wait_st();
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&sm.a_ready[b]);
      if (!kIsLo) {
        rx[c] += static_cast<double>(rx_t);
        nk[c] += static_cast<double>(nk_t);
      }
    }
  }
  if (!kIsLo) {
#pragma unroll
    for (int c = 0; c < kGroup; ++c) {
      partial_rx[((static_cast<int64_t>(blockIdx.x) * kRowParts + khalf) * kGroup + c) * kFeat + half * 32 + lane] = rx[c];
      if (q == 0 && lane == 0)
        partial_nk[(static_cast<int64_t>(blockIdx.x) * kRowParts + khalf) * kGroup + c] = nk[c];
    }
  }
}

// grid = n_groups * n_splits; CTA (g, s) = blockIdx.x % n_groups, blockIdx.x / n_groups
__global__ void __launch_bounds__(kThreads, 1)
weighted_stats_tc_kernel(const __grid_constant__ CUtensorMap x_map,
                         const __grid_constant__ CUtensorMap r_map, int64_t n_tiles, int n_groups,
                         int n_splits, double* __restrict__ partial_g,    // [grid][4][128][64]
                         double* __restrict__ partial_rx,                 // [grid][kRowParts][4][64]
                         double* __restrict__ partial_nk) {               // [grid][kRowParts][4]
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  SmemLayout& sm = *reinterpret_cast<SmemLayout*>(smem_raw);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int group = blockIdx.x % n_groups;
  const int split = blockIdx.x / n_groups;
  const int64_t tile_begin = n_tiles * split / n_splits;
  const int64_t tile_end = n_tiles * (split + 1) / n_splits;
  const int my_tiles = static_cast<int>(tile_end - tile_begin);

  if (warp == kMmaWarp) {
    if (lane == 0) {
      for (int s = 0; s < kStages; ++s) {
        ptx::mbar_init(&sm.full[s], 1);
        ptx::mbar_init(&sm.empty[s], 1);   // the MMA commit after the last component of a tile
      }
      for (int b = 0; b < 2; ++b) {
        ptx::mbar_init(&sm.a_ready[b], kSplitWarps);     // one elected arrival per split warp
        ptx::mbar_init(&sm.a_free[b], 1);
      }
      ptx::mbar_init(&sm.acc_full, 1);
      ptx::mbar_init(&sm.acc_empty, kEpiWarps);
      ptx::fence_mbar_init();
    }
    __syncwarp();
    ptx::tmem_alloc(&sm.tmem_base, kTmemCols);
  } else if (warp == kTmaWarp && lane == 0) {
    ptx::prefetch_tensormap(&x_map);
    ptx::prefetch_tensormap(&r_map);
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem = sm.tmem_base;

  if (warp == kTmaWarp) {
    if (ptx::elect_one()) {
      for (int i = 0; i < my_tiles; ++i) {
        const int s = i % kStages;
        ptx::mbar_wait(&sm.empty[s], ((i / kStages) & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(&sm.full[s], kStageBytes);
        const int32_t row0 = static_cast<int32_t>((tile_begin + i) * kTileRows);
        ptx::tma_load_2d(sm.stage[s], &x_map, &sm.full[s], 0, row0);
        ptx::tma_load_2d(sm.stage[s] + kHalfBytes, &x_map, &sm.full[s], kBoxCols, row0);
        ptx::tma_load_2d(&sm.rtile[s][0][0], &r_map, &sm.full[s], group * kGroup, row0);
      }
    }
  } else if (warp == kMmaWarp) {
    if (ptx::elect_one()) {
      int slot = 0;
      for (int i = 0; i < my_tiles; ++i) {
        const int s = i % kStages;
        const int interval = i / kFlushTiles;
        const bool first_in_interval = (i % kFlushTiles) == 0;
        if (first_in_interval && interval > 0) ptx::mbar_wait(&sm.acc_empty, (interval - 1) & 1);
        ptx::mbar_wait(&sm.full[s], (i / kStages) & 1);
        const uint32_t stage_addr = ptx::smem_u32(sm.stage[s]);
        for (int c = 0; c < kGroup; ++c, ++slot) {
          const int b = slot & 1;
          ptx::mbar_wait(&sm.a_ready[b], (slot >> 1) & 1);
          ptx::tc_fence_after_sync();
          const uint32_t d_tmem = tmem + kTmemAcc + c * kFeat;
          const uint32_t a_tmem = tmem + kTmemA + b * kTileRows;
#pragma unroll
          for (int kb = 0; kb < kKBlocks; ++kb) {
            const uint64_t b_desc = ptx::make_smem_desc(stage_addr + kb * 1024, kHalfBytes, 512,
                                                        ptx::kLayoutSwizzle128B32BAtom);
            ptx::mma_tf32_ts(d_tmem, a_tmem + kb * 8, b_desc, kIdesc,
                             (first_in_interval && kb == 0) ? 0u : 1u);
          }
          ptx::mma_commit(&sm.a_free[b]);
        }
        ptx::mma_commit(&sm.empty[s]);
        if ((i % kFlushTiles) == kFlushTiles - 1 || i == my_tiles - 1) ptx::mma_commit(&sm.acc_full);
      }
    }
  } else if (warp < kSplitWarps) {
    const int q = warp & 3;
    if (q < 2) split_loop<false>(sm, tmem, q, warp >> 2, lane, my_tiles, partial_rx, partial_nk);
    else split_loop<true>(sm, tmem, q, warp >> 2, lane, my_tiles, partial_rx, partial_nk);
  } else {
    // epilogue: quadrant q, all 4 components (256 accumulator columns)
    const int q = warp & 3;
    const int n_intervals = (my_tiles + kFlushTiles - 1) / kFlushTiles;
    double* base = partial_g + (static_cast<int64_t>(blockIdx.x) * kGroup * 128 + q * 32 + lane) * kFeat;
    for (int interval = 0; interval < n_intervals; ++interval) {
      ptx::mbar_wait(&sm.acc_full, interval & 1);
      ptx::tc_fence_after_sync();
#pragma unroll 1
      for (int part = 0; part < kGroup * kFeat / 16; ++part) {
        const int col = part * 16;
        uint32_t v[16];
        ptx::tmem_ld_32x32b_x16(tmem + (static_cast<uint32_t>(q * 32) << 16) + kTmemAcc + col, v);
        const int c = col / kFeat, e0 = col % kFeat;
        double* dst = base + static_cast<int64_t>(c) * 128 * kFeat + e0;
        double old[16];
        if (interval != 0) {
#pragma unroll
          for (int j = 0; j < 16; ++j) old[j] = dst[j];            // 16 L2 reads in flight
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) old[j] = 0.0;
        }
        ptx::tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 16; ++j) dst[j] = old[j] + static_cast<double>(__uint_as_float(v[j]));
      }
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&sm.acc_empty);
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == kMmaWarp) ptx::tmem_dealloc(tmem, kTmemCols);
}

// Srxx[k,d,e] = 1/2 sum_splits ( P[d][e] + P[64+d][e] + P[e][d] + P[64+e][d] ), plus Srx and Nk.
// grid = (k, 16): block (k, j) handles 256 of the d*d outputs of component k.
__global__ void __launch_bounds__(256)
weighted_finalize_kernel(const double* __restrict__ partial_g, const double* __restrict__ partial_rx,
                         const double* __restrict__ partial_nk, int n_groups, int n_splits, int d,
                         int k_total, double* __restrict__ nk, double* __restrict__ srx,
                         double* __restrict__ srxx) {
  const int k = blockIdx.x;
  const int g = k / kGroup, c = k % kGroup;
  const int idx = blockIdx.y * 256 + threadIdx.x;
  if (idx < d * d) {
    const int r = idx / d, e = idx % d;
    double acc = 0.0;
    for (int s = 0; s < n_splits; ++s) {
      const double* P = partial_g + ((static_cast<int64_t>(s) * n_groups + g) * kGroup + c) * 128 * kFeat;
      acc += P[r * kFeat + e] + P[(kFeat + r) * kFeat + e] + P[e * kFeat + r] + P[(kFeat + e) * kFeat + r];
    }
    srxx[(static_cast<int64_t>(k) * d + r) * d + e] = 0.5 * acc;
  }
  if (blockIdx.y == 0) {
    if (srx != nullptr && threadIdx.x < d) {
      double acc = 0.0;
      for (int s = 0; s < kRowParts * n_splits; ++s) {
        const int split = s / kRowParts, khalf = s % kRowParts;
        acc += partial_rx[(((static_cast<int64_t>(split) * n_groups + g) * kRowParts + khalf) * kGroup + c) * kFeat + threadIdx.x];
      }
      srx[static_cast<int64_t>(k) * d + threadIdx.x] = acc;
    }
    if (nk != nullptr && threadIdx.x == 0) {
      double acc = 0.0;
      for (int s = 0; s < kRowParts * n_splits; ++s) {
        const int split = s / kRowParts, khalf = s % kRowParts;
        acc += partial_nk[((static_cast<int64_t>(split) * n_groups + g) * kRowParts + khalf) * kGroup + c];
      }
      nk[k] = acc;
    }
  }
  (void)k_total;
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
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

void plan_grid(int64_t n, int k, int* n_groups, int* n_splits) {
  *n_groups = (k + kGroup - 1) / kGroup;
  const int64_t tiles = (n + kTileRows - 1) / kTileRows;
  int sms = device_sm_count();
  if (sms <= 0) sms = 148;
  int64_t splits = std::max<int64_t>(1, sms / *n_groups);
  splits = std::min<int64_t>(splits, std::max<int64_t>(1, tiles));
  *n_splits = static_cast<int>(splits);
}

}  // namespace

bool weighted_tc_supported(int64_t n, int d, int k, const void* x, const void* r) {
  return n > 0 && d >= 4 && d <= kFeat && d % 4 == 0 && k >= 1 && k % 4 == 0 &&
         reinterpret_cast<uintptr_t>(x) % 16 == 0 && reinterpret_cast<uintptr_t>(r) % 16 == 0 &&
         n < (int64_t(1) << 31) - kTileRows && (k + kGroup - 1) / kGroup <= 65535;
}

int64_t weighted_tc_workspace(int64_t n, int k) {
  int n_groups, n_splits;
  plan_grid(n, k, &n_groups, &n_splits);
  const int64_t grid = static_cast<int64_t>(n_groups) * n_splits;
  return grid * (kGroup * 128 * kFeat + kRowParts * kGroup * kFeat + kRowParts * kGroup) * static_cast<int64_t>(sizeof(double)) + 512;
}

int launch_weighted_stats_tc(const float* x, const float* r, int64_t n, int d, int k, double* nk,
                             double* sum_rx, double* sum_rxx, void* workspace,
                             int64_t workspace_bytes, cudaStream_t stream) {
  if (!weighted_tc_supported(n, d, k, x, r)) {
    set_error("weighted_stats_tc: unsupported shape n=%lld d=%d k=%d", static_cast<long long>(n), d, k);
    return BB_ERR_UNSUPPORTED;
  }
  if (workspace == nullptr || workspace_bytes < weighted_tc_workspace(n, k)) {
    set_error("weighted_stats_tc: workspace %lld < %lld bytes", static_cast<long long>(workspace_bytes),
              static_cast<long long>(weighted_tc_workspace(n, k)));
    return BB_ERR_WORKSPACE;
  }
  EncodeTiledFn encode = get_encode_tiled();
  if (encode == nullptr) {
    set_error("cuTensorMapEncodeTiled unavailable from the driver");
    return BB_ERR_CUDA;
  }
  CUtensorMap x_map, r_map;
  {
    const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(d), static_cast<cuuint64_t>(n)};
    const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(d) * sizeof(float)};
    const cuuint32_t box[2] = {kBoxCols, kTileRows};
    const cuuint32_t estride[2] = {1, 1};
    CUresult cr = encode(&x_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(x), gdim, gstride,
                         box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(X) failed: %d", static_cast<int>(cr)); return BB_ERR_CUDA; }
  }
  {
    const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(k), static_cast<cuuint64_t>(n)};
    const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(k) * sizeof(float)};
    const cuuint32_t box[2] = {kGroup, kTileRows};
    const cuuint32_t estride[2] = {1, 1};
    CUresult cr = encode(&r_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(r), gdim, gstride,
                         box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(R) failed: %d", static_cast<int>(cr)); return BB_ERR_CUDA; }
  }
  int n_groups, n_splits;
  plan_grid(n, k, &n_groups, &n_splits);
  const int grid = n_groups * n_splits;
  const int64_t tiles = (n + kTileRows - 1) / kTileRows;
  double* partial_g = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~static_cast<uintptr_t>(255));
  double* partial_rx = partial_g + static_cast<int64_t>(grid) * kGroup * 128 * kFeat;
  double* partial_nk = partial_rx + static_cast<int64_t>(grid) * kRowParts * kGroup * kFeat;
  const int smem_bytes = static_cast<int>(sizeof(SmemLayout));
  static SmemOptIn smem_opt_in_0;
  BB_CUDA_OK(smem_opt_in_0.ensure(weighted_stats_tc_kernel, smem_bytes));
  weighted_stats_tc_kernel<<<grid, kThreads, smem_bytes, stream>>>(x_map, r_map, tiles, n_groups, n_splits,
                                                                   partial_g, partial_rx, partial_nk);
  BB_CHECK_LAUNCH("weighted_stats_tc_kernel");
  dim3 fgrid(k, (d * d + 255) / 256);
  weighted_finalize_kernel<<<fgrid, 256, 0, stream>>>(partial_g, partial_rx, partial_nk, n_groups, n_splits, d,
                                                      k, nk, sum_rx, sum_rxx);
  BB_CHECK_LAUNCH("weighted_finalize_kernel");
  return BB_OK;
}

}  // namespace bb
