// Large-D Gram statistics on tcgen05 CTA pairs (D % 4 == 0, D > 64; the feature axis is padded with zeros to a
// multiple of 256 inside the converter -- nothing is padded in memory):
//     XtX[d,e] = sum_n x_nd x_ne        Xty[d] = sum_n x_nd y_n        yty = sum_n y_n^2
// over X[n, d] float32 row-major and (optionally) y[n].
//
// What it replaces: the reference evaluates these as the plans
//     _tensordot(_dimshuffle(X,1,0), X, [1],[0]),  _tensordot(_dimshuffle(X,1,0), y, [1],[0]),
//     _tensordot(y, y, [0],[0])
// (bayesic/algebra.py:527-551 -> 1347-1351), i.e. BLAS sgemm / sgemv over the data axis; they are
// the minibatch statistics of the conjugate natural-gradient SVI step for Bayesian linear
// regression / factor analysis (BASELINE cfg4: D = 1024, minibatch 1 Mi; README.md:69-80).
//
// Design (tensor-pipe bound: 2 D^2 flop/row against 4 D bytes/row):
//   * output tiled in 256 x 256 blocks, upper triangle only (XtX is symmetric); one CTA PAIR
//     (cluster of 2, tcgen05 cta_group::2, M = 256, N = 256) per (block, row-split): CTA r of
//     the pair owns output rows [128 r, +128) in its TMEM and supplies 128 features of the A
//     side and 128 features of the B side, so the pair reads 512 floats per data row -- the
//     2-CTA MMA halves the L2 -> SM traffic per flop compared with two independent CTAs;
//   * error-compensated BF16 ("BF16x3"): x = b1 + b2 + O(2^-17 x), b1 = bf16(x),
//     b2 = bf16(x - b1), and  x_d x_e ~= b1_d b1_e + b1_d b2_e + b2_d b1_e  (dropped terms
//     <= 2^-16 relative, unbiased because both splits round to nearest).  Three kind::f16 MMAs
//     cost 1.5x one TF32 pass (3xTF32 would cost 3x).  A DIAGONAL block issues two: its operand tile
//     holds 2 b2 (exact), S = b1^T b1 + b1^T (2 b2) = P + 2 Q, and the finalize takes (S + S^T) / 2 =
//     P + Q + Q^T -- the same three products;
//   * no fp32 staging: 16 converter warps load X with coalesced 128-bit loads straight into
//     registers (each warp group two stages ahead), split, and store b1/b2 into shared memory in the
//     UMMA MN-major SWIZZLE_128B canonical layout (the data axis is the MMA K axis, so X's
//     row-major layout IS MN-major: no transpose).  Shared memory therefore carries only the
//     bf16 tiles (1 KB/row written, 1.5 KB/row read by the tensor core);
//   * FP32 accumulation in TMEM (double-buffered, 2 x 256 columns) is drained every 2048 rows
//     by four epilogue warps into the pair's private fp32 partial block (read-modify-write
//     through L2) so the tensor core's truncating accumulate never sees a long chain;
//   * Xty / yty ride along in the converter warps of the diagonal blocks (they already hold
//     every x of their feature range in registers), accumulated in float64;
//   * a finalize kernel adds the row-splits in float64 in a fixed order (deterministic) and
//     mirrors the upper triangle.
//
// Algorithmic traffic: 4 D bytes per row; algorithmic flops: D (D + 1) per row (symmetric half).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"
#include "sm100_pair.cuh"
#include "sm100_ptx.cuh"

namespace bb {

namespace {

using namespace pair;

constexpr int kBlock = 256;               // output block edge (features) per CTA pair
constexpr int kHalf = 128;                // features per CTA per operand side
constexpr int kStageRows = 32;            // data rows per pipeline stage (2 MMA K-steps of 16)
constexpr int kStages = 6;
constexpr int kTileBytes = kHalf * kStageRows * 2;      // one bf16 operand tile: 8 KB
constexpr int kStageBytes = 4 * kTileBytes;             // A.b1, A.b2, B.b1, B.b2: 32 KB
constexpr int kFlushIters = 64 / BB_CHAIN_DIV;           // TMEM accumulators drained every 64 stages = 2048 rows
constexpr int kConvWarps = 16;
constexpr int kConvGroups = 2;           // warp groups taking alternate stages
constexpr int kRowsPerWarp = kStageRows / (kConvWarps / kConvGroups);   // 4
constexpr int kXtyFlushIters = 8;         // X^T y: fp32 partial sums over 8 x 4 rows, then float64
constexpr int kEpiWarps = 4;
constexpr int kMmaWarp = kConvWarps + kEpiWarps;        // warp 20 (20 % 4 == 0 is irrelevant for it)
constexpr int kThreads = (kMmaWarp + 1) * 32;           // 672
constexpr int kTmemCols = 512;            // 2 accumulator buffers x 256 columns

struct __align__(1024) SmemLayout {
  uint8_t stage[kStages][kStageBytes];
  double xty[kHalf];
  double yty;
  uint64_t full[kStages];        // leader: 2 CTAs x (kConvWarps / kConvGroups) arrivals (one warp group per stage)
  uint64_t empty[kStages];       // each CTA: one multicast MMA commit
  uint64_t acc_full[2];          // each CTA: one multicast MMA commit
  uint64_t acc_empty[2];         // leader: 2 CTAs x kEpiWarps arrivals
  uint32_t tmem_base;
};

// kind::f16, BF16 x BF16 -> FP32, both operands MN-major, M = 256 (pair), N = 256
constexpr uint32_t kIdesc = ptx::make_idesc(256, 256, /*bf16*/ 1, /*A MN-major*/ 1, /*B MN-major*/ 1);

__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ float4 ldg_f4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ void sts_u2(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

// x = b1 + b2 + O(2^-17 x): four consecutive features -> two packed bf16x2 words each.  kTwice stores 2 b2
// instead of b2 (exact: a power of two) -- the diagonal blocks' second product, see the MMA issuer.
template <bool kTwice = false>
__device__ __forceinline__ void split_bf16(const float4& x, uint32_t (&b1)[2], uint32_t (&b2)[2]) {
  __nv_bfloat162 p0 = __floats2bfloat162_rn(x.x, x.y);
  __nv_bfloat162 p1 = __floats2bfloat162_rn(x.z, x.w);
  b1[0] = *reinterpret_cast<uint32_t*>(&p0);
  b1[1] = *reinterpret_cast<uint32_t*>(&p1);
  float rx = x.x - __uint_as_float(b1[0] << 16);
  float ry = x.y - __uint_as_float(b1[0] & 0xFFFF0000u);
  float rz = x.z - __uint_as_float(b1[1] << 16);
  float rw = x.w - __uint_as_float(b1[1] & 0xFFFF0000u);
  if (kTwice) {
    rx += rx;
    ry += ry;
    rz += rz;
    rw += rw;
  }
  __nv_bfloat162 q0 = __floats2bfloat162_rn(rx, ry);
  __nv_bfloat162 q1 = __floats2bfloat162_rn(rz, rw);
  b2[0] = *reinterpret_cast<uint32_t*>(&q0);
  b2[1] = *reinterpret_cast<uint32_t*>(&q1);
}


struct ConvArgs {
  int prefetch_iters;      // L2 prefetch distance in converter iterations (0 = off)
  bool early;              // issue the next stage's loads before the proxy fence of the current one: their latency
                           // overlaps the fence instead of following it (the converters never wait on the stage barrier)
  const float* x;
  const float* y;
  int64_t n, row_begin, row_end;
  int d, feat_a, feat_b, n_iters, warp, lane;
  bool want_yty;
};

// Converter warps.  Two groups of eight warps; group g = warp & 1 owns every 2nd stage
// (it = g, g + 2, ...), warp wi = warp >> 1 of the group owns four rows of that stage.  A warp
// issues the loads of its NEXT stage right after publishing the current one, so they have two
// stage periods to land and no load is in flight across the fence of the arrive (a release with
// loads outstanding stalls until they return -- measured: 3.6x slower).  Lane l holds A features
// [4l, 4l+4) and B features [4l, 4l+4) of this CTA's 128-feature halves.
// kAlias: diagonal block, the B tiles ARE the A tiles (nothing loaded or stored for B); the b2 tile holds 2 b2.
// kXty:   also accumulate X^T y (and y^T y) for this CTA's A features.
// The loop is issue-bound (the first version spent 349 instructions per 4 rows, 698 issue cycles
// per stage against a 768-cycle MMA budget), hence the hoisted pointers, the bounds-check-free
// fast path and the precomputed swizzled offsets.
template <bool kAlias, bool kXty>
__device__ __forceinline__ void converter_loop(SmemLayout& sm, const ConvArgs& ca) {
  const int lane = ca.lane;
  const int group = ca.warp & (kConvGroups - 1);
  const int wi = ca.warp / kConvGroups;
  const int64_t d = ca.d;
  // shared-memory offset of this lane's 8-byte piece inside an operand tile, for row k = 4 wi + j:
  // [mn block (64 features) 4 KB][k group (k >> 3) 1 KB][k & 7 -> 128 B][16-byte chunk ^ (k & 7)]
  uint32_t soff[kRowsPerWarp];
#pragma unroll
  for (int j = 0; j < kRowsPerWarp; ++j) {
    const int k = wi * kRowsPerWarp + j;
    soff[j] = (lane >> 4) * 4096 + (k >> 3) * 1024 + (k & 7) * 128 +
              ((((lane >> 1) & 7) ^ (k & 7)) << 4) + (lane & 1) * 8;
  }
  const uint32_t stage0 = ptx::smem_u32(sm.stage[0]);
  // first row of this warp in its first stage, and the pointers that walk from there
  int64_t row0 = ca.row_begin + static_cast<int64_t>(group) * kStageRows + wi * kRowsPerWarp;
  const float* pa = ca.x + row0 * d + ca.feat_a + lane * 4;
  const float* pb = ca.x + row0 * d + ca.feat_b + lane * 4;
  // features at or beyond d (the zero padding up to a multiple of 256) are never loaded; d % 4 == 0, so a
  // lane's four features are all inside or all outside
  const bool a_in = ca.feat_a + lane * 4 < ca.d;
  const bool b_in = ca.feat_b + lane * 4 < ca.d;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const float* py = ca.y + row0;
  const int64_t step = static_cast<int64_t>(kConvGroups) * kStageRows * d;
  const bool do_yty = kXty && ca.want_yty && lane == 0;

  // register buffers of loaded rows: one set for the off-diagonal blocks (A and B rows), kDepth = 2 sets for the
  // diagonal ones (A rows only, so the second set is free) -- their loads are issued two of the warp's stages ahead
  constexpr int kDepth = kAlias ? 2 : 1;
  float4 ra[kDepth][kRowsPerWarp], rb[kAlias ? 1 : kRowsPerWarp];
  float ry[kDepth][kRowsPerWarp];
  auto load_rows = [&](float4 (&ra)[kRowsPerWarp], float (&ry)[kRowsPerWarp]) {
    if (row0 + kRowsPerWarp <= ca.n) {          // whole group of rows in range: no per-row checks
#pragma unroll
      for (int j = 0; j < kRowsPerWarp; ++j) {
        ra[j] = a_in ? ldg_f4(pa + j * d) : zero4;
        if (!kAlias) rb[j] = b_in ? ldg_f4(pb + j * d) : zero4;
        if (kXty) ry[j] = __ldg(py + j);
      }
    } else {
#pragma unroll
      for (int j = 0; j < kRowsPerWarp; ++j) {
        const bool ok = row0 + j < ca.n;
        ra[j] = ok && a_in ? ldg_f4(pa + j * d) : zero4;
        if (!kAlias) rb[j] = ok && b_in ? ldg_f4(pb + j * d) : zero4;
        if (kXty) ry[j] = ok ? __ldg(py + j) : 0.f;
      }
    }
    // L2 prefetch hint for the rows this warp converts `prefetch_iters` iterations from now:
    // 4 rows x (4 lines of A + 4 lines of B) = one 128-byte line per lane
    if (ca.prefetch_iters > 0) {
      const int64_t prow = row0 + static_cast<int64_t>(ca.prefetch_iters) * kConvGroups * kStageRows + (lane >> 3);
      if (prow < ca.row_end && (!kAlias || (lane & 4) == 0)) {
        const int feat = ((lane & 4) ? ca.feat_b : ca.feat_a) + (lane & 3) * 32;
        if (feat < ca.d) asm volatile("prefetch.global.L2 [%0];" ::"l"(ca.x + prow * d + feat));
      }
    }
    row0 += kConvGroups * kStageRows;
    pa += step;
    pb += step;
    py += kConvGroups * kStageRows;
  };

  double xty_acc[4] = {0.0, 0.0, 0.0, 0.0};
  double yty_acc = 0.0;
  float xty_f[4] = {0.f, 0.f, 0.f, 0.f};
  float yty_f = 0.f;
  int since_flush = 0;

  // one stage: convert the rows in (ra, ry), publish, and refill the same registers with the rows kDepth of this
  // warp's stages further on
  auto convert_stage = [&](int it, float4 (&ra)[kRowsPerWarp], float (&ry)[kRowsPerWarp]) {
    const int s = it % kStages;
    ptx::mbar_wait(&sm.empty[s], ((it / kStages) & 1) ^ 1);
    const uint32_t stage_addr = stage0 + s * kStageBytes;
#pragma unroll
    for (int j = 0; j < kRowsPerWarp; ++j) {
      const uint32_t addr = stage_addr + soff[j];
      uint32_t b1[2], b2[2];
      split_bf16<kAlias>(ra[j], b1, b2);
      sts_u2(addr, b1[0], b1[1]);
      sts_u2(addr + kTileBytes, b2[0], b2[1]);
      if (!kAlias) {
        split_bf16(rb[j], b1, b2);
        sts_u2(addr + 2 * kTileBytes, b1[0], b1[1]);
        sts_u2(addr + 3 * kTileBytes, b2[0], b2[1]);
      }
      if (kXty) {
        xty_f[0] = fmaf(ra[j].x, ry[j], xty_f[0]);
        xty_f[1] = fmaf(ra[j].y, ry[j], xty_f[1]);
        xty_f[2] = fmaf(ra[j].z, ry[j], xty_f[2]);
        xty_f[3] = fmaf(ra[j].w, ry[j], xty_f[3]);
        yty_f = fmaf(ry[j], ry[j], yty_f);
      }
    }
    const bool more = it + kDepth * kConvGroups < ca.n_iters;
    if (ca.early && more) load_rows(ra, ry);
    fence_proxy_async_smem();      // generic-proxy stores -> visible to the tensor core (async proxy)
    __syncwarp();
    if (lane == 0) mbar_arrive_cluster_relaxed(&sm.full[s], 0);
    if (!ca.early && more) load_rows(ra, ry);
    if (kXty && ++since_flush == kXtyFlushIters) {     // fp32 over 32 rows, then float64 (DADD is slow)
      since_flush = 0;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        xty_acc[c] += static_cast<double>(xty_f[c]);
        xty_f[c] = 0.f;
      }
      yty_acc += static_cast<double>(yty_f);
      yty_f = 0.f;
    }
  };

#pragma unroll
  for (int b = 0; b < kDepth; ++b)
    if (group + b * kConvGroups < ca.n_iters) load_rows(ra[b], ry[b]);
  for (int it = group; it < ca.n_iters; it += kDepth * kConvGroups) {
#pragma unroll
    for (int b = 0; b < kDepth; ++b)
      if (it + b * kConvGroups < ca.n_iters) convert_stage(it + b * kConvGroups, ra[b], ry[b]);
  }
  if (kXty) {
#pragma unroll
    for (int c = 0; c < 4; ++c) atomicAdd(&sm.xty[lane * 4 + c], xty_acc[c] + static_cast<double>(xty_f[c]));
    if (do_yty) atomicAdd(&sm.yty, yty_acc + static_cast<double>(yty_f));
  }
}

// grid = 2 * (nb * splits_diag + n_off * splits_off) CTAs; pairs [0, nb * splits_diag) own the diagonal blocks
// (block = p % nb, split = p / nb), the rest the off-diagonal ones in the same way: pairs of one split walk the
// same rows at the same time -> L2 reuse.  A diagonal block costs two products per K step, an off-diagonal one
// three, so the diagonal blocks get fewer, longer row ranges (GramGrid).
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gram_pair_kernel(const float* __restrict__ x, const float* __restrict__ y, int64_t n, int d,
                 int splits_diag, int splits_off, int ablate,
                 float* __restrict__ partial,          // [pair][2][256 cols][128 rows] fp32
                 double* __restrict__ partial_xty,     // [diagonal split][d]
                 double* __restrict__ partial_yty) {   // [diagonal split]
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  SmemLayout& sm = *reinterpret_cast<SmemLayout*>(smem_raw);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int nb = (d + kBlock - 1) / kBlock;
  const int n_off = nb * (nb - 1) / 2;
  const bool diag = pair < nb * splits_diag;
  int2 ij;
  int split, n_splits;
  if (diag) {
    ij.x = ij.y = pair % nb;
    split = pair / nb;
    n_splits = splits_diag;
  } else {
    const int q = pair - nb * splits_diag;
    split = q / n_off;
    n_splits = splits_off;
    // strict upper-triangle block index -> (i, j), row-major over i < j
    int i = 0, rem = q % n_off;
    while (rem >= nb - 1 - i) {
      rem -= nb - 1 - i;
      ++i;
    }
    ij.x = i;
    ij.y = i + 1 + rem;
  }
  const bool alias_b = diag;                       // diagonal block: the B tiles ARE the A tiles
  const int feat_a = ij.x * kBlock + static_cast<int>(rank) * kHalf;
  const int feat_b = ij.y * kBlock + static_cast<int>(rank) * kHalf;
  // rows of this split, in whole stages
  const int64_t total_iters = (n + kStageRows - 1) / kStageRows;
  const int64_t it_begin = total_iters * split / n_splits;
  const int64_t it_end = total_iters * (split + 1) / n_splits;
  const int n_iters = static_cast<int>(it_end - it_begin);
  const int64_t row_begin = it_begin * kStageRows;

#ifdef BB_GRAM_TIMELINE
  // developer timeline (variant build): every pair's leader prints its kind, stage count and duration
  unsigned long long tl_begin = 0;
  if (threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tl_begin));
#endif
  if (warp == kMmaWarp) {
    if (lane == 0) {
      for (int s = 0; s < kStages; ++s) {
        ptx::mbar_init(&sm.full[s], 2 * kConvWarps / kConvGroups);
        ptx::mbar_init(&sm.empty[s], 1);
      }
      for (int b = 0; b < 2; ++b) {
        ptx::mbar_init(&sm.acc_full[b], 1);
        ptx::mbar_init(&sm.acc_empty[b], 2 * kEpiWarps);
      }
      ptx::fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc_pair(&sm.tmem_base, kTmemCols);
  }
  if (threadIdx.x < kHalf) sm.xty[threadIdx.x] = 0.0;
  if (threadIdx.x == 0) sm.yty = 0.0;
  ptx::tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();          // both CTAs' barriers initialised before any remote arrive
  ptx::tc_fence_after_sync();
  const uint32_t tmem = sm.tmem_base;

  if (warp < kConvWarps) {
    // ---------------- converter warps: global fp32 -> registers -> bf16 b1/b2 tiles ----------------
    ConvArgs ca;
    ca.prefetch_iters = ablate >> 8;     // BB_GRAM_PREFETCH, default 0: L2 hints make it slower, from these warps or a dedicated one
    ca.early = (ablate & 128) == 0;      // default on (3.20 -> 3.08 ms at cfg4); BB_GRAM_ABLATE=128 restores loads-after-arrive
    ca.row_end = it_end * kStageRows < n ? it_end * kStageRows : n;
    ca.x = x; ca.y = y; ca.n = n; ca.d = d; ca.feat_a = feat_a; ca.feat_b = feat_b;
    ca.row_begin = row_begin; ca.n_iters = n_iters; ca.warp = warp; ca.lane = lane;
    ca.want_yty = diag && ij.x == 0 && rank == 0;
    if (!diag) converter_loop<false, false>(sm, ca);
    else if (y == nullptr) converter_loop<true, false>(sm, ca);
    else converter_loop<true, true>(sm, ca);
  } else if (warp < kMmaWarp) {
    // ---------------- epilogue warps: TMEM fp32 -> fp32 partial block (RMW through L2) ----------------
    const int q = warp & 3;                     // TMEM lane quadrant this warp may access
    const int n_intervals = (n_iters + kFlushIters - 1) / kFlushIters;
    // partial block layout [256 columns][128 rows]: for a fixed column the 32 lanes of a warp
    // touch 128 contiguous bytes (one fully used line per access)
    float* my_partial = partial + (static_cast<int64_t>(pair) * 2 + rank) * kHalf * kBlock + q * 32 + lane;
    if (n_intervals == 0) {
      for (int c = 0; c < kBlock; ++c) my_partial[c * kHalf] = 0.f;
    }
    for (int interval = 0; interval < n_intervals; ++interval) {
      const int buf = interval & 1;
      ptx::mbar_wait_sleep(&sm.acc_full[buf], (interval >> 1) & 1);
      ptx::tc_fence_after_sync();
      const uint32_t t_addr = tmem + (static_cast<uint32_t>(q * 32) << 16) + buf * kBlock;
#pragma unroll 1
      for (int cc = 0; cc < kBlock / 32; ++cc) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(t_addr + cc * 32, v);
        float* dst = my_partial + cc * 32 * kHalf;
        float old[32];
        if (interval != 0 && !(ablate & 32)) {
#pragma unroll
          for (int j = 0; j < 32; ++j) old[j] = dst[j * kHalf];
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) old[j] = 0.f;
        }
        ptx::tmem_wait_ld();
        if (!(ablate & 32)) {
#pragma unroll
          for (int j = 0; j < 32; ++j) dst[j * kHalf] = old[j] + __uint_as_float(v[j]);
        }
      }
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster_relaxed(&sm.acc_empty[buf], 0);   // TMEM reads done (wait::ld above)
    }
  } else if (rank == 0) {
    // ---------------- MMA issuer (leader CTA, one elected thread) ----------------
    if (ptx::elect_one()) {
      for (int it = 0; it < n_iters; ++it) {
        const int s = it % kStages;
        const int interval = it / kFlushIters;
        const int buf = interval & 1;
        const bool first = (it % kFlushIters) == 0;
        if (first) mbar_wait_cluster(&sm.acc_empty[buf], ((interval >> 1) & 1) ^ 1);
        ptx::mbar_wait(&sm.full[s], (it / kStages) & 1);   // the data is read by the async proxy, not by this thread
        ptx::tc_fence_after_sync();
        const uint32_t stage_addr = ptx::smem_u32(sm.stage[s]);
        const uint32_t d_tmem = tmem + buf * kBlock;
#pragma unroll
        for (int ks = 0; ks < kStageRows / 16; ++ks) {
          // one K = 16 step = two 8-row k groups (SBO = 1 KB apart); 64-feature MN blocks LBO = 4 KB apart
          const uint32_t base = stage_addr + ks * 2048;
          const uint64_t a1 = ptx::make_smem_desc(base, 4096, 1024, ptx::kLayoutSwizzle128B);
          const uint64_t a2 = ptx::make_smem_desc(base + kTileBytes, 4096, 1024, ptx::kLayoutSwizzle128B);
          const uint64_t b1 = alias_b ? a1 : ptx::make_smem_desc(base + 2 * kTileBytes, 4096, 1024, ptx::kLayoutSwizzle128B);
          const uint64_t b2 = alias_b ? a2 : ptx::make_smem_desc(base + 3 * kTileBytes, 4096, 1024, ptx::kLayoutSwizzle128B);
          if (ablate & 1) continue;
          if (diag) {
            // a1^T a1 + a1^T (2 a2) = P + 2 Q: the finalize takes (S + S^T) / 2 = P + Q + Q^T, i.e. the same three
            // products as an off-diagonal block for two MMAs
            mma_bf16_pair_a_fill(d_tmem, a1, a1, kIdesc, (first && ks == 0) ? 0u : 1u);
            mma_bf16_pair_a_lastuse(d_tmem, a1, a2, kIdesc, 1u);
            continue;
          }
          if (!(ablate & 64)) {               // A-collector hints: the second MMA takes A = a1 from the collector buffer, not from
                                              // shared memory (measured 3.200 -> 3.15 ms at cfg4; BB_GRAM_ABLATE=64 switches it off)
            mma_bf16_pair_a_fill(d_tmem, a1, b1, kIdesc, (first && ks == 0) ? 0u : 1u);
            mma_bf16_pair_a_lastuse(d_tmem, a1, b2, kIdesc, 1u);
            mma_bf16_pair(d_tmem, a2, b1, kIdesc, 1u);
            continue;
          }
          mma_bf16_pair(d_tmem, a1, b1, kIdesc, (first && ks == 0) ? 0u : 1u);
          if (ablate & 16) continue;
          mma_bf16_pair(d_tmem, a1, b2, kIdesc, 1u);
          mma_bf16_pair(d_tmem, a2, b1, kIdesc, 1u);
        }
        mma_commit_pair(&sm.empty[s]);
        if ((it % kFlushIters) == kFlushIters - 1 || it == n_iters - 1) mma_commit_pair(&sm.acc_full[buf]);
      }
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (y != nullptr && diag) {
    if (threadIdx.x < kHalf && feat_a + static_cast<int>(threadIdx.x) < d)
      partial_xty[static_cast<int64_t>(split) * d + feat_a + threadIdx.x] = sm.xty[threadIdx.x];
    if (ij.x == 0 && rank == 0 && threadIdx.x == 0) partial_yty[split] = sm.yty;
  }
  cluster_sync_all();          // the peer's shared memory / barriers stay alive until both are done
  if (warp == kMmaWarp) tmem_dealloc_pair(tmem, kTmemCols);
#ifdef BB_GRAM_TIMELINE
  if (threadIdx.x == 0 && rank == 0) {
    unsigned long long tl_end;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tl_end));
    printf("gram pair %3d %s (%d,%d) split %d/%d stages %d  begin %llu  us %.1f  ns/stage %.0f\n", pair,
           diag ? "diag" : "off ", ij.x, ij.y, split, n_splits, n_iters, tl_begin % 100000000ull,
           (tl_end - tl_begin) * 1e-3, n_iters > 0 ? static_cast<double>(tl_end - tl_begin) / n_iters : 0.0);
  }
#endif
}

// XtX[d,e] (float64, full symmetric matrix) = sum over splits of the upper-triangle block partials (a diagonal
// block holds S = P + 2 Q, its element is (S[r,c] + S[c,r]) / 2); Xty / yty likewise.  One thread per output element.
__global__ void __launch_bounds__(256)
gram_finalize_kernel(const float* __restrict__ partial, const double* __restrict__ partial_xty,
                     const double* __restrict__ partial_yty, int d, int splits_diag, int splits_off,
                     double* __restrict__ xtx, double* __restrict__ xty, double* __restrict__ yty) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int nb = (d + kBlock - 1) / kBlock;
  const int n_off = nb * (nb - 1) / 2;
  if (idx < static_cast<int64_t>(d) * d) {
    const int row = static_cast<int>(idx / d), col = static_cast<int>(idx % d);
    const int r = row < col ? row : col, c = row < col ? col : row;
    const int bi = r / kBlock, bj = c / kBlock;
    const int rr = r % kBlock, cc = c % kBlock;
    auto at = [&](int64_t pair, int a, int b) {
      return static_cast<double>(partial[((pair * 2 + a / kHalf) * kBlock + b) * kHalf + a % kHalf]);
    };
    double acc = 0.0;
    if (bi == bj) {
      for (int s = 0; s < splits_diag; ++s) {
        const int64_t pair = static_cast<int64_t>(s) * nb + bi;
        acc += at(pair, rr, cc) + at(pair, cc, rr);
      }
      acc *= 0.5;
    } else {
      const int off = bi * (nb - 1) - bi * (bi - 1) / 2 + (bj - bi - 1);
      for (int s = 0; s < splits_off; ++s)
        acc += at(static_cast<int64_t>(nb) * splits_diag + static_cast<int64_t>(s) * n_off + off, rr, cc);
    }
    xtx[idx] = acc;
  }
  if (xty != nullptr && idx < d) {
    double acc = 0.0;
    for (int s = 0; s < splits_diag; ++s) acc += partial_xty[static_cast<int64_t>(s) * d + idx];
    xty[idx] = acc;
  }
  if (yty != nullptr && idx == 0) {
    double acc = 0.0;
    for (int s = 0; s < splits_diag; ++s) acc += partial_yty[s];
    *yty = acc;
  }
}

struct GramGrid {
  int nb, n_off, splits_diag, splits_off;
  int64_t pairs() const { return static_cast<int64_t>(nb) * splits_diag + static_cast<int64_t>(n_off) * splits_off; }
};

// Row splits per block kind.  Default: the same number of splits for every block, so that all pairs of a split walk
// the same rows at the same time (D = 1024: 7 x 10 = 70 pairs).  BB_GRAM_BALANCE=1 (timing experiments) weighs the
// splits by MMA count instead -- a diagonal block issues 2 products per K step, an off-diagonal one 3; pick
// (splits_diag, splits_off) with nb * sd + n_off * so <= SMs / 2 minimising max(2 / sd, 3 / so): D = 1024 gives
// 4 x 6 + 6 x 8 = 72 pairs.  Measured SLOWER (3.46 against 2.87 ms): a pair's pace is ~630 ns per 32-row stage
// whatever its MMA count (profiles/r02_gram_pair_timeline.txt), so the longer diagonal row ranges just take longer,
// and the diagonal and off-diagonal read fronts no longer share X through L2.
GramGrid plan_gram(int64_t n, int d) {
  GramGrid g;
  g.nb = (d + kBlock - 1) / kBlock;
  g.n_off = g.nb * (g.nb - 1) / 2;
  int sms = device_sm_count();
  if (sms <= 0) sms = 148;
  const int P = sms / 2;
  const int64_t iters = std::max<int64_t>(1, (n + kStageRows - 1) / kStageRows);
  const int cap = static_cast<int>(std::min<int64_t>(iters, P));
  static const bool balance = getenv("BB_GRAM_BALANCE") && atoi(getenv("BB_GRAM_BALANCE")) != 0;
  const int equal = std::max(1, std::min(cap, P / (g.nb + g.n_off)));
  g.splits_diag = g.splits_off = equal;
  if (!balance) return g;
  if (g.n_off == 0) {
    g.splits_diag = std::max(1, std::min(cap, P / g.nb));
    return g;
  }
  int64_t best_num = 3, best_den = equal;            // cost of the equal plan: 3 / equal
  for (int so = 1; so <= cap; ++so) {
    const int left = P - g.n_off * so;
    if (left < g.nb) break;
    const int sd = std::min(cap, left / g.nb);
    // cost = max(2 / sd, 3 / so) as a fraction
    int64_t num = 2, den = sd;
    if (3 * static_cast<int64_t>(sd) > 2 * static_cast<int64_t>(so)) {
      num = 3;
      den = so;
    }
    if (num * best_den < best_num * den) {
      best_num = num;
      best_den = den;
      g.splits_off = so;
      // no more diagonal splits than the cost needs: fewer read fronts over X
      int sd_min = static_cast<int>((2 * den + num - 1) / num);
      g.splits_diag = std::max(1, std::min(sd, sd_min));
    }
  }
  return g;
}

}  // namespace

bool gram_tc_supported(int64_t n, int d, const void* x) {
  return n > 0 && d > 64 && d % 4 == 0 && d <= 4096 && reinterpret_cast<uintptr_t>(x) % 16 == 0;
}

int64_t gram_tc_workspace(int64_t n, int d) {
  const GramGrid g = plan_gram(n, d);
  return g.pairs() * 2 * kHalf * kBlock * static_cast<int64_t>(sizeof(float)) +
         static_cast<int64_t>(g.splits_diag) * (d + 1) * static_cast<int64_t>(sizeof(double)) + 1024;
}

// xtx [d, d] float64 (required); xty [d], yty [1] float64 (both or neither, with y).
int launch_gram_tc(const float* x, const float* y, int64_t n, int d, double* xtx, double* xty,
                   double* yty, void* workspace, int64_t workspace_bytes, cudaStream_t stream) {
  if (!gram_tc_supported(n, d, x)) {
    set_error("gram_tc: unsupported shape n=%lld d=%d", static_cast<long long>(n), d);
    return BB_ERR_UNSUPPORTED;
  }
  if ((y == nullptr) != (xty == nullptr) || (y == nullptr) != (yty == nullptr)) {
    set_error("gram_tc: y, xty and yty must be given together");
    return BB_ERR_INVALID;
  }
  if (workspace == nullptr || workspace_bytes < gram_tc_workspace(n, d)) {
    set_error("gram_tc: workspace %lld < %lld bytes", static_cast<long long>(workspace_bytes),
              static_cast<long long>(gram_tc_workspace(n, d)));
    return BB_ERR_WORKSPACE;
  }
  const GramGrid g = plan_gram(n, d);
  const int64_t pairs = g.pairs();
  uint8_t* ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~static_cast<uintptr_t>(255));
  float* partial = reinterpret_cast<float*>(ws);
  ws += pairs * 2 * kHalf * kBlock * sizeof(float);
  double* partial_xty = reinterpret_cast<double*>(ws);
  ws += static_cast<int64_t>(g.splits_diag) * d * sizeof(double);
  double* partial_yty = reinterpret_cast<double*>(ws);
  const int smem_bytes = static_cast<int>(sizeof(SmemLayout));
  static SmemOptIn smem_opt_in_0;
  BB_CUDA_OK(smem_opt_in_0.ensure(gram_pair_kernel, smem_bytes));
  const int grid = static_cast<int>(2 * pairs);
  // developer ablation switches (timing experiments only; results are wrong when set)
  static const int ablate = ((getenv("BB_GRAM_ABLATE") ? atoi(getenv("BB_GRAM_ABLATE")) : 0) & 0xff) |
                            ((getenv("BB_GRAM_PREFETCH") ? atoi(getenv("BB_GRAM_PREFETCH")) : 0) << 8);
  gram_pair_kernel<<<grid, kThreads, smem_bytes, stream>>>(x, y, n, d, g.splits_diag, g.splits_off, ablate, partial,
                                                           partial_xty, partial_yty);
  BB_CHECK_LAUNCH("gram_pair_kernel");
  const int64_t total = static_cast<int64_t>(d) * d;
  gram_finalize_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, stream>>>(
      partial, partial_xty, partial_yty, d, g.splits_diag, g.splits_off, xtx, xty, yty);
  BB_CHECK_LAUNCH("gram_finalize_kernel");
  return BB_OK;
}

}  // namespace bb
