// Responsibility-weighted sufficient statistics as ONE tensor-core contraction over the data axis
// (D % 8 == 0, D <= 64, K % 4 == 0, K <= 256):
//     Nk[k] = sum_n r_nk      Srx[k,d] = sum_n r_nk x_nd      Srxx[k,d,e] = sum_n r_nk x_nd x_ne
//
// What it replaces: the reference's plan for sum_n R[n,k] X[n,d] X[n,e],
//     _tensordot(_mul(_dimshuffle(R,1,'x',0), _dimshuffle(X,'x',1,0)), X, [2],[0])
// (bayesic/algebra.py:741-765 -> 1297-1306 -> 1347-1351), which groups the
// factors as (R (x) X) . X and materialises a K x D x N tensor.  The same einsum
// (algebra.py:314-346 declares only the index pattern) regrouped as  R^T . (X (x) X):
//     Srxx[k, (d,e)] = sum_n R[n,k] Phi[n,(d,e)],      Phi[n,(d,e)] = x_nd x_ne
// is a plain GEMM with M = K, contraction over n, whose B operand Phi does not depend on k -- so
// it is generated once per row instead of once per (row, component) -- and whose symmetric half
// (d <= e by 8 x 8 feature blocks: 36 blocks of 64 columns at D = 64) is all that is needed.
// weighted_sm100.cu (the (r x) . x grouping, TF32, N = 64 tiles) is ALU-bound on forming the A
// operand; this grouping is tensor-pipe bound:
//   * CTA = (column tile of up to 4 pair blocks = 256 columns, contiguous range of rows);
//     M = K components in up to two 128-lane blocks -> 2 x 256 TMEM columns (all of TMEM);
//   * error-compensated BF16 (r = r1 + r2, phi = p1 + p2; r1 p1 + r1 p2 + r2 p1; see
//     gram_sm100.cu), both operands MN-major SWIZZLE_128B as they are produced: 16 converter warps
//     load R with coalesced 128-bit loads, form Phi in fp32 from the row of X, split, store;
//   * the linear statistics ride along as one more 64-column block (Phi = x_d), Nk in the
//     converter warps of the first column tile;
//   * FP32 TMEM accumulation drained every 2048 rows into the CTA's fp32 partial block
//     (coalesced read-modify-write); finalize adds the row ranges in float64 and unpacks the
//     symmetric blocks.
// Algorithmic traffic: 4 (K + D) bytes/row; flops issued: 3 x 2 K x 64 x (#blocks) per row.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"
#include "sm100_pair.cuh"
#include "sm100_ptx.cuh"

namespace bb {

namespace {

constexpr int kStageRows = 16;
constexpr int kStages = 6;
constexpr int kMaxK = 256;                       // components (MMA M, two 128-lane blocks)
constexpr int kTileCols = 256;                   // pair-feature columns per CTA (MMA N)
constexpr int kRPart = kStageRows * kMaxK * 2;   // 8 KB: R tile, one bf16 part
constexpr int kPPart = kStageRows * kTileCols * 2;   // 8 KB: Phi tile, one bf16 part
constexpr int kStageBytes = 2 * kRPart + 2 * kPPart; // 32 KB
constexpr int kFlushIters = 128 / BB_CHAIN_DIV;                 // 2048 rows per TMEM accumulation chain
constexpr int kConvWarps = 16;
constexpr int kConvGroups = 2;
constexpr int kEpiWarps = 4;
constexpr int kMmaWarp = kConvWarps + kEpiWarps;
constexpr int kThreads = (kMmaWarp + 1) * 32;
constexpr int kProdWarp = kMmaWarp + 1;          // pre-split R mode only: one more warp, its elected lane issues the bulk copies
constexpr int kThreadsSplit = kThreads + 32;
constexpr int kTmemCols = 512;

struct __align__(1024) SmemLayout {
  uint8_t stage[kStages][kStageBytes];
  uint64_t full[kStages];
  uint64_t empty[kStages];
  uint64_t acc_full;
  uint64_t acc_empty;
  uint32_t tmem_base;
};

__device__ __forceinline__ void mma_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same MMA with an A-operand collector hint: the second of two consecutive MMAs that share A may take it from the
// tensor core's collector buffer instead of fetching it from shared memory again (the port is the bottleneck here)
__device__ __forceinline__ void mma_bf16_ss_a_fill(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                                   uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16.collector::a::fill [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_bf16_ss_a_lastuse(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                                      uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
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
__device__ __forceinline__ float4 ldg_f4_l1(const float* p) {      // through L1 (allocating): for lines an L1 prefetch brought in
  float4 v;
  asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];"
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
__device__ __forceinline__ void split_bf16(const float4& x, uint32_t (&b1)[2], uint32_t (&b2)[2]) {
  __nv_bfloat162 p0 = __floats2bfloat162_rn(x.x, x.y);
  __nv_bfloat162 p1 = __floats2bfloat162_rn(x.z, x.w);
  b1[0] = *reinterpret_cast<uint32_t*>(&p0);
  b1[1] = *reinterpret_cast<uint32_t*>(&p1);
  const float rx = x.x - __uint_as_float(b1[0] << 16);
  const float ry = x.y - __uint_as_float(b1[0] & 0xFFFF0000u);
  const float rz = x.z - __uint_as_float(b1[1] << 16);
  const float rw = x.w - __uint_as_float(b1[1] & 0xFFFF0000u);
  __nv_bfloat162 q0 = __floats2bfloat162_rn(rx, ry);
  __nv_bfloat162 q1 = __floats2bfloat162_rn(rz, rw);
  b2[0] = *reinterpret_cast<uint32_t*>(&q0);
  b2[1] = *reinterpret_cast<uint32_t*>(&q1);
}

// Column blocks of 64: block b < n_pair is the feature-group pair (ga, gb), ga <= gb, enumerated
// row-major over the upper triangle; column (i, j) -> x[8 ga + i] * x[8 gb + j].  Block n_pair is
// the linear block: column c -> x[c] (c < d).  Block n_pair + 1 is the ones block: column 0 -> 1
// (its contraction with R is Nk; at D = 64 the 38 blocks fill 10 tiles of N = 256 that are
// issued in full anyway, so Nk costs no extra MMA).
struct Geometry {
  int d, k, ldr, n_groups, n_pair, n_blocks, n_tiles;      // ldr: row stride of R (>= k: a launch may cover a slice of the components)
  int base, rem;      // tile t holds base + (t < rem) blocks, starting at block t * base + min(t, rem)
};

__host__ __device__ __forceinline__ int tile_first_block(const Geometry& g, int t) {
  return t * g.base + (t < g.rem ? t : g.rem);
}
__host__ __device__ __forceinline__ int tile_block_count(const Geometry& g, int t) { return g.base + (t < g.rem ? 1 : 0); }

__device__ __forceinline__ void pair_of(int b, int n_groups, int* ga, int* gb) {
  int a = 0, rem = b;
  while (rem >= n_groups - a) {
    rem -= n_groups - a;
    ++a;
  }
  *ga = a;
  *gb = a + rem;
}

// grid = n_tiles * n_splits: CTA -> (tile = blockIdx / n_splits, row range = blockIdx % n_splits).
// The blocks are spread evenly over the tiles (3 or 4 each at D = 64): every CTA converts the same
// R rows whatever its column count, so equal row ranges keep the CTAs in step.
// kMode 0: r = responsibilities R[n, k] (float32).  kMode 1: r = logits, the weights are exp(logit - lse[row]).
// kMode 2: r = the responsibilities ALREADY split into the BF16 (b1 | b2) operand tiles, one 16 KB image per
// 16-row stage in exactly the shared-memory layout of a stage's R part (written by softmax_rows_split_kernel,
// mixture_kernels.cu): a producer thread bulk-copies them (cp.async.bulk) and the converter warps only form Phi.
// Converting R in every one of the ten column-tile CTAs of a row range is 45 % of the converters' instructions,
// and the converters, not the tensor pipe, bound modes 0 and 1 (ablation in DESIGN.md 4.3b).
template <int kMode>
__global__ void __launch_bounds__(kMode == 2 ? kThreadsSplit : kThreads, 1)
weighted_pairs_kernel(const float* __restrict__ x, const float* __restrict__ r, const float* __restrict__ lse,
                      int64_t n, Geometry g, int n_splits, int prefetch_iters,
                      float* __restrict__ partial,        // [cta][2 m-blocks][256 cols][128 lanes]
                      double* __restrict__ /*unused*/) {
  constexpr bool kFromLogits = kMode == 1;
  constexpr bool kSplitR = kMode == 2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  SmemLayout& sm = *reinterpret_cast<SmemLayout*>(smem_raw);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tile = blockIdx.x / n_splits;
  const int split = blockIdx.x % n_splits;
  const int block0 = tile_first_block(g, tile);
  const int tile_blocks = tile_block_count(g, tile);
  const int m_blocks = (g.k + 127) / 128;
  const int64_t total_iters = (n + kStageRows - 1) / kStageRows;
  const int64_t it_begin = total_iters * split / n_splits;
  const int64_t it_end = total_iters * (split + 1) / n_splits;
  const int n_iters = static_cast<int>(it_end - it_begin);
  const int64_t row_begin = it_begin * kStageRows;
  const int64_t row_end = min(n, it_end * kStageRows);

  if (warp == kMmaWarp) {
    if (lane == 0) {
      for (int s = 0; s < kStages; ++s) {
        ptx::mbar_init(&sm.full[s], kConvWarps / kConvGroups + (kSplitR ? 1 : 0));
        ptx::mbar_init(&sm.empty[s], 1);
      }
      ptx::mbar_init(&sm.acc_full, 1);
      ptx::mbar_init(&sm.acc_empty, kEpiWarps);
      ptx::fence_mbar_init();
    }
    __syncwarp();
    ptx::tmem_alloc(&sm.tmem_base, kTmemCols);
  }
  // Every tile issues N = 256 MMAs so that all CTAs of a row range run at the same pace and share
  // the R rows through L2; the unused blocks of a 3-block tile are never written and stay zero.
  for (int i = threadIdx.x; i < kStages * kStageBytes / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(&sm.stage[0][0])[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem = sm.tmem_base;

  if (warp < kConvWarps) {
    // ---------------- converter warps ----------------
    // group = warp & 1 takes every 2nd stage; warp wi of the group owns rows 2 wi, 2 wi + 1.
    // R: lane l holds components [4 l + 128 h, +4), h = 0, 1, of both rows.
    // Phi: item (row j, block slot t2, lane) -> 4 columns: 16 lanes per 64-column block row
    //      (i = (lane & 15) >> 1, j0 = 4 (lane & 1)); lanes >= 16 take the odd block of a pair.
    // The SM's issue slots are the budget here (the first version spent ~2 300 warp instructions
    // per stage against ~3 000 slots in the stage's MMA time), so: every per-lane quantity is
    // hoisted, the products are formed as (x_a m_a + c_a) * x_b so that pair / linear / ones /
    // padding columns share one code path, and the ragged tail is a separate (cold) path.
    const int group = warp & (kConvGroups - 1);
    const int wi = warp / kConvGroups;
    uint32_t off[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int kk = 2 * wi + j;      // row within the stage = MMA k index
      off[j] = (kk >> 3) * 1024 + (kk & 7) * 128 + ((((lane & 15) >> 1) ^ (kk & 7)) << 4) + (lane & 1) * 8;
    }
    // left factor a = x[src_a] * ma + ca; right factors x[src_b .. +3] * mb + (cb, 0, 0, 0):
    //   pair block   ma = 1, ca = 0, mb = 1, cb = 0        (x_d x_e)
    //   linear block ma = 0, ca = 1, mb = 1 (0 past d)     (x_e)
    //   ones block   ma = 0, ca = 1, mb = 0, cb = 1 on column 0  (1: gives Nk)
    //   unused slot  all zero
    int src_a[2], src_b[2];
    float ma[2], ca[2], mb[2], cb[2];
#pragma unroll
    for (int t2 = 0; t2 < 2; ++t2) {
      const int bl = 2 * t2 + (lane >> 4);
      const int b = block0 + bl;
      const int i = (lane & 15) >> 1, j0 = (lane & 1) * 4;
      src_a[t2] = 0; src_b[t2] = 0;
      ma[t2] = 0.f; ca[t2] = 0.f; mb[t2] = 0.f; cb[t2] = 0.f;
      if (bl < tile_blocks) {
        if (b < g.n_pair) {
          int ga, gb;
          pair_of(b, g.n_groups, &ga, &gb);
          src_a[t2] = 8 * ga + i;
          src_b[t2] = 8 * gb + j0;
          ma[t2] = 1.f;
          mb[t2] = 1.f;
        } else if (b == g.n_pair) {
          ca[t2] = 1.f;
          if (8 * i + j0 < g.d) {        // linear block: column c = 8 i + j0 .. +3
            src_b[t2] = 8 * i + j0;
            mb[t2] = 1.f;
          }
        } else {                         // ones block: column 0 = 1
          ca[t2] = 1.f;
          cb[t2] = (lane & 15) == 0 ? 1.f : 0.f;
        }
      }
    }
    const uint32_t stage0 = ptx::smem_u32(sm.stage[0]);
    int64_t row0 = row_begin + static_cast<int64_t>(group) * kStageRows + 2 * wi;
    const int kc0 = min(4 * lane, g.k - 4), kc1 = min(4 * lane + 128, g.k - 4);
    const float km0 = 4 * lane < g.k ? 1.f : 0.f, km1 = 4 * lane + 128 < g.k ? 1.f : 0.f;
    const bool k_partial = g.k < kMaxK;                  // some lanes hold components past k
    // walking pointers to this warp's first row of the next stage it converts
    const float* rp = r + row0 * g.ldr;
    const float* xp = x + row0 * g.d;
    const float* lp = kFromLogits ? lse + row0 : nullptr;
    const int64_t r_step = static_cast<int64_t>(kConvGroups) * kStageRows * g.ldr;
    const int64_t x_step = static_cast<int64_t>(kConvGroups) * kStageRows * g.d;
    float4 rr[2][2];
    float row_lse[2] = {0.f, 0.f};   // logits mode: r = exp(logit - lse[row])
    float xa[2][2];                  // [row][block slot]
    float4 xb[2][2];
    bool rows_ok = true;             // both rows of the prefetched stage are inside [0, n)
    // everything a stage needs is loaded one iteration ahead (nothing is fetched between the
    // barrier wait and the stores: dependent loads there serialise on the L2 latency)
    const bool use_l1 = (prefetch_iters & 0x200) != 0;     // BB_WP_L1=1: prefetch into L1 and load through it
    // part 1: R (+ lse) of the next stage this warp converts; part 2: X of that stage, then advance the pointers
    // timing experiments (BB_WP_ABLATE; results are WRONG when set): 1 no global loads, 2 no converter stores,
    // 4 no MMAs, 8 no proxy fence
    const int ablate = prefetch_iters >> 12;
    auto load_part = [&](int part) {
      if (ablate & 1) {
        if (part == 2) row0 += kConvGroups * kStageRows;
        return;
      }
      const bool ok = row0 + 2 <= n;
      // ragged last stage: a row past the end re-reads row n - 1 (its weights are zeroed later)
      int jj[2] = {0, 1};
      if (!ok) {
        jj[0] = row0 < n ? 0 : static_cast<int>((n - 1) - row0);
        jj[1] = row0 + 1 < n ? 1 : static_cast<int>((n - 1) - row0);
      }
      if (part == 1) {
        if constexpr (!kSplitR) {
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const float* rrow = rp + jj[j] * g.ldr;
            rr[j][0] = use_l1 ? ldg_f4_l1(rrow + kc0) : ldg_f4(rrow + kc0);
            rr[j][1] = use_l1 ? ldg_f4_l1(rrow + kc1) : ldg_f4(rrow + kc1);
            if (kFromLogits) row_lse[j] = __ldg(lp + jj[j]);
          }
        }
        return;
      }
      rows_ok = ok;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float* xrow = xp + jj[j] * g.d;
#pragma unroll
        for (int t2 = 0; t2 < 2; ++t2) {
          xa[j][t2] = __ldg(xrow + src_a[t2]);
          xb[j][t2] = __ldg(reinterpret_cast<const float4*>(xrow + src_b[t2]));
        }
      }
      // L2 / L1 prefetch of the rows this warp converts `pf` iterations from now (BB_WP_PREFETCH): 2 rows x
      // (k * 4 / 128 lines of R) -- lanes 0-15 take one 128-byte line each at K = 256.  Measured neutral.
      const int pf = kSplitR ? 0 : (prefetch_iters & 0xff);
      if (pf > 0) {
        const int64_t prow = row0 + static_cast<int64_t>(pf) * kConvGroups * kStageRows + (lane >> 4);
        const int col = (lane & 15) * 32;
        if (prow < row_end && col < g.k) {
          if (use_l1) asm volatile("prefetch.global.L1 [%0];" ::"l"(r + prow * g.ldr + col));
          else asm volatile("prefetch.global.L2 [%0];" ::"l"(r + prow * g.ldr + col));
        }
      }
      row0 += kConvGroups * kStageRows;
      rp += r_step;
      xp += x_step;
      if (kFromLogits) lp += kConvGroups * kStageRows;
    };
    auto load_all = [&]() {
      load_part(1);
      load_part(2);
    };
    // Default (BB_WP_EARLY=0 restores the old order): issue the next stage's loads as soon as the registers they
    // land in are free (R right after the R tile is stored, X right after the Phi tile) instead of after the
    // hand-over: when the converters are the bottleneck the stage barrier never makes them wait, so loads issued
    // after the arrive are needed at once and their whole latency is exposed; issued early it overlaps the Phi
    // conversion and the proxy fence (measured at 1 Mi rows, K = 256: 3.55 -> 3.30 ms)
    const bool early = (prefetch_iters & 0x400) != 0;
#ifdef BB_WP_TIMELINE
    long long tl[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tl_prev = clock64();
#define BB_WP_TL(i) { const long long now_ = clock64(); tl[i] += now_ - tl_prev; tl_prev = now_; }
#else
#define BB_WP_TL(i)
#endif
    if (group < n_iters) load_all();
    for (int it = group; it < n_iters; it += kConvGroups) {
      const int s = it % kStages;
      BB_WP_TL(0)
      ptx::mbar_wait(&sm.empty[s], ((it / kStages) & 1) ^ 1);
      BB_WP_TL(1)
      const uint32_t stage_addr = stage0 + s * kStageBytes;
      if (kFromLogits) {                                  // responsibilities from logits, on the fly
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            rr[j][h].x = __expf(rr[j][h].x - row_lse[j]);
            rr[j][h].y = __expf(rr[j][h].y - row_lse[j]);
            rr[j][h].z = __expf(rr[j][h].z - row_lse[j]);
            rr[j][h].w = __expf(rr[j][h].w - row_lse[j]);
          }
      }
      if (!kSplitR) {
        if (!rows_ok || k_partial) {                        // cold: zero what lies past n or past k
          const int64_t first = row0 - kConvGroups * kStageRows;   // row0 was advanced by load_all
  #pragma unroll
          for (int j = 0; j < 2; ++j) {
            const float rowm = first + j < n ? 1.f : 0.f;
            const float m0 = rowm * km0, m1 = rowm * km1;
            rr[j][0] = make_float4(rr[j][0].x * m0, rr[j][0].y * m0, rr[j][0].z * m0, rr[j][0].w * m0);
            rr[j][1] = make_float4(rr[j][1].x * m1, rr[j][1].y * m1, rr[j][1].z * m1, rr[j][1].w * m1);
          }
        }
        // ---- R tile: [m block (64 components) 2 KB][k group][8][128 B] ----
  #pragma unroll
        for (int j = 0; j < 2; ++j)
  #pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t b1[2], b2[2];
            split_bf16(rr[j][h], b1, b2);
            const uint32_t addr = stage_addr + (2 * h + (lane >> 4)) * 2048 + off[j];
            if (ablate & 2) continue;
            sts_u2(addr, b1[0], b1[1]);
            sts_u2(addr + kRPart, b2[0], b2[1]);
          }
      }
      BB_WP_TL(2)
      const bool more = it + kConvGroups < n_iters;
      if (early && more) load_part(1);
      BB_WP_TL(3)
      // ---- Phi tile: [block (64 columns) 2 KB][k group][8][128 B] ----
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int t2 = 0; t2 < 2; ++t2) {
          const float a = fmaf(xa[j][t2], ma[t2], ca[t2]);
          const float am = a * mb[t2];
          const float4 p = make_float4(fmaf(am, xb[j][t2].x, a * cb[t2]), am * xb[j][t2].y, am * xb[j][t2].z,
                                       am * xb[j][t2].w);
          uint32_t b1[2], b2[2];
          split_bf16(p, b1, b2);
          const uint32_t addr = stage_addr + 2 * kRPart + (2 * t2 + (lane >> 4)) * 2048 + off[j];
          if (ablate & 2) continue;
          sts_u2(addr, b1[0], b1[1]);
          sts_u2(addr + kPPart, b2[0], b2[1]);
        }
      BB_WP_TL(4)
      if (early && more) load_part(2);
      BB_WP_TL(5)
      if (!(ablate & 8)) fence_proxy_async_smem();
      __syncwarp();
      BB_WP_TL(6)
      if (lane == 0) ptx::mbar_arrive(&sm.full[s]);
      if (!early && more) load_all();
      BB_WP_TL(7)
    }
#ifdef BB_WP_TIMELINE
    if (blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 15)) {
      const int its = (n_iters - group + kConvGroups - 1) / kConvGroups;
      printf("converter warp %d (CTA 0, %d of its stages), clock64 cycles per stage: loop top %lld  wait empty %lld  "
             "R convert+store %lld  R loads issued %lld  Phi convert+store %lld  X loads issued %lld  fence+syncwarp %lld  "
             "arrive %lld | total %lld\n", warp, its, tl[0] / its, tl[1] / its, tl[2] / its, tl[3] / its, tl[4] / its,
             tl[5] / its, tl[6] / its, tl[7] / its, (tl[0] + tl[1] + tl[2] + tl[3] + tl[4] + tl[5] + tl[6] + tl[7]) / its);
    }
#endif
  } else if (warp < kMmaWarp) {
    // ---------------- epilogue warps: TMEM fp32 -> fp32 partial (coalesced RMW), single-buffered ----------------
    const int qd = warp & 3;
    const int n_intervals = (n_iters + kFlushIters - 1) / kFlushIters;
    const int cols = tile_blocks * 64;
    float* my_partial = partial + static_cast<int64_t>(blockIdx.x) * 2 * kTileCols * 128 + qd * 32 + lane;
    if (n_intervals == 0) {
      for (int c = 0; c < 2 * kTileCols; ++c) my_partial[c * 128] = 0.f;
    }
    for (int interval = 0; interval < n_intervals; ++interval) {
      ptx::mbar_wait_sleep(&sm.acc_full, interval & 1);
      ptx::tc_fence_after_sync();
      for (int mb = 0; mb < m_blocks; ++mb) {
        const uint32_t t_addr = tmem + (static_cast<uint32_t>(qd * 32) << 16) + mb * kTileCols;
#pragma unroll 1
        for (int cc = 0; cc < cols / 32; ++cc) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(t_addr + cc * 32, v);
          float* dst = my_partial + (mb * kTileCols + cc * 32) * 128;
          float old[32];
          if (interval != 0) {
#pragma unroll
            for (int j = 0; j < 32; ++j) old[j] = dst[j * 128];
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) old[j] = 0.f;
          }
          ptx::tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) dst[j * 128] = old[j] + __uint_as_float(v[j]);
        }
      }
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&sm.acc_empty);
    }
  } else if (kSplitR && warp == kProdWarp) {
    // ---------------- R producer (pre-split mode): one 16 KB bulk copy per stage ----------------
    if (ptx::elect_one()) {
      uint64_t keep;
      asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(keep));     // re-read by the other column tiles
      const uint8_t* src = reinterpret_cast<const uint8_t*>(r) + it_begin * (2 * kRPart);
      for (int it = 0; it < n_iters; ++it) {
        const int s = it % kStages;
        ptx::mbar_wait(&sm.empty[s], ((it / kStages) & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(&sm.full[s], 2 * kRPart);
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
            ::"r"(ptx::smem_u32(sm.stage[s])), "l"(src + static_cast<int64_t>(it) * (2 * kRPart)), "r"(2 * kRPart),
              "r"(ptx::smem_u32(&sm.full[s])), "l"(keep)
            : "memory");
      }
    }
  } else if (warp == kMmaWarp) {
    // ---------------- MMA issuer ----------------
    if (ptx::elect_one()) {
      const uint32_t idesc = ptx::make_idesc(128, kTileCols, /*bf16*/ 1, 1, 1);
#ifdef BB_WP_TIMELINE
      long long m_wait_acc = 0, m_wait_full = 0, m_t0 = clock64(), m_prev = m_t0;
#endif
      for (int it = 0; it < n_iters; ++it) {
        const int s = it % kStages;
        const int interval = it / kFlushIters;
        const bool first = (it % kFlushIters) == 0;
#ifdef BB_WP_TIMELINE
        m_prev = clock64();
#endif
        if (first && interval > 0) ptx::mbar_wait(&sm.acc_empty, (interval - 1) & 1);
#ifdef BB_WP_TIMELINE
        { const long long now_ = clock64(); m_wait_acc += now_ - m_prev; m_prev = now_; }
#endif
        ptx::mbar_wait(&sm.full[s], (it / kStages) & 1);
#ifdef BB_WP_TIMELINE
        { const long long now_ = clock64(); m_wait_full += now_ - m_prev; m_prev = now_; }
#endif
        ptx::tc_fence_after_sync();
        const uint32_t base = ptx::smem_u32(sm.stage[s]);
        const uint64_t p1 = ptx::make_smem_desc(base + 2 * kRPart, 2048, 1024, ptx::kLayoutSwizzle128B);
        const uint64_t p2 = ptx::make_smem_desc(base + 2 * kRPart + kPPart, 2048, 1024, ptx::kLayoutSwizzle128B);
        for (int mb = 0; mb < m_blocks; ++mb) {
          if ((prefetch_iters >> 12) & 4) break;
          const uint64_t r1 = ptx::make_smem_desc(base + mb * 4096, 2048, 1024, ptx::kLayoutSwizzle128B);
          const uint64_t r2 = ptx::make_smem_desc(base + kRPart + mb * 4096, 2048, 1024, ptx::kLayoutSwizzle128B);
          const uint32_t d_tmem = tmem + mb * kTileCols;
          if (prefetch_iters & 0x100) {        // developer switch BB_WP_COLLECTOR=1 (bit 8 of the packed knobs)
            mma_bf16_ss_a_fill(d_tmem, r1, p1, idesc, first ? 0u : 1u);
            mma_bf16_ss_a_lastuse(d_tmem, r1, p2, idesc, 1u);
          } else {
            mma_bf16_ss(d_tmem, r1, p1, idesc, first ? 0u : 1u);
            mma_bf16_ss(d_tmem, r1, p2, idesc, 1u);
          }
          mma_bf16_ss(d_tmem, r2, p1, idesc, 1u);
        }
        ptx::mma_commit(&sm.empty[s]);
        if ((it % kFlushIters) == kFlushIters - 1 || it == n_iters - 1) ptx::mma_commit(&sm.acc_full);
      }
#ifdef BB_WP_TIMELINE
      if (blockIdx.x == 0)
        printf("MMA issuer (CTA 0, %d stages), clock64 cycles per stage: waiting for a full stage %lld  waiting for the "
               "accumulator drain %lld | total %lld (768 = the six MMAs of a stage)\n", n_iters, m_wait_full / n_iters,
               m_wait_acc / n_iters, (clock64() - m_t0) / n_iters);
#endif
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == kMmaWarp) ptx::tmem_dealloc(tmem, kTmemCols);
}

// One thread per (k, d, e): add the row ranges of the CTA tile that owns the pair block (min, max),
// in float64; blocks (ga, ga) hold both (i, j) and (j, i) -- use i <= j so the result is symmetric.
__global__ void __launch_bounds__(256)
weighted_pairs_finalize_kernel(const float* __restrict__ partial, const double* __restrict__ partial_nk, Geometry g,
                               int n_splits, double* __restrict__ nk,
                               double* __restrict__ srx, double* __restrict__ srxx) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  auto column_sum = [&](int k, int block, int col) -> double {
    // tile that owns `block`: the first `rem` tiles hold base + 1 blocks
    const int big = g.rem * (g.base + 1);
    const int tile = block < big ? block / (g.base + 1) : g.rem + (block - big) / g.base;
    const int within = block - tile_first_block(g, tile);
    const int64_t off = (static_cast<int64_t>(k / 128) * kTileCols + within * 64 + col) * 128 + k % 128;
    double acc = 0.0;
    for (int s = 0; s < n_splits; ++s)
      acc += static_cast<double>(partial[static_cast<int64_t>(tile * n_splits + s) * 2 * kTileCols * 128 + off]);
    return acc;
  };
  const int64_t n_xx = static_cast<int64_t>(g.k) * g.d * g.d;
  if (idx < n_xx) {
    const int k = static_cast<int>(idx / (g.d * g.d));
    const int de = static_cast<int>(idx % (g.d * g.d));
    const int dd = de / g.d, ee = de % g.d;
    const int lo = dd < ee ? dd : ee, hi = dd < ee ? ee : dd;
    const int ga = lo / 8, gb = hi / 8;
    const int block = ga * g.n_groups - ga * (ga - 1) / 2 + (gb - ga);
    srxx[idx] = column_sum(k, block, (lo % 8) * 8 + hi % 8);
  } else if (idx < n_xx + static_cast<int64_t>(g.k) * g.d) {
    const int64_t j = idx - n_xx;
    if (srx != nullptr) srx[j] = column_sum(static_cast<int>(j / g.d), g.n_pair, static_cast<int>(j % g.d));
  } else if (idx < n_xx + static_cast<int64_t>(g.k) * g.d + g.k) {
    const int k = static_cast<int>(idx - n_xx - static_cast<int64_t>(g.k) * g.d);
    if (nk != nullptr) nk[k] = column_sum(k, g.n_pair + 1, 0);
  }
}

// ---- CTA-pair version for pre-split responsibilities (K = 256 per slice) -----------------------------------
// The single-CTA kernel above asks 135 B/clk of the 128 B/clk shared-memory port at full tensor rate (per 16-row
// stage the six M128 x N256 MMAs fetch 72 KB of operands, the converters and copies write 32 KB), which is what
// keeps its tensor pipe at ~60 % whatever the converters are relieved of (pre-split R: 6.8 -> 6.4 ms only).  Here
// one CTA PAIR (cluster of 2, tcgen05 cta_group::2, M = 256, N = 256) owns a (column tile, row range): CTA r holds
// components [128 r, +128) in its TMEM and supplies 128 components of R (bulk-copied operand tiles) and 128 columns
// of Phi (two 64-column blocks, formed by its converter warps), so each SM fetches 8 KB per MMA instead of 12 and
// writes 16 KB per K step instead of 32: 104 B/clk.  The 256 TMEM columns this frees double-buffer the accumulators,
// so the drain every 2048 rows no longer stalls the MMAs.
constexpr int kP2StageRows = 32;                  // two K = 16 steps per stage
constexpr int kP2Part = 128 * kP2StageRows * 2;   // one bf16 part of one operand of one CTA: 8 KB
constexpr int kP2StageBytes = 4 * kP2Part;        // R.b1, R.b2, Phi.b1, Phi.b2: 32 KB
constexpr int kP2FlushIters = 64 / BB_CHAIN_DIV;  // 2048 rows per TMEM accumulation chain

struct __align__(1024) Pair2Smem {
  uint8_t stage[kStages][kP2StageBytes];
  uint64_t full[kStages];        // leader: 2 x 8 converter warps + rank 1's "my R tiles landed"
  uint64_t rfull[kStages];       // each CTA: its own bulk copies of the stage (expect_tx)
  uint64_t empty[kStages];       // each CTA: one multicast MMA commit
  uint64_t acc_full[2];          // each CTA: one multicast MMA commit
  uint64_t acc_empty[2];         // leader: 2 x kEpiWarps arrivals
  uint32_t tmem_base;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreadsSplit, 1)
weighted_pairs2_kernel(const float* __restrict__ x, const uint8_t* __restrict__ tiles, int64_t n, Geometry g,
                       int n_splits, float* __restrict__ partial) {      // partial: [pair][2 ranks][256 cols][128 lanes]
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  Pair2Smem& sm = *reinterpret_cast<Pair2Smem*>(smem_raw);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = pair::cluster_ctarank();
  const int pr = blockIdx.x >> 1;
  const int tile = pr / n_splits;
  const int split = pr % n_splits;
  const int block0 = tile_first_block(g, tile);
  const int tile_blocks = tile_block_count(g, tile);
  const int64_t total_iters = (n + kP2StageRows - 1) / kP2StageRows;
  const int64_t it_begin = total_iters * split / n_splits;
  const int64_t it_end = total_iters * (split + 1) / n_splits;
  const int n_iters = static_cast<int>(it_end - it_begin);
  const int64_t row_begin = it_begin * kP2StageRows;

  if (warp == kMmaWarp) {
    if (lane == 0) {
      for (int s = 0; s < kStages; ++s) {
        ptx::mbar_init(&sm.full[s], 2 * (kConvWarps / kConvGroups) + 1);
        ptx::mbar_init(&sm.rfull[s], 1);
        ptx::mbar_init(&sm.empty[s], 1);
      }
      for (int b = 0; b < 2; ++b) {
        ptx::mbar_init(&sm.acc_full[b], 1);
        ptx::mbar_init(&sm.acc_empty[b], 2 * kEpiWarps);
      }
      ptx::fence_mbar_init();
    }
    __syncwarp();
    pair::tmem_alloc_pair(&sm.tmem_base, kTmemCols);
  }
  // the unused blocks of a 3-block tile are never written and stay zero (every tile issues N = 256 MMAs)
  for (int i = threadIdx.x; i < kStages * kP2StageBytes / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(&sm.stage[0][0])[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  ptx::tc_fence_before_sync();
  __syncthreads();
  pair::cluster_sync_all();        // both CTAs' barriers initialised before any remote arrive
  ptx::tc_fence_after_sync();
  const uint32_t tmem = sm.tmem_base;

  if (warp < kConvWarps) {
    // ---------------- converter warps: Phi only ----------------
    // group = warp & 1 takes every 2nd stage; warp wi of the group owns rows 4 wi .. 4 wi + 3 of the stage; lane ->
    // (block j = lane >> 4 of this CTA's two, 4 columns of it: i = (lane & 15) >> 1, j0 = 4 (lane & 1))
    const int group = warp & (kConvGroups - 1);
    const int wi = warp / kConvGroups;
    uint32_t off[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int kk = 4 * wi + q;
      off[q] = 2 * kP2Part + (lane >> 4) * 4096 + (kk >> 3) * 1024 + (kk & 7) * 128 +
               ((((lane & 15) >> 1) ^ (kk & 7)) << 4) + (lane & 1) * 8;
    }
    int src_a = 0, src_b = 0;
    float ma = 0.f, ca = 0.f, mb = 0.f, cb = 0.f;
    {
      const int bl = 2 * static_cast<int>(rank) + (lane >> 4);
      const int b = block0 + bl;
      const int i = (lane & 15) >> 1, j0 = (lane & 1) * 4;
      if (bl < tile_blocks) {
        if (b < g.n_pair) {
          int ga, gb;
          pair_of(b, g.n_groups, &ga, &gb);
          src_a = 8 * ga + i;
          src_b = 8 * gb + j0;
          ma = 1.f;
          mb = 1.f;
        } else if (b == g.n_pair) {          // linear block: column c = 8 i + j0 .. + 3
          ca = 1.f;
          if (8 * i + j0 < g.d) {
            src_b = 8 * i + j0;
            mb = 1.f;
          }
        } else {                             // ones block: column 0 = 1 (its contraction with R is Nk)
          ca = 1.f;
          cb = (lane & 15) == 0 ? 1.f : 0.f;
        }
      }
    }
    const uint32_t stage0 = ptx::smem_u32(sm.stage[0]);
    int64_t row0 = row_begin + static_cast<int64_t>(group) * kP2StageRows + 4 * wi;
    float xa[4];
    float4 xb[4];
    auto load_rows = [&]() {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        // rows past n: their R operand rows are zero, so any finite Phi will do -- re-read row n - 1
        const int64_t row = row0 + q < n ? row0 + q : n - 1;
        const float* xrow = x + row * g.d;
        xa[q] = __ldg(xrow + src_a);
        xb[q] = __ldg(reinterpret_cast<const float4*>(xrow + src_b));
      }
      row0 += kConvGroups * kP2StageRows;
    };
    if (group < n_iters) load_rows();
    for (int it = group; it < n_iters; it += kConvGroups) {
      const int s = it % kStages;
      ptx::mbar_wait(&sm.empty[s], ((it / kStages) & 1) ^ 1);
      const uint32_t stage_addr = stage0 + s * kP2StageBytes;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float a = fmaf(xa[q], ma, ca);
        const float am = a * mb;
        const float4 ph = make_float4(fmaf(am, xb[q].x, a * cb), am * xb[q].y, am * xb[q].z, am * xb[q].w);
        uint32_t b1[2], b2[2];
        split_bf16(ph, b1, b2);
        sts_u2(stage_addr + off[q], b1[0], b1[1]);
        sts_u2(stage_addr + off[q] + kP2Part, b2[0], b2[1]);
      }
      if (it + kConvGroups < n_iters) load_rows();       // before the fence: their latency overlaps it
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) pair::mbar_arrive_cluster_relaxed(&sm.full[s], 0);
    }
  } else if (warp < kMmaWarp) {
    // ---------------- epilogue warps: this CTA's TMEM fp32 -> fp32 partial block (RMW through L2) ----------------
    const int qd = warp & 3;
    const int n_intervals = (n_iters + kP2FlushIters - 1) / kP2FlushIters;
    const int cols = tile_blocks * 64;
    float* my_partial = partial + (static_cast<int64_t>(pr) * 2 + rank) * kTileCols * 128 + qd * 32 + lane;
    if (n_intervals == 0) {
      for (int c = 0; c < kTileCols; ++c) my_partial[c * 128] = 0.f;
    }
    for (int interval = 0; interval < n_intervals; ++interval) {
      const int buf = interval & 1;
      ptx::mbar_wait_sleep(&sm.acc_full[buf], (interval >> 1) & 1);
      ptx::tc_fence_after_sync();
      const uint32_t t_addr = tmem + (static_cast<uint32_t>(qd * 32) << 16) + buf * kTileCols;
#pragma unroll 1
      for (int cc = 0; cc < cols / 32; ++cc) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(t_addr + cc * 32, v);
        float* dst = my_partial + cc * 32 * 128;
        float old[32];
        if (interval != 0) {
#pragma unroll
          for (int j = 0; j < 32; ++j) old[j] = dst[j * 128];
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) old[j] = 0.f;
        }
        ptx::tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) dst[j * 128] = old[j] + __uint_as_float(v[j]);
      }
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) pair::mbar_arrive_cluster_relaxed(&sm.acc_empty[buf], 0);
    }
  } else if (warp == kProdWarp) {
    // ---------------- R producer: this CTA's 128 components of every stage, 8 bulk copies of 2 KB ----------------
    // a 32-row stage = two 16-row tile images; per image and BF16 part this CTA's two 64-component blocks are
    // 2 KB each ([8-row group (2)][8 rows x 128 B]) and land 4 KB apart (a block holds 4 row groups here)
    if (ptx::elect_one()) {
      uint64_t keep;
      asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(keep));     // re-read by the other column tiles
      const uint8_t* src0 = tiles + it_begin * 2 * (2 * kRPart) + rank * 4096;
      for (int it = 0; it < n_iters; ++it) {
        const int s = it % kStages;
        ptx::mbar_wait(&sm.empty[s], ((it / kStages) & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(&sm.rfull[s], 2 * kP2Part);
        const uint32_t dst0 = ptx::smem_u32(sm.stage[s]);
        const uint32_t bar = ptx::smem_u32(&sm.rfull[s]);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int half = c & 1, blk = (c >> 1) & 1, part = c >> 2;
          const uint8_t* src = src0 + (static_cast<int64_t>(it) * 2 + half) * (2 * kRPart) + part * kRPart + blk * 2048;
          const uint32_t dst = dst0 + part * kP2Part + blk * 4096 + half * 2048;
          asm volatile(
              "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
              ::"r"(dst), "l"(src), "r"(2048), "r"(bar), "l"(keep)
              : "memory");
        }
      }
    }
  } else if (rank == 0) {
    // ---------------- MMA issuer (leader CTA, one elected thread) ----------------
    if (ptx::elect_one()) {
      const uint32_t idesc = ptx::make_idesc(256, kTileCols, /*bf16*/ 1, /*A MN-major*/ 1, /*B MN-major*/ 1);
      for (int it = 0; it < n_iters; ++it) {
        const int s = it % kStages;
        const int interval = it / kP2FlushIters;
        const int buf = interval & 1;
        const bool first = (it % kP2FlushIters) == 0;
        if (first) pair::mbar_wait_cluster(&sm.acc_empty[buf], ((interval >> 1) & 1) ^ 1);
        ptx::mbar_wait(&sm.full[s], (it / kStages) & 1);        // both CTAs' Phi tiles and rank 1's R tiles
        ptx::mbar_wait(&sm.rfull[s], (it / kStages) & 1);       // this CTA's R tiles
        ptx::tc_fence_after_sync();
        const uint32_t base = ptx::smem_u32(sm.stage[s]);
        const uint32_t d_tmem = tmem + buf * kTileCols;
#pragma unroll
        for (int ks = 0; ks < kP2StageRows / 16; ++ks) {
          // one K = 16 step = two 8-row groups (SBO = 1 KB); 64-element MN blocks 4 KB apart (LBO)
          const uint32_t b = base + ks * 2048;
          const uint64_t r1 = ptx::make_smem_desc(b, 4096, 1024, ptx::kLayoutSwizzle128B);
          const uint64_t r2 = ptx::make_smem_desc(b + kP2Part, 4096, 1024, ptx::kLayoutSwizzle128B);
          const uint64_t p1 = ptx::make_smem_desc(b + 2 * kP2Part, 4096, 1024, ptx::kLayoutSwizzle128B);
          const uint64_t p2 = ptx::make_smem_desc(b + 3 * kP2Part, 4096, 1024, ptx::kLayoutSwizzle128B);
          pair::mma_bf16_pair(d_tmem, r1, p1, idesc, (first && ks == 0) ? 0u : 1u);
          pair::mma_bf16_pair(d_tmem, r1, p2, idesc, 1u);
          pair::mma_bf16_pair(d_tmem, r2, p1, idesc, 1u);
        }
        pair::mma_commit_pair(&sm.empty[s]);
        if ((it % kP2FlushIters) == kP2FlushIters - 1 || it == n_iters - 1) pair::mma_commit_pair(&sm.acc_full[buf]);
      }
    }
  } else {
    // ---------------- rank 1, idle MMA warp: tell the leader when this CTA's R tiles have landed ----------------
    if (ptx::elect_one()) {
      for (int it = 0; it < n_iters; ++it) {
        const int s = it % kStages;
        ptx::mbar_wait(&sm.rfull[s], (it / kStages) & 1);
        pair::mbar_arrive_cluster_relaxed(&sm.full[s], 0);
      }
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  pair::cluster_sync_all();        // the peer's shared memory / barriers stay alive until both are done
  if (warp == kMmaWarp) pair::tmem_dealloc_pair(tmem, kTmemCols);
}

struct PairsPlan {
  Geometry g;
  int n_splits, grid;
};

PairsPlan plan_pairs(int64_t n, int d, int k) {
  PairsPlan p;
  p.g.d = d;
  p.g.k = k;
  p.g.ldr = k;
  p.g.n_groups = d / 8;
  p.g.n_pair = p.g.n_groups * (p.g.n_groups + 1) / 2;
  p.g.n_blocks = p.g.n_pair + 2;      // pair blocks, the linear block, the ones block (Nk)
  p.g.n_tiles = (p.g.n_blocks + 3) / 4;
  p.g.base = p.g.n_blocks / p.g.n_tiles;
  p.g.rem = p.g.n_blocks % p.g.n_tiles;
  int sms = device_sm_count();
  if (sms <= 0) sms = 148;
  const int64_t iters = (n + kStageRows - 1) / kStageRows;
  p.n_splits = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(sms / p.g.n_tiles, iters)));
  p.grid = p.g.n_tiles * p.n_splits;
  return p;
}

}  // namespace

bool weighted_pairs_supported(int64_t n, int d, int k, const void* x, const void* r) {
  // more than 256 components: one launch per slice of 256 (the kernel holds M = 256 in TMEM)
  return n > 0 && d >= 8 && d <= 64 && d % 8 == 0 && k >= 4 && k <= 16 * kMaxK && k % 4 == 0 &&
         reinterpret_cast<uintptr_t>(x) % 16 == 0 && reinterpret_cast<uintptr_t>(r) % 16 == 0;
}

int64_t weighted_pairs_workspace(int64_t n, int d, int k) {
  const PairsPlan p = plan_pairs(n, d, k < kMaxK ? k : kMaxK);
  return static_cast<int64_t>(p.grid) * 2 * kTileCols * 128 * static_cast<int64_t>(sizeof(float)) +
         static_cast<int64_t>(p.n_splits) * k * static_cast<int64_t>(sizeof(double)) + 1024;
}

// Pre-split responsibilities (mode 2): per slice of 256 components and per 16-row stage one 16 KB image
// [b1 8 KB | b2 8 KB], each [64-component block (4) 2 KB][8-row group (2) 1 KB][8 rows x 128 B, 16-byte chunks
// XOR-swizzled with the row] -- the shared-memory layout of a stage's R part.  Needs k % 256 == 0.
bool weighted_pairs_split_supported(int64_t n, int d, int k, const void* x, const void* rsplit) {
  return n > 0 && d >= 8 && d <= 64 && d % 8 == 0 && k >= kMaxK && k % kMaxK == 0 && k <= 16 * kMaxK &&
         reinterpret_cast<uintptr_t>(x) % 16 == 0 && reinterpret_cast<uintptr_t>(rsplit) % 16 == 0;
}
// 16-row tile images per slice: an even number, so that the 32-row stages of the CTA-pair kernel never read past it
int64_t weighted_pairs_split_stages(int64_t n) { return 2 * ((n + 31) / 32); }
int64_t weighted_pairs_split_bytes(int64_t n, int k) {
  return static_cast<int64_t>(k / kMaxK) * weighted_pairs_split_stages(n) * (2 * kRPart);
}

// r: responsibilities [n, k]; or, with lse != nullptr, logits [n, k] and r = exp(logit - lse[row]); or, with
// pre_split, the operand tiles described above (lse unused)
static int launch_weighted_pairs_any(const float* x, const float* r, const float* lse, bool pre_split, int64_t n, int d,
                                     int k, double* nk, double* sum_rx, double* sum_rxx, void* workspace,
                                     int64_t workspace_bytes, cudaStream_t stream) {
  if (pre_split ? !weighted_pairs_split_supported(n, d, k, x, r) : !weighted_pairs_supported(n, d, k, x, r)) {
    set_error("weighted_pairs: unsupported shape n=%lld d=%d k=%d%s", static_cast<long long>(n), d, k,
              pre_split ? " (pre-split responsibilities need k % 256 == 0)" : "");
    return BB_ERR_UNSUPPORTED;
  }
  if (workspace == nullptr || workspace_bytes < weighted_pairs_workspace(n, d, k)) {
    set_error("weighted_pairs: workspace %lld < %lld bytes", static_cast<long long>(workspace_bytes),
              static_cast<long long>(weighted_pairs_workspace(n, d, k)));
    return BB_ERR_WORKSPACE;
  }
  uint8_t* ws0 = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~static_cast<uintptr_t>(255));
  const int smem_bytes = static_cast<int>(sizeof(SmemLayout));
  static SmemOptIn smem_opt_in_0;
  BB_CUDA_OK(smem_opt_in_0.ensure(weighted_pairs_kernel<0>, smem_bytes));
  static SmemOptIn smem_opt_in_1;
  BB_CUDA_OK(smem_opt_in_1.ensure(weighted_pairs_kernel<1>, smem_bytes));
  static SmemOptIn smem_opt_in_2;
  BB_CUDA_OK(smem_opt_in_2.ensure(weighted_pairs_kernel<2>, smem_bytes));
  static const int prefetch_iters = (getenv("BB_WP_PREFETCH") ? atoi(getenv("BB_WP_PREFETCH")) & 0xff : 0) |
                                    ((getenv("BB_WP_COLLECTOR") ? atoi(getenv("BB_WP_COLLECTOR")) : 0) ? 0x100 : 0) |
                                    ((getenv("BB_WP_L1") ? atoi(getenv("BB_WP_L1")) : 0) ? 0x200 : 0) |
                                    ((getenv("BB_WP_EARLY") ? atoi(getenv("BB_WP_EARLY")) : 1) ? 0x400 : 0) |
                                    ((getenv("BB_WP_ABLATE") ? atoi(getenv("BB_WP_ABLATE")) & 15 : 0) << 12);
  const int64_t stages = weighted_pairs_split_stages(n);
  static const bool use_pairs = !(getenv("BB_WP_PAIR") && atoi(getenv("BB_WP_PAIR")) == 0);
  static SmemOptIn smem_opt_in_p2;
  BB_CUDA_OK(smem_opt_in_p2.ensure(weighted_pairs2_kernel, static_cast<int>(sizeof(Pair2Smem))));
  for (int k0 = 0; k0 < k; k0 += kMaxK) {                 // slices of at most 256 components (stream-ordered)
    const int kc = k - k0 < kMaxK ? k - k0 : kMaxK;
    PairsPlan p = plan_pairs(n, d, kc);
    p.g.ldr = k;
    float* partial = reinterpret_cast<float*>(ws0);
    double* partial_nk = reinterpret_cast<double*>(ws0 + static_cast<int64_t>(p.grid) * 2 * kTileCols * 128 * sizeof(float));
    if (pre_split && use_pairs) {
      // CTA pairs: (column tile, row range) per pair, row ranges in 32-row stages
      const uint8_t* tiles = reinterpret_cast<const uint8_t*>(r) + static_cast<int64_t>(k0 / kMaxK) * stages * (2 * kRPart);
      int sms = device_sm_count();
      if (sms <= 0) sms = 148;
      const int64_t iters32 = (n + kP2StageRows - 1) / kP2StageRows;
      p.n_splits = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((sms / 2) / p.g.n_tiles, iters32)));
      p.grid = 2 * p.g.n_tiles * p.n_splits;
      weighted_pairs2_kernel<<<p.grid, kThreadsSplit, static_cast<int>(sizeof(Pair2Smem)), stream>>>(x, tiles, n, p.g,
                                                                                                    p.n_splits, partial);
    } else if (pre_split) {
      const float* tiles = reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(r) +
                                                          static_cast<int64_t>(k0 / kMaxK) * stages * (2 * kRPart));
      weighted_pairs_kernel<2><<<p.grid, kThreadsSplit, smem_bytes, stream>>>(x, tiles, nullptr, n, p.g, p.n_splits,
                                                                               prefetch_iters, partial, partial_nk);
    } else if (lse != nullptr) {
      weighted_pairs_kernel<1><<<p.grid, kThreads, smem_bytes, stream>>>(x, r + k0, lse, n, p.g, p.n_splits, prefetch_iters,
                                                                         partial, partial_nk);
    } else {
      weighted_pairs_kernel<0><<<p.grid, kThreads, smem_bytes, stream>>>(x, r + k0, lse, n, p.g, p.n_splits, prefetch_iters,
                                                                         partial, partial_nk);
    }
    BB_CHECK_LAUNCH("weighted_pairs_kernel");
    const int64_t total = static_cast<int64_t>(kc) * d * d + static_cast<int64_t>(kc) * d + kc;
    weighted_pairs_finalize_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, stream>>>(
        partial, partial_nk, p.g, p.n_splits, nk != nullptr ? nk + k0 : nullptr,
        sum_rx != nullptr ? sum_rx + static_cast<int64_t>(k0) * d : nullptr, sum_rxx + static_cast<int64_t>(k0) * d * d);
    BB_CHECK_LAUNCH("weighted_pairs_finalize_kernel");
  }
  return BB_OK;
}

int launch_weighted_pairs(const float* x, const float* r, const float* lse, int64_t n, int d, int k, double* nk,
                          double* sum_rx, double* sum_rxx, void* workspace, int64_t workspace_bytes,
                          cudaStream_t stream) {
  return launch_weighted_pairs_any(x, r, lse, false, n, d, k, nk, sum_rx, sum_rxx, workspace, workspace_bytes, stream);
}

int launch_weighted_pairs_split(const float* x, const void* rsplit, int64_t n, int d, int k, double* nk,
                                double* sum_rx, double* sum_rxx, void* workspace, int64_t workspace_bytes,
                                cudaStream_t stream) {
  return launch_weighted_pairs_any(x, static_cast<const float*>(rsplit), nullptr, true, n, d, k, nk, sum_rx, sum_rxx,
                                   workspace, workspace_bytes, stream);
}

}  // namespace bb
