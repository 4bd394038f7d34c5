// Large-D Gram statistics on tcgen05 CTA pairs (D % 256 == 0):
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
//     cost 1.5x one TF32 pass (3xTF32 would cost 3x);
//   * no fp32 staging: 16 converter warps load X with coalesced 128-bit loads straight into
//     registers (prefetched one stage ahead), split, and store b1/b2 into shared memory in the
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

#include "common.cuh"
#include "kernels.h"
#include "sm100_ptx.cuh"

namespace bb {

namespace {

constexpr int kBlock = 256;               // output block edge (features) per CTA pair
constexpr int kHalf = 128;                // features per CTA per operand side
constexpr int kStageRows = 32;            // data rows per pipeline stage (2 MMA K-steps of 16)
constexpr int kStages = 6;
constexpr int kTileBytes = kHalf * kStageRows * 2;      // one bf16 operand tile: 8 KB
constexpr int kStageBytes = 4 * kTileBytes;             // A.b1, A.b2, B.b1, B.b2: 32 KB
constexpr int kFlushIters = 64;           // TMEM accumulators drained every 64 stages = 2048 rows
constexpr int kConvWarps = 16;
constexpr int kEpiWarps = 4;
constexpr int kMmaWarp = kConvWarps + kEpiWarps;        // warp 20 (20 % 4 == 0 is irrelevant for it)
constexpr int kThreads = (kMmaWarp + 1) * 32;           // 672
constexpr int kTmemCols = 512;            // 2 accumulator buffers x 256 columns
constexpr int kPrefetchDist = 6;          // stages ahead for the L2 prefetch hints

struct __align__(1024) SmemLayout {
  uint8_t stage[kStages][kStageBytes];
  double xty[kHalf];
  double yty;
  uint64_t full[kStages];        // leader: 2 CTAs x kConvWarps arrivals
  uint64_t empty[kStages];       // each CTA: one multicast MMA commit
  uint64_t acc_full[2];          // each CTA: one multicast MMA commit
  uint64_t acc_empty[2];         // leader: 2 CTAs x kEpiWarps arrivals
  uint32_t tmem_base;
};

// kind::f16, BF16 x BF16 -> FP32, both operands MN-major, M = 256 (pair), N = 256
constexpr uint32_t kIdesc = ptx::make_idesc(256, 256, /*bf16*/ 1, /*A MN-major*/ 1, /*B MN-major*/ 1);

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive (release at cluster scope) on the barrier at the same shared-memory offset in CTA `rank`
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}\n"
      ::"r"(ptx::smem_u32(bar)), "r"(rank)
      : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(ptx::smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_slot, uint32_t n_cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   ptx::smem_u32(smem_slot)),
               "r"(n_cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t n_cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(n_cols)
               : "memory");
}
// D[tmem, both CTAs] (+)= A[smem desc, both CTAs] * B[smem desc, both CTAs], issued by one
// thread of the leader CTA.
__device__ __forceinline__ void mma_bf16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                              uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this offset in BOTH CTAs once all MMAs issued so far completed
__device__ __forceinline__ void mma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
      "[%0], %1;" ::"r"(ptx::smem_u32(bar)),
      "h"(static_cast<uint16_t>(3))
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
__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
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

// x = b1 + b2 + O(2^-17 x): four consecutive features -> two packed bf16x2 words each
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

struct RowRegs {
  float4 a, b;
  float y;
};

// grid = 2 * n_blocks * n_splits CTAs; pair p = blockIdx.x / 2: block = p % n_blocks,
// split = p / n_blocks (pairs of one split walk the same rows at the same time -> L2 reuse).
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gram_pair_kernel(const float* __restrict__ x, const float* __restrict__ y, int64_t n, int d,
                 int n_blocks, int n_splits,
                 float* __restrict__ partial,          // [pair][2][128][256] fp32
                 double* __restrict__ partial_xty,     // [split][d]
                 double* __restrict__ partial_yty) {   // [split]
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  SmemLayout& sm = *reinterpret_cast<SmemLayout*>(smem_raw);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int blk = pair % n_blocks;
  const int split = pair / n_blocks;
  // upper-triangle block index -> (i, j), row-major over i <= j
  int2 ij;
  {
    const int nb = d / kBlock;
    int i = 0, rem = blk;
    while (rem >= nb - i) {
      rem -= nb - i;
      ++i;
    }
    ij.x = i;
    ij.y = i + rem;
  }
  const bool diag = ij.x == ij.y;
  const int feat_a = ij.x * kBlock + static_cast<int>(rank) * kHalf;
  const int feat_b = ij.y * kBlock + static_cast<int>(rank) * kHalf;
  // rows of this split, in whole stages
  const int64_t total_iters = (n + kStageRows - 1) / kStageRows;
  const int64_t it_begin = total_iters * split / n_splits;
  const int64_t it_end = total_iters * (split + 1) / n_splits;
  const int n_iters = static_cast<int>(it_end - it_begin);
  const int64_t row_begin = it_begin * kStageRows;

  if (warp == kMmaWarp) {
    if (lane == 0) {
      for (int s = 0; s < kStages; ++s) {
        ptx::mbar_init(&sm.full[s], 2 * kConvWarps);
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
    // warp w handles rows w and w + 16 of every stage; lane l holds A features [4l, 4l+4) and
    // B features [4l, 4l+4) of this CTA's 128-feature halves.
    const bool do_xty = (y != nullptr) && diag;
    const bool do_yty = do_xty && blk == 0 && rank == 0 && lane == 0;
    const float* xa = x + feat_a + lane * 4;
    const float* xb = x + feat_b + lane * 4;
    // shared-memory offset of this lane's 8-byte piece inside an operand tile, for row k:
    // [mn block (64 features) 4 KB][k group 1 KB][k & 7 -> 128 B][16-byte chunk ^ (k & 7)]
    const uint32_t mn_off = (lane >> 4) * 4096;
    const uint32_t chunk = (lane >> 1) & 7;
    const uint32_t half8 = (lane & 1) * 8;
    double xty_acc[4] = {0.0, 0.0, 0.0, 0.0};
    double yty_acc = 0.0;
    float xty_f[4] = {0.f, 0.f, 0.f, 0.f};
    float yty_f = 0.f;

    auto load_rows = [&](int it, RowRegs (&regs)[2]) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int64_t row = row_begin + static_cast<int64_t>(it) * kStageRows + warp + 16 * j;
        if (it < n_iters && row < n) {
          regs[j].a = ldg_f4(xa + row * d);
          regs[j].b = ldg_f4(xb + row * d);
          regs[j].y = do_xty ? __ldg(y + row) : 0.f;
        } else {
          regs[j].a = make_float4(0.f, 0.f, 0.f, 0.f);
          regs[j].b = make_float4(0.f, 0.f, 0.f, 0.f);
          regs[j].y = 0.f;
        }
      }
      // L2 prefetch hint for a stage further ahead (one 128-byte line per 8 lanes)
      const int pit = it + kPrefetchDist;
      if ((lane & 7) == 0 && pit < n_iters) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int64_t row = row_begin + static_cast<int64_t>(pit) * kStageRows + warp + 16 * j;
          if (row < n) {
            prefetch_l2(xa + row * d);
            prefetch_l2(xb + row * d);
          }
        }
      }
    };
    auto convert_store = [&](int it, const RowRegs (&regs)[2]) {
      const int s = it % kStages;
      ptx::mbar_wait(&sm.empty[s], ((it / kStages) & 1) ^ 1);
      const uint32_t stage_addr = ptx::smem_u32(sm.stage[s]);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int k = warp + 16 * j;
        const uint32_t off = mn_off + (k >> 3) * 1024 + (k & 7) * 128 + ((chunk ^ (k & 7)) << 4) + half8;
        uint32_t b1[2], b2[2];
        split_bf16(regs[j].a, b1, b2);
        sts_u2(stage_addr + off, b1[0], b1[1]);
        sts_u2(stage_addr + kTileBytes + off, b2[0], b2[1]);
        split_bf16(regs[j].b, b1, b2);
        sts_u2(stage_addr + 2 * kTileBytes + off, b1[0], b1[1]);
        sts_u2(stage_addr + 3 * kTileBytes + off, b2[0], b2[1]);
        if (do_xty) {
          xty_f[0] = fmaf(regs[j].a.x, regs[j].y, xty_f[0]);
          xty_f[1] = fmaf(regs[j].a.y, regs[j].y, xty_f[1]);
          xty_f[2] = fmaf(regs[j].a.z, regs[j].y, xty_f[2]);
          xty_f[3] = fmaf(regs[j].a.w, regs[j].y, xty_f[3]);
          yty_f = fmaf(regs[j].y, regs[j].y, yty_f);
        }
      }
      fence_proxy_async_smem();      // generic-proxy stores -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&sm.full[s], 0);
      if (do_xty && (it & 15) == 15) {          // fp32 partial sums over 32 rows, then float64
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          xty_acc[c] += static_cast<double>(xty_f[c]);
          xty_f[c] = 0.f;
        }
        yty_acc += static_cast<double>(yty_f);
        yty_f = 0.f;
      }
    };

    RowRegs r0[2], r1[2];
    load_rows(0, r0);
    for (int it = 0; it < n_iters; it += 2) {
      load_rows(it + 1, r1);
      convert_store(it, r0);
      if (it + 1 < n_iters) {
        load_rows(it + 2, r0);
        convert_store(it + 1, r1);
      }
    }
    if (do_xty) {
#pragma unroll
      for (int c = 0; c < 4; ++c) atomicAdd(&sm.xty[lane * 4 + c], xty_acc[c] + static_cast<double>(xty_f[c]));
      if (do_yty) atomicAdd(&sm.yty, yty_acc + static_cast<double>(yty_f));
    }
  } else if (warp < kMmaWarp) {
    // ---------------- epilogue warps: TMEM fp32 -> fp32 partial block (RMW through L2) ----------------
    const int q = warp & 3;                     // TMEM lane quadrant this warp may access
    const int n_intervals = (n_iters + kFlushIters - 1) / kFlushIters;
    float* my_partial = partial + ((static_cast<int64_t>(pair) * 2 + rank) * kHalf + q * 32 + lane) * kBlock;
    if (n_intervals == 0) {
      for (int c = 0; c < kBlock; c += 4)
        *reinterpret_cast<float4*>(my_partial + c) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int interval = 0; interval < n_intervals; ++interval) {
      const int buf = interval & 1;
      ptx::mbar_wait(&sm.acc_full[buf], (interval >> 1) & 1);
      ptx::tc_fence_after_sync();
      const uint32_t t_addr = tmem + (static_cast<uint32_t>(q * 32) << 16) + buf * kBlock;
#pragma unroll 1
      for (int cc = 0; cc < kBlock / 32; ++cc) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(t_addr + cc * 32, v);
        float4* dst = reinterpret_cast<float4*>(my_partial + cc * 32);
        float4 old[8];
        if (interval != 0) {
#pragma unroll
          for (int j = 0; j < 8; ++j) old[j] = dst[j];
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) old[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        ptx::tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 o = old[j];
          o.x += __uint_as_float(v[4 * j + 0]);
          o.y += __uint_as_float(v[4 * j + 1]);
          o.z += __uint_as_float(v[4 * j + 2]);
          o.w += __uint_as_float(v[4 * j + 3]);
          dst[j] = o;
        }
      }
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&sm.acc_empty[buf], 0);
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
        mbar_wait_cluster(&sm.full[s], (it / kStages) & 1);
        ptx::tc_fence_after_sync();
        const uint32_t stage_addr = ptx::smem_u32(sm.stage[s]);
        const uint32_t d_tmem = tmem + buf * kBlock;
#pragma unroll
        for (int ks = 0; ks < kStageRows / 16; ++ks) {
          // one K = 16 step = two 8-row k groups (SBO = 1 KB apart); 64-feature MN blocks LBO = 4 KB apart
          const uint32_t base = stage_addr + ks * 2048;
          const uint64_t a1 = ptx::make_smem_desc(base, 4096, 1024, ptx::kLayoutSwizzle128B);
          const uint64_t a2 = ptx::make_smem_desc(base + kTileBytes, 4096, 1024, ptx::kLayoutSwizzle128B);
          const uint64_t b1 = ptx::make_smem_desc(base + 2 * kTileBytes, 4096, 1024, ptx::kLayoutSwizzle128B);
          const uint64_t b2 = ptx::make_smem_desc(base + 3 * kTileBytes, 4096, 1024, ptx::kLayoutSwizzle128B);
          mma_bf16_pair(d_tmem, a1, b1, kIdesc, (first && ks == 0) ? 0u : 1u);
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
    if (threadIdx.x < kHalf)
      partial_xty[static_cast<int64_t>(split) * d + feat_a + threadIdx.x] = sm.xty[threadIdx.x];
    if (blk == 0 && rank == 0 && threadIdx.x == 0) partial_yty[split] = sm.yty;
  }
  cluster_sync_all();          // the peer's shared memory / barriers stay alive until both are done
  if (warp == kMmaWarp) tmem_dealloc_pair(tmem, kTmemCols);
}

// XtX[d,e] (float64, full symmetric matrix) = sum over splits of the upper-triangle block
// partials; Xty / yty likewise.  One thread per output element.
__global__ void __launch_bounds__(256)
gram_finalize_kernel(const float* __restrict__ partial, const double* __restrict__ partial_xty,
                     const double* __restrict__ partial_yty, int d, int n_blocks, int n_splits,
                     double* __restrict__ xtx, double* __restrict__ xty, double* __restrict__ yty) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int nb = d / kBlock;
  if (idx < static_cast<int64_t>(d) * d) {
    const int row = static_cast<int>(idx / d), col = static_cast<int>(idx % d);
    const int r = row < col ? row : col, c = row < col ? col : row;
    const int bi = r / kBlock, bj = c / kBlock;
    const int blk = bi * nb - bi * (bi - 1) / 2 + (bj - bi);
    const int rr = r % kBlock, cc = c % kBlock;
    double acc = 0.0;
    for (int s = 0; s < n_splits; ++s) {
      const int64_t pair = static_cast<int64_t>(s) * n_blocks + blk;
      acc += static_cast<double>(partial[((pair * 2 + rr / kHalf) * kHalf + rr % kHalf) * kBlock + cc]);
    }
    xtx[idx] = acc;
  }
  if (xty != nullptr && idx < d) {
    double acc = 0.0;
    for (int s = 0; s < n_splits; ++s) acc += partial_xty[static_cast<int64_t>(s) * d + idx];
    xty[idx] = acc;
  }
  if (yty != nullptr && idx == 0) {
    double acc = 0.0;
    for (int s = 0; s < n_splits; ++s) acc += partial_yty[s];
    *yty = acc;
  }
}

struct GramGrid {
  int nb, n_blocks, n_splits;
};

GramGrid plan_gram(int64_t n, int d) {
  GramGrid g;
  g.nb = d / kBlock;
  g.n_blocks = g.nb * (g.nb + 1) / 2;
  int sms = device_sm_count();
  if (sms <= 0) sms = 148;
  const int64_t iters = (n + kStageRows - 1) / kStageRows;
  int64_t splits = std::max<int64_t>(1, (sms / 2) / g.n_blocks);
  splits = std::min<int64_t>(splits, std::max<int64_t>(1, iters));
  g.n_splits = static_cast<int>(splits);
  return g;
}

}  // namespace

bool gram_tc_supported(int64_t n, int d, const void* x) {
  return n > 0 && d >= kBlock && d % kBlock == 0 && d <= 4096 &&
         reinterpret_cast<uintptr_t>(x) % 16 == 0;
}

int64_t gram_tc_workspace(int64_t n, int d) {
  const GramGrid g = plan_gram(n, d);
  const int64_t pairs = static_cast<int64_t>(g.n_blocks) * g.n_splits;
  return pairs * 2 * kHalf * kBlock * static_cast<int64_t>(sizeof(float)) +
         static_cast<int64_t>(g.n_splits) * (d + 1) * static_cast<int64_t>(sizeof(double)) + 1024;
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
  const int64_t pairs = static_cast<int64_t>(g.n_blocks) * g.n_splits;
  uint8_t* ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~static_cast<uintptr_t>(255));
  float* partial = reinterpret_cast<float*>(ws);
  ws += pairs * 2 * kHalf * kBlock * sizeof(float);
  double* partial_xty = reinterpret_cast<double*>(ws);
  ws += static_cast<int64_t>(g.n_splits) * d * sizeof(double);
  double* partial_yty = reinterpret_cast<double*>(ws);
  const int smem_bytes = static_cast<int>(sizeof(SmemLayout));
  static bool attr_set = false;
  if (!attr_set) {
    BB_CUDA_OK(cudaFuncSetAttribute(gram_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    attr_set = true;
  }
  const int grid = static_cast<int>(2 * pairs);
  gram_pair_kernel<<<grid, kThreads, smem_bytes, stream>>>(x, y, n, d, g.n_blocks, g.n_splits, partial,
                                                           partial_xty, partial_yty);
  BB_CHECK_LAUNCH("gram_pair_kernel");
  const int64_t total = static_cast<int64_t>(d) * d;
  gram_finalize_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, stream>>>(
      partial, partial_xty, partial_yty, d, g.n_blocks, g.n_splits, xtx, xty, yty);
  BB_CHECK_LAUNCH("gram_finalize_kernel");
  return BB_OK;
}

}  // namespace bb
