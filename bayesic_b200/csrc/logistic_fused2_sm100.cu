// Reparameterised-gradient pass of Bayesian logistic regression (BASELINE cfg5), second design of
// the single-kernel pass: every element of X is loaded from HBM once AND converted once.
//     Z = X W^T;   loglik[s] = sum_n y_n z_ns - log(1 + exp z_ns);   G[d,s] = sum_n x_nd (y_n - sigmoid(z_ns))
// (plans of sum(ycol * Z - log(1 + exp(Z)), 0) and dot(X.T, ycol - (1 + exp(-1 * Z)) ** -1),
// Z = dot(X, Wm.T); bayesic/algebra.py:1435-1448 vocabulary, README.md:47-51.)
//
// logistic_fused_sm100.cu keeps W resident and streams X through a staging ring twice (K-major for
// Z, MN-major for G, the second time from L2); its converter warps are issue-bound (ncu: 63 % issue
// utilisation, every element split to BF16 twice).  Here the roles are swapped:
//   * a 64-row X tile is split once into error-compensated BF16 (x = b1 + b2) and stays in shared
//     memory (128 KB at D = 512) in ONE SWIZZLE_128B image, [64-feature chunk][b1 | b2][64 rows x 128 B],
//     that is both the K-major operand of the first contraction (N = 128: the 64 rows of b1, then of
//     b2) and the MN-major operand of the second (features x rows);
//   * W streams instead: a pre-kernel stacks the two BF16 parts of the 64 draws (= 128 MMA rows, in
//     blocks of 16 draws: W1 rows then W2 rows) per 64-feature chunk in the UMMA layout, and a producer
//     thread bulk-copies the L2-resident 16 KB chunks through a 4-stage mbarrier ring -- except the
//     first four chunks, which sit in the 128 TMEM columns Z and G leave free and feed their MMAs as
//     the A operand straight from TMEM (no stream, no shared-memory fetch);
//   * Z^T[(part, draw), (b1 | b2, row)] = [W1; W2] . [X1; X2]^T : ONE M128 x N128 MMA per K step gives
//     all four partial products.  Sixteen worker warps are converter and epilogue in turn (the
//     converter is idle exactly while the epilogue has work): as epilogue, a warp reads its TMEM lane
//     quadrant, adds the b1 / b2 column halves and, with one shuffle (lanes l and l ^ 16), the two W
//     parts, forms loglik terms and the residual, and writes the residual as the K-major B operand
//     (one 128-byte row per draw);
//   * G[128-feature segment] += X^T R: three M128 x N64 MMAs per 16-row K step straight from the
//     resident tile; a segment is released to the converter as its MMAs retire, and the two converter
//     groups (even / odd segments) alternate so one group's store + proxy fence overlaps the other's.
// TMEM: Z 128 columns, G up to 4 x 64 columns; accumulation chains of 2048 rows drained to fp32
// partial blocks (gram_sm100.cu explains the chain limit), float64 across CTAs in the finalize.
//
// What bounds it (tests/cuda/mma_ss_rate.cu, 1 B200): an SS-mode M128 x N x K16 BF16 MMA costs
// max(N / 2, (4096 + 32 N) / 128) cycles -- 48 for N = 64 (operand fetch from shared memory, not
// math), 64 for N = 128, 128 for N = 256.  With S = 64 draws the second contraction cannot have
// N > 64 (TMEM holds D x 64 accumulators), so a 64-row tile costs 32 x 64 + 48 x 48 = 4352 tensor
// cycles plus the converter's 128 KB of stores and the 128 KB W stream through the same
// shared-memory port: the pass runs at about half the HBM roofline, 1.2x the W-resident design.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"
#include "sm100_ptx.cuh"

namespace bb {

namespace {

constexpr int kTileRows = 64;
constexpr int kS = 64;                          // parameter draws
constexpr int kChunkBytes = 8192;               // 64 rows x 64 features, one bf16 part
constexpr int kWChunkBytes = 16384;             // [W1; W2]: 128 rows x 64 features bf16
constexpr int kWStages = 4;
constexpr int kResidPart = 8192;                // 64 draws x 64 rows bf16
constexpr int kChainTiles = 32 / BB_CHAIN_DIV;                 // G accumulators drained every 32 tiles = 2048 rows
constexpr int kWorkerWarps = 16;                // converter + epilogue
constexpr int kMmaWarp = kWorkerWarps;
constexpr int kTmaWarp = kMmaWarp + 1;
constexpr int kThreads = (kTmaWarp + 1) * 32;   // 576
constexpr int kTmemCols = 512;
constexpr int kTmemZ = 0;                       // 128 columns: [W1; W2] X1^T | [W1; W2] X2^T
constexpr int kTmemG = 128;                     // up to 4 x 64 columns
constexpr int kTmemW = 384;                     // up to 4 stacked-W chunks x 32 columns (A operand held in TMEM)
constexpr int kMaxTmemChunks = 4;

template <int kNSeg>
struct __align__(1024) Smem {
  uint8_t x[2 * kNSeg][2][kChunkBytes];         // [64-feature chunk][bf16 part]: a chunk's (b1, b2) = 128 MMA rows
  uint8_t w[kWStages][kWChunkBytes];
  uint8_t resid[2][kResidPart];
  uint64_t x_full[4], x_free[4];
  uint64_t w_full[kWStages], w_empty[kWStages];
  uint64_t z_full, z_empty, r_full, g_full, g_empty;
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
// Same MMA with an A-operand collector hint (SASS .A_KEEP / .A_REUSE): consecutive MMAs that share
// the A descriptor may keep A in the tensor core's collector buffer (fill on the first, lastuse on
// the last).  These M128 x N64 x K16 MMAs read 6 KB of operands for 32 cycles of math and run at 48
// cycles (operand fetch at 128 B/cycle); measured, the hint brings that to 45-47 cycles
// (tests/cuda/mma_ss_rate.cu) and the kernel gains 1-2 %.
__device__ __forceinline__ void mma_bf16_ss_fill(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                                 uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16.collector::a::fill [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_bf16_ss_lastuse(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                                    uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A operand from TMEM (lane = MMA row, 32-bit column j = the K pair (2 j, 2 j + 1), even element in
// the low half; tests/cuda/ts_bf16_probe.cu): no shared-memory fetch for A.
__device__ __forceinline__ void mma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
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
__device__ __forceinline__ void sts_u2(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void sts_u4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void bulk_load_keep(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar,
                                               uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(ptx::smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(ptx::smem_u32(bar)), "l"(policy)
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
// exp2 / log2 / reciprocal approximations without the range fix-ups of __expf / __logf / __fdividef:
// the arguments here are in ranges that need none (exponent <= 0, 1 <= 1 + e <= 2)
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct Fused2Params {
  int w_tmem_chunks;                // leading 64-feature chunks of the stacked W kept in TMEM (BB_FUSED2_W_TMEM, default 4)
  int prefetch;                     // L2 prefetch one converter step ahead of the register loads (BB_FUSED2_PREFETCH, default on)
  int collector;                    // A-operand collector reuse between MMAs that share A (BB_FUSED2_COLLECTOR, default on)
  int ablate;                       // developer timing experiments (BB_FUSED2_ABLATE; results are WRONG when set): 1 no global
                                    // loads, 2 no converter stores, 4 no W stream, 8 no Z MMAs, 16 no G MMAs, 32 no epilogue math;
                                    // 64 (results valid): next segment's loads issued before the proxy fence
  const float* x;
  const float* y;
  const uint8_t* wprep;             // [d / 64 chunks][16 KB UMMA image of [W1; W2]]
  float* partial_g;                 // [cta][d / 128][64 draws][128 features] fp32
  double* partial_ll;               // [cta][kWorkerWarps][32]
  int64_t n;
};

template <int kNSeg>       // d / 128
__global__ void __launch_bounds__(kThreads, 1) logistic_fused2_kernel(const Fused2Params p) {
  constexpr int kD = kNSeg * 128;
  constexpr int kChunks = 2 * kNSeg;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  Smem<kNSeg>& sm = *reinterpret_cast<Smem<kNSeg>*>(smem_raw);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t n_tiles = (p.n + kTileRows - 1) / kTileRows;
  const int64_t tile_begin = n_tiles * blockIdx.x / gridDim.x;
  const int64_t tile_end = n_tiles * (blockIdx.x + 1) / gridDim.x;
  const int T = static_cast<int>(tile_end - tile_begin);

  if (warp == kMmaWarp) {
    if (lane == 0) {
      for (int s = 0; s < 4; ++s) {
        ptx::mbar_init(&sm.x_full[s], kWorkerWarps / 2);
        ptx::mbar_init(&sm.x_free[s], 1);
      }
      for (int s = 0; s < kWStages; ++s) {
        ptx::mbar_init(&sm.w_full[s], 1);
        ptx::mbar_init(&sm.w_empty[s], 1);
      }
      ptx::mbar_init(&sm.z_full, 1);
      ptx::mbar_init(&sm.z_empty, kWorkerWarps);
      ptx::mbar_init(&sm.r_full, kWorkerWarps);
      ptx::mbar_init(&sm.g_full, 1);
      ptx::mbar_init(&sm.g_empty, kWorkerWarps);
      ptx::fence_mbar_init();
    }
    __syncwarp();
    ptx::tmem_alloc(&sm.tmem_base, kTmemCols);
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem = sm.tmem_base;
  // The first n_tm chunks of the stacked W live in TMEM for the whole kernel (128 free columns = 4
  // chunks): their MMAs take A from TMEM, so they neither stream through the ring nor fetch A from
  // shared memory.  Warp q < 4 fills its lane quadrant from the UMMA image (row = lane).
  const int n_tm = min(min(p.w_tmem_chunks, kMaxTmemChunks), kChunks);
  if (warp < 4) {
    const int r = warp * 32 + lane;
    for (int c = 0; c < n_tm; ++c) {
      const uint8_t* row = p.wprep + static_cast<int64_t>(c) * kWChunkBytes + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        const uint4 lo = __ldg(reinterpret_cast<const uint4*>(row + ((j ^ (r & 7)) << 4)));
        const uint4 hi = __ldg(reinterpret_cast<const uint4*>(row + (((j + 1) ^ (r & 7)) << 4)));
        const uint32_t v[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
        ptx::tmem_st_32x32b_x8(tmem + (static_cast<uint32_t>(warp * 32) << 16) + kTmemW + c * 32 + j * 4, v);
      }
    }
    ptx::tmem_wait_st();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();

  if (warp < kWorkerWarps) {
    // ---------------- worker warps: converter AND epilogue of every tile ----------------
    // As converter, warp w owns tile rows 4 w .. 4 w + 3: per 128-feature segment a thread handles
    // 4 float4 (rows 2 i + sub, two 64-feature halves), loaded kNB segments ahead into registers,
    // split to BF16 in place, stored once the previous tile's second contraction has released the
    // segment.  As epilogue, warp w reads TMEM lane quadrant q = w & 3 (lane = (W part, draw)) and the
    // 16 columns (tile rows) 16 cg .., cg = w >> 2: the two W parts of a draw sit in quadrants q and
    // q ^ 2, so the pair of warps swaps half of its columns through shared memory and each thread
    // finishes 8 (draw, row) values.  The converter is idle exactly while the epilogue has work
    // (all segments stored, waiting for the tile to be released), so one set of warps does both.
    const int sub = lane >> 4, c4 = lane & 15;
    const int group = warp >> 3, wi = warp & 7;      // converter group (segments seg % 2 == group), rows 8 wi .. 8 wi + 7
    uint32_t soff[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = wi * 8 + 2 * i + sub;
      soff[i] = ptx::smem_u32(sm.x[0][0]) + (r >> 3) * 1024 + (r & 7) * 128 + (((c4 >> 1) ^ (r & 7)) << 4) + (c4 & 1) * 8;
    }
    const int n_own = (kNSeg + 1 - group) / 2;       // segments of a tile this group converts
    const int total = T * n_own;
    const int64_t first_row = tile_begin * kTileRows + wi * 8 + sub;
    // running load pointer over this group's segments: +256 floats to the next one, then on to the next tile
    const float* ld_ptr = p.x + first_row * kD + group * 128 + c4 * 4;
    int64_t ld_rows_left = p.n - first_row;            // row 2 i of this thread is valid iff 2 i < ld_rows_left
    int ld_seg = group, ld_g = 0;
    uint32_t rx[8][4];                                 // [i * 2 + half][fp32 x 4, then b1[2], b2[2]]
    auto load = [&]() {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (2 * i < ld_rows_left && !(p.ablate & 1)) v = ldg_f4(ld_ptr + static_cast<int64_t>(2 * i) * kD + half * 64);
          rx[i * 2 + half][0] = __float_as_uint(v.x);
          rx[i * 2 + half][1] = __float_as_uint(v.y);
          rx[i * 2 + half][2] = __float_as_uint(v.z);
          rx[i * 2 + half][3] = __float_as_uint(v.w);
        }
      ++ld_g;
      if (ld_seg + 2 < kNSeg) {
        ld_seg += 2;
        ld_ptr += 256;
      } else {
        ld_ptr += kTileRows * kD - (ld_seg - group) * 128;
        ld_seg = group;
        ld_rows_left -= kTileRows;
      }
      // L2 prefetch of the step after this one (ld_ptr now points at it): the warp's 8 rows x 4 lines
      if (p.prefetch && !(p.ablate & 1) && ld_g < total && (lane >> 2) < ld_rows_left + sub)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(ld_ptr - (sub * kD + c4 * 4) + (lane >> 2) * kD + (lane & 3) * 32));
    };
    auto split_in_place = [&]() {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        uint32_t b1[2], b2[2];
        split_bf16(make_float4(__uint_as_float(rx[k][0]), __uint_as_float(rx[k][1]), __uint_as_float(rx[k][2]),
                               __uint_as_float(rx[k][3])), b1, b2);
        rx[k][0] = b1[0]; rx[k][1] = b1[1]; rx[k][2] = b2[0]; rx[k][3] = b2[1];
      }
    };
    auto store = [&](int seg) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const uint32_t addr = soff[i] + (2 * seg + half) * (2 * kChunkBytes);
          sts_u2(addr, rx[i * 2 + half][0], rx[i * 2 + half][1]);
          sts_u2(addr + kChunkBytes, rx[i * 2 + half][2], rx[i * 2 + half][3]);
        }
    };
    // epilogue geometry: the stacked W rows are ordered in blocks of 16 draws, (W1 of draws 16 b .. 16 b + 15,
    // then W2 of the same draws), so the two parts of a draw sit in lanes l and l ^ 16 of one TMEM lane
    // quadrant and one shuffle adds them; the two lanes then share the 16 columns (8 each)
    const int q = warp & 3, cg = warp >> 2;
    const int half = lane >> 4;
    const int s = 16 * q + (lane & 15);           // draw
    const int row_base = 16 * cg + 8 * half;      // the 8 tile rows this thread finishes = one 16-byte chunk
    const uint32_t resid_addr = ptx::smem_u32(sm.resid[0]) + (s >> 3) * 1024 + (s & 7) * 128 +
                                ((static_cast<uint32_t>(row_base >> 3) ^ (s & 7)) << 4);
    float* my_partial = p.partial_g + static_cast<int64_t>(blockIdx.x) * kNSeg * kS * 128 + q * 32 + lane;
    double ll = 0.0;
    int chains = 0;
    if (T == 0) {
      for (int c = cg * kNSeg * 16; c < (cg + 1) * kNSeg * 16; ++c) my_partial[c * 128] = 0.f;
    }
    if (total > 0) load();
    for (int t = 0; t < T; ++t) {
      // ---- converter part: this group's segments of tile t (the two groups alternate, so one group's
      //      store / proxy fence / hand-over overlaps the other's) ----
#pragma unroll
      for (int seg = 0; seg < kNSeg; ++seg) {
        if ((seg & 1) == group) {
          split_in_place();
          ptx::mbar_wait_parked(&sm.x_free[seg], (static_cast<uint32_t>(t) & 1) ^ 1);
          if (!(p.ablate & 2)) store(seg);
          if ((p.ablate & 64) && ld_g < total) load();        // experiment: next loads before the proxy fence
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&sm.x_full[seg]);
          if (!(p.ablate & 64) && ld_g < total) load();
        }
      }
      // ---- epilogue part ----
      const int64_t row0 = (tile_begin + t) * kTileRows + row_base;
      const int64_t rows_left = p.n - row0;
      const int n_valid = rows_left >= 8 ? 8 : (rows_left > 0 ? static_cast<int>(rows_left) : 0);
      float yv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) yv[j] = j < n_valid ? __ldg(p.y + row0 + j) : 0.f;   // uniform loads
      ptx::mbar_wait_parked(&sm.z_full, static_cast<uint32_t>(t) & 1);
      ptx::tc_fence_after_sync();
      float zv[8];
      {
        uint32_t v[16], v2[16];                     // products with X1 (columns 0..63) and with X2 (64..127)
        ptx::tmem_ld_32x32b_x16(tmem + (static_cast<uint32_t>(q * 32) << 16) + kTmemZ + 16 * cg, v);
        ptx::tmem_ld_32x32b_x16(tmem + (static_cast<uint32_t>(q * 32) << 16) + kTmemZ + kTileRows + 16 * cg, v2);
        ptx::tmem_wait_ld();
        ptx::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&sm.z_empty);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float lo = __uint_as_float(v[j]) + __uint_as_float(v2[j]);
          const float hi = __uint_as_float(v[8 + j]) + __uint_as_float(v2[8 + j]);
          const float lo_t = lo + __shfl_xor_sync(0xffffffffu, lo, 16);      // W1 part + W2 part
          const float hi_t = hi + __shfl_xor_sync(0xffffffffu, hi, 16);
          zv[j] = half ? hi_t : lo_t;
        }
      }
      float ll_t = 0.f;
      uint32_t rb1[4], rb2[4];
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        float res[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const bool valid = j + u < n_valid;
          const float z = zv[j + u];
          float ez = 0.5f, ope = 1.5f, softplus = z, rcp = 0.6f;
          if (!(p.ablate & 32)) {
            ez = ex2_approx(-1.4426950408889634f * fabsf(z));      // exp(-|z|) in (0, 1]
            ope = 1.f + ez;
            softplus = fmaf(0.6931471805599453f, lg2_approx(ope), fmaxf(z, 0.f));
            rcp = rcp_approx(ope);
          }
          const float sig = z >= 0.f ? rcp : ez * rcp;
          ll_t += valid ? fmaf(yv[j + u], z, -softplus) : 0.f;
          res[u] = valid ? yv[j + u] - sig : 0.f;        // rows past n contribute nothing to G
        }
        __nv_bfloat162 hi = __floats2bfloat162_rn(res[0], res[1]);
        const uint32_t hb = *reinterpret_cast<uint32_t*>(&hi);
        __nv_bfloat162 lo = __floats2bfloat162_rn(res[0] - __uint_as_float(hb << 16),
                                                 res[1] - __uint_as_float(hb & 0xFFFF0000u));
        rb1[j >> 1] = hb;
        rb2[j >> 1] = *reinterpret_cast<uint32_t*>(&lo);
      }
      ll += static_cast<double>(ll_t);
      // residual tile, K-major: row = draw (128 bytes = 64 tile rows); this thread's 8 rows = one 16-byte chunk
      sts_u4(resid_addr, rb1[0], rb1[1], rb1[2], rb1[3]);
      sts_u4(resid_addr + kResidPart, rb2[0], rb2[1], rb2[2], rb2[3]);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&sm.r_full);
      // drain G at the end of an accumulation chain (and after the last tile)
      if ((t % kChainTiles) == kChainTiles - 1 || t == T - 1) {
        ptx::mbar_wait_parked(&sm.g_full, static_cast<uint32_t>(chains) & 1);
        ptx::tc_fence_after_sync();
        const uint32_t g_addr = tmem + (static_cast<uint32_t>(q * 32) << 16) + kTmemG + cg * kNSeg * 16;
        float* dst0 = my_partial + static_cast<int64_t>(cg) * kNSeg * 16 * 128;
#pragma unroll 1
        for (int cc = 0; cc < kNSeg; ++cc) {
          uint32_t v[16];
          ptx::tmem_ld_32x32b_x16(g_addr + cc * 16, v);
          float* dst = dst0 + cc * 16 * 128;
          float old[16];
          if (chains != 0) {
#pragma unroll
            for (int j = 0; j < 16; ++j) old[j] = dst[j * 128];
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) old[j] = 0.f;
          }
          ptx::tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 16; ++j) dst[j * 128] = old[j] + __uint_as_float(v[j]);
        }
        ptx::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&sm.g_empty);
        ++chains;
      }
    }
    p.partial_ll[(static_cast<int64_t>(blockIdx.x) * kWorkerWarps + warp) * 32 + lane] = ll;
  } else if (warp == kMmaWarp) {
    // ---------------- MMA issuer ----------------
    if (ptx::elect_one()) {
      // first contraction: N = 128 = the tile's 64 rows twice (b1 part, then b2 part, contiguous per chunk)
      const uint32_t idesc_a = ptx::make_idesc(128, 2 * kTileRows, /*bf16*/ 1, /*A K-major*/ 0, /*B K-major*/ 0);
      const uint32_t idesc_b = ptx::make_idesc(128, kS, /*bf16*/ 1, /*A MN-major*/ 1, /*B K-major*/ 0);
      const uint32_t x1 = ptx::smem_u32(sm.x[0][0]), x2 = x1 + kChunkBytes;
      const uint32_t r1 = ptx::smem_u32(sm.resid[0]), r2 = r1 + kResidPart;
      int64_t it = 0;
      int chain = 0;
      for (int t = 0; t < T; ++t) {
        // ---- A(t): Z^T = [W1; W2] (X1 + X2)^T ----
        ptx::mbar_wait_parked(&sm.z_empty, (static_cast<uint32_t>(t) & 1) ^ 1);
        for (int seg = 0; seg < kNSeg; ++seg) {
          ptx::mbar_wait_parked(&sm.x_full[seg], static_cast<uint32_t>(t) & 1);
          for (int half = 0; half < 2; ++half) {
            const int c = 2 * seg + half;
            if (c < n_tm) {                       // A = stacked W chunk resident in TMEM
              ptx::tc_fence_after_sync();
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                const uint64_t b12 = ptx::make_smem_desc(x1 + c * (2 * kChunkBytes) + ks * 32, 16, 1024, ptx::kLayoutSwizzle128B);
                if (!(p.ablate & 8)) mma_bf16_ts(tmem + kTmemZ, tmem + kTmemW + c * 32 + ks * 8, b12, idesc_a, (c == 0 && ks == 0) ? 0u : 1u);
              }
              continue;
            }
            const int ws = (p.ablate & 4) ? 0 : static_cast<int>(it % kWStages);
            if (!(p.ablate & 4)) ptx::mbar_wait_parked(&sm.w_full[ws], static_cast<uint32_t>(it / kWStages) & 1);
            ptx::tc_fence_after_sync();
            const uint32_t wbase = ptx::smem_u32(sm.w[ws]);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint64_t a = ptx::make_smem_desc(wbase + ks * 32, 16, 1024, ptx::kLayoutSwizzle128B);
              const uint64_t b12 = ptx::make_smem_desc(x1 + c * (2 * kChunkBytes) + ks * 32, 16, 1024, ptx::kLayoutSwizzle128B);
              if (!(p.ablate & 8)) mma_bf16_ss(tmem + kTmemZ, a, b12, idesc_a, (c == 0 && ks == 0) ? 0u : 1u);
            }
            if (!(p.ablate & 4)) ptx::mma_commit(&sm.w_empty[ws]);
            ++it;
          }
        }
        ptx::mma_commit(&sm.z_full);
        // ---- B(t): G += X^T R ----
        const bool first_in_chain = (t % kChainTiles) == 0;
        if (first_in_chain && chain > 0) ptx::mbar_wait_parked(&sm.g_empty, static_cast<uint32_t>(chain - 1) & 1);
        ptx::mbar_wait_parked(&sm.r_full, static_cast<uint32_t>(t) & 1);
        ptx::tc_fence_after_sync();
        for (int seg = 0; seg < kNSeg; ++seg) {
          const uint32_t d_tmem = tmem + kTmemG + seg * kS;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            // features: two 64-wide MN atoms (chunks 2 seg, 2 seg + 1, 16 KB apart); rows: 8-row groups 1 KB apart
            const uint64_t a1 = ptx::make_smem_desc(x1 + 2 * seg * (2 * kChunkBytes) + ks * 2048, 2 * kChunkBytes, 1024, ptx::kLayoutSwizzle128B);
            const uint64_t a2 = ptx::make_smem_desc(x2 + 2 * seg * (2 * kChunkBytes) + ks * 2048, 2 * kChunkBytes, 1024, ptx::kLayoutSwizzle128B);
            const uint64_t rb1 = ptx::make_smem_desc(r1 + ks * 32, 16, 1024, ptx::kLayoutSwizzle128B);
            const uint64_t rb2 = ptx::make_smem_desc(r2 + ks * 32, 16, 1024, ptx::kLayoutSwizzle128B);
            if (p.ablate & 16) continue;
            if (p.collector) {
              mma_bf16_ss_fill(d_tmem, a1, rb1, idesc_b, (first_in_chain && ks == 0) ? 0u : 1u);
              mma_bf16_ss_lastuse(d_tmem, a1, rb2, idesc_b, 1u);
            } else {
              mma_bf16_ss(d_tmem, a1, rb1, idesc_b, (first_in_chain && ks == 0) ? 0u : 1u);
              mma_bf16_ss(d_tmem, a1, rb2, idesc_b, 1u);
            }
            mma_bf16_ss(d_tmem, a2, rb1, idesc_b, 1u);
          }
          ptx::mma_commit(&sm.x_free[seg]);
        }
        if ((t % kChainTiles) == kChainTiles - 1 || t == T - 1) {
          ptx::mma_commit(&sm.g_full);
          ++chain;
        }
      }
    }
  } else {
    // ---------------- W producer: bulk copies of the L2-resident stacked chunks ----------------
    if (lane == 0 && !(p.ablate & 4)) {
      uint64_t keep;
      asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(keep));
      int64_t it = 0;
      for (int t = 0; t < T; ++t)
        for (int c = n_tm; c < kChunks; ++c, ++it) {
          const int ws = static_cast<int>(it % kWStages);
          ptx::mbar_wait_parked(&sm.w_empty[ws], (static_cast<uint32_t>(it / kWStages) & 1) ^ 1);
          ptx::mbar_arrive_expect_tx(&sm.w_full[ws], kWChunkBytes);
          bulk_load_keep(sm.w[ws], p.wprep + static_cast<int64_t>(c) * kWChunkBytes, kWChunkBytes, &sm.w_full[ws], keep);
        }
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == kMmaWarp) ptx::tmem_dealloc(tmem, kTmemCols);
}

// W[s, d] float32 -> per 64-feature chunk the UMMA image of [W1; W2] (128 rows x 128 bytes,
// K-major SWIZZLE_128B): row = (draw / 16) * 32 + part * 16 + draw % 16
__global__ void prep_w_fused2_kernel(const float* __restrict__ w, int d, uint8_t* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;       // one thread per (draw, 8 features)
  const int groups = d / 8;
  if (idx >= kS * groups) return;
  const int s = idx / groups, g8 = idx - s * groups;
  const int c = g8 >> 3, j = g8 & 7;
  uint32_t b1[4], b2[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const float x0 = w[static_cast<int64_t>(s) * d + g8 * 8 + 2 * u], x1 = w[static_cast<int64_t>(s) * d + g8 * 8 + 2 * u + 1];
    const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
    const __nv_bfloat16 l0 = __float2bfloat16_rn(x0 - __bfloat162float(h0)), l1 = __float2bfloat16_rn(x1 - __bfloat162float(h1));
    b1[u] = static_cast<uint32_t>(__bfloat16_as_ushort(h0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(h1)) << 16);
    b2[u] = static_cast<uint32_t>(__bfloat16_as_ushort(l0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(l1)) << 16);
  }
  uint8_t* chunk = out + static_cast<int64_t>(c) * kWChunkBytes;
  const int r1 = (s >> 4) * 32 + (s & 15), r2 = r1 + 16;      // blocks of 16 draws: W1 rows, then W2 rows
  *reinterpret_cast<uint4*>(chunk + (r1 >> 3) * 1024 + (r1 & 7) * 128 + ((j ^ (r1 & 7)) << 4)) = make_uint4(b1[0], b1[1], b1[2], b1[3]);
  *reinterpret_cast<uint4*>(chunk + (r2 >> 3) * 1024 + (r2 & 7) * 128 + ((j ^ (r2 & 7)) << 4)) = make_uint4(b2[0], b2[1], b2[2], b2[3]);
}

// G[d, s] (float64) = sum over CTAs of partial_g[cta][d / 128][s][d % 128];
// loglik[s] = sum over CTAs and over the four worker warps of the draw's TMEM quadrant, two lanes each
__global__ void __launch_bounds__(256)
logistic_fused2_finalize_kernel(const float* __restrict__ partial_g, const double* __restrict__ partial_ll,
                                int n_ctas, int d, double* __restrict__ g_out, double* __restrict__ ll_out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < d * kS) {
    const int col = idx / d, row = idx % d;
    const int64_t per_cta = static_cast<int64_t>(d) * kS;
    const int64_t off = (static_cast<int64_t>(row / 128) * kS + col) * 128 + row % 128;
    double acc = 0.0;
    for (int c = 0; c < n_ctas; ++c) acc += static_cast<double>(partial_g[c * per_cta + off]);
    g_out[static_cast<int64_t>(row) * kS + col] = acc;
  } else if (idx < d * kS + kS) {
    const int s = idx - d * kS;
    double acc = 0.0;
    for (int c = 0; c < n_ctas; ++c)
      for (int w = (s >> 4); w < kWorkerWarps; w += 4)
        acc += partial_ll[(static_cast<int64_t>(c) * kWorkerWarps + w) * 32 + (s & 15)] +
               partial_ll[(static_cast<int64_t>(c) * kWorkerWarps + w) * 32 + (s & 15) + 16];
    ll_out[s] = acc;
  }
}

int fused2_grid(int64_t n) {
  int sms = device_sm_count();
  if (sms <= 0) sms = 148;
  const int64_t tiles = (n + kTileRows - 1) / kTileRows;
  return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(sms, tiles)));
}

template <int kNSeg>
int launch_fused2_instance(const Fused2Params& p, int grid, cudaStream_t stream) {
  const int smem_bytes = static_cast<int>(sizeof(Smem<kNSeg>));
  static SmemOptIn smem_opt_in;
  BB_CUDA_OK(smem_opt_in.ensure(logistic_fused2_kernel<kNSeg>, smem_bytes));
  logistic_fused2_kernel<kNSeg><<<grid, kThreads, smem_bytes, stream>>>(p);
  BB_CHECK_LAUNCH("logistic_fused2_kernel");
  return BB_OK;
}

}  // namespace

bool logistic_fused2_supported(int64_t n, int d, int s, const void* x) {
  return n > 0 && s == kS && d >= 128 && d % 128 == 0 && d <= 512 && reinterpret_cast<uintptr_t>(x) % 16 == 0;
}

int64_t logistic_fused2_workspace(int64_t n, int d, int s) {
  const int grid = fused2_grid(n);
  return align_up(static_cast<int64_t>(d / 64) * kWChunkBytes, 256) + static_cast<int64_t>(grid) * d * s * 4 +
         static_cast<int64_t>(grid) * kWorkerWarps * 32 * 8 + 1024;
}

int launch_logistic_fused2(const float* x, const float* y, const float* w, int64_t n, int d, int s, double* loglik,
                           double* g, void* workspace, int64_t workspace_bytes, cudaStream_t stream) {
  if (!logistic_fused2_supported(n, d, s, x)) {
    set_error("logistic_fused2: unsupported shape n=%lld d=%d s=%d", static_cast<long long>(n), d, s);
    return BB_ERR_UNSUPPORTED;
  }
  if (workspace == nullptr || workspace_bytes < logistic_fused2_workspace(n, d, s)) {
    set_error("logistic_fused2: workspace %lld < %lld bytes", static_cast<long long>(workspace_bytes),
              static_cast<long long>(logistic_fused2_workspace(n, d, s)));
    return BB_ERR_WORKSPACE;
  }
  const int grid = fused2_grid(n);
  uint8_t* ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~static_cast<uintptr_t>(255));
  uint8_t* wprep = ws;
  ws += align_up(static_cast<int64_t>(d / 64) * kWChunkBytes, 256);
  float* partial_g = reinterpret_cast<float*>(ws);
  ws += static_cast<int64_t>(grid) * d * s * 4;
  double* partial_ll = reinterpret_cast<double*>(ws);
  const int prep_threads = s * d / 8;
  prep_w_fused2_kernel<<<(prep_threads + 255) / 256, 256, 0, stream>>>(w, d, wprep);
  BB_CHECK_LAUNCH("prep_w_fused2_kernel");
  static const int collector = getenv("BB_FUSED2_COLLECTOR") ? atoi(getenv("BB_FUSED2_COLLECTOR")) : 1;
  Fused2Params p;
  p.collector = collector;
  static const int w_tmem = getenv("BB_FUSED2_W_TMEM") ? atoi(getenv("BB_FUSED2_W_TMEM")) : 4;
  p.w_tmem_chunks = w_tmem;
  static const int prefetch = getenv("BB_FUSED2_PREFETCH") ? atoi(getenv("BB_FUSED2_PREFETCH")) : 1;
  p.prefetch = prefetch;
  static const int ablate = getenv("BB_FUSED2_ABLATE") ? atoi(getenv("BB_FUSED2_ABLATE")) : 0;
  p.ablate = ablate;
  p.x = x; p.y = y; p.wprep = wprep; p.partial_g = partial_g; p.partial_ll = partial_ll; p.n = n;
  switch (d / 128) {
    case 1: BB_TRY(launch_fused2_instance<1>(p, grid, stream)); break;
    case 2: BB_TRY(launch_fused2_instance<2>(p, grid, stream)); break;
    case 3: BB_TRY(launch_fused2_instance<3>(p, grid, stream)); break;
    default: BB_TRY(launch_fused2_instance<4>(p, grid, stream)); break;
  }
  logistic_fused2_finalize_kernel<<<(d * s + s + 255) / 256, 256, 0, stream>>>(partial_g, partial_ll, grid, d, g, loglik);
  BB_CHECK_LAUNCH("logistic_fused2_finalize_kernel");
  return BB_OK;
}

}  // namespace bb
