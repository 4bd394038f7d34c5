// Reparameterised-gradient pass of Bayesian logistic regression (BASELINE cfg5), third design: the same 64 data
// rows on a CTA PAIR that splits the FEATURE axis, two tiles in flight.
//     Z = X W^T;   loglik[s] = sum_n y_n z_ns - log(1 + exp z_ns);   G[d,s] = sum_n x_nd (y_n - sigmoid(z_ns))
// (plans of sum(ycol * Z - log(1 + exp(Z)), 0) and dot(X.T, ycol - (1 + exp(-1 * Z)) ** -1),
// Z = dot(X, Wm.T); bayesic/algebra.py:1435-1448 vocabulary, README.md:47-51.)
//
// What was wrong with logistic_fused2_sm100.cu (one CTA per 64-row tile, all D features): the
// second contraction needs the residual of the WHOLE first one and the 128 KB X tile is single-buffered, so one tile
// is in flight on a chain of ~8 dependent hand-offs (2.5 us per tile with all work removed), and a tile pushes
// 688 KB through the 128 B/clk shared-memory port -- 94 % of its HBM time.  Here CTA r of a cluster of two owns
// features [D/2 r, +D/2) of the SAME rows:
//   * its X half-tile is 64 KB, so TWO fit: the converter fills tile t+1 while tile t is in its epilogue and
//     second contraction, and Z is double-buffered in TMEM (2 x 128 columns) -- MMA issue order A(t+1), B(t);
//   * its half of the stacked W (4 chunks x 32 columns) lives entirely in TMEM: every first-contraction MMA takes A
//     from TMEM, there is no W stream and no producer warp; G needs 128 columns instead of 256;
//   * the first contraction runs over half of the features, so each CTA holds a PARTIAL Z: the 16 worker warps
//     push their partial sums (64 draws x 64 rows x 4 B = 16 KB) into the peer's shared memory (st.async with
//     mbarrier complete_tx: no fences), wait for the peer's, and add -- a + b on one side, b + a on the other, so
//     both CTAs see bit-identical Z and form the same residual tile locally (the epilogue math is done twice; the
//     MUFU pipe has the room);
//   * shared-memory traffic per CTA and tile: 64 KB of B operands for the first contraction, 144 KB for the second,
//     64 + 16 KB of converter / residual stores, 32 KB of exchange: 320 KB against 2 850 cycles of HBM time for the
//     half tile (88 % of the port at 100 % efficiency per 2 500 cycles -- was 94 % of 5 700 with nothing overlapped).
// Shapes: S = 64 draws, D in {256, 512}.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"
#include "sm100_pair.cuh"
#include "sm100_ptx.cuh"

namespace bb {

namespace {

constexpr int kTileRows = 64;
constexpr int kS = 64;                          // parameter draws
constexpr int kChunkBytes = 8192;               // 64 rows x 64 features, one bf16 part
constexpr int kWChunkBytes = 16384;             // [W1; W2]: 128 rows x 64 features bf16 (prep image, global)
constexpr int kResidPart = 8192;                // 64 draws x 64 rows bf16
constexpr int kChainTiles = 32 / BB_CHAIN_DIV;  // G accumulators drained every 32 tiles = 2048 rows
constexpr int kConvWarps = 8;                   // warps 0-7: converter
constexpr int kWorkerWarps = 8;                 // warps 8-15: epilogue
constexpr int kMmaWarp = kConvWarps + kWorkerWarps;
constexpr int kThreads = (kMmaWarp + 1) * 32;   // 544
constexpr int kTmemCols = 512;
constexpr int kTmemZ = 0;                       // 2 x 128 columns: [W1; W2] X1^T | [W1; W2] X2^T, double-buffered
constexpr int kTmemG = 256;                     // kSegC x 64 columns
constexpr int kTmemW = 384;                     // 2 kSegC stacked-W chunks x 32 columns (A operand held in TMEM)

template <int kSegC>      // 128-feature segments per CTA = d / 256
struct __align__(1024) Smem {
  uint8_t x[2][2 * kSegC][2][kChunkBytes];      // [tile parity][64-feature chunk][bf16 part]
  uint8_t resid[2][2][kResidPart];              // [tile parity][bf16 part]
  float zx[2][kS][kTileRows];                   // [tile parity]: the PEER's partial Z (draw-major), written by the peer
  uint64_t x_full[2][kSegC], x_free[2][kSegC];
  uint64_t z_full[2], z_empty[2], r_full[2], zx_full[2], zx_free[2];
  uint64_t g_full, g_empty;
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
// A operand from TMEM (lane = MMA row, 32-bit column j = the K pair (2 j, 2 j + 1); tests/cuda/ts_bf16_probe.cu)
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
// Four floats into the PEER CTA's shared memory at the same offset as `local_addr`, as an ASYNC store that
// completes 16 bytes of the transaction count of the peer's mbarrier at the same offset as `local_bar`: the
// receiver's barrier phase completes when all 16 KB of a tile's partial Z have landed -- no fence on either side
// (a cluster-scope fence here is a MEMBAR.ALL.GPU that waits for the converter's global loads in flight: the
// first version of this kernel ran at 3.8 ms against 2.4 for logistic_fused2).
__device__ __forceinline__ void st_peer_async_f4(uint32_t local_addr, uint32_t local_bar, uint32_t peer, float a, float b,
                                                 float c, float d) {
  asm volatile(
      "{\n\t.reg .b32 ra, rb;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %2;\n\t"
      "mapa.shared::cluster.u32 rb, %1, %2;\n\t"
      "st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [ra], {%3, %4, %5, %6}, [rb];\n\t}\n"
      ::"r"(local_addr), "r"(local_bar), "r"(peer), "r"(__float_as_uint(a)), "r"(__float_as_uint(b)),
        "r"(__float_as_uint(c)), "r"(__float_as_uint(d))
      : "memory");
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
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

struct Fused3Params {
  const float* x;
  const float* y;
  const uint8_t* wprep;             // [d / 64 chunks][16 KB UMMA image of [W1; W2]]
  float* partial_g;                 // [pair][d / 128][64 draws][128 features] fp32
  double* partial_ll;               // [pair][kWorkerWarps][32]
  int64_t n;
  int ablate;                       // timing experiments (BB_FUSED3_ABLATE; results WRONG when set): 1 no Z exchange, 2 no global
                                    // loads, 4 no G MMAs, 8 no Z MMAs, 16 no converter stores, 32 no epilogue math
};

template <int kSegC>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1) logistic_fused3_kernel(const Fused3Params p) {
  constexpr int kD = kSegC * 256;                 // all features
  constexpr int kChunksC = 2 * kSegC;             // 64-feature chunks of this CTA
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  Smem<kSegC>& sm = *reinterpret_cast<Smem<kSegC>*>(smem_raw);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = pair::cluster_ctarank();
  const int pr = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int64_t n_tiles = (p.n + kTileRows - 1) / kTileRows;
  const int64_t tile_begin = n_tiles * pr / n_pairs;
  const int64_t tile_end = n_tiles * (pr + 1) / n_pairs;
  const int T = static_cast<int>(tile_end - tile_begin);
  const int feat0 = static_cast<int>(rank) * (kD / 2);       // this CTA's first feature

  if (warp == kMmaWarp) {
    if (lane == 0) {
      for (int b = 0; b < 2; ++b) {
        for (int s = 0; s < kSegC; ++s) {
          ptx::mbar_init(&sm.x_full[b][s], kConvWarps);
          ptx::mbar_init(&sm.x_free[b][s], 1);
        }
        ptx::mbar_init(&sm.z_full[b], 1);
        ptx::mbar_init(&sm.z_empty[b], kWorkerWarps);
        ptx::mbar_init(&sm.r_full[b], kWorkerWarps);
        ptx::mbar_init(&sm.zx_full[b], 1);                  // one local arming arrive + 16 KB of the peer's async stores
        ptx::mbar_init(&sm.zx_free[b], kWorkerWarps);       // the PEER's epilogue warps: "your partial has been read"
      }
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
  // this CTA's stacked-W chunks into TMEM (row = lane): they stay for the whole kernel
  if (warp < 4) {
    const int r = warp * 32 + lane;
    for (int c = 0; c < kChunksC; ++c) {
      const uint8_t* row = p.wprep + static_cast<int64_t>(rank * kChunksC + c) * kWChunkBytes + (r >> 3) * 1024 + (r & 7) * 128;
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
  pair::cluster_sync_all();        // both CTAs' barriers initialised before any remote arrive / remote store
  ptx::tc_fence_after_sync();

  if (warp < kConvWarps) {
    // ---------------- converter warps (8): every segment of every tile, one segment ahead in registers ----------------
    // a thread handles rows 8 warp + 2 i + sub (i < 4) and two 64-feature halves of a 128-feature segment: 8 float4
    const int sub = lane >> 4, c4 = lane & 15;
    const int wi = warp;
    uint32_t soff[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = wi * 8 + 2 * i + sub;
      soff[i] = (r >> 3) * 1024 + (r & 7) * 128 + (((c4 >> 1) ^ (r & 7)) << 4) + (c4 & 1) * 8;
    }
    const uint32_t x_base = ptx::smem_u32(sm.x[0][0][0]);
    constexpr uint32_t kXBuf = 2 * kSegC * 2 * kChunkBytes;      // bytes of one tile buffer
    const int total = T * kSegC;
    const int64_t first_row = tile_begin * kTileRows + wi * 8 + sub;
    const float* ld_ptr = p.x + first_row * kD + feat0 + c4 * 4;
    int64_t ld_rows_left = p.n - first_row;
    int ld_seg = 0, ld_g = 0;
    uint32_t rx[8][4];
    auto load = [&]() {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (2 * i < ld_rows_left && !(p.ablate & 2)) v = ldg_f4(ld_ptr + static_cast<int64_t>(2 * i) * kD + half * 64);
          rx[i * 2 + half][0] = __float_as_uint(v.x);
          rx[i * 2 + half][1] = __float_as_uint(v.y);
          rx[i * 2 + half][2] = __float_as_uint(v.z);
          rx[i * 2 + half][3] = __float_as_uint(v.w);
        }
      ++ld_g;
      if (ld_seg + 1 < kSegC) {
        ld_seg += 1;
        ld_ptr += 128;
      } else {
        ld_ptr += kTileRows * kD - ld_seg * 128;
        ld_seg = 0;
        ld_rows_left -= kTileRows;
      }
      if (ld_g < total && !(p.ablate & 2) && (lane >> 2) < ld_rows_left + sub)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(ld_ptr - (sub * kD + c4 * 4) + (lane >> 2) * kD + (lane & 3) * 32));
    };
    if (total > 0) load();
    for (int t = 0; t < T; ++t) {
      const int b = t & 1;
      const uint32_t use = static_cast<uint32_t>(t >> 1);
#pragma unroll
      for (int seg = 0; seg < kSegC; ++seg) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          uint32_t b1[2], b2[2];
          split_bf16(make_float4(__uint_as_float(rx[k][0]), __uint_as_float(rx[k][1]), __uint_as_float(rx[k][2]),
                                 __uint_as_float(rx[k][3])), b1, b2);
          rx[k][0] = b1[0]; rx[k][1] = b1[1]; rx[k][2] = b2[0]; rx[k][3] = b2[1];
        }
        ptx::mbar_wait_parked(&sm.x_free[b][seg], (use & 1) ^ 1);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const uint32_t addr = x_base + b * kXBuf + (2 * seg + half) * (2 * kChunkBytes) + soff[i];
            if (p.ablate & 16) continue;
            sts_u2(addr, rx[i * 2 + half][0], rx[i * 2 + half][1]);
            sts_u2(addr + kChunkBytes, rx[i * 2 + half][2], rx[i * 2 + half][3]);
          }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&sm.x_full[b][seg]);
        if (ld_g < total) load();
      }
    }
  } else if (warp < kMmaWarp) {
    // ---------------- epilogue warps (8): send(t + 1), then finish(t) ----------------
    // warp -> (TMEM lane quadrant q, tile rows 32 cgh .. + 31 in two passes of 16); after one shuffle a thread owns
    // draw s and 8 tile rows of the pass
    const int ew = warp - kConvWarps;
    const int q = ew & 3, cgh = ew >> 2;
    const int half = lane >> 4;
    const int s = 16 * q + (lane & 15);
    const uint32_t resid_base = ptx::smem_u32(sm.resid[0][0]);
    const uint32_t zx0 = ptx::smem_u32(&sm.zx[0][0][0]);
    float* my_partial = p.partial_g + (static_cast<int64_t>(pr) * (kD / 128) + rank * kSegC) * kS * 128 + q * 32 + lane;
    double ll = 0.0;
    int chains = 0;
    if (T == 0) {
      for (int c = 2 * cgh * kSegC * 16; c < (2 * cgh + 2) * kSegC * 16; ++c) my_partial[c * 128] = 0.f;
    }
    // this CTA's partial Z of (draw s, 8 rows of pass cgi) out of TMEM buffer b
    auto partial_z = [&](int b, int cg, float (&zv)[8]) {
      uint32_t v[16], v2[16];
      const uint32_t z_addr = tmem + (static_cast<uint32_t>(q * 32) << 16) + kTmemZ + b * 128 + 16 * cg;
      ptx::tmem_ld_32x32b_x16(z_addr, v);
      ptx::tmem_ld_32x32b_x16(z_addr + kTileRows, v2);
      ptx::tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float lo = __uint_as_float(v[j]) + __uint_as_float(v2[j]);
        const float hi = __uint_as_float(v[8 + j]) + __uint_as_float(v2[8 + j]);
        const float lo_t = lo + __shfl_xor_sync(0xffffffffu, lo, 16);      // W1 part + W2 part
        const float hi_t = hi + __shfl_xor_sync(0xffffffffu, hi, 16);
        zv[j] = half ? hi_t : lo_t;
      }
    };
    // first half of a tile's epilogue: the partial Z on its way to the peer (st.async, ~1.7 us round trip, hidden
    // behind the second half of the previous tile)
    auto send = [&](int t) {
      const int b = t & 1;
      const uint32_t use = static_cast<uint32_t>(t >> 1);
      // the peer has taken tile t - 2's partial out of its zx[b] (its finish(t - 2))
      if (t >= 2 && !(p.ablate & 1)) pair::mbar_wait_cluster(&sm.zx_free[b], (use - 1) & 1);
      ptx::mbar_wait_parked(&sm.z_full[b], use & 1);
      ptx::tc_fence_after_sync();
      if (p.ablate & 1) return;
      const uint32_t bar = ptx::smem_u32(&sm.zx_full[b]);
      if (ew == 0 && lane == 0) ptx::mbar_arrive_expect_tx(&sm.zx_full[b], kS * kTileRows * 4);    // arm this phase
#pragma unroll
      for (int cgi = 0; cgi < 2; ++cgi) {
        const int cg = 2 * cgh + cgi;
        float zv[8];
        partial_z(b, cg, zv);
        const uint32_t mine = zx0 + b * (kS * kTileRows * 4) + (s * kTileRows + 16 * cg + 8 * half) * 4;
        st_peer_async_f4(mine, bar, rank ^ 1u, zv[0], zv[1], zv[2], zv[3]);
        st_peer_async_f4(mine + 16, bar, rank ^ 1u, zv[4], zv[5], zv[6], zv[7]);
      }
    };
    if (T > 0) send(0);
    for (int t = 0; t < T; ++t) {
      if (t + 1 < T) send(t + 1);
      // ---- second half of tile t: Z = this CTA's partial (read from TMEM again) + the peer's; both sides add the
      //      same two numbers (a + b == b + a exactly), so the two CTAs form the same residual tile ----
      const int b = t & 1;
      const uint32_t use = static_cast<uint32_t>(t >> 1);
      if (!(p.ablate & 1)) pair::mbar_wait_cluster(&sm.zx_full[b], use & 1);
      ptx::tc_fence_after_sync();
      float ll_t = 0.f;
#pragma unroll
      for (int cgi = 0; cgi < 2; ++cgi) {
        const int cg = 2 * cgh + cgi;
        const int row_base = 16 * cg + 8 * half;
        const int64_t row0 = (tile_begin + t) * kTileRows + row_base;
        const int64_t rows_left = p.n - row0;
        const int n_valid = rows_left >= 8 ? 8 : (rows_left > 0 ? static_cast<int>(rows_left) : 0);
        float yv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) yv[j] = j < n_valid ? __ldg(p.y + row0 + j) : 0.f;
        float zv[8];
        partial_z(b, cg, zv);
        if (!(p.ablate & 1)) {
          const uint32_t mine = zx0 + b * (kS * kTileRows * 4) + (s * kTileRows + row_base) * 4;
          const float4 p0 = lds_f4(mine);
          const float4 p1 = lds_f4(mine + 16);
          zv[0] += p0.x; zv[1] += p0.y; zv[2] += p0.z; zv[3] += p0.w;
          zv[4] += p1.x; zv[5] += p1.y; zv[6] += p1.z; zv[7] += p1.w;
        }
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
          __nv_bfloat162 hi2 = __floats2bfloat162_rn(res[0], res[1]);
          const uint32_t hb = *reinterpret_cast<uint32_t*>(&hi2);
          __nv_bfloat162 lo2 = __floats2bfloat162_rn(res[0] - __uint_as_float(hb << 16),
                                                    res[1] - __uint_as_float(hb & 0xFFFF0000u));
          rb1[j >> 1] = hb;
          rb2[j >> 1] = *reinterpret_cast<uint32_t*>(&lo2);
        }
        // residual tile of parity b (free: the second contraction of tile t - 2 retired before z_full of tile t)
        const uint32_t raddr = resid_base + b * (2 * kResidPart) + (s >> 3) * 1024 + (s & 7) * 128 +
                               ((static_cast<uint32_t>(row_base >> 3) ^ (s & 7)) << 4);
        sts_u4(raddr, rb1[0], rb1[1], rb1[2], rb1[3]);
        sts_u4(raddr + kResidPart, rb2[0], rb2[1], rb2[2], rb2[3]);
      }
      ll += static_cast<double>(ll_t);
      ptx::tc_fence_before_sync();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive(&sm.z_empty[b]);                                      // Z buffer b read for the last time
        ptx::mbar_arrive(&sm.r_full[b]);
        if (!(p.ablate & 1)) pair::mbar_arrive_cluster_relaxed(&sm.zx_free[b], rank ^ 1u);   // the peer may overwrite my zx[b]
      }
      // drain G at the end of an accumulation chain (and after the last tile)
      if ((t % kChainTiles) == kChainTiles - 1 || t == T - 1) {
        ptx::mbar_wait_parked(&sm.g_full, static_cast<uint32_t>(chains) & 1);
        ptx::tc_fence_after_sync();
#pragma unroll 1
        for (int cgi = 0; cgi < 2; ++cgi) {
          const int cg = 2 * cgh + cgi;
          const uint32_t g_addr = tmem + (static_cast<uint32_t>(q * 32) << 16) + kTmemG + cg * kSegC * 16;
          float* dst0 = my_partial + static_cast<int64_t>(cg) * kSegC * 16 * 128;
#pragma unroll 1
          for (int cc = 0; cc < kSegC; ++cc) {
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
        }
        ptx::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&sm.g_empty);
        ++chains;
      }
    }
    if (rank == 0) p.partial_ll[(static_cast<int64_t>(pr) * kWorkerWarps + ew) * 32 + lane] = ll;
  } else {
    // ---------------- MMA issuer: A(0); then per tile A(t + 1), B(t) ----------------
    if (ptx::elect_one()) {
      const uint32_t idesc_a = ptx::make_idesc(128, 2 * kTileRows, /*bf16*/ 1, /*A K-major*/ 0, /*B K-major*/ 0);
      const uint32_t idesc_b = ptx::make_idesc(128, kS, /*bf16*/ 1, /*A MN-major*/ 1, /*B K-major*/ 0);
      const uint32_t x_base = ptx::smem_u32(sm.x[0][0][0]);
      constexpr uint32_t kXBuf = 2 * kSegC * 2 * kChunkBytes;
      const uint32_t r_base = ptx::smem_u32(sm.resid[0][0]);
      auto first_contraction = [&](int t) {        // Z^T (this CTA's features) = [W1; W2] (X1 + X2)^T, A from TMEM
        const int b = t & 1;
        const uint32_t use = static_cast<uint32_t>(t >> 1);
        ptx::mbar_wait_parked(&sm.z_empty[b], (use & 1) ^ 1);
        for (int seg = 0; seg < kSegC; ++seg) {
          ptx::mbar_wait_parked(&sm.x_full[b][seg], use & 1);
          ptx::tc_fence_after_sync();
          for (int hf = 0; hf < 2; ++hf) {
            const int c = 2 * seg + hf;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint64_t b12 = ptx::make_smem_desc(x_base + b * kXBuf + c * (2 * kChunkBytes) + ks * 32, 16, 1024,
                                                       ptx::kLayoutSwizzle128B);
              if (!(p.ablate & 8))
                mma_bf16_ts(tmem + kTmemZ + b * 128, tmem + kTmemW + c * 32 + ks * 8, b12, idesc_a,
                            (c == 0 && ks == 0) ? 0u : 1u);
            }
          }
        }
        ptx::mma_commit(&sm.z_full[b]);
      };
      int chain = 0;
      if (T > 0) first_contraction(0);
      for (int t = 0; t < T; ++t) {
        if (t + 1 < T) first_contraction(t + 1);
        // ---- B(t): G (this CTA's features) += X^T R ----
        const int b = t & 1;
        const uint32_t use = static_cast<uint32_t>(t >> 1);
        const bool first_in_chain = (t % kChainTiles) == 0;
        if (first_in_chain && chain > 0) ptx::mbar_wait_parked(&sm.g_empty, static_cast<uint32_t>(chain - 1) & 1);
        ptx::mbar_wait_parked(&sm.r_full[b], use & 1);
        ptx::tc_fence_after_sync();
        const uint32_t x1 = x_base + b * kXBuf, x2 = x1 + kChunkBytes;
        const uint32_t r1 = r_base + b * (2 * kResidPart), r2 = r1 + kResidPart;
        for (int seg = 0; seg < kSegC; ++seg) {
          const uint32_t d_tmem = tmem + kTmemG + seg * kS;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            // features: two 64-wide MN atoms (chunks 2 seg, 2 seg + 1, 16 KB apart); rows: 8-row groups 1 KB apart
            const uint64_t a1 = ptx::make_smem_desc(x1 + 2 * seg * (2 * kChunkBytes) + ks * 2048, 2 * kChunkBytes, 1024, ptx::kLayoutSwizzle128B);
            const uint64_t a2 = ptx::make_smem_desc(x2 + 2 * seg * (2 * kChunkBytes) + ks * 2048, 2 * kChunkBytes, 1024, ptx::kLayoutSwizzle128B);
            const uint64_t rb1 = ptx::make_smem_desc(r1 + ks * 32, 16, 1024, ptx::kLayoutSwizzle128B);
            const uint64_t rb2 = ptx::make_smem_desc(r2 + ks * 32, 16, 1024, ptx::kLayoutSwizzle128B);
            if (p.ablate & 4) continue;
            mma_bf16_ss_fill(d_tmem, a1, rb1, idesc_b, (first_in_chain && ks == 0) ? 0u : 1u);
            mma_bf16_ss_lastuse(d_tmem, a1, rb2, idesc_b, 1u);
            mma_bf16_ss(d_tmem, a2, rb1, idesc_b, 1u);
          }
          ptx::mma_commit(&sm.x_free[b][seg]);
        }
        if ((t % kChainTiles) == kChainTiles - 1 || t == T - 1) {
          ptx::mma_commit(&sm.g_full);
          ++chain;
        }
      }
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  pair::cluster_sync_all();        // the peer may still be writing into this CTA's exchange buffer / barriers
  if (warp == kMmaWarp) ptx::tmem_dealloc(tmem, kTmemCols);
}

// G[d, s] (float64) = sum over pairs of partial_g[pair][d / 128][s][d % 128];
// loglik[s] = sum over pairs and over the four worker warps of the draw's TMEM quadrant, two lanes each
__global__ void __launch_bounds__(256)
logistic_fused3_finalize_kernel(const float* __restrict__ partial_g, const double* __restrict__ partial_ll,
                                int n_pairs, int d, double* __restrict__ g_out, double* __restrict__ ll_out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < d * kS) {
    const int col = idx / d, row = idx % d;
    const int64_t per_pair = static_cast<int64_t>(d) * kS;
    const int64_t off = (static_cast<int64_t>(row / 128) * kS + col) * 128 + row % 128;
    double acc = 0.0;
    for (int c = 0; c < n_pairs; ++c) acc += static_cast<double>(partial_g[c * per_pair + off]);
    g_out[static_cast<int64_t>(row) * kS + col] = acc;
  } else if (idx < d * kS + kS) {
    const int s = idx - d * kS;
    double acc = 0.0;
    for (int c = 0; c < n_pairs; ++c)
      for (int w = (s >> 4); w < kWorkerWarps; w += 4)
        acc += partial_ll[(static_cast<int64_t>(c) * kWorkerWarps + w) * 32 + (s & 15)] +
               partial_ll[(static_cast<int64_t>(c) * kWorkerWarps + w) * 32 + (s & 15) + 16];
    ll_out[s] = acc;
  }
}

// W[s, d] float32 -> per 64-feature chunk the UMMA image of [W1; W2] (128 rows x 128 bytes,
// K-major SWIZZLE_128B): row = (draw / 16) * 32 + part * 16 + draw % 16   (same image as logistic_fused2_sm100.cu)
__global__ void prep_w_fused3_kernel(const float* __restrict__ w, int d, uint8_t* __restrict__ out) {
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
  const int r1 = (s >> 4) * 32 + (s & 15), r2 = r1 + 16;
  *reinterpret_cast<uint4*>(chunk + (r1 >> 3) * 1024 + (r1 & 7) * 128 + ((j ^ (r1 & 7)) << 4)) = make_uint4(b1[0], b1[1], b1[2], b1[3]);
  *reinterpret_cast<uint4*>(chunk + (r2 >> 3) * 1024 + (r2 & 7) * 128 + ((j ^ (r2 & 7)) << 4)) = make_uint4(b2[0], b2[1], b2[2], b2[3]);
}

int fused3_pairs(int64_t n) {
  int sms = device_sm_count();
  if (sms <= 0) sms = 148;
  const int64_t tiles = (n + kTileRows - 1) / kTileRows;
  return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(sms / 2, tiles)));
}

template <int kSegC>
int launch_fused3_instance(const Fused3Params& p, int pairs, cudaStream_t stream) {
  const int smem_bytes = static_cast<int>(sizeof(Smem<kSegC>));
  static SmemOptIn smem_opt_in;
  BB_CUDA_OK(smem_opt_in.ensure(logistic_fused3_kernel<kSegC>, smem_bytes));
  logistic_fused3_kernel<kSegC><<<2 * pairs, kThreads, smem_bytes, stream>>>(p);
  BB_CHECK_LAUNCH("logistic_fused3_kernel");
  return BB_OK;
}

}  // namespace

bool logistic_fused3_supported(int64_t n, int d, int s, const void* x) {
  return n > 0 && s == kS && (d == 256 || d == 512) && reinterpret_cast<uintptr_t>(x) % 16 == 0;
}

int64_t logistic_fused3_workspace(int64_t n, int d, int s) {
  const int pairs = fused3_pairs(n);
  return align_up(static_cast<int64_t>(d / 64) * kWChunkBytes, 256) + static_cast<int64_t>(pairs) * d * s * 4 +
         static_cast<int64_t>(pairs) * kWorkerWarps * 32 * 8 + 1024;
}

int launch_logistic_fused3(const float* x, const float* y, const float* w, int64_t n, int d, int s, double* loglik,
                           double* g, void* workspace, int64_t workspace_bytes, cudaStream_t stream) {
  if (!logistic_fused3_supported(n, d, s, x)) {
    set_error("logistic_fused3: unsupported shape n=%lld d=%d s=%d", static_cast<long long>(n), d, s);
    return BB_ERR_UNSUPPORTED;
  }
  if (workspace == nullptr || workspace_bytes < logistic_fused3_workspace(n, d, s)) {
    set_error("logistic_fused3: workspace %lld < %lld bytes", static_cast<long long>(workspace_bytes),
              static_cast<long long>(logistic_fused3_workspace(n, d, s)));
    return BB_ERR_WORKSPACE;
  }
  const int pairs = fused3_pairs(n);
  uint8_t* ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~static_cast<uintptr_t>(255));
  uint8_t* wprep = ws;
  ws += align_up(static_cast<int64_t>(d / 64) * kWChunkBytes, 256);
  float* partial_g = reinterpret_cast<float*>(ws);
  ws += static_cast<int64_t>(pairs) * d * s * 4;
  double* partial_ll = reinterpret_cast<double*>(ws);
  const int prep_threads = s * d / 8;
  prep_w_fused3_kernel<<<(prep_threads + 255) / 256, 256, 0, stream>>>(w, d, wprep);
  BB_CHECK_LAUNCH("prep_w_fused3_kernel");
  Fused3Params p;
  p.x = x; p.y = y; p.wprep = wprep; p.partial_g = partial_g; p.partial_ll = partial_ll; p.n = n;
  static const int ablate = getenv("BB_FUSED3_ABLATE") ? atoi(getenv("BB_FUSED3_ABLATE")) : 0;
  p.ablate = ablate;
  if (d == 256) BB_TRY(launch_fused3_instance<1>(p, pairs, stream));
  else BB_TRY(launch_fused3_instance<2>(p, pairs, stream));
  logistic_fused3_finalize_kernel<<<(d * s + s + 255) / 256, 256, 0, stream>>>(partial_g, partial_ll, pairs, d, g, loglik);
  BB_CHECK_LAUNCH("logistic_fused3_finalize_kernel");
  return BB_OK;
}

}  // namespace bb
