// Reparameterised-gradient pass of Bayesian logistic regression in ONE kernel that reads X from
// HBM once (D = 512, S = 64 -- BASELINE cfg5 -- and, generally, D % 128 == 0, D <= 512, S == 64):
//     Z = X W^T;   loglik[s] = sum_n y_n z_ns - log(1 + exp z_ns);   G[d,s] = sum_n x_nd (y_n - sigmoid(z_ns))
// Same value as rowproj_sm100.cu (logistic epilogue) followed by colproj_sm100.cu -- the plans of
// sum(ycol * Z - log(1 + exp(Z)), 0) and dot(X.T, ycol - (1 + exp(-1 * Z)) ** -1), Z = dot(X, Wm.T)
// (bayesic/algebra.py:1435-1448 vocabulary; README.md:47-51) -- but the N x S residual never leaves
// the SM and the second contraction re-reads the 128-row X tile from L2 right after the first has
// pulled it in from HBM.
//
// Per persistent CTA, per 128-row tile t (software-pipelined A(t+1) | epilogue(t) | B(t)):
//   A(t)  Z tile = X_tile W^T: 8 K-major 64-feature chunks, W resident in shared memory (128 KB),
//         M128 x N64 accumulator in TMEM (double-buffered)                        [= rowproj]
//   E(t)  four epilogue warps: Z -> loglik terms (column sums), residual -> error-compensated
//         bf16 pair written to shared memory in the MN-major SWIZZLE_128B layout (rows = K)
//   B(t)  G += X_tile^T resid: the tile again as 8 MN-major 16-row stages (L2 hits), B operand =
//         the residual tile in shared memory, 4 x (M128 x N64) accumulators in TMEM, drained every
//         16 tiles (2048 rows) to the CTA's fp32 partial block                      [= colproj]
// The 16 converter warps serve both phases through one 2-stage ring of 32 KB stages (BF16x3
// split in registers, see gram_sm100.cu).  Shared memory: 128 (W) + 64 (ring) + 32 (residual) KB.
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

constexpr int kTileRows = 128;
constexpr int kS = 64;                          // parameter draws (MMA N in both phases)
constexpr int kStageBytes = 32 * 1024;          // A: 128 rows x 64 feats x (b1, b2); B: 16 rows x 512 feats x (b1, b2)
constexpr int kPartBytes = kStageBytes / 2;
constexpr int kStages = 2;
constexpr int kWBytes = 128 * 1024;
constexpr int kResidPart = kTileRows * 128;     // 16 KB: 128 rows (K) x 64 draws, one bf16 part
constexpr int kFlushTiles = 16;                 // G accumulators drained every 16 tiles = 2048 rows
constexpr int kConvWarps = 16;
constexpr int kConvGroups = 2;
constexpr int kEpiWarps = 4;
constexpr int kMmaWarp = kConvWarps + kEpiWarps;
constexpr int kThreads = (kMmaWarp + 1) * 32;
constexpr int kTmemCols = 512;
constexpr int kTmemZ = 0;                       // 2 x 64 columns
constexpr int kTmemG = 128;                     // up to 4 x 64 columns

struct __align__(1024) SmemLayout {
  uint8_t stage[kStages][kStageBytes];
  uint8_t w[kWBytes];
  uint8_t resid[2 * kResidPart];
  uint64_t full[kStages], empty[kStages];
  uint64_t z_full[2], z_empty[2];
  uint64_t resid_full, resid_free;
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
// loads with an L2 eviction policy: the A phase reads a tile with evict_last (it will be read
// again by the B phase ~1.5 tiles later; across 148 CTAs that is ~76 MB of other traffic in
// between -- measured: without the hint the second read misses L2, DRAM traffic 15.5 GB instead
// of 8.6 GB), the B phase with evict_first (dead after this use)
__device__ __forceinline__ uint64_t make_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t make_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ float4 ldg_f4_hint(const float* p, uint64_t policy) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p), "l"(policy));
  return v;
}
__device__ __forceinline__ void sts_u2(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void sts_u4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
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
__device__ __forceinline__ void warp_transpose_reduce(float (&e)[32], int lane) {
#pragma unroll
  for (int w = 16; w >= 1; w >>= 1) {
    const bool upper = (lane & w) != 0;
#pragma unroll
    for (int j = 0; j < w; ++j) {
      const float send = upper ? e[j] : e[j + w];
      const float keep = upper ? e[j + w] : e[j];
      e[j] = keep + __shfl_xor_sync(0xffffffffu, send, w);
    }
  }
}

struct FusedParams {
  int prefetch_iters;               // L2 prefetch distance in converter iterations (0 = off)
  int l2_hints;                     // eviction-priority hints on the two reads of a tile
  const float* x;
  const float* y;
  const __nv_bfloat16* wsplit;      // [2][64][d]
  float* partial_g;                 // [cta][d / 128][64 cols][128 rows] fp32
  double* partial_ll;               // [cta][kEpiWarps][64]
  int64_t n;
  int d;
};

// The stage sequence both the converters and the MMA issuer walk: for step = 0 .. T:
//   A(step) if step < T  (chunks_a stages), then B(step - 1) if step >= 1 (8 stages).
template <int kNSeg>       // d / 128
__global__ void __launch_bounds__(kThreads, 1) logistic_fused_kernel(const FusedParams p) {
  constexpr int kD = kNSeg * 128;
  constexpr int kChunksA = kD / 64;           // K-major 64-feature chunks per tile
  constexpr int kStagesB = kTileRows / 16;    // 16-row MN-major stages per tile
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  SmemLayout& sm = *reinterpret_cast<SmemLayout*>(smem_raw);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t n_tiles = (p.n + kTileRows - 1) / kTileRows;
  // contiguous range of tiles per CTA (the B phase of a tile follows its A phase closely in time,
  // so the second read of the tile hits L2)
  const int64_t tile_begin = n_tiles * blockIdx.x / gridDim.x;
  const int64_t tile_end = n_tiles * (blockIdx.x + 1) / gridDim.x;
  const int T = static_cast<int>(tile_end - tile_begin);

  if (warp == kMmaWarp) {
    if (lane == 0) {
      for (int s = 0; s < kStages; ++s) {
        ptx::mbar_init(&sm.full[s], kConvWarps / kConvGroups);
        ptx::mbar_init(&sm.empty[s], 1);
      }
      for (int b = 0; b < 2; ++b) {
        ptx::mbar_init(&sm.z_full[b], 1);
        ptx::mbar_init(&sm.z_empty[b], kEpiWarps);
      }
      ptx::mbar_init(&sm.resid_full, kEpiWarps);
      ptx::mbar_init(&sm.resid_free, 1);
      ptx::mbar_init(&sm.g_full, 1);
      ptx::mbar_init(&sm.g_empty, kEpiWarps);
      ptx::fence_mbar_init();
    }
    __syncwarp();
    ptx::tmem_alloc(&sm.tmem_base, kTmemCols);
  }
  {  // resident W: wsplit[part][row][d] -> [part][kc][row >> 3][row & 7][16-byte chunk ^ (row & 7)]
    constexpr int chunks_per_row = kD / 8;
    constexpr int total = 2 * kS * chunks_per_row;
    const uint4* src = reinterpret_cast<const uint4*>(p.wsplit);
    for (int i = threadIdx.x; i < total; i += kThreads) {
      const int part = i / (kS * chunks_per_row);
      const int rem = i - part * kS * chunks_per_row;
      const int row = rem / chunks_per_row;
      const int c = rem - row * chunks_per_row;
      const int kc = c >> 3, j = c & 7;
      const uint32_t off = static_cast<uint32_t>(part) * kS * kD * 2 + static_cast<uint32_t>(kc) * kS * 128 +
                           (row >> 3) * 1024 + (row & 7) * 128 + ((j ^ (row & 7)) << 4);
      *reinterpret_cast<uint4*>(sm.w + off) = __ldg(src + i);
    }
  }
  fence_proxy_async_smem();
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem = sm.tmem_base;

  if (warp < kConvWarps) {
    // ---------------- converter warps (both phases) ----------------
    const int group = warp & (kConvGroups - 1);
    const int wi = warp / kConvGroups;
    const int sub = lane >> 4, c4 = lane & 15;
    // phase A (K-major chunk, 128 rows x 64 feats): load i covers row 16 wi + 2 i + sub, float4 c4
    // phase B (MN-major stage, 16 rows x kD feats): load (j, seg) covers row 2 wi + j, feats 128 seg + 4 lane
    uint32_t soff_a[8], soff_b[2];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = wi * 16 + 2 * i + sub;
      soff_a[i] = (r >> 3) * 1024 + (r & 7) * 128 + (((c4 >> 1) ^ (r & 7)) << 4) + (c4 & 1) * 8;
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int k = 2 * wi + j;
      soff_b[j] = (lane >> 4) * 2048 + (k >> 3) * 1024 + (k & 7) * 128 + ((((lane & 15) >> 1) ^ (k & 7)) << 4) + (lane & 1) * 8;
    }
    const uint32_t stage0 = ptx::smem_u32(sm.stage[0]);
    // stage index -> (phase, tile, sub-stage); stages of step: A(step) [kChunksA] then B(step-1) [kStagesB]
    const int per_mid = kChunksA + kStagesB;
    const int64_t total_stages = static_cast<int64_t>(T) * per_mid;
    auto decode = [&](int64_t it, bool* is_b, int* tile, int* sub_stage) {
      // it < kChunksA: A(0).  Then blocks of per_mid: A(step) B(step-1) for step = 1..T-1; tail: B(T-1).
      if (it < kChunksA) { *is_b = false; *tile = 0; *sub_stage = static_cast<int>(it); return; }
      const int64_t r = it - kChunksA;
      const int step = static_cast<int>(r / per_mid) + 1;
      const int w = static_cast<int>(r % per_mid);
      if (step < T) {
        if (w < kChunksA) { *is_b = false; *tile = step; *sub_stage = w; }
        else { *is_b = true; *tile = step - 1; *sub_stage = w - kChunksA; }
      } else {            // step == T: only B(T-1) remains
        *is_b = true; *tile = T - 1; *sub_stage = w;
      }
    };
    const uint64_t pol_keep = make_policy_evict_last(), pol_drop = make_policy_evict_first();
    const bool hints = p.l2_hints != 0;
    float4 rx[8];
    bool cur_b = false;
    auto load = [&](int64_t it) {
      int tile, ss;
      decode(it, &cur_b, &tile, &ss);
      const int64_t tile_row0 = (tile_begin + tile) * kTileRows;
      if (!cur_b) {
        const int64_t row0 = tile_row0 + wi * 16 + sub;
        const float* base = p.x + row0 * kD + ss * 64 + c4 * 4;
        if (row0 + 15 < p.n) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            rx[i] = hints ? ldg_f4_hint(base + static_cast<int64_t>(2 * i) * kD, pol_keep)
                          : ldg_f4(base + static_cast<int64_t>(2 * i) * kD);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            rx[i] = (row0 + 2 * i < p.n) ? ldg_f4(base + static_cast<int64_t>(2 * i) * kD) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      } else {
        const int64_t row0 = tile_row0 + ss * 16 + 2 * wi;
        const float* base = p.x + row0 * kD + lane * 4;
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int seg = 0; seg < 4; ++seg) {
            if (seg < kNSeg)
              rx[j * 4 + seg] = (row0 + j < p.n) ? (hints ? ldg_f4_hint(base + j * kD + seg * 128, pol_drop)
                                                          : ldg_f4(base + j * kD + seg * 128))
                                                 : make_float4(0.f, 0.f, 0.f, 0.f);
          }
      }
    };
    // L2 prefetch hint for an A-phase chunk a few iterations ahead (B-phase stages re-read the
    // tile and hit L2 anyway): 16 rows x 2 lines per warp = one 128-byte line per lane
    auto prefetch = [&](int64_t it) {
      if (it >= total_stages) return;
      bool pb;
      int tile, ss;
      decode(it, &pb, &tile, &ss);
      if (pb) return;
      const int64_t row = (tile_begin + tile) * kTileRows + wi * 16 + (lane >> 1);
      if (row < p.n) {
        if (hints) asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(p.x + row * kD + ss * 64 + (lane & 1) * 32));
        else asm volatile("prefetch.global.L2 [%0];" ::"l"(p.x + row * kD + ss * 64 + (lane & 1) * 32));
      }
    };
    if (group < total_stages) load(group);
    for (int64_t it = group; it < total_stages; it += kConvGroups) {
      const int s = static_cast<int>(it % kStages);
      const bool this_b = cur_b;
      if (p.prefetch_iters > 0) prefetch(it + static_cast<int64_t>(p.prefetch_iters + 1) * kConvGroups);
      ptx::mbar_wait(&sm.empty[s], (static_cast<uint32_t>(it / kStages) & 1) ^ 1);
      const uint32_t stage_addr = stage0 + s * kStageBytes;
      if (!this_b) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          uint32_t b1[2], b2[2];
          split_bf16(rx[i], b1, b2);
          sts_u2(stage_addr + soff_a[i], b1[0], b1[1]);
          sts_u2(stage_addr + kPartBytes + soff_a[i], b2[0], b2[1]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int seg = 0; seg < 4; ++seg) {
            if (seg < kNSeg) {
              uint32_t b1[2], b2[2];
              split_bf16(rx[j * 4 + seg], b1, b2);
              const uint32_t addr = stage_addr + seg * 4096 + soff_b[j];
              sts_u2(addr, b1[0], b1[1]);
              sts_u2(addr + kPartBytes, b2[0], b2[1]);
            }
          }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&sm.full[s]);
      if (it + kConvGroups < total_stages) load(it + kConvGroups);
    }
  } else if (warp < kMmaWarp) {
    // ---------------- epilogue warps: lane = data row of the tile ----------------
    const int qd = warp & 3;
    double colsum[2] = {0.0, 0.0};
    const int r = qd * 32 + lane;                 // row within the tile = K index of the residual tile
    const uint32_t resid_base = ptx::smem_u32(sm.resid) + (r >> 3) * 1024 + (r & 7) * 128;
    float* my_partial = p.partial_g + static_cast<int64_t>(blockIdx.x) * kNSeg * kS * 128 + qd * 32 + lane;
    int flushes = 0;
    if (T == 0) {
      for (int c = 0; c < kNSeg * kS; ++c) my_partial[c * 128] = 0.f;
    }
    for (int t = 0; t < T; ++t) {
      const int zb = t & 1;
      ptx::mbar_wait(&sm.z_full[zb], static_cast<uint32_t>(t >> 1) & 1);
      ptx::tc_fence_after_sync();
      const int64_t row = (tile_begin + t) * kTileRows + r;
      const bool valid = row < p.n;
      const float yv = valid ? __ldg(p.y + row) : 0.f;
      const uint32_t t_addr = tmem + (static_cast<uint32_t>(qd * 32) << 16) + kTmemZ + zb * kS;
      // the residual tile is free once B(t-1) has read it (checked before the Z loop so that each
      // half can be stored as soon as it is formed: keeps the live registers low)
      ptx::mbar_wait(&sm.resid_free, (static_cast<uint32_t>(t) & 1) ^ 1);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(t_addr + half * 32, v);
        ptx::tmem_wait_ld();
        float e[32];
        uint32_t rb1[16], rb2[16];                // packed bf16x2 residual parts of 32 draws
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          float res[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const float z = __uint_as_float(v[j + u]);
            const float ez = __expf(-fabsf(z));
            const float softplus = fmaxf(z, 0.f) + __logf(1.f + ez);
            const float sig = __fdividef(z >= 0.f ? 1.f : ez, 1.f + ez);
            e[j + u] = valid ? fmaf(yv, z, -softplus) : 0.f;
            res[u] = valid ? yv - sig : 0.f;      // rows past n contribute nothing to G
          }
          __nv_bfloat162 h = __floats2bfloat162_rn(res[0], res[1]);
          const uint32_t hb = *reinterpret_cast<uint32_t*>(&h);
          __nv_bfloat162 l = __floats2bfloat162_rn(res[0] - __uint_as_float(hb << 16),
                                                   res[1] - __uint_as_float(hb & 0xFFFF0000u));
          rb1[j >> 1] = hb;
          rb2[j >> 1] = *reinterpret_cast<uint32_t*>(&l);
        }
        // residual tile (MN-major, K = row): chunks of 8 draws; this half holds chunks 4 half .. +3
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint32_t addr = resid_base + (((half * 4 + c) ^ (r & 7)) << 4);
          sts_u4(addr, rb1[4 * c], rb1[4 * c + 1], rb1[4 * c + 2], rb1[4 * c + 3]);
          sts_u4(addr + kResidPart, rb2[4 * c], rb2[4 * c + 1], rb2[4 * c + 2], rb2[4 * c + 3]);
        }
        warp_transpose_reduce(e, lane);
        colsum[half] += static_cast<double>(e[0]);
      }
      ptx::tc_fence_before_sync();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive(&sm.z_empty[zb]);
        ptx::mbar_arrive(&sm.resid_full);
      }
      // drain G every kFlushTiles tiles (and after the last tile)
      if ((t % kFlushTiles) == kFlushTiles - 1 || t == T - 1) {
        ptx::mbar_wait_sleep(&sm.g_full, static_cast<uint32_t>(flushes) & 1);
        ptx::tc_fence_after_sync();
        const uint32_t g_addr = tmem + (static_cast<uint32_t>(qd * 32) << 16) + kTmemG;
#pragma unroll 1
        for (int cc = 0; cc < kNSeg * kS / 32; ++cc) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(g_addr + cc * 32, v);
          float* dst = my_partial + cc * 32 * 128;
          float old[32];
          if (flushes != 0) {
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
        if (lane == 0) ptx::mbar_arrive(&sm.g_empty);
        ++flushes;
      }
    }
    double* out = p.partial_ll + (static_cast<int64_t>(blockIdx.x) * kEpiWarps + qd) * kS;
    out[lane] = colsum[0];
    out[32 + lane] = colsum[1];
  } else {
    // ---------------- MMA issuer ----------------
    if (ptx::elect_one()) {
      const uint32_t idesc_a = ptx::make_idesc(128, kS, /*bf16*/ 1, /*A K-major*/ 0, /*B K-major*/ 0);
      const uint32_t idesc_b = ptx::make_idesc(128, kS, /*bf16*/ 1, /*A MN-major*/ 1, /*B MN-major*/ 1);
      const uint32_t w1 = ptx::smem_u32(sm.w), w2 = w1 + kS * kD * 2;
      const uint32_t resid1 = ptx::smem_u32(sm.resid), resid2 = resid1 + kResidPart;
      int64_t it = 0;
      int flush_idx = 0;
      for (int step = 0; step <= T; ++step) {
        if (step < T) {               // ---- A(step): Z tile ----
          const int zb = step & 1;
          ptx::mbar_wait(&sm.z_empty[zb], (static_cast<uint32_t>(step >> 1) & 1) ^ 1);
          const uint32_t d_tmem = tmem + kTmemZ + zb * kS;
          for (int kc = 0; kc < kChunksA; ++kc, ++it) {
            const int s = static_cast<int>(it % kStages);
            ptx::mbar_wait(&sm.full[s], static_cast<uint32_t>(it / kStages) & 1);
            ptx::tc_fence_after_sync();
            const uint32_t a_base = ptx::smem_u32(sm.stage[s]);
            const uint32_t w_off = static_cast<uint32_t>(kc) * kS * 128;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint64_t a1 = ptx::make_smem_desc(a_base + ks * 32, 16, 1024, ptx::kLayoutSwizzle128B);
              const uint64_t a2 = ptx::make_smem_desc(a_base + kPartBytes + ks * 32, 16, 1024, ptx::kLayoutSwizzle128B);
              const uint64_t b1 = ptx::make_smem_desc(w1 + w_off + ks * 32, 16, 1024, ptx::kLayoutSwizzle128B);
              const uint64_t b2 = ptx::make_smem_desc(w2 + w_off + ks * 32, 16, 1024, ptx::kLayoutSwizzle128B);
              mma_bf16_ss(d_tmem, a1, b1, idesc_a, (kc == 0 && ks == 0) ? 0u : 1u);
              mma_bf16_ss(d_tmem, a1, b2, idesc_a, 1u);
              mma_bf16_ss(d_tmem, a2, b1, idesc_a, 1u);
            }
            ptx::mma_commit(&sm.empty[s]);
          }
          ptx::mma_commit(&sm.z_full[zb]);
        }
        if (step >= 1) {              // ---- B(step - 1): G += X_tile^T resid ----
          const int t = step - 1;
          const bool first_in_interval = (t % kFlushTiles) == 0;
          if (first_in_interval && flush_idx > 0) ptx::mbar_wait(&sm.g_empty, static_cast<uint32_t>(flush_idx - 1) & 1);
          ptx::mbar_wait(&sm.resid_full, static_cast<uint32_t>(t) & 1);
          for (int sb = 0; sb < kStagesB; ++sb, ++it) {
            const int s = static_cast<int>(it % kStages);
            ptx::mbar_wait(&sm.full[s], static_cast<uint32_t>(it / kStages) & 1);
            ptx::tc_fence_after_sync();
            const uint32_t base = ptx::smem_u32(sm.stage[s]);
            // residual rows 16 sb .. +15 = k groups 2 sb, 2 sb + 1 (1 KB each): one 64-wide MN block
            const uint64_t r1 = ptx::make_smem_desc(resid1 + sb * 2048, 2048, 1024, ptx::kLayoutSwizzle128B);
            const uint64_t r2 = ptx::make_smem_desc(resid2 + sb * 2048, 2048, 1024, ptx::kLayoutSwizzle128B);
#pragma unroll
            for (int blk = 0; blk < kNSeg; ++blk) {
              const uint64_t x1 = ptx::make_smem_desc(base + blk * 4096, 2048, 1024, ptx::kLayoutSwizzle128B);
              const uint64_t x2 = ptx::make_smem_desc(base + kPartBytes + blk * 4096, 2048, 1024, ptx::kLayoutSwizzle128B);
              const uint32_t d_tmem = tmem + kTmemG + blk * kS;
              mma_bf16_ss(d_tmem, x1, r1, idesc_b, (first_in_interval && sb == 0) ? 0u : 1u);
              mma_bf16_ss(d_tmem, x1, r2, idesc_b, 1u);
              mma_bf16_ss(d_tmem, x2, r1, idesc_b, 1u);
            }
            ptx::mma_commit(&sm.empty[s]);
          }
          ptx::mma_commit(&sm.resid_free);
          if ((t % kFlushTiles) == kFlushTiles - 1 || t == T - 1) {
            ptx::mma_commit(&sm.g_full);
            ++flush_idx;
          }
        }
      }
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == kMmaWarp) ptx::tmem_dealloc(tmem, kTmemCols);
}

// W[q, d] float32 -> wsplit[part][q][d] bf16
__global__ void split_w_fused_kernel(const float* __restrict__ w, int64_t count, __nv_bfloat16* __restrict__ out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const float x = w[i];
  const __nv_bfloat16 b1 = __float2bfloat16_rn(x);
  out[i] = b1;
  out[count + i] = __float2bfloat16_rn(x - __bfloat162float(b1));
}

// G[d, s] (float64) = sum over CTAs of partial_g[cta][d / 128][s][d % 128]; loglik[s] likewise
__global__ void __launch_bounds__(256)
logistic_fused_finalize_kernel(const float* __restrict__ partial_g, const double* __restrict__ partial_ll,
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
    for (int c = 0; c < n_ctas * kEpiWarps; ++c) acc += partial_ll[static_cast<int64_t>(c) * kS + s];
    ll_out[s] = acc;
  }
}

int fused_grid(int64_t n) {
  int sms = device_sm_count();
  if (sms <= 0) sms = 148;
  const int64_t tiles = (n + kTileRows - 1) / kTileRows;
  return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(sms, tiles)));
}

template <int kNSeg>
int launch_fused_instance(const FusedParams& p, int grid, cudaStream_t stream) {
  const int smem_bytes = static_cast<int>(sizeof(SmemLayout));
  static SmemOptIn smem_opt_in_0;
  BB_CUDA_OK(smem_opt_in_0.ensure(logistic_fused_kernel<kNSeg>, smem_bytes));
  logistic_fused_kernel<kNSeg><<<grid, kThreads, smem_bytes, stream>>>(p);
  BB_CHECK_LAUNCH("logistic_fused_kernel");
  return BB_OK;
}

}  // namespace

bool logistic_fused_supported(int64_t n, int d, int s, const void* x) {
  return n > 0 && s == kS && d >= 128 && d % 128 == 0 && d <= 512 && reinterpret_cast<uintptr_t>(x) % 16 == 0;
}

int64_t logistic_fused_workspace(int64_t n, int d, int s) {
  const int grid = fused_grid(n);
  return align_up(static_cast<int64_t>(2) * s * d * 2, 256) + static_cast<int64_t>(grid) * d * s * 4 +
         static_cast<int64_t>(grid) * kEpiWarps * s * 8 + 1024;
}

int launch_logistic_fused(const float* x, const float* y, const float* w, int64_t n, int d, int s, double* loglik,
                          double* g, void* workspace, int64_t workspace_bytes, cudaStream_t stream) {
  if (!logistic_fused_supported(n, d, s, x)) {
    set_error("logistic_fused: unsupported shape n=%lld d=%d s=%d", static_cast<long long>(n), d, s);
    return BB_ERR_UNSUPPORTED;
  }
  if (workspace == nullptr || workspace_bytes < logistic_fused_workspace(n, d, s)) {
    set_error("logistic_fused: workspace %lld < %lld bytes", static_cast<long long>(workspace_bytes),
              static_cast<long long>(logistic_fused_workspace(n, d, s)));
    return BB_ERR_WORKSPACE;
  }
  const int grid = fused_grid(n);
  uint8_t* ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~static_cast<uintptr_t>(255));
  __nv_bfloat16* wsplit = reinterpret_cast<__nv_bfloat16*>(ws);
  ws += align_up(static_cast<int64_t>(2) * s * d * 2, 256);
  float* partial_g = reinterpret_cast<float*>(ws);
  ws += static_cast<int64_t>(grid) * d * s * 4;
  double* partial_ll = reinterpret_cast<double*>(ws);
  const int64_t count = static_cast<int64_t>(s) * d;
  split_w_fused_kernel<<<static_cast<int>((count + 255) / 256), 256, 0, stream>>>(w, count, wsplit);
  BB_CHECK_LAUNCH("split_w_fused_kernel");
  static const int prefetch_iters = getenv("BB_FUSED_PREFETCH") ? atoi(getenv("BB_FUSED_PREFETCH")) : 2;
  FusedParams p;
  static const int l2_hints = getenv("BB_FUSED_L2HINTS") ? atoi(getenv("BB_FUSED_L2HINTS")) : 1;
  p.prefetch_iters = prefetch_iters;
  p.l2_hints = l2_hints;
  p.x = x; p.y = y; p.wsplit = wsplit; p.partial_g = partial_g; p.partial_ll = partial_ll; p.n = n; p.d = d;
  switch (d / 128) {
    case 1: BB_TRY(launch_fused_instance<1>(p, grid, stream)); break;
    case 2: BB_TRY(launch_fused_instance<2>(p, grid, stream)); break;
    case 3: BB_TRY(launch_fused_instance<3>(p, grid, stream)); break;
    default: BB_TRY(launch_fused_instance<4>(p, grid, stream)); break;
  }
  logistic_fused_finalize_kernel<<<(d * s + s + 255) / 256, 256, 0, stream>>>(partial_g, partial_ll, grid, d, g, loglik);
  BB_CHECK_LAUNCH("logistic_fused_finalize_kernel");
  return BB_OK;
}

}  // namespace bb
