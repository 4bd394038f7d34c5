// Per-row projection on tcgen05 (D % 64 == 0, Q % 16 == 0, Q <= 256, Q * D <= 32768):
//     Z[n, q] = sum_d X[n, d] W[q, d]
// with the consumer fused into the TMEM epilogue so Z never travels to HBM when it is only an
// intermediate.  What it replaces: the plan  _tensordot(X, _dimshuffle(W,1,0), [1],[0])
// (bayesic/algebra.py:527-551 -> 1347-1351; "dot(X, M.T)" in SURVEY.md section 3.2) -- one
// Theano/BLAS sgemm -- followed, in the reparameterised-gradient pass of Bayesian logistic
// regression (BASELINE cfg5; README.md:47-51), by the elementwise chain built from the
// reference's vocabulary (algebra.py:1435-1448):
//     loglik[s] = sum_n ( y_n z_ns - log(1 + exp(z_ns)) ),     resid[n, s] = y_n - (1 + exp(-z_ns))^-1
//
// Design (HBM-bound: 4 D bytes/row in, 4 Q bytes/row out, 2 D Q flop/row):
//   * persistent CTAs, one per SM, over 128-row tiles; the data rows are the MMA M axis, the
//     feature axis D is K, so X's row-major layout is K-major: 16 converter warps load X with
//     coalesced 128-bit loads straight into registers, split it into error-compensated BF16
//     (x = b1 + b2, see gram_sm100.cu) and store the two tiles in the UMMA K-major SWIZZLE_128B
//     layout, 64 features (one 128-byte swizzle row) per pipeline stage;
//   * W (tiny, reused by every tile) is split once by a pre-kernel and kept RESIDENT in shared
//     memory for the whole kernel (2 x Q x D bf16 <= 128 KB), also K-major SWIZZLE_128B;
//   * three kind::f16 MMAs per K = 16 step (b1 w1 + b1 w2 + b2 w1), M = 128, N = Q, FP32
//     accumulation in TMEM, double-buffered across tiles (2 x Q columns); a tile's chain is only
//     3 D / 16 accumulate steps long, so the truncating accumulate stays at the 1e-6 level;
//   * four epilogue warps read a finished tile from TMEM (lane = data row) and either store Z
//     (plan executor) or apply the logistic epilogue: per-row terms in fp32 with expf / log1pf,
//     resid stored row-major, the column sums of the log-likelihood terms reduced across the 32
//     rows of a warp with a 31-shuffle transpose-reduce and accumulated in float64.
//
// Measured on B200 (round 1, cfg5: n = 4 Mi, D = 512, Q = 64): plain projection 2.05 ms
// (4.7 TB/s of X read + Z write).
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
constexpr int kChunk = 64;                      // features per stage (one 128-byte bf16 swizzle row)
constexpr int kStages = 3;
constexpr int kPartBytes = kTileRows * kChunk * 2;       // one bf16 tile: 16 KB
constexpr int kStageBytes = 2 * kPartBytes;              // b1, b2
constexpr int kMaxWBytes = 128 * 1024;                   // resident W (w1 and w2)
constexpr int kConvWarps = 16;
constexpr int kConvGroups = 2;
constexpr int kEpiWarps = 4;
constexpr int kMmaWarp = kConvWarps + kEpiWarps;
constexpr int kThreads = (kMmaWarp + 1) * 32;            // 672

struct __align__(1024) SmemLayout {
  uint8_t stage[kStages][kStageBytes];
  uint8_t w[kMaxWBytes];                         // [part 2][D / 64][Q rows x 128 B]
  uint64_t full[kStages];
  uint64_t empty[kStages];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint32_t tmem_base;
};

enum { kEpiStore = 0, kEpiLogistic = 1 };

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

// W[q, d] float32 -> wsplit[part][q][d] bf16 (part 0 = b1, part 1 = b2)
__global__ void split_w_kernel(const float* __restrict__ w, int64_t count, __nv_bfloat16* __restrict__ out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const float x = w[i];
  const __nv_bfloat16 b1 = __float2bfloat16_rn(x);
  out[i] = b1;
  out[count + i] = __float2bfloat16_rn(x - __bfloat162float(b1));
}

// Sum over the 32 lanes of a warp of 32 per-lane values each: afterwards e[0] of lane l holds the
// total of element l (31 shuffles).
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

struct RowProjParams {
  int prefetch_iters;               // L2 prefetch distance in converter iterations (0 = off)
  const float* x;
  const __nv_bfloat16* wsplit;      // [2][q][d]
  const float* y;                   // logistic epilogue: labels [n]
  float* out;                       // Z or resid, [n, q] row-major
  double* partial_colsum;           // logistic epilogue: [grid][kEpiWarps][q]
  int64_t n;
  int d, q;
};

template <int kEpi>
__global__ void __launch_bounds__(kThreads, 1) rowproj_kernel(const RowProjParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  SmemLayout& sm = *reinterpret_cast<SmemLayout*>(smem_raw);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int d = p.d, q = p.q;
  const int kc_count = d / kChunk;
  const int64_t n_tiles = (p.n + kTileRows - 1) / kTileRows;
  const int64_t my_tiles = n_tiles > blockIdx.x ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int64_t n_iters = my_tiles * kc_count;
  const uint32_t tmem_cols = q <= 16 ? 32u : (q <= 32 ? 64u : (q <= 64 ? 128u : (q <= 128 ? 256u : 512u)));

  if (warp == kMmaWarp) {
    if (lane == 0) {
      for (int s = 0; s < kStages; ++s) {
        ptx::mbar_init(&sm.full[s], kConvWarps / kConvGroups);
        ptx::mbar_init(&sm.empty[s], 1);
      }
      for (int b = 0; b < 2; ++b) {
        ptx::mbar_init(&sm.acc_full[b], 1);
        ptx::mbar_init(&sm.acc_empty[b], kEpiWarps);
      }
      ptx::fence_mbar_init();
    }
    __syncwarp();
    ptx::tmem_alloc(&sm.tmem_base, tmem_cols);
  }
  // resident W: wsplit[part][row][d] -> [part][kc][row >> 3][row & 7][16-byte chunk ^ (row & 7)]
  {
    const int chunks_per_row = d / 8;                        // 16-byte chunks (8 bf16)
    const int total = 2 * q * chunks_per_row;
    const uint4* src = reinterpret_cast<const uint4*>(p.wsplit);
    for (int i = threadIdx.x; i < total; i += kThreads) {
      const int part = i / (q * chunks_per_row);
      const int rem = i - part * q * chunks_per_row;
      const int row = rem / chunks_per_row;
      const int c = rem - row * chunks_per_row;
      const int kc = c >> 3, j = c & 7;
      const uint32_t off = static_cast<uint32_t>(part) * q * d * 2 + static_cast<uint32_t>(kc) * q * 128 +
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
    // ---------------- converter warps ----------------
    // group g = warp & 1 takes every 2nd stage; warp wi of the group owns 16 rows of the stage:
    // load i covers rows 16 wi + 2 i + (lane >> 4), lane & 15 = float4 within the 64-feature chunk.
    const int group = warp & (kConvGroups - 1);
    const int wi = warp / kConvGroups;
    const int sub = lane >> 4, c4 = lane & 15;
    uint32_t soff[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = wi * 16 + 2 * i + sub;
      soff[i] = (r >> 3) * 1024 + (r & 7) * 128 + (((c4 >> 1) ^ (r & 7)) << 4) + (c4 & 1) * 8;
    }
    const uint32_t stage0 = ptx::smem_u32(sm.stage[0]);
    float4 rx[8];
    auto load = [&](int64_t it) {
      const int64_t tile = blockIdx.x + (it / kc_count) * gridDim.x;
      const int kc = static_cast<int>(it % kc_count);
      const int64_t row0 = tile * kTileRows + wi * 16 + sub;
      const float* base = p.x + row0 * d + kc * kChunk + c4 * 4;
      if (row0 + 15 < p.n) {
#pragma unroll
        for (int i = 0; i < 8; ++i) rx[i] = ldg_f4(base + static_cast<int64_t>(2 * i) * d);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          rx[i] = (row0 + 2 * i < p.n) ? ldg_f4(base + static_cast<int64_t>(2 * i) * d)
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    // L2 prefetch hint for the chunk this group converts `prefetch_iters` iterations from now: the
    // register-resident loads keep only ~64 KB per SM in flight, which covers the L2 latency but
    // not the HBM latency.  16 rows x 2 lines of 128 B per warp = one line per lane.
    auto prefetch = [&](int64_t it) {
      if (it >= n_iters) return;
      const int64_t tile = blockIdx.x + (it / kc_count) * gridDim.x;
      const int kc = static_cast<int>(it % kc_count);
      const int64_t row = tile * kTileRows + wi * 16 + (lane >> 1);
      if (row < p.n) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.x + row * d + kc * kChunk + (lane & 1) * 32));
    };
    if (group < n_iters) load(group);
    for (int64_t it = group; it < n_iters; it += kConvGroups) {
      const int s = static_cast<int>(it % kStages);
      if (p.prefetch_iters > 0) prefetch(it + static_cast<int64_t>(p.prefetch_iters + 1) * kConvGroups);
      ptx::mbar_wait(&sm.empty[s], (static_cast<uint32_t>(it / kStages) & 1) ^ 1);
      const uint32_t stage_addr = stage0 + s * kStageBytes;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        uint32_t b1[2], b2[2];
        split_bf16(rx[i], b1, b2);
        sts_u2(stage_addr + soff[i], b1[0], b1[1]);
        sts_u2(stage_addr + kPartBytes + soff[i], b2[0], b2[1]);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&sm.full[s]);
      if (it + kConvGroups < n_iters) load(it + kConvGroups);
    }
  } else if (warp < kMmaWarp) {
    // ---------------- epilogue warps: lane = data row ----------------
    const int qd = warp & 3;
    double colsum[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) colsum[c] = 0.0;
    for (int64_t t = 0; t < my_tiles; ++t) {
      const int buf = static_cast<int>(t & 1);
      ptx::mbar_wait(&sm.acc_full[buf], static_cast<uint32_t>(t >> 1) & 1);
      ptx::tc_fence_after_sync();
      const int64_t tile = blockIdx.x + t * gridDim.x;
      const int64_t row = tile * kTileRows + qd * 32 + lane;
      const bool valid = row < p.n;
      float yv = 0.f;
      if (kEpi == kEpiLogistic && valid) yv = __ldg(p.y + row);
      const uint32_t t_addr = tmem + (static_cast<uint32_t>(qd * 32) << 16) + buf * q;
      float* dst = p.out + row * q;
      for (int cc = 0; cc < q; cc += 32) {
        uint32_t v[32];
        if (q - cc >= 32) {
          tmem_ld_32x32b_x32(t_addr + cc, v);
        } else {                                  // q % 32 == 16 tail
          uint32_t h[16];
          ptx::tmem_ld_32x32b_x16(t_addr + cc, h);
#pragma unroll
          for (int j = 0; j < 16; ++j) { v[j] = h[j]; v[16 + j] = 0u; }
        }
        ptx::tmem_wait_ld();
        const int width = (q - cc >= 32) ? 32 : 16;
        if (kEpi == kEpiLogistic) {
          float e[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float z = __uint_as_float(v[j]);
            // fast intrinsics: ez in (0, 1], so 1 + ez in (1, 2] and the absolute error of
            // __logf / __fdividef stays ~1e-7, far inside the 1e-4 bar after the sums
            const float ez = __expf(-fabsf(z));
            const float softplus = fmaxf(z, 0.f) + __logf(1.f + ez);          // log(1 + exp(z))
            const float sig = __fdividef(z >= 0.f ? 1.f : ez, 1.f + ez);      // (1 + exp(-z))^-1
            e[j] = (valid && j < width) ? fmaf(yv, z, -softplus) : 0.f;
            v[j] = __float_as_uint(yv - sig);
          }
          warp_transpose_reduce(e, lane);
          colsum[cc >> 5] += static_cast<double>(e[0]);
        }
        if (valid) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (4 * j < width)
              *reinterpret_cast<uint4*>(dst + cc + 4 * j) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
      }
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&sm.acc_empty[buf]);
    }
    if (kEpi == kEpiLogistic) {
      double* out = p.partial_colsum + (static_cast<int64_t>(blockIdx.x) * kEpiWarps + qd) * q;
      for (int cc = 0; cc < q; cc += 32)
        if (cc + lane < q) out[cc + lane] = colsum[cc >> 5];
    }
  } else {
    // ---------------- MMA issuer ----------------
    if (ptx::elect_one()) {
      const uint32_t idesc = ptx::make_idesc(128, static_cast<uint32_t>(q), /*bf16*/ 1, /*A K-major*/ 0, /*B K-major*/ 0);
      const uint32_t w1 = ptx::smem_u32(sm.w), w2 = w1 + static_cast<uint32_t>(q) * d * 2;
      int64_t it = 0;
      for (int64_t t = 0; t < my_tiles; ++t) {
        const int buf = static_cast<int>(t & 1);
        ptx::mbar_wait(&sm.acc_empty[buf], (static_cast<uint32_t>(t >> 1) & 1) ^ 1);
        const uint32_t d_tmem = tmem + buf * q;
        for (int kc = 0; kc < kc_count; ++kc, ++it) {
          const int s = static_cast<int>(it % kStages);
          ptx::mbar_wait(&sm.full[s], static_cast<uint32_t>(it / kStages) & 1);
          ptx::tc_fence_after_sync();
          const uint32_t a_base = ptx::smem_u32(sm.stage[s]);
          const uint32_t w_off = static_cast<uint32_t>(kc) * q * 128;
#pragma unroll
          for (int ks = 0; ks < kChunk / 16; ++ks) {
            // K-major SWIZZLE_128B: 8-row groups 1024 B apart (SBO); a K = 16 step is 32 B along the row
            const uint64_t a1 = ptx::make_smem_desc(a_base + ks * 32, 16, 1024, ptx::kLayoutSwizzle128B);
            const uint64_t a2 = ptx::make_smem_desc(a_base + kPartBytes + ks * 32, 16, 1024, ptx::kLayoutSwizzle128B);
            const uint64_t b1 = ptx::make_smem_desc(w1 + w_off + ks * 32, 16, 1024, ptx::kLayoutSwizzle128B);
            const uint64_t b2 = ptx::make_smem_desc(w2 + w_off + ks * 32, 16, 1024, ptx::kLayoutSwizzle128B);
            mma_bf16_ss(d_tmem, a1, b1, idesc, (kc == 0 && ks == 0) ? 0u : 1u);
            mma_bf16_ss(d_tmem, a1, b2, idesc, 1u);
            mma_bf16_ss(d_tmem, a2, b1, idesc, 1u);
          }
          ptx::mma_commit(&sm.empty[s]);
        }
        ptx::mma_commit(&sm.acc_full[buf]);
      }
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == kMmaWarp) ptx::tmem_dealloc(tmem, tmem_cols);
}

__global__ void colsum_finalize_kernel(const double* __restrict__ partial, int n_partials, int q,
                                       double* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= q) return;
  double acc = 0.0;
  for (int i = 0; i < n_partials; ++i) acc += partial[static_cast<int64_t>(i) * q + c];
  out[c] = acc;
}

int rowproj_grid(int64_t n) {
  int sms = device_sm_count();
  if (sms <= 0) sms = 148;
  const int64_t tiles = (n + kTileRows - 1) / kTileRows;
  return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(sms, tiles)));
}

}  // namespace

bool rowproj_tc_supported(int64_t n, int d, int q, const void* x) {
  return n > 0 && d >= kChunk && d % kChunk == 0 && q >= 16 && q % 16 == 0 && q <= 256 &&
         static_cast<int64_t>(q) * d * 4 <= kMaxWBytes && reinterpret_cast<uintptr_t>(x) % 16 == 0;
}

// scratch: split W (2 q d bf16) + per-CTA column-sum partials
int64_t rowproj_tc_workspace(int64_t n, int d, int q) {
  return align_up(static_cast<int64_t>(2) * q * d * 2, 256) +
         static_cast<int64_t>(rowproj_grid(n)) * kEpiWarps * q * static_cast<int64_t>(sizeof(double)) + 512;
}

// out[n, q] = X W^T (y == nullptr), or the logistic epilogue: out = resid, colsum[q] (float64) =
// sum_n (y z - log(1 + exp z)).
int launch_rowproj_tc(const float* x, const float* w, const float* y, int64_t n, int d, int q, float* out,
                      double* colsum, void* workspace, int64_t workspace_bytes, cudaStream_t stream) {
  if (!rowproj_tc_supported(n, d, q, x) || reinterpret_cast<uintptr_t>(out) % 16 != 0) {
    set_error("rowproj_tc: unsupported shape n=%lld d=%d q=%d", static_cast<long long>(n), d, q);
    return BB_ERR_UNSUPPORTED;
  }
  if ((y == nullptr) != (colsum == nullptr)) {
    set_error("rowproj_tc: y and colsum go together");
    return BB_ERR_INVALID;
  }
  if (workspace == nullptr || workspace_bytes < rowproj_tc_workspace(n, d, q)) {
    set_error("rowproj_tc: workspace %lld < %lld bytes", static_cast<long long>(workspace_bytes),
              static_cast<long long>(rowproj_tc_workspace(n, d, q)));
    return BB_ERR_WORKSPACE;
  }
  uint8_t* ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~static_cast<uintptr_t>(255));
  __nv_bfloat16* wsplit = reinterpret_cast<__nv_bfloat16*>(ws);
  ws += align_up(static_cast<int64_t>(2) * q * d * 2, 256);
  double* partial = reinterpret_cast<double*>(ws);
  const int64_t count = static_cast<int64_t>(q) * d;
  split_w_kernel<<<static_cast<int>((count + 255) / 256), 256, 0, stream>>>(w, count, wsplit);
  BB_CHECK_LAUNCH("split_w_kernel");
  static const int prefetch_iters = getenv("BB_ROWPROJ_PREFETCH") ? atoi(getenv("BB_ROWPROJ_PREFETCH")) : 2;
  RowProjParams p;
  p.prefetch_iters = prefetch_iters;
  p.x = x; p.wsplit = wsplit; p.y = y; p.out = out; p.partial_colsum = partial; p.n = n; p.d = d; p.q = q;
  const int grid = rowproj_grid(n);
  const int smem_bytes = static_cast<int>(sizeof(SmemLayout));
  static SmemOptIn smem_opt_in_0;
  BB_CUDA_OK(smem_opt_in_0.ensure(rowproj_kernel<kEpiStore>, smem_bytes));
  static SmemOptIn smem_opt_in_1;
  BB_CUDA_OK(smem_opt_in_1.ensure(rowproj_kernel<kEpiLogistic>, smem_bytes));
  if (y == nullptr) {
    rowproj_kernel<kEpiStore><<<grid, kThreads, smem_bytes, stream>>>(p);
    BB_CHECK_LAUNCH("rowproj_kernel<store>");
  } else {
    rowproj_kernel<kEpiLogistic><<<grid, kThreads, smem_bytes, stream>>>(p);
    BB_CHECK_LAUNCH("rowproj_kernel<logistic>");
    colsum_finalize_kernel<<<(q + 63) / 64, 64, 0, stream>>>(partial, grid * kEpiWarps, q, colsum);
    BB_CHECK_LAUNCH("colsum_finalize_kernel");
  }
  return BB_OK;
}

}  // namespace bb
