// Data-axis contraction of two matrices on tcgen05 (D % 128 == 0, Q % 64 == 0, (D/128)(Q/64) <= 4):
//     G[d, q] = sum_n X[n, d] R[n, q]
// What it replaces: the plan  _tensordot(_dimshuffle(X,1,0), R, [1],[0])  (bayesic/algebra.py:
// 527-551 -> 1347-1351; "dot(X.T, R)"), one Theano/BLAS sgemm over the data axis -- e.g. the
// gradient G = X^T (y - sigmoid(Z)) of the reparameterised logistic-regression ELBO (BASELINE
// cfg5; README.md:47-51).
//
// Design (HBM-bound: 4 (D + Q) bytes/row against 2 D Q flop/row):
//   * the whole D x Q result fits one CTA's TMEM ((D/128) blocks of 128 lanes x Q columns, double
//     buffered), so every CTA streams its own contiguous range of rows exactly once -- no
//     cross-CTA operand sharing is needed;
//   * the data axis is the MMA K axis, so both X and R are MN-major as they lie in memory: 16
//     converter warps load them with coalesced 128-bit loads straight into registers, split them
//     into error-compensated BF16 (x = b1 + b2, see gram_sm100.cu) and store the tiles in the
//     UMMA MN-major SWIZZLE_128B layout, 16 rows (one K step) per pipeline stage;
//   * per stage and 128-feature block three kind::f16 MMAs (x1 r1 + x1 r2 + x2 r1), M = 128,
//     N = Q; FP32 accumulation in TMEM drained every 2048 rows into the CTA's fp32 partial block
//     (coalesced read-modify-write through L2), as in gram_sm100.cu;
//   * a finalize kernel adds the per-CTA partials in float64 in a fixed order.
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

constexpr int kStageRows = 16;                 // one K = 16 step per stage
constexpr int kFlushIters = 128 / BB_CHAIN_DIV;               // 2048 rows per TMEM accumulation chain
constexpr int kConvWarps = 16;
constexpr int kConvGroups = 2;
constexpr int kEpiWarps = 4;
constexpr int kMmaWarp = kConvWarps + kEpiWarps;
constexpr int kThreads = (kMmaWarp + 1) * 32;  // 672
constexpr int kMaxStages = 8;
constexpr int kSmemBudget = 220 * 1024;

struct Bars {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
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

// Stage layout: [X b1 | X b2 | R b1 | R b2]; each part is (extent / 64) MN blocks of 2 KB:
// [k group (row >> 3) 1 KB][row & 7 -> 128 B][16-byte chunk ^ (row & 7)].
template <int kNSeg, int kNQ>
__global__ void __launch_bounds__(kThreads, 1)
colproj_kernel(const float* __restrict__ x, const float* __restrict__ r, int64_t n, int n_stages, int prefetch_iters,
               float* __restrict__ partial) {         // [cta][kNSeg][Q cols][128 rows] fp32
  constexpr int kD = kNSeg * 128, kQ = kNQ * 64;
  constexpr int kXPart = kD * 32, kRPart = kQ * 32;             // bytes per bf16 part per stage
  constexpr int kStageBytes = 2 * kXPart + 2 * kRPart;
  constexpr int kAccCols = kNSeg * kQ;                          // <= 256
  constexpr uint32_t kTmemCols = 2 * kAccCols <= 32 ? 32 : (2 * kAccCols <= 64 ? 64 : (2 * kAccCols <= 128 ? 128 : (2 * kAccCols <= 256 ? 256 : 512)));
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* stages = smem_raw;
  Bars& bar = *reinterpret_cast<Bars*>(smem_raw + static_cast<size_t>(n_stages) * kStageBytes);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t total_iters = (n + kStageRows - 1) / kStageRows;
  const int64_t it_begin = total_iters * blockIdx.x / gridDim.x;
  const int64_t it_end = total_iters * (blockIdx.x + 1) / gridDim.x;
  const int n_iters = static_cast<int>(it_end - it_begin);
  const int64_t row_begin = it_begin * kStageRows;

  if (warp == kMmaWarp) {
    if (lane == 0) {
      for (int s = 0; s < n_stages; ++s) {
        ptx::mbar_init(&bar.full[s], kConvWarps / kConvGroups);
        ptx::mbar_init(&bar.empty[s], 1);
      }
      for (int b = 0; b < 2; ++b) {
        ptx::mbar_init(&bar.acc_full[b], 1);
        ptx::mbar_init(&bar.acc_empty[b], kEpiWarps);
      }
      ptx::fence_mbar_init();
    }
    __syncwarp();
    ptx::tmem_alloc(&bar.tmem_base, kTmemCols);
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem = bar.tmem_base;

  if (warp < kConvWarps) {
    // ---------------- converter warps ----------------
    // group g = warp & 1 takes every 2nd stage; warp wi of the group owns rows 2 wi, 2 wi + 1.
    // X: lane l holds features [128 seg + 4 l, +4) of both rows; R: lane l holds columns
    // [64 qs + 4 (l & 15), +4) of row 2 wi + (l >> 4).
    const int group = warp & (kConvGroups - 1);
    const int wi = warp / kConvGroups;
    uint32_t xoff[2], roff;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int k = 2 * wi + j;
      xoff[j] = (lane >> 4) * 2048 + (k >> 3) * 1024 + (k & 7) * 128 + ((((lane & 15) >> 1) ^ (k & 7)) << 4) + (lane & 1) * 8;
    }
    {
      const int k = 2 * wi + (lane >> 4);
      roff = (k >> 3) * 1024 + (k & 7) * 128 + ((((lane & 15) >> 1) ^ (k & 7)) << 4) + (lane & 1) * 8;
    }
    const uint32_t stage0 = ptx::smem_u32(stages);
    int64_t row0 = row_begin + static_cast<int64_t>(group) * kStageRows + 2 * wi;
    const float* px = x + row0 * kD + lane * 4;
    const float* pr = r + (row0 + (lane >> 4)) * kQ + (lane & 15) * 4;
    float4 rx[2][kNSeg], rr[kNQ];
    auto load = [&]() {
      if (row0 + 2 <= n) {
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int seg = 0; seg < kNSeg; ++seg) rx[j][seg] = ldg_f4(px + j * kD + seg * 128);
#pragma unroll
        for (int qs = 0; qs < kNQ; ++qs) rr[qs] = ldg_f4(pr + qs * 64);
      } else {
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int seg = 0; seg < kNSeg; ++seg)
            rx[j][seg] = (row0 + j < n) ? ldg_f4(px + j * kD + seg * 128) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int qs = 0; qs < kNQ; ++qs)
          rr[qs] = (row0 + (lane >> 4) < n) ? ldg_f4(pr + qs * 64) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      // L2 prefetch hints a few stages ahead: the register-resident loads above can keep only
      // ~64 KB per SM in flight, which covers the L2 latency but not the HBM latency
      if (prefetch_iters > 0) {
        const int64_t prow = row0 + static_cast<int64_t>(prefetch_iters) * kConvGroups * kStageRows;   // this warp's 2 rows, later
        constexpr int kXLines = kD * 4 / 128;                  // 128-byte lines per row of X
        const int64_t row_limit = (it_end < total_iters ? it_end * kStageRows : n);
        if (lane < 2 * kXLines) {
          const int64_t pr_row = prow + lane / kXLines;
          if (pr_row < row_limit) asm volatile("prefetch.global.L2 [%0];" ::"l"(x + pr_row * kD + (lane % kXLines) * 32));
        }
        constexpr int kRLines = kQ * 4 / 128;
        if (lane >= 32 - 2 * kRLines) {
          const int l2 = lane - (32 - 2 * kRLines);
          const int64_t pr_row = prow + l2 / kRLines;
          if (pr_row < row_limit) asm volatile("prefetch.global.L2 [%0];" ::"l"(r + pr_row * kQ + (l2 % kRLines) * 32));
        }
      }
      row0 += kConvGroups * kStageRows;
      px += static_cast<int64_t>(kConvGroups) * kStageRows * kD;
      pr += static_cast<int64_t>(kConvGroups) * kStageRows * kQ;
    };
    if (group < n_iters) load();
    for (int it = group; it < n_iters; it += kConvGroups) {
      const int s = it % n_stages;
      ptx::mbar_wait(&bar.empty[s], ((it / n_stages) & 1) ^ 1);
      const uint32_t stage_addr = stage0 + s * kStageBytes;
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int seg = 0; seg < kNSeg; ++seg) {
          uint32_t b1[2], b2[2];
          split_bf16(rx[j][seg], b1, b2);
          const uint32_t addr = stage_addr + seg * 4096 + xoff[j];
          sts_u2(addr, b1[0], b1[1]);
          sts_u2(addr + kXPart, b2[0], b2[1]);
        }
#pragma unroll
      for (int qs = 0; qs < kNQ; ++qs) {
        uint32_t b1[2], b2[2];
        split_bf16(rr[qs], b1, b2);
        const uint32_t addr = stage_addr + 2 * kXPart + qs * 2048 + roff;
        sts_u2(addr, b1[0], b1[1]);
        sts_u2(addr + kRPart, b2[0], b2[1]);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bar.full[s]);
      if (it + kConvGroups < n_iters) load();
    }
  } else if (warp < kMmaWarp) {
    // ---------------- epilogue warps: TMEM fp32 -> fp32 partial block (coalesced RMW) ----------------
    const int qd = warp & 3;
    const int n_intervals = (n_iters + kFlushIters - 1) / kFlushIters;
    float* my_partial = partial + static_cast<int64_t>(blockIdx.x) * kNSeg * kQ * 128 + qd * 32 + lane;
    if (n_intervals == 0) {
      for (int c = 0; c < kNSeg * kQ; ++c) my_partial[c * 128] = 0.f;
    }
    for (int interval = 0; interval < n_intervals; ++interval) {
      const int buf = interval & 1;
      ptx::mbar_wait_sleep(&bar.acc_full[buf], (interval >> 1) & 1);
      ptx::tc_fence_after_sync();
      const uint32_t t_addr = tmem + (static_cast<uint32_t>(qd * 32) << 16) + buf * kAccCols;
#pragma unroll 1
      for (int cc = 0; cc < kAccCols / 32; ++cc) {
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
      if (lane == 0) ptx::mbar_arrive(&bar.acc_empty[buf]);
    }
  } else {
    // ---------------- MMA issuer ----------------
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::make_idesc(128, kQ, /*bf16*/ 1, /*A MN-major*/ 1, /*B MN-major*/ 1);
      for (int it = 0; it < n_iters; ++it) {
        const int s = it % n_stages;
        const int interval = it / kFlushIters;
        const int buf = interval & 1;
        const bool first = (it % kFlushIters) == 0;
        if (first) ptx::mbar_wait(&bar.acc_empty[buf], ((interval >> 1) & 1) ^ 1);
        ptx::mbar_wait(&bar.full[s], (it / n_stages) & 1);
        ptx::tc_fence_after_sync();
        const uint32_t base = ptx::smem_u32(stages) + s * kStageBytes;
        // 64-wide MN blocks are 2 KB apart (LBO), the two 8-row k groups of the K = 16 step 1 KB (SBO)
        const uint64_t r1 = ptx::make_smem_desc(base + 2 * kXPart, 2048, 1024, ptx::kLayoutSwizzle128B);
        const uint64_t r2 = ptx::make_smem_desc(base + 2 * kXPart + kRPart, 2048, 1024, ptx::kLayoutSwizzle128B);
#pragma unroll
        for (int blk = 0; blk < kNSeg; ++blk) {
          const uint64_t x1 = ptx::make_smem_desc(base + blk * 4096, 2048, 1024, ptx::kLayoutSwizzle128B);
          const uint64_t x2 = ptx::make_smem_desc(base + kXPart + blk * 4096, 2048, 1024, ptx::kLayoutSwizzle128B);
          const uint32_t d_tmem = tmem + buf * kAccCols + blk * kQ;
          mma_bf16_ss(d_tmem, x1, r1, idesc, first ? 0u : 1u);
          mma_bf16_ss(d_tmem, x1, r2, idesc, 1u);
          mma_bf16_ss(d_tmem, x2, r1, idesc, 1u);
        }
        ptx::mma_commit(&bar.empty[s]);
        if ((it % kFlushIters) == kFlushIters - 1 || it == n_iters - 1) ptx::mma_commit(&bar.acc_full[buf]);
      }
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == kMmaWarp) ptx::tmem_dealloc(tmem, kTmemCols);
}

// G[d, q] (float64) = sum over CTAs of partial[cta][d / 128][q][d % 128]
__global__ void __launch_bounds__(256)
colproj_finalize_kernel(const float* __restrict__ partial, int n_ctas, int d, int q, double* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;          // consecutive threads: consecutive d
  if (idx >= d * q) return;
  const int col = idx / d, row = idx % d;
  const int64_t per_cta = static_cast<int64_t>(d) * q;
  const int64_t off = (static_cast<int64_t>(row / 128) * q + col) * 128 + row % 128;
  double acc = 0.0;
  for (int c = 0; c < n_ctas; ++c) acc += static_cast<double>(partial[c * per_cta + off]);
  out[static_cast<int64_t>(row) * q + col] = acc;
}

struct ColProjPlan {
  int grid, n_stages, smem_bytes;
};

ColProjPlan plan_colproj(int64_t n, int d, int q) {
  ColProjPlan p;
  int sms = device_sm_count();
  if (sms <= 0) sms = 148;
  const int64_t iters = (n + kStageRows - 1) / kStageRows;
  p.grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(sms, iters)));
  const int stage_bytes = (d + q) * 64;
  p.n_stages = std::max(2, std::min(kMaxStages, kSmemBudget / stage_bytes));
  p.smem_bytes = p.n_stages * stage_bytes + static_cast<int>(sizeof(Bars)) + 64;
  return p;
}

template <int kNSeg, int kNQ>
int launch_instance(const float* x, const float* r, int64_t n, const ColProjPlan& p, float* partial, cudaStream_t stream) {
  static SmemOptIn smem_opt_in;
  BB_CUDA_OK(smem_opt_in.ensure(colproj_kernel<kNSeg, kNQ>, kSmemBudget + 4096));
  // measured on B200 (cfg5): 0 -> 1.674 ms, 2 -> 1.575 ms, 4 -> 1.595 ms, 8 -> 2.08 ms
  static const int prefetch_iters = getenv("BB_COLPROJ_PREFETCH") ? atoi(getenv("BB_COLPROJ_PREFETCH")) : 2;
  colproj_kernel<kNSeg, kNQ><<<p.grid, kThreads, p.smem_bytes, stream>>>(x, r, n, p.n_stages, prefetch_iters, partial);
  BB_CHECK_LAUNCH("colproj_kernel");
  return BB_OK;
}

}  // namespace

bool colproj_tc_supported(int64_t n, int d, int q, const void* x, const void* r) {
  return n > 0 && d >= 128 && d % 128 == 0 && q >= 64 && q % 64 == 0 && (d / 128) * (q / 64) <= 4 &&
         reinterpret_cast<uintptr_t>(x) % 16 == 0 && reinterpret_cast<uintptr_t>(r) % 16 == 0;
}

int64_t colproj_tc_workspace(int64_t n, int d, int q) {
  const ColProjPlan p = plan_colproj(n, d, q);
  return static_cast<int64_t>(p.grid) * d * q * static_cast<int64_t>(sizeof(float)) + 512;
}

// out[d, q] float64 = X^T R
int launch_colproj_tc(const float* x, const float* r, int64_t n, int d, int q, double* out, void* workspace,
                      int64_t workspace_bytes, cudaStream_t stream) {
  if (!colproj_tc_supported(n, d, q, x, r)) {
    set_error("colproj_tc: unsupported shape n=%lld d=%d q=%d", static_cast<long long>(n), d, q);
    return BB_ERR_UNSUPPORTED;
  }
  if (workspace == nullptr || workspace_bytes < colproj_tc_workspace(n, d, q)) {
    set_error("colproj_tc: workspace %lld < %lld bytes", static_cast<long long>(workspace_bytes),
              static_cast<long long>(colproj_tc_workspace(n, d, q)));
    return BB_ERR_WORKSPACE;
  }
  const ColProjPlan p = plan_colproj(n, d, q);
  float* partial = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~static_cast<uintptr_t>(255));
  const int nseg = d / 128, nq = q / 64;
  int st = BB_ERR_UNSUPPORTED;
  if (nseg == 1 && nq == 1) st = launch_instance<1, 1>(x, r, n, p, partial, stream);
  else if (nseg == 2 && nq == 1) st = launch_instance<2, 1>(x, r, n, p, partial, stream);
  else if (nseg == 3 && nq == 1) st = launch_instance<3, 1>(x, r, n, p, partial, stream);
  else if (nseg == 4 && nq == 1) st = launch_instance<4, 1>(x, r, n, p, partial, stream);
  else if (nseg == 1 && nq == 2) st = launch_instance<1, 2>(x, r, n, p, partial, stream);
  else if (nseg == 2 && nq == 2) st = launch_instance<2, 2>(x, r, n, p, partial, stream);
  else if (nseg == 1 && nq == 3) st = launch_instance<1, 3>(x, r, n, p, partial, stream);
  else if (nseg == 1 && nq == 4) st = launch_instance<1, 4>(x, r, n, p, partial, stream);
  BB_TRY(st);
  colproj_finalize_kernel<<<(d * q + 255) / 256, 256, 0, stream>>>(partial, p.grid, d, q, out);
  BB_CHECK_LAUNCH("colproj_finalize_kernel");
  return BB_OK;
}

}  // namespace bb
