// Gaussian-mixture expected log-densities (the "logits" of the responsibility softmax) on tcgen05
// (D % 8 == 0, D <= 64 -- the feature axis is zero-padded to a multiple of 16 on chip --, K % 4 == 0):
//     logit[n, k] = c_k - 1/2 || U_k x_n - t_k ||^2          (+ lse[n] = logsumexp_k logit[n, :])
// i.e.  c'_k + x.b_k - 1/2 x^T A_k x  with A_k = U_k^T U_k (Cholesky), t_k = U_k m_k -- the
// expression  dot(X, bk.T) + (-0.5) * einsum(X_nd Ak_kde X_ne) + ck  that a user of the reference
// writes for the VMP local step (README.md:30-37).  The reference plans the quadratic form as
//     _tensordot(_tensordot(X, Ak', ...), X^T, batch axes ...)          
// whose batched evaluation is broken (bayesic/algebra.py:1370-1373, :1380) and which would
// materialise an N x K x D intermediate; in the whitened form the same value is ONE projection
// Z = X [U_1; ...; U_K]^T (2 K D^2 flop/row) whose N x (K D) result is consumed on chip.
//
// Design (tensor-pipe bound: 2 K D^2 flop/row against 4 D bytes in, 4 K bytes out per row):
//   * persistent CTAs over 128-row tiles; the X tile is split once into error-compensated BF16
//     (x = b1 + b2, see gram_sm100.cu), K-major SWIZZLE_128B, and stays in shared memory while
//     the CTA sweeps all component groups (4 components = 256 columns per MMA group);
//   * the stacked factors are split and laid out in the UMMA shared-memory layout ONCE by a
//     pre-kernel, so a group's 64 KB operand tile arrives with two plain bulk copies
//     (cp.async.bulk, no tensor map) through a 2-stage mbarrier ring; it is L2-resident (4 MB);
//   * per group 3 x D/16 kind::f16 MMAs (b1 w1 + b1 w2 + b2 w1), M = 128, N = 256, FP32 in TMEM,
//     double-buffered (2 x 256 columns) so the epilogue of group g overlaps the MMAs of g + 1;
//   * eight epilogue warps (lane = data row, two warps per TMEM lane quadrant, 2 components each)
//     read Z from TMEM, subtract t (fetched coalesced, broadcast by shuffle), square-accumulate,
//     write the logit and keep an online log-sum-exp per row.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <algorithm>

#include "common.cuh"
#include "kernels.h"
#include "sm100_ptx.cuh"

namespace bb {

namespace {

constexpr int kTileRows = 128;
constexpr int kGroupCols = 256;                    // = 4 x 64 padded factor rows
constexpr int kAPart = kTileRows * 128;            // 16 KB: one bf16 part of the X tile (128-byte rows)
constexpr int kABytes = 2 * kAPart;                // 32 KB
constexpr int kWPart = kGroupCols * 128;           // 32 KB: one bf16 part of a 256-column factor tile
constexpr int kWBytes = 2 * kWPart;                // 64 KB per tile: [half 0: b1 | b2][half 1: b1 | b2]
constexpr int kWHalfPart = kWPart / 2;             // 16 KB: 128 columns of one part
constexpr int kWHalfBytes = kWBytes / 2;           // 32 KB: what one CTA of the pair loads per tile
constexpr int kWStages = 4;
constexpr int kEpiWarps = 16;                      // warps 0-15: (TMEM lane quadrant, 64-column slice)
constexpr int kProdWarp = 16;
constexpr int kMmaWarp = 17;
constexpr int kConvWarp0 = 18;                     // warps 18-21
constexpr int kConvWarps = 4;
constexpr int kThreads = (kConvWarp0 + kConvWarps) * 32;   // 704
constexpr int kTmemCols = 512;

struct __align__(1024) SmemLayout {
  uint8_t a[2][kABytes];
  uint8_t w[kWStages][kWHalfBytes];
  float lse_m[2][3][kTileRows];    // partial (max, sum) of column slices 1-3, per row
  float lse_s[2][3][kTileRows];
  float tscratch[kEpiWarps][32];   // per-warp staging of 32 t values (read back as broadcast pairs)
  double sum_lse;
  uint64_t a_full[2], a_empty[2];
  uint64_t w_full[kWStages], w_empty[kWStages];
  uint64_t w_peer[kWStages];       // leader: the peer CTA's half of the factor tile has landed
  uint64_t acc_full[2], acc_empty[2];
  uint64_t lse_ready[2], lse_taken[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` (no release fence: the
// caller has made its writes visible with fence.proxy.async / tcgen05 fences + __syncwarp)
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t}\n"
      ::"r"(ptx::smem_u32(bar)), "r"(rank)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_slot, uint32_t n_cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(smem_slot)),
               "r"(n_cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t n_cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(n_cols) : "memory");
}
__device__ __forceinline__ void mma_bf16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                              uint32_t accumulate) {
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
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(ptx::smem_u32(bar)), "h"(static_cast<uint16_t>(3))
      : "memory");
}
// plain (non-tensor) bulk copy global -> shared, completion counted on an mbarrier
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(ptx::smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(ptx::smem_u32(bar))
      : "memory");
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

// Column order of the projection: MMA tile (cg, jr) holds, for the 16 components 16 cg .. 16 cg + 15,
// rows j in [16 jr, 16 jr + 16) of their factors: column c * 16 + jj <-> (component 16 cg + c,
// row j = 16 jr + jj).  For upper-triangular factors (Cholesky) row j is zero left of feature j, so
// tile jr only needs the K steps ks >= jr: 10 of 16 MMA K-steps at D = 64.
// U[k, j, i] float32 -> wprep[tile = 4 cg + jr][half][part][128 rows x 128 B] bf16, K-major SWIZZLE_128B
// (components >= k, rows j >= d and features i >= d are zero); t[k, j] -> tprep[tile][256].
__global__ void prep_factors_kernel(const float* __restrict__ u, const float* __restrict__ t, int k_total, int k16,
                                    int d, uint8_t* __restrict__ wprep, float* __restrict__ tprep) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;   // (k, j, chunk of 8 features)
  const int64_t total = static_cast<int64_t>(k16) * 64 * 8;
  if (idx >= total) return;
  const int c8 = static_cast<int>(idx % 8);
  const int j = static_cast<int>((idx / 8) % 64);
  const int k = static_cast<int>(idx / (8 * 64));
  __align__(16) __nv_bfloat16 b1[8], b2[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int i = c8 * 8 + e;
    const float x = (k < k_total && j < d && i < d) ? u[(static_cast<int64_t>(k) * d + j) * d + i] : 0.f;
    b1[e] = __float2bfloat16_rn(x);
    b2[e] = __float2bfloat16_rn(x - __bfloat162float(b1[e]));
  }
  const int tile = (k / 16) * 4 + j / 16, r = (k % 16) * 16 + j % 16;
  // the tile's 256 columns are split between the two CTAs of a pair: half = r / 128
  const int rh = r & 127;
  const uint32_t off = (rh >> 3) * 1024 + (rh & 7) * 128 + ((c8 ^ (rh & 7)) << 4);
  uint8_t* base = wprep + static_cast<int64_t>(tile) * kWBytes + (r >> 7) * kWHalfBytes;
  *reinterpret_cast<uint4*>(base + off) = *reinterpret_cast<const uint4*>(b1);
  *reinterpret_cast<uint4*>(base + kWHalfPart + off) = *reinterpret_cast<const uint4*>(b2);
  if (c8 == 0) tprep[static_cast<int64_t>(tile) * 256 + r] = (k < k_total && j < d) ? t[static_cast<int64_t>(k) * d + j] : 0.f;
}

struct LogitParams {
  const float* x;
  const uint8_t* wprep;
  const float* tprep;       // [tile][256]
  const float* c;           // [k]
  float* logits;            // [n, k]
  float* lse;               // [n] or nullptr
  double* partial_sum_lse;  // [grid] or nullptr
  int64_t n;
  int d, k;
  int triangular;           // factors are upper triangular: tile jr skips the K steps below jr
};

// One CTA PAIR (cluster of 2, tcgen05 cta_group::2, M = 256, N = 256) per 256-row tile: CTA r owns
// rows [128 r, +128) (A operand, accumulator, epilogue) and loads half of every factor tile, so
// the L2 -> SM traffic of the factor stream -- what bounded the single-CTA version at 8.4 TB/s --
// is halved.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1) mixture_logits_kernel(const LogitParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  SmemLayout& sm = *reinterpret_cast<SmemLayout*>(smem_raw);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int n_cg = (p.k + 15) / 16;          // component groups of 16
  const int k_steps = (p.d + 15) / 16;       // = number of j ranges with data (d % 16 == 8: the last one is half zeros)
  const int n_seq = n_cg * k_steps;          // MMA tiles per row tile, in order (cg, jr)
  const int64_t n_ptiles = (p.n + 2 * kTileRows - 1) / (2 * kTileRows);      // 256-row pair tiles
  const int64_t my_tiles = n_ptiles > pair ? (n_ptiles - pair + n_pairs - 1) / n_pairs : 0;

  if (warp == kMmaWarp) {
    if (lane == 0) {
      for (int b = 0; b < 2; ++b) {
        ptx::mbar_init(&sm.a_full[b], 2 * kConvWarps);        // leader: both CTAs' converter warps
        ptx::mbar_init(&sm.a_empty[b], 1);
        ptx::mbar_init(&sm.acc_full[b], 1);
        ptx::mbar_init(&sm.acc_empty[b], 2 * kEpiWarps);      // leader: both CTAs' epilogue warps
        ptx::mbar_init(&sm.lse_ready[b], 12);
        ptx::mbar_init(&sm.lse_taken[b], 4);
      }
      for (int s = 0; s < kWStages; ++s) {
        ptx::mbar_init(&sm.w_full[s], 1);
        ptx::mbar_init(&sm.w_empty[s], 1);
        ptx::mbar_init(&sm.w_peer[s], 1);
      }
      ptx::fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc_pair(&sm.tmem_base, kTmemCols);
  }
  if (threadIdx.x == 0) sm.sum_lse = 0.0;
  // zero the padding features of both X tile buffers once (d < 64: chunks beyond d stay zero)
  for (int i = threadIdx.x; i < 2 * kABytes / 16; i += kThreads)
    reinterpret_cast<uint4*>(&sm.a[0][0])[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  ptx::tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();          // both CTAs' barriers initialised before any remote arrive
  ptx::tc_fence_after_sync();
  const uint32_t tmem = sm.tmem_base;

  if (warp < kEpiWarps) {
    // ---------------- epilogue: lane = data row; warp = (quadrant q, 64-column slice h) ----------------
    // Per value one packed FADD2 / FFMA2 half-instruction: t is staged per warp in shared memory and
    // read back as broadcast 64-bit pairs (instead of one shuffle per value), (z - t)^2 is
    // accumulated in f32x2 pairs.  16 warps (4 per scheduler) hide the TMEM-load and ALU latencies.
    const int q = warp & 3, h = warp >> 2;
    float* scratch = sm.tscratch[warp];
    const uint32_t scratch_addr = ptx::smem_u32(scratch);
    double sum_lse = 0.0;
    int64_t gi = 0;
    // t values of a tile are fetched one tile ahead (a load issued right before its use stalled the
    // epilogue on the L2 latency -- the top stall in the first ncu profile)
    float tl_next[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) tl_next[i] = __ldg(p.tprep + h * 64 + i * 32 + lane);       // tile (0, 0)
    for (int64_t t = 0; t < my_tiles; ++t) {
      const int64_t ptile = pair + t * n_pairs;
      const int64_t row = (ptile * 2 + rank) * kTileRows + q * 32 + lane;
      const bool valid = row < p.n;
      float run_m = -INFINITY, run_s = 0.f;
      for (int cg = 0; cg < n_cg; ++cg) {
        // this warp's 4 components of the group: 16 cg + 4 h + (0 .. 3)
        const int comp0 = cg * 16 + h * 4;
        float cc[4];
        uint64_t acc2[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          cc[c] = comp0 + c < p.k ? __ldg(p.c + comp0 + c) : 0.f;
          acc2[c] = 0ull;
        }
        for (int jr = 0; jr < k_steps; ++jr, ++gi) {
          float tl[2];
          tl[0] = tl_next[0];
          tl[1] = tl_next[1];
          {
            int ncg = cg, njr = jr + 1;
            if (njr == k_steps) { njr = 0; ncg = (cg + 1 == n_cg) ? 0 : cg + 1; }
            const float* src = p.tprep + static_cast<int64_t>(ncg * 4 + njr) * 256 + h * 64 + lane;
            tl_next[0] = __ldg(src);
            tl_next[1] = __ldg(src + 32);
          }
          const int ab = static_cast<int>(gi & 1);
          ptx::mbar_wait(&sm.acc_full[ab], static_cast<uint32_t>(gi >> 1) & 1);
          ptx::tc_fence_after_sync();
          const uint32_t t_addr = tmem + (static_cast<uint32_t>(q * 32) << 16) + ab * kGroupCols + h * 64;
#pragma unroll
          for (int quarter = 0; quarter < 2; ++quarter) {        // 32 columns = 2 components x 16 rows j
            uint32_t v[32];
            tmem_ld_32x32b_x32(t_addr + quarter * 32, v);
            scratch[lane] = tl[quarter];
            __syncwarp();
            ptx::tmem_wait_ld();
#pragma unroll
            for (int m = 0; m < 16; m += 2) {                    // two pairs per 128-bit broadcast read
              uint64_t t01, t23;
              asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(t01), "=l"(t23) : "r"(scratch_addr + m * 8));
              uint64_t z01, z23, d01, d23;
              asm("mov.b64 %0, {%1, %2};" : "=l"(z01) : "r"(v[2 * m]), "r"(v[2 * m + 1]));
              asm("mov.b64 %0, {%1, %2};" : "=l"(z23) : "r"(v[2 * m + 2]), "r"(v[2 * m + 3]));
              asm("sub.f32x2 %0, %1, %2;" : "=l"(d01) : "l"(z01), "l"(t01));
              asm("sub.f32x2 %0, %1, %2;" : "=l"(d23) : "l"(z23), "l"(t23));
              uint64_t& acc = acc2[quarter * 2 + (m >> 3)];
              asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(acc) : "l"(d01));
              asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(acc) : "l"(d23));
            }
            __syncwarp();                                        // scratch is rewritten by the next quarter
          }
          ptx::tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster_relaxed(&sm.acc_empty[ab], 0);      // TMEM reads done (wait::ld above)
        }
        // the 4 logits of this row are complete: store, fold into the online log-sum-exp
        if (comp0 < p.k) {
          float lg[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float sq = __uint_as_float(static_cast<uint32_t>(acc2[c])) + __uint_as_float(static_cast<uint32_t>(acc2[c] >> 32));
            lg[c] = comp0 + c < p.k ? fmaf(-0.5f, sq, cc[c]) : -INFINITY;
          }
          if (valid) *reinterpret_cast<float4*>(p.logits + row * p.k + comp0) = make_float4(lg[0], lg[1], lg[2], lg[3]);
          float m_new = run_m;
#pragma unroll
          for (int c = 0; c < 4; ++c) m_new = fmaxf(m_new, lg[c]);
          float sum = run_s * __expf(run_m - m_new);
#pragma unroll
          for (int c = 0; c < 4; ++c) sum += __expf(lg[c] - m_new);      // exp(-inf) = 0 for padding
          run_s = sum;
          run_m = m_new;
        }
      }
      // combine the four column slices of each row: h = 1..3 hand (m, s) to h = 0 through shared memory
      const int tb = static_cast<int>(t & 1);
      const uint32_t ph = static_cast<uint32_t>(t >> 1) & 1;
      if (h != 0) {
        ptx::mbar_wait(&sm.lse_taken[tb], ph ^ 1);
        sm.lse_m[tb][h - 1][q * 32 + lane] = run_m;
        sm.lse_s[tb][h - 1][q * 32 + lane] = run_s;
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&sm.lse_ready[tb]);
      } else {
        ptx::mbar_wait(&sm.lse_ready[tb], ph);
        float m = run_m;
#pragma unroll
        for (int o = 0; o < 3; ++o) m = fmaxf(m, sm.lse_m[tb][o][q * 32 + lane]);
        float ssum = run_s * __expf(run_m - m);
#pragma unroll
        for (int o = 0; o < 3; ++o) ssum += sm.lse_s[tb][o][q * 32 + lane] * __expf(sm.lse_m[tb][o][q * 32 + lane] - m);
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&sm.lse_taken[tb]);
        const float lse = m + __logf(ssum);
        if (valid) {
          if (p.lse != nullptr) p.lse[row] = lse;
          sum_lse += static_cast<double>(lse);
        }
      }
    }
    if (h == 0 && p.partial_sum_lse != nullptr) atomicAdd(&sm.sum_lse, sum_lse);
  } else if (warp == kProdWarp) {
    // ---------------- factor-tile producer: two 32 KB bulk copies per group ----------------
    if (ptx::elect_one()) {
      int64_t gi = 0;
      for (int64_t t = 0; t < my_tiles; ++t)
        for (int g = 0; g < n_seq; ++g, ++gi) {
          const int s = static_cast<int>(gi % kWStages);
          ptx::mbar_wait(&sm.w_empty[s], (static_cast<uint32_t>(gi / kWStages) & 1) ^ 1);
          ptx::mbar_arrive_expect_tx(&sm.w_full[s], kWHalfBytes);
          const int wtile = (g / k_steps) * 4 + g % k_steps;
          const uint8_t* src = p.wprep + static_cast<int64_t>(wtile) * kWBytes + rank * kWHalfBytes;
          bulk_load(sm.w[s], src, kWHalfPart, &sm.w_full[s]);
          bulk_load(sm.w[s] + kWHalfPart, src + kWHalfPart, kWHalfPart, &sm.w_full[s]);
        }
    }
  } else if (warp == kMmaWarp && rank == 1) {
    // ---------------- peer relay: tell the leader when this CTA's half of a factor tile has landed ----------------
    if (ptx::elect_one()) {
      int64_t gi = 0;
      for (int64_t t = 0; t < my_tiles; ++t)
        for (int g = 0; g < n_seq; ++g, ++gi) {
          const int s = static_cast<int>(gi % kWStages);
          ptx::mbar_wait(&sm.w_full[s], static_cast<uint32_t>(gi / kWStages) & 1);
          mbar_arrive_cluster_relaxed(&sm.w_peer[s], 0);
        }
    }
  } else if (warp == kMmaWarp) {          // rank == 0 (the peer's MMA warp took the relay branch above)
    // ---------------- MMA issuer (leader CTA, one elected thread) ----------------
    if (ptx::elect_one()) {
      const uint32_t idesc = ptx::make_idesc(256, kGroupCols, /*bf16*/ 1, /*A K-major*/ 0, /*B K-major*/ 0);
      int64_t gi = 0;
      for (int64_t t = 0; t < my_tiles; ++t) {
        const int tb = static_cast<int>(t & 1);
        ptx::mbar_wait(&sm.a_full[tb], static_cast<uint32_t>(t >> 1) & 1);
        const uint32_t a_base = ptx::smem_u32(sm.a[tb]);
        for (int g = 0; g < n_seq; ++g, ++gi) {
          const int s = static_cast<int>(gi % kWStages);
          const int ab = static_cast<int>(gi & 1);
          const int ks_begin = p.triangular ? g % k_steps : 0;     // tile jr: rows j >= 16 jr are zero left of feature 16 jr
          ptx::mbar_wait(&sm.w_full[s], static_cast<uint32_t>(gi / kWStages) & 1);
          ptx::mbar_wait(&sm.w_peer[s], static_cast<uint32_t>(gi / kWStages) & 1);
          ptx::mbar_wait(&sm.acc_empty[ab], (static_cast<uint32_t>(gi >> 1) & 1) ^ 1);
          ptx::tc_fence_after_sync();
          const uint32_t w_base = ptx::smem_u32(sm.w[s]);
          const uint32_t d_tmem = tmem + ab * kGroupCols;
          for (int ks = ks_begin; ks < k_steps; ++ks) {
            const uint64_t a1 = ptx::make_smem_desc(a_base + ks * 32, 16, 1024, ptx::kLayoutSwizzle128B);
            const uint64_t a2 = ptx::make_smem_desc(a_base + kAPart + ks * 32, 16, 1024, ptx::kLayoutSwizzle128B);
            const uint64_t b1 = ptx::make_smem_desc(w_base + ks * 32, 16, 1024, ptx::kLayoutSwizzle128B);
            const uint64_t b2 = ptx::make_smem_desc(w_base + kWHalfPart + ks * 32, 16, 1024, ptx::kLayoutSwizzle128B);
            mma_bf16_pair(d_tmem, a1, b1, idesc, ks == ks_begin ? 0u : 1u);
            mma_bf16_pair(d_tmem, a1, b2, idesc, 1u);
            mma_bf16_pair(d_tmem, a2, b1, idesc, 1u);
          }
          mma_commit_pair(&sm.w_empty[s]);
          mma_commit_pair(&sm.acc_full[ab]);
        }
        mma_commit_pair(&sm.a_empty[tb]);
      }
    }
  } else {
    // ---------------- X tile converter: 4 warps x 32 rows, lane covers float4 (lane & 15) of 2 rows ----------------
    // (two halves of 16 rows per tile: the kernel runs 704 threads, so 80 registers per thread)
    const int wi = warp - kConvWarp0;
    const int sub = lane >> 4, c4 = lane & 15;
    const bool col_ok = c4 * 4 < p.d;
    const bool col_stored = c4 * 4 < 16 * k_steps;      // features d .. 16 k_steps - 1 are operand columns too: zeros
    for (int64_t t = 0; t < my_tiles; ++t) {
      const int tb = static_cast<int>(t & 1);
      const int64_t tile = (pair + t * n_pairs) * 2 + rank;
      const uint32_t a_base = ptx::smem_u32(sm.a[tb]);
#pragma unroll 1
      for (int hf = 0; hf < 2; ++hf) {
        float4 rx[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int64_t row = tile * kTileRows + wi * 32 + hf * 16 + 2 * i + sub;
          rx[i] = (col_ok && row < p.n) ? ldg_f4(p.x + row * p.d + c4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        // a tile lasts ~100k cycles: back off between polls instead of spinning over the epilogue warps
        if (hf == 0) ptx::mbar_wait_sleep(&sm.a_empty[tb], (static_cast<uint32_t>(t >> 1) & 1) ^ 1, 512);
        if (col_stored) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = wi * 32 + hf * 16 + 2 * i + sub;
            const uint32_t off = (r >> 3) * 1024 + (r & 7) * 128 + (((c4 >> 1) ^ (r & 7)) << 4) + (c4 & 1) * 8;
            uint32_t b1[2], b2[2];
            split_bf16(rx[i], b1, b2);
            sts_u2(a_base + off, b1[0], b1[1]);
            sts_u2(a_base + kAPart + off, b2[0], b2[1]);
          }
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster_relaxed(&sm.a_full[tb], 0);
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x == 0 && p.partial_sum_lse != nullptr) p.partial_sum_lse[blockIdx.x] = sm.sum_lse;
  cluster_sync_all();          // the peer's shared memory / barriers stay alive until both are done
  if (warp == kMmaWarp) tmem_dealloc_pair(tmem, kTmemCols);
}

__global__ void sum_partials_kernel(const double* __restrict__ partial, int n, double* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double acc = 0.0;
    for (int i = 0; i < n; ++i) acc += partial[i];
    *out = acc;
  }
}

int logits_grid(int64_t n) {           // CTAs (two per pair)
  int sms = device_sm_count();
  if (sms <= 0) sms = 148;
  const int64_t ptiles = (n + 2 * kTileRows - 1) / (2 * kTileRows);
  return 2 * static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(sms / 2, ptiles)));
}

}  // namespace

bool mixture_logits_supported(int64_t n, int d, int k, const void* x) {
  return n > 0 && d >= 8 && d <= 64 && d % 8 == 0 && k >= 4 && k % 4 == 0 && k <= 4096 &&
         reinterpret_cast<uintptr_t>(x) % 16 == 0;
}

int64_t mixture_logits_workspace(int64_t n, int d, int k) {
  (void)d;
  const int64_t n_cg = (k + 15) / 16;
  return n_cg * 4 * kWBytes + n_cg * 4 * 256 * 4 + static_cast<int64_t>(logits_grid(n)) * 8 + 1024;
}

// u [k, d, d] (upper Cholesky factors, row-major), t [k, d], c [k]: device float32.
// logits [n, k] float32 (required), lse [n] float32 and sum_lse (float64) optional.
int launch_mixture_logits(const float* x, const float* u, const float* t, const float* c, int64_t n, int d, int k,
                          int upper_triangular, float* logits, float* lse, double* sum_lse, void* workspace,
                          int64_t workspace_bytes, cudaStream_t stream) {
  if (!mixture_logits_supported(n, d, k, x) || reinterpret_cast<uintptr_t>(logits) % 8 != 0) {
    set_error("mixture_logits: unsupported shape n=%lld d=%d k=%d", static_cast<long long>(n), d, k);
    return BB_ERR_UNSUPPORTED;
  }
  if (workspace == nullptr || workspace_bytes < mixture_logits_workspace(n, d, k)) {
    set_error("mixture_logits: workspace %lld < %lld bytes", static_cast<long long>(workspace_bytes),
              static_cast<long long>(mixture_logits_workspace(n, d, k)));
    return BB_ERR_WORKSPACE;
  }
  uint8_t* ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~static_cast<uintptr_t>(255));
  const int n_cg = (k + 15) / 16;
  uint8_t* wprep = ws;
  ws += static_cast<int64_t>(n_cg) * 4 * kWBytes;
  float* tprep = reinterpret_cast<float*>(ws);
  ws += static_cast<int64_t>(n_cg) * 4 * 256 * 4;
  double* partial = reinterpret_cast<double*>(ws);
  const int64_t prep_items = static_cast<int64_t>(n_cg) * 16 * 64 * 8;
  prep_factors_kernel<<<static_cast<int>((prep_items + 255) / 256), 256, 0, stream>>>(u, t, k, n_cg * 16, d, wprep, tprep);
  BB_CHECK_LAUNCH("prep_factors_kernel");
  LogitParams p;
  p.x = x; p.wprep = wprep; p.tprep = tprep; p.c = c; p.logits = logits; p.lse = lse;
  p.partial_sum_lse = sum_lse != nullptr ? partial : nullptr;
  p.n = n; p.d = d; p.k = k; p.triangular = upper_triangular ? 1 : 0;
  const int grid = logits_grid(n);
  const int smem_bytes = static_cast<int>(sizeof(SmemLayout));
  static SmemOptIn smem_opt_in_0;
  BB_CUDA_OK(smem_opt_in_0.ensure(mixture_logits_kernel, smem_bytes));
  mixture_logits_kernel<<<grid, kThreads, smem_bytes, stream>>>(p);
  BB_CHECK_LAUNCH("mixture_logits_kernel");
  if (sum_lse != nullptr) {
    sum_partials_kernel<<<1, 32, 0, stream>>>(partial, grid, sum_lse);
    BB_CHECK_LAUNCH("sum_partials_kernel");
  }
  return BB_OK;
}

}  // namespace bb
