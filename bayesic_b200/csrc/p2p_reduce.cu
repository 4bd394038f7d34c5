// All-reduce (sum) of the per-rank partial statistics over NVLink peer memory -- the "combine" step the
// reference's iid summation implies once the data axis is sharded (bayesic/distribution/base.py:328-332;
// SURVEY.md 8(e)).  Payloads of this path: cfg2 33 KB (fused into the statistics kernel itself,
// suffstats_sm100.cu), cfg5 260 KB, cfg4 8.4 MB, cfg3 8.5 MB of float64.
//
// One kernel per rank, no library call, no host round trip: a two-shot all-reduce in which every CTA
// talks only to its counterpart CTA on the peers, so there is no grid-wide barrier anywhere.
//
//   payload [count] is cut into `world` chunks (chunk r is reduced BY rank r) and every chunk into G
//   sub-slices, G = gridDim.x (a function of count only, so all ranks agree).  CTA j of rank q:
//     phase 0  publishes "rank q's input is complete" to CTA j of every peer (st.release.sys into the
//              peer's flag array) and waits for the same from every peer -- the input was written by
//              earlier kernels of the same stream, so any CTA can vouch for it;
//     phase 1  PULLS sub-slice j of chunk q from every rank's input buffer (peer loads, 128-bit), adds
//              them in rank order (bit-identical result on every rank) and PUSHES the sums into every
//              rank's output buffer; fence.sys; publishes "chunk q / sub-slice j delivered";
//     phase 2  waits until every rank has delivered sub-slice j of its chunk.
//   When the kernel completes on rank q, its output buffer holds the whole reduced payload and no peer
//   reads rank q's input any more (a peer's delivery flag is set after its pulls), so the input buffer
//   may be overwritten at once: no double buffering.  The output buffer of epoch e is overwritten by a
//   peer only after this rank has published phase 0 of epoch e + 1, i.e. after everything this rank
//   enqueued before its next all-reduce has run -- consumers must be stream-ordered before that call.
//
// Flags hold epochs (monotonic, compared as signed distances), so they are never reset.  A spin limit
// turns a lost peer into a status flag AND poisons the output with NaN -- never a partial sum.
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "kernels.h"

namespace bb {
namespace {

constexpr int kCommThreads = 512;

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double2 ld_sys_f64x2(const double* p) {
  double2 v;
  asm volatile("ld.relaxed.sys.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_sys_f64(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

struct CommParams {
  const double* const* in;      // [world] input buffers (own + peers' mapped)
  double* const* out;           // [world] output buffers
  uint32_t* const* flags;       // [world] flag arrays: [2 phases][world][kCommMaxCtas] uint32
  int rank, world;
  int64_t count, chunk;         // chunk = elements per rank (even)
  uint32_t* epoch_dev;          // device word: epochs completed (this launch is stored + 1; graph-replayable)
  unsigned int* ticket;         // CTA ticket, zero between launches (the last CTA stores the epoch and resets it)
  long long spin_limit;
  int* status;
};

// Thread t < world: tell peer t (slot [phase][rank][cta]) and wait for peer t's word in my own array.
__device__ __forceinline__ bool handshake(const CommParams& p, uint32_t epoch, int phase, int t, int cta) {
  uint32_t* theirs = p.flags[t] + (static_cast<int64_t>(phase) * p.world + p.rank) * BB_COMM_MAX_CTAS + cta;
  const uint32_t* mine = p.flags[p.rank] + (static_cast<int64_t>(phase) * p.world + t) * BB_COMM_MAX_CTAS + cta;
  st_release_sys(theirs, epoch);
  const long long t0 = clock64();
  while (static_cast<int32_t>(ld_acquire_sys(mine) - epoch) < 0) {
    if (clock64() - t0 > p.spin_limit) {
      atomicMax(p.status, t + 1);
      return false;
    }
  }
  return true;
}

__global__ void __launch_bounds__(kCommThreads) p2p_allreduce_kernel(const CommParams p) {
  __shared__ int lost;
  __shared__ uint32_t epoch_s;
  const int t = threadIdx.x, j = blockIdx.x, G = gridDim.x;
  if (t == 0) {
    lost = 0;
    epoch_s = __ldcg(p.epoch_dev) + 1u;      // stored back by the last CTA, which needs every CTA's ticket first
  }
  __syncthreads();
  const uint32_t epoch = epoch_s;
  if (t < p.world) {
    __threadfence_system();
    if (!handshake(p, epoch, 0, t, j)) lost = 1;
  }
  __syncthreads();
  const int64_t sub = ((p.chunk + G - 1) / G + 1) & ~int64_t(1);                 // even
  const int64_t chunk_lo = p.rank * p.chunk;
  const int64_t chunk_hi = min(p.count, chunk_lo + p.chunk);
  const int64_t lo = chunk_lo + j * sub;
  const int64_t hi = min(chunk_hi, lo + sub);
  const bool poisoned = lost != 0;
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  for (int64_t i = lo + 2 * t; i < hi; i += 2 * kCommThreads) {
    if (i + 1 < hi) {
      double2 acc = make_double2(0.0, 0.0);
      for (int r = 0; r < p.world; ++r) {
        const double2 v = ld_sys_f64x2(p.in[r] + i);
        acc.x += v.x;
        acc.y += v.y;
      }
      if (poisoned) acc = make_double2(nan, nan);
      for (int r = 0; r < p.world; ++r) *reinterpret_cast<double2*>(p.out[r] + i) = acc;
    } else {
      double acc = 0.0;
      for (int r = 0; r < p.world; ++r) acc += ld_sys_f64(p.in[r] + i);
      if (poisoned) acc = nan;
      for (int r = 0; r < p.world; ++r) p.out[r][i] = acc;
    }
  }
  __threadfence_system();
  __syncthreads();
  if (t < p.world && !handshake(p, epoch, 1, t, j)) lost = 1;
  __syncthreads();
  if (lost) {
    // a peer never answered: its chunk of the local output is stale (or never written) -- poison all of it
    for (int64_t i = static_cast<int64_t>(j) * kCommThreads + t; i < p.count; i += static_cast<int64_t>(G) * kCommThreads)
      p.out[p.rank][i] = nan;
  }
  __syncthreads();
  if (t == 0 && atomicAdd(p.ticket, 1u) == static_cast<unsigned int>(G) - 1u) {
    *p.epoch_dev = epoch;
    __threadfence();
    *p.ticket = 0u;
  }
}

}  // namespace

int comm_grid_for(int64_t count, int world) {
  // a function of (count, world) only: every rank must launch the same grid
  const int64_t chunk = (count + world - 1) / world;
  int64_t g = (chunk + 4095) / 4096;
  if (g < 1) g = 1;
  if (g > BB_COMM_MAX_CTAS) g = BB_COMM_MAX_CTAS;
  return static_cast<int>(g);
}

int launch_p2p_allreduce(const double* const* in, double* const* out, uint32_t* const* flags, int rank, int world,
                         int64_t count, uint32_t* epoch_dev, unsigned int* ticket, double spin_limit_ms, int* status,
                         cudaStream_t stream) {
  if (world < 1 || world > kCommThreads || rank < 0 || rank >= world || count < 0 || !in || !out || !flags || !status ||
      !epoch_dev || !ticket) {
    set_error("p2p_allreduce: bad rank/world/count (rank %d world %d count %lld)", rank, world,
              static_cast<long long>(count));
    return BB_ERR_INVALID;
  }
  if (count == 0) return BB_OK;
  CommParams p;
  p.in = in; p.out = out; p.flags = flags; p.rank = rank; p.world = world; p.count = count;
  p.chunk = (((count + world - 1) / world) + 1) & ~int64_t(1);
  p.epoch_dev = epoch_dev;
  p.ticket = ticket;
  p.spin_limit = static_cast<long long>(spin_limit_ms * 2.0e6);      // ~2 GHz ticks
  p.status = status;
  p2p_allreduce_kernel<<<comm_grid_for(count, world), kCommThreads, 0, stream>>>(p);
  BB_CHECK_LAUNCH("p2p_allreduce_kernel");
  return BB_OK;
}

}  // namespace bb
