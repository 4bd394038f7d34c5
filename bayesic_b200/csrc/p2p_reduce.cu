// One-shot all-reduce over NVLink peer memory, fused with the consumer of the reduced statistics.
//
// The per-minibatch exchange of this path is tiny (cfg2: 4 161 float64 = 33 KB; cfg5: 33 K float64):
// its cost is latency, not bandwidth.  Instead of a collective library call followed by another
// kernel, ONE single-CTA kernel per rank
//   1. publishes "my partial statistics for epoch e are complete" by a system-scope release store of
//      e into slot [rank] of every peer's flag array (peer memory mapped into this process:
//      CUDA IPC / torch symmetric memory -- plumbing done by the caller),
//   2. waits (acquire loads) until its own flag array holds >= e from every peer,
//   3. sums all ranks' partial buffers with direct peer loads, in rank order (bit-identical
//      result on every rank), into local memory, and
//   4. (optionally) evaluates the expected log-likelihood from the reduced statistics in the
//      same kernel (stats_kernels.cu has the stand-alone version).
// Partial buffers are double-buffered by epoch parity, so one flag round per step suffices: a
// peer can only overwrite slot e & 1 for epoch e + 2 after this rank has published e + 1, i.e.
// after this rank's epoch-e kernel (and its reads) finished.  A spin limit turns a lost peer into a
// status flag instead of a hang.
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "kernels.h"

namespace bb {
namespace {

constexpr int kP2PThreads = 1024;

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

struct P2PParams {
  const double* const* bufs;     // [world] peer buffers, 2 * slot_stride doubles each
  uint32_t* const* flags;        // [world] peer flag arrays, >= world entries each
  int rank, world;
  int64_t count, slot_stride;
  uint32_t epoch;
  long long spin_limit;          // clock64 ticks
  double* out;                   // [count] reduced statistics (local)
  int* status;                   // set to 1 + peer index on a spin timeout
  // optional fused consumer: expected log-likelihood from the packed layout [S2 (d*d) | S1 (d) | count]
  const double* e_lambda;
  const double* e_lambda_mu;
  double e_mu_l_mu, e_logdet;
  int d;
  double* elbo;
};

__global__ void __launch_bounds__(kP2PThreads) p2p_allreduce_kernel(const P2PParams p) {
  __shared__ double part[kP2PThreads / 32];
  const int tid = threadIdx.x;
  if (tid < p.world) {
    __threadfence_system();
    st_release_sys(p.flags[tid] + p.rank, p.epoch);
    const long long t0 = clock64();
    // epochs are compared as signed distances so that the counter may wrap
    while (static_cast<int32_t>(ld_acquire_sys(p.flags[p.rank] + tid) - p.epoch) < 0) {
      if (clock64() - t0 > p.spin_limit) {
        atomicMax(p.status, tid + 1);
        break;
      }
    }
  }
  __syncthreads();
  const int64_t slot = static_cast<int64_t>(p.epoch & 1u) * p.slot_stride;
  for (int64_t i = tid; i < p.count; i += kP2PThreads) {
    double acc = 0.0;
    for (int r = 0; r < p.world; ++r) acc += __ldcv(p.bufs[r] + slot + i);
    p.out[i] = acc;
  }
  if (p.elbo == nullptr) return;
  __syncthreads();                                   // out[] was written by this CTA
  const int d = p.d;
  const double* s2 = p.out;
  const double* s1 = p.out + static_cast<int64_t>(d) * d;
  double acc = 0.0;
  for (int i = tid; i < d * d; i += kP2PThreads) acc -= 0.5 * p.e_lambda[i] * s2[i];
  for (int i = tid; i < d; i += kP2PThreads) acc += s1[i] * p.e_lambda_mu[i];
  for (int off = 16; off >= 1; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if ((tid & 31) == 0) part[tid >> 5] = acc;
  __syncthreads();
  if (tid == 0) {
    double total = 0.0;
    for (int w = 0; w < kP2PThreads / 32; ++w) total += part[w];
    const double n = s1[d];                          // reduced row count
    const double log_2pi = 1.8378770664093454835606594728112;
    p.elbo[0] = total - 0.5 * n * d * log_2pi + 0.5 * n * p.e_logdet - 0.5 * n * p.e_mu_l_mu;
  }
}

}  // namespace

int launch_p2p_allreduce(const double* const* bufs, uint32_t* const* flags, int rank, int world, int64_t count,
                         int64_t slot_stride, uint32_t epoch, double spin_limit_ms, double* out, int* status,
                         const double* e_lambda, const double* e_lambda_mu, double e_mu_l_mu, double e_logdet, int d,
                         double* elbo, cudaStream_t stream) {
  if (world < 1 || world > kP2PThreads || rank < 0 || rank >= world || count < 0 || slot_stride < count) {
    set_error("p2p_allreduce: bad rank/world/count (rank %d world %d count %lld)", rank, world,
              static_cast<long long>(count));
    return BB_ERR_INVALID;
  }
  if (elbo != nullptr && (d < 1 || count < static_cast<int64_t>(d) * d + d + 1)) {
    set_error("p2p_allreduce: the fused expected log-likelihood needs the packed layout [S2 | S1 | count]");
    return BB_ERR_SHAPE;
  }
  P2PParams p;
  p.bufs = bufs; p.flags = flags; p.rank = rank; p.world = world; p.count = count; p.slot_stride = slot_stride;
  p.epoch = epoch;
  p.spin_limit = static_cast<long long>(spin_limit_ms * 2.0e6);      // ~2 GHz ticks
  p.out = out; p.status = status; p.e_lambda = e_lambda; p.e_lambda_mu = e_lambda_mu; p.e_mu_l_mu = e_mu_l_mu;
  p.e_logdet = e_logdet; p.d = d; p.elbo = elbo;
  p2p_allreduce_kernel<<<1, kP2PThreads, 0, stream>>>(p);
  BB_CHECK_LAUNCH("p2p_allreduce_kernel");
  return BB_OK;
}

}  // namespace bb
