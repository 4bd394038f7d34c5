// Developer microbenchmark: cycles per tcgen05.mma kind::f16 (BF16, M = 128, K = 16) with BOTH
// operands in shared memory (SWIZZLE_128B, K-major), for N = 64 / 128 / 256, one or several
// accumulators in rotation, distinct or repeated operand tiles, and the A-collector hints.
// One CTA; the point is the per-SM issue/fetch/compute floor of the small-N MMAs the fused
// logistic pass is made of.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../bayesic_b200/csrc/sm100_ptx.cuh"
using namespace bb;

__device__ __forceinline__ void mma_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc, int mode) {
  if (mode == 1)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16.collector::a::fill [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else if (mode == 2)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// n_acc accumulators in rotation; a_tiles / b_tiles distinct operand tiles in rotation (1 = always the same bytes);
// collector: 0 none, 1 = pairs (fill, lastuse) sharing A
template <int N, int kAcc, int kATiles, int kBTiles, int kCollector>      // tile / accumulator counts: powers of two
__global__ void rate_kernel(int reps, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int t = threadIdx.x, warp = t >> 5;
  for (int i = t; i < 192 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (warp == 0) {
    if (t == 0) { ptx::mbar_init(&bar, 1); ptx::fence_mbar_init(); }
    __syncwarp();
    ptx::tmem_alloc(&tmem_slot, 512);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  if (t == 0) {
    const uint32_t sA = ptx::smem_u32(smem), sB = ptx::smem_u32(smem + 64 * 1024);
    const uint32_t idesc = ptx::make_idesc(128, N, 1, 0, 0);
    const long long t0 = clock64();
    for (int r = 0; r < reps; r += 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = r + u;
        // A tile: 128 rows x 128 B = 16 KB (one 64-wide K chunk; K step = 32 B inside it); B tile: N rows x 128 B
        const uint32_t a_addr = sA + ((kCollector ? (u >> 1) : u) & (kATiles - 1)) * 16384 + (u & 3) * 32;
        const uint32_t b_addr = sB + (u & (kBTiles - 1)) * (N * 128) + (u & 3) * 32;
        const uint64_t a = ptx::make_smem_desc(a_addr, 16, 1024, ptx::kLayoutSwizzle128B);
        const uint64_t b = ptx::make_smem_desc(b_addr, 16, 1024, ptx::kLayoutSwizzle128B);
        mma_f16(tmem + (u & (kAcc - 1)) * N, a, b, idesc, i >= kAcc ? 1u : 0u, kCollector ? 1 + (u & 1) : 0);
      }
    }
    ptx::mma_commit(&bar);
    ptx::mbar_wait(&bar, 0);
    cycles[0] = clock64() - t0;
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 512);
}

template <int N, int kAcc, int kATiles, int kBTiles, int kCollector>
void run(long long* d, int reps) {
  auto kernel = rate_kernel<N, kAcc, kATiles, kBTiles, kCollector>;
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  long long h = 0;
  for (int it = 0; it < 2; ++it) {
    kernel<<<1, 128, 192 * 1024>>>(reps, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return; }
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  }
  printf("SS bf16 M=128 N=%3d K=16  acc=%d a_tiles=%d b_tiles=%d collector=%d: %6.1f cycles/MMA  (math floor %3d, operand bytes %5d)\n",
         N, kAcc, kATiles, kBTiles, kCollector, double(h) / reps, N / 2, 4096 + N * 32);
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  const int reps = 8192;
  run<64, 1, 4, 2, 0>(d, reps);
  run<64, 4, 4, 2, 0>(d, reps);
  run<64, 1, 1, 1, 0>(d, reps);
  run<64, 1, 4, 2, 1>(d, reps);
  run<64, 4, 4, 2, 1>(d, reps);
  run<128, 1, 4, 2, 0>(d, reps);
  run<128, 2, 4, 2, 0>(d, reps);
  run<128, 1, 4, 2, 1>(d, reps);
  run<256, 1, 4, 2, 0>(d, reps);
  run<256, 2, 4, 2, 0>(d, reps);
  run<256, 1, 4, 2, 1>(d, reps);
  return 0;
}
