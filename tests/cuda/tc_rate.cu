// Developer microbenchmark: cycles per tcgen05.mma (kind::tf32, M=128) for several N, with A from
// TMEM (TS) or SMEM (SS) and B MN-major (SW128_32B atom) or K-major (no swizzle).  One CTA.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../bayesic_b200/csrc/sm100_ptx.cuh"
using namespace bb;

template <int N_ACC, int N_COLS>
__global__ void rate_kernel(int ts, int b_mn, int reps, long long* cycles) {
  constexpr int n_cols = N_COLS;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int t = threadIdx.x, warp = t >> 5;
  for (int i = t; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 1.0f;
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
  {
    uint32_t v[8];
    for (int k = 0; k < 8; ++k) v[k] = __float_as_uint(1.0f);
    ptx::tmem_st_32x32b_x8(tmem + (static_cast<uint32_t>(warp * 32) << 16) + 256, v);
    ptx::tmem_wait_st();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  if (t == 0) {
    const uint32_t sB = ptx::smem_u32(smem), sA = ptx::smem_u32(smem + 128 * 1024);
    uint64_t b_desc; uint32_t idesc;
    if (b_mn) { b_desc = ptx::make_smem_desc(sB, 16384, 512, 1); idesc = ptx::make_idesc(128, n_cols, 2, 0, 1); }
    else { b_desc = ptx::make_smem_desc(sB, 128, 256, 0); idesc = ptx::make_idesc(128, n_cols, 2, 0, 0); }
    const uint64_t a_desc = ptx::make_smem_desc(sA, 128, 256, 0);
    const long long t0 = clock64();
    if (ts) {
      for (int r = 0; r < reps; r += 16) {
#pragma unroll
        for (int u = 0; u < 16; ++u)      // rotate over N_ACC independent accumulators
          ptx::mma_tf32_ts(tmem + (u % N_ACC) * N_COLS, tmem + 256, b_desc, idesc, (r > 0 || u >= N_ACC) ? 1u : 0u);
      }
    } else {
      for (int r = 0; r < reps; r += 16) {
#pragma unroll
        for (int u = 0; u < 16; ++u)
          ptx::mma_tf32_ss(tmem + (u % N_ACC) * N_COLS, a_desc, b_desc, idesc, (r > 0 || u >= N_ACC) ? 1u : 0u);
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

int main() {
  long long* d; cudaMalloc(&d, 8);
  const int reps = 4096;
  auto run = [&](auto kernel, int n_acc, int n, int ts) {
        const int b_mn = 1;
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        long long h = 0;
        for (int it = 0; it < 2; ++it) {
          kernel<<<1, 128, 192 * 1024>>>(ts, b_mn, reps, d);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return; }
          cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        }
        printf("%s  accumulators=%d M=128 N=%3d K=8 tf32: %.1f cycles/MMA  -> %.0f MAC/clk/SM\n", ts ? "TS" : "SS",
               n_acc, n, double(h) / reps, 128.0 * n * 8 * reps / double(h));
  };
  for (int ts = 1; ts >= 0; --ts) {
    run(rate_kernel<1, 64>, 1, 64, ts);
    run(rate_kernel<2, 64>, 2, 64, ts);
    run(rate_kernel<4, 64>, 4, 64, ts);
    run(rate_kernel<1, 128>, 1, 128, ts);
    run(rate_kernel<2, 128>, 2, 128, ts);
    run(rate_kernel<1, 256>, 1, 256, ts);
  }
  return 0;
}
