// Developer probe: tcgen05.mma kind::f16 (BF16) with the A operand in TMEM (TS mode), M = 128,
// N = 64, K = 16.  Establishes the TMEM layout of a 16-bit A operand: lane = row m, 32-bit column j
// holds the K pair (2 j, 2 j + 1) -- which half is the even element is what `order` selects.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o ts_bf16_probe ts_bf16_probe.cu && ./ts_bf16_probe
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include "../../bayesic_b200/csrc/sm100_ptx.cuh"
using namespace bb;

__device__ __forceinline__ void mma_f16_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
               ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

__host__ __device__ inline float a_val(int m, int k) { return static_cast<float>((m * 3 + k * 5) % 7 - 3); }
__host__ __device__ inline float b_val(int n, int k) { return static_cast<float>((n + 2 * k) % 5 - 2); }

__global__ void probe(int order, int k_steps, float* __restrict__ D) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  // B[n][k], K-major SWIZZLE_128B: row n = 128 bytes = 64 k values, 16-byte chunks XOR (n & 7)
  for (int idx = t; idx < 64 * 64; idx += blockDim.x) {
    const int n = idx / 64, k = idx % 64;
    const uint32_t off = (n >> 3) * 1024 + (n & 7) * 128 + (((k >> 3) ^ (n & 7)) << 4) + (k & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(smem + off) = __float2bfloat16(k < 16 * k_steps ? b_val(n, k) : 0.f);
  }
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
  const uint32_t a_col = 256;
  {  // A rows of this warp's lane quadrant: 8 columns per K step of 16
    const int m = warp * 32 + lane;
    for (int ks = 0; ks < k_steps; ++ks) {
      uint32_t v[8];
      for (int j = 0; j < 8; ++j) {
        const uint32_t e = __bfloat16_as_ushort(__float2bfloat16(a_val(m, 16 * ks + 2 * j)));
        const uint32_t o = __bfloat16_as_ushort(__float2bfloat16(a_val(m, 16 * ks + 2 * j + 1)));
        v[j] = order == 0 ? (e | (o << 16)) : (o | (e << 16));
      }
      ptx::tmem_st_32x32b_x8(tmem + (static_cast<uint32_t>(warp * 32) << 16) + a_col + 8 * ks, v);
    }
    ptx::tmem_wait_st();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  if (t == 0) {
    const uint32_t idesc = ptx::make_idesc(128, 64, 1, 0, 0);
    for (int ks = 0; ks < k_steps; ++ks) {
      const uint64_t b = ptx::make_smem_desc(ptx::smem_u32(smem) + ks * 32, 16, 1024, ptx::kLayoutSwizzle128B);
      mma_f16_ts(tmem, tmem + a_col + 8 * ks, b, idesc, ks > 0 ? 1u : 0u);
    }
    ptx::mma_commit(&bar);
  }
  ptx::mbar_wait(&bar, 0);
  ptx::tc_fence_after_sync();
  for (int c = 0; c < 64; c += 16) {
    uint32_t v[16];
    ptx::tmem_ld_32x32b_x16(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c, v);
    ptx::tmem_wait_ld();
    for (int j = 0; j < 16; ++j) D[(warp * 32 + lane) * 64 + c + j] = __uint_as_float(v[j]);
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 512);
}

int main() {
  float* d; cudaMalloc(&d, 128 * 64 * 4);
  static float h[128 * 64];
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 1024);
  for (int k_steps = 1; k_steps <= 2; ++k_steps)
    for (int order = 0; order < 2; ++order) {
      probe<<<1, 128, 8 * 1024>>>(order, k_steps, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      double worst = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 64; ++n) {
          double ref = 0;
          for (int k = 0; k < 16 * k_steps; ++k) ref += a_val(m, k) * b_val(n, k);
          worst = fmax(worst, fabs(ref - h[m * 64 + n]));
        }
      printf("TS bf16 A-in-TMEM, k_steps=%d, pair order %s: max |D - ref| = %g  %s\n", k_steps,
             order == 0 ? "even element in the low half" : "even element in the high half", worst,
             worst == 0 ? "EXACT" : "mismatch");
    }
  return 0;
}
