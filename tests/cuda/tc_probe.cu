// Standalone probe of tcgen05 semantics on sm_100a (developer tool, not part of the library):
// one CTA, one K=8 TF32 MMA, A from TMEM or SMEM, B in several SMEM layouts; dumps D.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_probe tc_probe.cu && ./tc_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../bayesic_b200/csrc/sm100_ptx.cuh"

using namespace bb;

// modes: bit0: A from SMEM (SS) instead of TMEM (TS); bit1: B K-major no-swizzle instead of
// MN-major SW128; bit2: B MN-major no-swizzle (interleave)
__global__ void probe_kernel(int mode, const float* __restrict__ Ah,  // [128][8]
                             const float* __restrict__ Bh,            // [8][64]  (k, n)
                             float* __restrict__ D, float* __restrict__ Aback) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sB = smem;               // 32 KB region
  uint8_t* sA = smem + 32768;       // 8 KB region
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int t = threadIdx.x, warp = t >> 5;
  for (int i = t; i < 40960 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.f;
  __syncthreads();
  const bool ss = mode & 1, b_kmajor = mode & 2, b_mn_noswz = mode & 4, b_mn_32b = mode & 8;
  // ---- B ----
  for (int idx = t; idx < 8 * 64; idx += blockDim.x) {
    const int k = idx / 64, n = idx % 64;
    uint32_t off;
    if (b_kmajor) {
      // K-major, no swizzle: core matrix 8 (n) x 16 B; LBO = 128 B between k-halves, SBO = 256 B
      off = (n % 8) * 16 + (k / 4) * 128 + (n / 8) * 256 + (k % 4) * 4;
    } else if (b_mn_32b) {
      // MN-major SWIZZLE_128B_BASE32B: row k = 128 B (32 n); 32B chunks XOR-swizzled by k%4;
      // 4-row groups 512 B apart (SBO); n >= 32 at +16 KB (LBO)
      const int h = n / 32, c32 = (n % 32) / 8;
      off = h * 16384 + k * 128 + (((c32 ^ (k & 3)) & 3) << 5) + (n % 8) * 4;
    } else if (b_mn_noswz) {
      // MN-major, no swizzle: ((T,1,m),(8,k)):((1,T,SBO),(1T,LBO)), T = 4 elements (16 B).
      // core matrix = 8 (k) rows x 16 B (4 n); n-blocks of 4 strided by SBO, k-groups by LBO
      off = (n % 4) * 4 + k * 16 + (n / 4) * 128;       // SBO = 128 B
    } else {
      // MN-major SW128: row k = 128 B holding 32 n; 16B chunks XOR-swizzled by k%8; n >= 32 at +16 KB
      const int h = n / 32, c16 = (n % 32) / 4;
      off = h * 16384 + k * 128 + (((c16 ^ (k & 7)) & 7) << 4) + (n % 4) * 4;
    }
    *reinterpret_cast<float*>(sB + off) = Bh[idx];
  }
  // ---- A in SMEM (for SS): K-major no swizzle, 128 rows (m) x 8 k
  if (ss) {
    for (int idx = t; idx < 128 * 8; idx += blockDim.x) {
      const int m = idx / 8, k = idx % 8;
      const uint32_t off = (m % 8) * 16 + (k / 4) * 128 + (m / 8) * 256 + (k % 4) * 4;
      *reinterpret_cast<float*>(sA + off) = Ah[idx];
    }
  }
  if (warp == 0) {
    if (t == 0) { ptx::mbar_init(&bar, 1); ptx::fence_mbar_init(); }
    __syncwarp();
    ptx::tmem_alloc(&tmem_slot, 512);
  }
  // make generic-proxy smem writes visible to the async (tensor core) proxy
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
  // A into TMEM columns 128..135
  {
    uint32_t v[8];
    for (int k = 0; k < 8; ++k) v[k] = __float_as_uint(Ah[t * 8 + k]);
    ptx::tmem_st_32x32b_x8(tmem + lane_base + 128, v);
    ptx::tmem_wait_st();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  if (t == 0) {
    uint64_t b_desc;
    uint32_t idesc;
    if (b_kmajor) {
      b_desc = ptx::make_smem_desc(ptx::smem_u32(sB), 128, 256, 0);
      idesc = ptx::make_idesc(128, 64, 2, 0, 0);
    } else if (b_mn_32b) {
      b_desc = ptx::make_smem_desc(ptx::smem_u32(sB), 16384, 512, 1);
      idesc = ptx::make_idesc(128, 64, 2, 0, 1);
    } else if (b_mn_noswz) {
      b_desc = ptx::make_smem_desc(ptx::smem_u32(sB), /*LBO (k-groups)*/ 2048, /*SBO (n-blocks)*/ 128, 0);
      idesc = ptx::make_idesc(128, 64, 2, 0, 1);
    } else {
      b_desc = ptx::make_smem_desc(ptx::smem_u32(sB), 16384, 1024, ptx::kLayoutSwizzle128B);
      idesc = ptx::make_idesc(128, 64, 2, 0, 1);
    }
    if (ss) {
      const uint64_t a_desc = ptx::make_smem_desc(ptx::smem_u32(sA), 128, 256, 0);
      ptx::mma_tf32_ss(tmem + 0, a_desc, b_desc, idesc, 0);
    } else {
      ptx::mma_tf32_ts(tmem + 0, tmem + 128, b_desc, idesc, 0);
    }
    ptx::mma_commit(&bar);
  }
  ptx::mbar_wait(&bar, 0);
  ptx::tc_fence_after_sync();
  for (int part = 0; part < 4; ++part) {
    uint32_t v[16];
    ptx::tmem_ld_32x32b_x16(tmem + lane_base + part * 16, v);
    ptx::tmem_wait_ld();
    for (int j = 0; j < 16; ++j) D[t * 64 + part * 16 + j] = __uint_as_float(v[j]);
  }
  {
    uint32_t v[16];
    ptx::tmem_ld_32x32b_x16(tmem + lane_base + 128, v);
    ptx::tmem_wait_ld();
    for (int j = 0; j < 8; ++j) Aback[t * 8 + j] = __uint_as_float(v[j]);
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 512);
}

int main() {
  float hA[128 * 8], hB[8 * 64], hD[128 * 64], hAb[128 * 8];
  for (int m = 0; m < 128; ++m)
    for (int k = 0; k < 8; ++k) hA[m * 8 + k] = float((m % 7) + k * 2 - 3);
  for (int k = 0; k < 8; ++k)
    for (int n = 0; n < 64; ++n) hB[k * 64 + n] = float((n % 5) - k + 1);
  float *dA, *dB, *dD, *dAb;
  cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dB, sizeof(hB)); cudaMalloc(&dD, sizeof(hD)); cudaMalloc(&dAb, sizeof(hAb));
  cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
  const char* names[] = {"TS  B=MN-major SW128", "SS  B=MN-major SW128", "TS  B=K-major noswz",
                         "SS  B=K-major noswz", "TS  B=MN-major noswz", "SS  B=MN-major noswz",
                         "TS  B=MN-major SW128_32B", "SS  B=MN-major SW128_32B"};
  const int modes[] = {0, 1, 2, 3, 4, 5, 8, 9};
  for (int c = 0; c < 8; ++c) {
    cudaMemset(dD, 0xff, sizeof(hD));
    cudaMemset(dAb, 0xff, sizeof(hAb));
    probe_kernel<<<1, 128, 44 * 1024>>>(modes[c], dA, dB, dD, dAb);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: CUDA error %s\n", names[c], cudaGetErrorString(e)); return 1; }
    cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost);
    cudaMemcpy(hAb, dAb, sizeof(hAb), cudaMemcpyDeviceToHost);
    double maxerr = 0, maxabs = 0, aerr = 0;
    int nonzero = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < 64; ++n) {
        double ref = 0;
        for (int k = 0; k < 8; ++k) ref += double(hA[m * 8 + k]) * hB[k * 64 + n];
        maxerr = fmax(maxerr, fabs(ref - hD[m * 64 + n]));
        maxabs = fmax(maxabs, fabs(ref));
        nonzero += hD[m * 64 + n] != 0.f;
      }
    for (int i = 0; i < 128 * 8; ++i) aerr = fmax(aerr, fabs(hA[i] - hAb[i]));
    printf("%-24s max|err| %.3g (max|ref| %.3g) nonzero %d/8192  A roundtrip err %.3g   D[0][0..3] = %g %g %g %g   D[70][5] = %g\n",
           names[c], maxerr, maxabs, nonzero, aerr, hD[0], hD[1], hD[2], hD[3], hD[70 * 64 + 5]);
  }
  return 0;
}
