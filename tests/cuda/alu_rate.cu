// Developer microbenchmark: issue rate per SM of the instructions the BF16x3 converters are made of -- F2FP
// (cvt.rn.bf16x2.f32), FSUB, LOP3/SHF, and the whole 4-element split -- with W warps per SM sub-partition and
// independent chains, so that the answer is throughput, not latency.  nvcc -arch=sm_100a -O3 alu_rate.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

template <int MODE>
__global__ void rate_kernel(int reps, float seed, long long* cycles, float* sink) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = seed * (threadIdx.x + 1 + i);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
      if (MODE == 0) {                 // F2FP only: two floats -> packed bf16x2
        __nv_bfloat162 p = __floats2bfloat162_rn(a[i], a[i + 1]);
        const uint32_t w = *reinterpret_cast<uint32_t*>(&p);
        acc ^= w;
        a[i] = __uint_as_float(w | 0x3f000000u);          // keep a dependency so nothing is hoisted
      } else if (MODE == 1) {          // FSUB + LOP only (no conversion)
        a[i] = a[i] - __uint_as_float(__float_as_uint(a[i + 1]) & 0xFFFF0000u);
        a[i + 1] = a[i + 1] - __uint_as_float(__float_as_uint(a[i]) << 16);
      } else {                         // the converters' split of a pair: hi = RN, lo = RN(x - hi)
        __nv_bfloat162 p = __floats2bfloat162_rn(a[i], a[i + 1]);
        const uint32_t w = *reinterpret_cast<uint32_t*>(&p);
        const float r0 = a[i] - __uint_as_float(w << 16);
        const float r1 = a[i + 1] - __uint_as_float(w & 0xFFFF0000u);
        __nv_bfloat162 q = __floats2bfloat162_rn(r0, r1);
        const uint32_t v = *reinterpret_cast<uint32_t*>(&q);
        acc ^= w ^ v;
        a[i] = r0 + 1.5f;
        a[i + 1] = r1 + 2.5f;
      }
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s + __uint_as_float(acc);
}

template <int MODE>
void run(const char* what, int per_iter_pairs, int threads) {
  long long* cyc;
  float* sink;
  cudaMalloc(&cyc, 148 * 8);
  cudaMalloc(&sink, 148 * 1024 * 4);
  const int reps = 20000;
  rate_kernel<MODE><<<148, threads>>>(reps, 1.0001f, cyc, sink);
  rate_kernel<MODE><<<148, threads>>>(reps, 1.0001f, cyc, sink);
  cudaDeviceSynchronize();
  long long h;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double pairs = double(reps) * per_iter_pairs * (threads / 32);        // warp-level "pair" operations per SM
  printf("%-44s %4d threads/SM: %.2f cycles per warp-pair-op per SM  (%lld cycles)\n", what, threads, h / pairs, h);
  cudaFree(cyc);
  cudaFree(sink);
}

int main() {
  for (int threads : {128, 256, 512, 1024}) {
    run<0>("F2FP.BF16 (2 floats -> bf16x2)", 4, threads);
    run<1>("2 x (LOP + FSUB)", 4, threads);
    run<2>("full split of a pair (2 F2FP, 2 LOP/SHF, 2 FSUB)", 4, threads);
  }
  return 0;
}
