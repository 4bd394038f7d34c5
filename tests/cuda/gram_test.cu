// Developer test + timing of the cta_group::2 Gram kernel (gram_sm100.cu) without Python:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o gram_test gram_test.cu \
//        ../../bayesic_b200/csrc/gram_sm100.cu ../../bayesic_b200/csrc/runtime.cu
//   ./gram_test [n] [d] [reps]
// Structured inputs first (so a wrong quadrant / layout shows up as a pattern), then random.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../../bayesic_b200/csrc/kernels.h"

namespace bb {
bool gram_tc_supported(int64_t n, int d, const void* x);
int64_t gram_tc_workspace(int64_t n, int d);
int launch_gram_tc(const float* x, const float* y, int64_t n, int d, double* xtx, double* xty,
                   double* yty, void* workspace, int64_t workspace_bytes, cudaStream_t stream);
}  // namespace bb

#define CK(e)                                                                       \
  do {                                                                              \
    cudaError_t _e = (e);                                                           \
    if (_e != cudaSuccess) {                                                        \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return 2;                                                                     \
    }                                                                               \
  } while (0)

static double urand(uint64_t& s) {
  s = s * 6364136223846793005ULL + 1442695040888963407ULL;
  return ((s >> 11) & ((1ULL << 53) - 1)) / double(1ULL << 53);
}
static double nrand(uint64_t& s) {
  const double u = urand(s) + 1e-300, v = urand(s);
  return sqrt(-2.0 * log(u)) * cos(6.283185307179586 * v);
}

int run_case(int64_t n, int d, int mode, bool check, int reps) {
  std::vector<float> hx(size_t(n) * d), hy(n);
  uint64_t seed = 1234 + n + d;
  for (int64_t i = 0; i < n; ++i) {
    for (int j = 0; j < d; ++j) {
      float v;
      if (mode == 0) v = (i == (j % n)) ? float(1 + j % 7) : 0.f;          // sparse pattern
      else if (mode == 1) v = float((i * 3 + j * 5) % 11) - 5.f;              // small integers (exact in bf16)
      else v = float(nrand(seed) * 1.3 + 0.2);
      hx[size_t(i) * d + j] = v;
    }
    hy[i] = mode == 2 ? float(nrand(seed)) : float(i % 5) - 2.f;
  }
  float *dx, *dy;
  double *dxtx, *dxty, *dyty;
  void* ws;
  const int64_t wsb = bb::gram_tc_workspace(n, d);
  CK(cudaMalloc(&dx, hx.size() * 4));
  CK(cudaMalloc(&dy, hy.size() * 4));
  CK(cudaMalloc(&dxtx, size_t(d) * d * 8));
  CK(cudaMalloc(&dxty, size_t(d) * 8));
  CK(cudaMalloc(&dyty, 8));
  CK(cudaMalloc(&ws, wsb));
  CK(cudaMemcpy(dx, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dy, hy.data(), hy.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dxtx, 0xff, size_t(d) * d * 8));
  int st = bb::launch_gram_tc(dx, dy, n, d, dxtx, dxty, dyty, ws, wsb, 0);
  if (st != 0) {
    printf("launch failed: %d %s\n", st, bb::get_error());
    return 2;
  }
  CK(cudaDeviceSynchronize());
  int bad = 0;
  if (check) {
    std::vector<double> g(size_t(d) * d), gy(d);
    double gyy = 0;
    CK(cudaMemcpy(g.data(), dxtx, g.size() * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(gy.data(), dxty, gy.size() * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&gyy, dyty, 8, cudaMemcpyDeviceToHost));
    // reference on a subset of entries (all entries when small)
    const int step = (int64_t(d) * d * n > (int64_t(1) << 31)) ? 37 : 1;
    double maxrel = 0, maxabs_ref = 0, maxerr = 0;
    int printed = 0;
    // natural scale of entry (a, b): sqrt(S_aa S_bb) (Cauchy-Schwarz bound of |S_ab|)
    std::vector<double> diag(d, 0.0);
    for (int64_t i = 0; i < n; ++i)
      for (int a = 0; a < d; ++a) diag[a] += double(hx[size_t(i) * d + a]) * hx[size_t(i) * d + a];
    for (int a = 0; a < d; a += step)
      for (int b = 0; b < d; b += (step == 1 ? 1 : 29)) {
        double ref = 0;
        for (int64_t i = 0; i < n; ++i) ref += double(hx[size_t(i) * d + a]) * hx[size_t(i) * d + b];
        const double got = g[size_t(a) * d + b];
        const double err = fabs(got - ref);
        maxerr = fmax(maxerr, err);
        maxabs_ref = fmax(maxabs_ref, fabs(ref));
        const double rel = err / fmax(sqrt(diag[a] * diag[b]), 1e-30);
        maxrel = fmax(maxrel, rel);
        if (rel > 3e-5 && printed < 12) {
          printf("   mismatch [%d,%d] got %.9g ref %.9g\n", a, b, got, ref);
          ++printed;
          ++bad;
        }
      }
    double yerr = 0, yref_max = 0;
    for (int a = 0; a < d; ++a) {
      double ref = 0;
      for (int64_t i = 0; i < n; ++i) ref += double(hx[size_t(i) * d + a]) * hy[i];
      yerr = fmax(yerr, fabs(gy[a] - ref));
      yref_max = fmax(yref_max, fabs(ref));
    }
    double yy = 0;
    for (int64_t i = 0; i < n; ++i) yy += double(hy[i]) * hy[i];
    printf("n=%lld d=%d mode=%d: XtX max|err| %.3g (max|ref| %.3g) max err/sqrt(Saa Sbb) %.3g | Xty max|err| %.3g (max|ref| %.3g) | yty %.9g ref %.9g  %s\n",
           (long long)n, d, mode, maxerr, maxabs_ref, maxrel, yerr, yref_max, gyy, yy, bad ? "FAIL" : "ok");
    if (yerr > 1e-5 * fmax(yref_max, 1.0) || fabs(gyy - yy) > 1e-6 * fmax(yy, 1.0)) {
      printf("   Xty / yty mismatch\n");
      ++bad;
    }
  }
  if (reps > 0) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) bb::launch_gram_tc(dx, dy, n, d, dxtx, dxty, dyty, ws, wsb, 0);
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) bb::launch_gram_tc(dx, dy, n, d, dxtx, dxty, dyty, ws, wsb, 0);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= reps;
    const double flops_sym = double(d) * (d + 1) * n;            // algorithmic (symmetric half)
    const double blocks = double(d / 256) * (d / 256 + 1) / 2;
    const double flops_issued = 3.0 * 2.0 * 256 * 256 * blocks * n;   // three bf16 MMAs per block
    printf("n=%lld d=%d: %.3f ms  %.1f M rows/s  useful %.1f TFLOP/s (sym)  issued bf16 %.1f TFLOP/s  HBM %.1f GB/s\n",
           (long long)n, d, ms, n / ms / 1e3, flops_sym / ms / 1e9, flops_issued / ms / 1e9, double(n) * d * 4 / ms / 1e6);
  }
  cudaFree(dx); cudaFree(dy); cudaFree(dxtx); cudaFree(dxty); cudaFree(dyty); cudaFree(ws);
  return bad ? 1 : 0;
}

int main(int argc, char** argv) {
  int fails = 0;
  if (argc >= 3) {
    const int64_t n = atoll(argv[1]);
    const int d = atoi(argv[2]);
    const int reps = argc >= 4 ? atoi(argv[3]) : 5;
    return run_case(n, d, 2, n * int64_t(d) * d <= (int64_t(1) << 36), reps);
  }
  fails += run_case(32, 256, 0, true, 0);
  fails += run_case(32, 256, 1, true, 0);
  fails += run_case(64, 256, 1, true, 0);
  fails += run_case(100, 512, 1, true, 0);
  fails += run_case(1000, 256, 2, true, 0);
  fails += run_case(5000, 512, 2, true, 0);
  fails += run_case(70001, 256, 2, true, 0);
  fails += run_case(150000, 1024, 2, true, 0);
  printf(fails ? "FAILED %d cases\n" : "all cases ok\n", fails);
  if (!fails) run_case(1 << 20, 1024, 2, false, 10);
  return fails ? 1 : 0;
}
