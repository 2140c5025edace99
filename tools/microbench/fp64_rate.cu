// FP64 issue-rate calibration on B200: lane-ops per clock per SM for DADD / DMUL / DFMA / mixed, vs resident warps.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE, int ILP>
__global__ void k(double* out, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = a + i + threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (MODE == 0) x[i] = __dadd_rn(x[i], b);
      else if (MODE == 1) x[i] = __dmul_rn(x[i], b);
      else if (MODE == 2) x[i] = __fma_rn(x[i], b, a);
      else { x[i] = __dadd_rn(x[i], b); x[i] = __dmul_rn(x[i], a); }   // 2 dependent ops
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  if (s == 12345.678) out[0] = s;
}

template <int MODE, int ILP>
void run(const char* name, int threads, int ctas_per_sm) {
  int sms = 148, iters = 4096;
  double* d; cudaMalloc(&d, 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE, ILP><<<sms * ctas_per_sm, threads>>>(d, 16, 1.0, 1.0000001);
  cudaEventRecord(e0);
  k<MODE, ILP><<<sms * ctas_per_sm, threads>>>(d, iters, 1.0, 1.0000001);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double ops = (double)sms * ctas_per_sm * threads * iters * ILP * (MODE == 3 ? 2 : 1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%-10s ILP=%d warps/SM=%2d  %.3f ms  %.1f lane-ops/clk/SM (at %.0f MHz nominal)\n", name, ILP, threads / 32 * ctas_per_sm, ms,
         ops / (ms * 1e-3) / (clk * 1e3) / sms, clk / 1e3);
  cudaFree(d);
}

int main() {
  run<0, 8>("DADD", 256, 2); run<1, 8>("DMUL", 256, 2); run<2, 8>("DFMA", 256, 2); run<3, 8>("DADD+DMUL", 256, 2);
  run<0, 8>("DADD", 256, 8); run<2, 8>("DFMA", 256, 8); run<3, 8>("DADD+DMUL", 256, 8);
  run<0, 2>("DADD", 256, 2); run<0, 4>("DADD", 256, 2); run<3, 4>("DADD+DMUL", 256, 2);
  run<0, 1>("DADD", 128, 1); run<0, 1>("DADD", 32, 1);
  return 0;
}
