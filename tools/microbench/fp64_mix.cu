// Does a non-FP64 instruction issue in the shadow of a DADD (16-lane pipe, 2 cycles per warp instruction)?
// K integer ops (independent LOP3/IADD chains) or K shared-memory loads per DADD, 12 warps/SM (3 per sub-partition).
#include <cstdio>
#include <cuda_runtime.h>

template <int KI, int KL>
__global__ void k(double* out, int iters, double b, unsigned m) {
  __shared__ double sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = i;
  __syncthreads();
  double x[8];
  unsigned u[8];
  double acc = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { x[i] = b + i + threadIdx.x; u[i] = threadIdx.x + i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      x[i] = __dadd_rn(x[i], b);
#pragma unroll
      for (int j = 0; j < KI; ++j) u[i] = (u[i] ^ m) + 0x9e3779b9u * (j + 1);
#pragma unroll
      for (int j = 0; j < KL; ++j) acc = __longlong_as_double(__double_as_longlong(acc) ^ __double_as_longlong(sm[(u[i] + it + 32 * j) & 1023]));
    }
  }
  double s = acc;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i] + u[i];
  if (s == 12345.678) out[0] = s;
}

template <int KI, int KL>
void run(const char* name) {
  int sms = 148, iters = 4096, threads = 384;
  double* d; cudaMalloc(&d, 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<KI, KL><<<sms, threads>>>(d, 16, 1.0000001, 5u);
  cudaEventRecord(e0);
  k<KI, KL><<<sms, threads>>>(d, iters, 1.0000001, 5u);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double dadd = (double)threads / 32 * iters * 8 / 4;   // DADD warp-instructions per sub-partition
  printf("%-28s %.3f ms  %.2f cycles per DADD per sub-partition\n", name, ms, ms * 1e-3 * clk * 1e3 / dadd);
  cudaFree(d);
}

int main() {
  run<0, 0>("DADD only");
  run<1, 0>("DADD + 2 int ops");      // xor + mad
  run<2, 0>("DADD + 4 int ops");
  run<4, 0>("DADD + 8 int ops");
  run<0, 1>("DADD + 1 LDS (+3 int)");
  return 0;
}
