// Debug / parity entry points that exercise device helpers on their own (include/lgx.h: lgx_debug_sqrt).
#include "lgx_internal.cuh"
#include "lgx_sqrt.cuh"

namespace lgx {
namespace {

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
  x += 0x9e3779b97f4a7c15ull;
  x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
  return x ^ (x >> 31);
}

// mode 0: uniformly random bit patterns of non-negative finite doubles (every exponent, incl. subnormals and 0-ish)
// mode 1: random mantissa, exponent uniform in [2^-120, 2^8): the magnitudes the ridge / Sauvola radicands have
// mode 2: explicit list `extra`
__device__ __forceinline__ double radicand(unsigned long long seed, unsigned long long idx, int mode, const double* extra,
                                           unsigned long long n_extra) {
  if (mode == 2) return extra[idx < n_extra ? idx : n_extra - 1];
  const unsigned long long r = splitmix64(seed + idx);
  if (mode == 0) {
    unsigned long long b = r & 0x7fffffffffffffffull;
    if ((b >> 52) == 0x7ffull) b &= 0x7fefffffffffffffull;      // inf / nan -> a large finite value
    return __longlong_as_double((long long)b);
  }
  const unsigned long long e = 1023ull - 120ull + (r >> 52) % 128ull;
  return __longlong_as_double((long long)((e << 52) | (r & 0x000fffffffffffffull)));
}

// out[0] = values whose bits differ from __dsqrt_rn (in range), out[1] = values flagged out of range,
// out[2] = out-of-range values that were NOT flagged (must be 0: range-check soundness), out[3] = bits of the first
// mismatching radicand
__global__ void __launch_bounds__(256) sqrt_check_kernel(unsigned long long seed, unsigned long long n, int mode,
                                                         const double* __restrict__ extra, unsigned long long* __restrict__ out) {
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x * 8ull;
  unsigned long long bad = 0, flagged = 0, unsound = 0;
  for (unsigned long long i0 = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) * 8ull; i0 < n; i0 += stride) {
    double X[8], R[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) X[i] = radicand(seed, i0 + i < n ? i0 + i : n - 1, mode, extra, n);
    unsigned worst = 0;
    sqrt_inrange<8>(X, R, worst);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i0 + i >= n) continue;
      unsigned w1 = 0;
      double x1[1] = {X[i]}, r1[1];
      sqrt_inrange<1>(x1, r1, w1);
      const bool in_range = X[i] >= 0x1p-970 && X[i] <= 1.7976931348623157e308;
      if (w1 >= kSqrtOutOfRange) {
        ++flagged;
      } else {
        if (!in_range) ++unsound;
        const double ref = __dsqrt_rn(X[i]);
        if (__double_as_longlong(ref) != __double_as_longlong(R[i]) || __double_as_longlong(ref) != __double_as_longlong(r1[0])) {
          if (bad == 0) atomicCAS(&out[3], 0ull, (unsigned long long)__double_as_longlong(X[i]) | (1ull << 63));
          ++bad;
        }
      }
    }
    (void)worst;
  }
  if (bad) atomicAdd(&out[0], bad);
  if (flagged) atomicAdd(&out[1], flagged);
  if (unsound) atomicAdd(&out[2], unsound);
}

}  // namespace
}  // namespace lgx

extern "C" int lgx_debug_sqrt(unsigned long long seed, unsigned long long n, int mode, const double* extra_host,
                              unsigned long long* out4) {
  if (!out4 || n == 0 || mode < 0 || mode > 2 || (mode == 2 && !extra_host)) return LGX_ERR_BAD_ARG;
  unsigned long long* d_out = nullptr;
  double* d_extra = nullptr;
  if (cudaMalloc((void**)&d_out, 4 * sizeof(unsigned long long)) != cudaSuccess) return LGX_ERR_CUDA;
  cudaMemset(d_out, 0, 4 * sizeof(unsigned long long));
  if (mode == 2) {
    if (cudaMalloc((void**)&d_extra, n * sizeof(double)) != cudaSuccess) { cudaFree(d_out); return LGX_ERR_CUDA; }
    cudaMemcpy(d_extra, extra_host, n * sizeof(double), cudaMemcpyHostToDevice);
  }
  const unsigned long long threads = (n + 7) / 8;
  unsigned long long blocks = (threads + 255) / 256;
  if (blocks > 148ull * 32ull) blocks = 148ull * 32ull;
  lgx::sqrt_check_kernel<<<(unsigned)blocks, 256>>>(seed, n, mode, d_extra, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaMemcpy(out4, d_out, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  cudaFree(d_out);
  if (d_extra) cudaFree(d_extra);
  return e == cudaSuccess ? LGX_OK : LGX_ERR_CUDA;
}

extern "C" int lgx_debug_fused_prof(unsigned long long* out32, int reset) {
  if (!out32) return LGX_ERR_BAD_ARG;
  lgx::fused_prof_read(out32, reset != 0);
  return cudaGetLastError() == cudaSuccess ? LGX_OK : LGX_ERR_CUDA;
}
