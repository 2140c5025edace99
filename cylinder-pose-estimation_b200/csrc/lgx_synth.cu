// Synthetic frames for benches and GPU tests: a noise-free float32 scene (rendered once on the host by
// synth.render_base) plus per-frame Gaussian sensor noise generated on the device, so
// that multi-GB batches never cross PCIe.  Counter-based (splitmix64 of seed and pixel index) so a frame
// is a pure function of (base, seed); frames are copied back to the host when the CPU oracle needs them.
#include "lgx_internal.cuh"

namespace lgx {
namespace {

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

template <typename PIX>
__global__ void render_kernel(const float* __restrict__ base, int n_base, size_t npix, float sigma, uint64_t seed0,
                              float scale, float vmax, PIX* __restrict__ out) {
  const int f = blockIdx.y;
  const float* b = base + (size_t)(f % n_base) * npix;
  PIX* o = out + (size_t)f * npix;
  const uint64_t key = splitmix64(seed0 + (uint64_t)f);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x) {
    uint64_t r = splitmix64(key ^ (i * 0xD6E8FEB86659FD93ull));
    float u1 = ((uint32_t)(r >> 40) + 1u) * (1.0f / 16777216.0f);   // (0,1]
    float u2 = (uint32_t)(r & 0xffffffu) * (1.0f / 16777216.0f);
    float n = sqrtf(-2.0f * __logf(u1)) * __cosf(6.28318530718f * u2);
    float v = rintf((b[i] + sigma * n) * scale);
    o[i] = (PIX)fminf(fmaxf(v, 0.0f), vmax);
  }
}

}  // namespace
}  // namespace lgx

extern "C" int lgx_render_noisy(const float* d_base, int n_base, int batch, int height, int width, float sigma,
                                uint64_t seed0, int bits, void* d_out, void* stream) {
  if (!d_base || !d_out || n_base < 1 || batch < 0 || height < 1 || width < 1 || (bits != 8 && bits != 16))
    return LGX_ERR_BAD_ARG;
  if (batch == 0) return LGX_OK;
  size_t npix = (size_t)height * width;
  dim3 grid(148 * 8, batch);
  cudaStream_t s = (cudaStream_t)stream;
  if (bits == 8)
    lgx::render_kernel<uint8_t><<<grid, 256, 0, s>>>(d_base, n_base, npix, sigma, seed0, 1.0f, 255.0f, (uint8_t*)d_out);
  else
    lgx::render_kernel<uint16_t><<<grid, 256, 0, s>>>(d_base, n_base, npix, sigma, seed0, 257.0f, 65535.0f, (uint16_t*)d_out);
  return cudaGetLastError() == cudaSuccess ? LGX_OK : LGX_ERR_CUDA;
}
