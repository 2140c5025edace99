// K2 "sauvola": vertical running sums of cv2.boxFilter (ColumnSum<double,double>), the Sauvola
// threshold and the compare, for both planes (b and b*b) at once.
//
// Replaces, bit for bit: the column pass of cv2.boxFilter at /root/reference/utils/util_cylinder.py:1755-1757,
// the NumPy expression at :1760-1765 and the threshold/invert at :1798-1800.  CPU twin:
// oracle/restate.py col_sums15 / sauvola_T / binarize.
//
// OpenCV's column sum is a serial chain from the top row (SUM += new; out = SUM*scale; SUM -= old),
// so one thread owns one image column and walks down; a warp covers 32 adjacent columns, which makes
// every f64 load a full 256-byte row segment and lets __ballot_sync emit the bit-packed binary row.
#include "lgx_internal.cuh"

namespace lgx {
namespace {

constexpr int kThreads = 128;
constexpr int kUnroll = 4;

__global__ void __launch_bounds__(kThreads) sauvola_kernel(const SauvolaParams p) {
  const int x = blockIdx.x * kThreads + threadIdx.x;
  const int frame = blockIdx.y;
  const int H = p.H, W = p.W, Wp = p.Wp;
  const bool active = x < W;
  const int xc = active ? x : W - 1;
  const size_t base = (size_t)frame * p.plane_stride + xc;
  const double* __restrict__ rb = p.rsb + base;
  const double* __restrict__ rq = p.rsb2 + base;
  const double* __restrict__ bp = p.b + base;
  uint8_t* __restrict__ obin = p.binary ? p.binary + (size_t)frame * H * W : nullptr;
  uint32_t* __restrict__ obits = p.bits + (size_t)frame * H * p.WW;
  double* __restrict__ oT = p.T ? p.T + (size_t)frame * p.plane_stride : nullptr;
  const int lane = threadIdx.x & 31;

  // SUM over the first 14 padded rows (padded row pr <-> image row clamp(pr-7))
  double sum_b = 0.0, sum_q = 0.0;
  for (int pr = 0; pr < 14; ++pr) {
    int r = min(max(pr - 7, 0), H - 1);
    sum_b = __dadd_rn(sum_b, rb[(size_t)r * Wp]);
    sum_q = __dadd_rn(sum_q, rq[(size_t)r * Wp]);
  }
  const double scale = 1.0 / 225;

  for (int y0 = 0; y0 < H; y0 += kUnroll) {
    double nb[kUnroll], nq[kUnroll], ob[kUnroll], oq[kUnroll], bv[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      int y = min(y0 + u, H - 1);
      int yn = min(y + 7, H - 1), yo = max(y - 7, 0);
      nb[u] = rb[(size_t)yn * Wp];
      nq[u] = rq[(size_t)yn * Wp];
      ob[u] = rb[(size_t)yo * Wp];
      oq[u] = rq[(size_t)yo * Wp];
      bv[u] = bp[(size_t)y * Wp];
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int y = y0 + u;
      if (y >= H) break;
      double s0 = __dadd_rn(sum_b, nb[u]);
      double m = __dmul_rn(s0, scale);
      sum_b = __dsub_rn(s0, ob[u]);
      double q0 = __dadd_rn(sum_q, nq[u]);
      double msq = __dmul_rn(q0, scale);
      sum_q = __dsub_rn(q0, oq[u]);
      double var = __dsub_rn(msq, __dmul_rn(m, m));
      if (var < 0) var = 0.0;
      double sd = __dsqrt_rn(var);
      // mean * (1 + k * ((std / R) - 1)),  k = 0.5, R = 128
      double T = __dmul_rn(m, __dadd_rn(1.0, __dmul_rn(0.5, __dsub_rn(__dmul_rn(sd, 0.0078125), 1.0))));
      const bool white = active && !(bv[u] > T);   // binary = 255 - 255*(b > T)
      const unsigned word = __ballot_sync(0xffffffffu, white);
      if (lane == 0 && x < W) obits[(size_t)y * p.WW + (x >> 5)] = word;
      if (active) {
        if (obin) obin[(size_t)y * W + x] = white ? 255 : 0;
        if (oT) oT[(size_t)y * Wp + x] = T;
      }
    }
  }
}

// binary u8 -> bit plane (entry for extract_joints on an arbitrary binary image)
__global__ void pack_bits_kernel(const uint8_t* __restrict__ binary, int H, int W, int WW, uint32_t* __restrict__ bits) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int frame = blockIdx.z;
  const bool on = (x < W) && binary[((size_t)frame * H + y) * W + x] != 0;
  const unsigned word = __ballot_sync(0xffffffffu, on);
  if ((threadIdx.x & 31) == 0 && x < W) bits[((size_t)frame * H + y) * WW + (x >> 5)] = word;
}

}  // namespace

cudaError_t launch_sauvola(const SauvolaParams& p, int batch, cudaStream_t stream) {
  dim3 grid((p.W + kThreads - 1) / kThreads, batch);
  sauvola_kernel<<<grid, kThreads, 0, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_pack_bits(const uint8_t* binary, int batch, int H, int W, uint32_t* bits, cudaStream_t stream) {
  dim3 grid((W + 127) / 128, H, batch);
  pack_bits_kernel<<<grid, 128, 0, stream>>>(binary, H, W, bits_pitch(W), bits);
  return cudaGetLastError();
}

}  // namespace lgx
