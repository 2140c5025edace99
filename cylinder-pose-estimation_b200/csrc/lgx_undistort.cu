// N3 "undistort": the per-frame part of cv2.undistort (/root/reference/utils/iotool.py:22-39, cv2.undistort at :38;
// callers python_grid_detection_cylinder.py:36-41 and iotool.py:62-65) on 8-bit images with 1 or 3 interleaved
// channels.
//
// cv2.undistort = initUndistortRectifyMap(CV_16SC2) + remap(INTER_LINEAR, BORDER_CONSTANT 0).  The maps depend on the
// camera only: the host computes them once per camera with OpenCV's own initUndistortRectifyMap, stripe by stripe as
// cv2.undistort does (iotool.py of this package), so they are the reference's maps by construction.  This kernel
// is the remap: for each destination pixel, (sx, sy) = map_xy, (fx, fy) = the two 5-bit fractions of map_frac, and
//     dst = (S00*(32-fy)*(32-fx)*32 + S01*(32-fy)*fx*32 + S10*fy*(32-fx)*32 + S11*fy*fx*32 + 2^14) >> 15
// with source samples outside the image = 0: OpenCV's fixed-point bilinear table (INTER_BITS = 5,
// INTER_REMAP_COEF_BITS = 15; the products are exact so its table needs no normalisation fix-up).  CPU twin:
// oracle/restate.py remap_bilinear_fixed, checked against cv2.remap and cv2.undistort.
//
// HBM-bound gather: per pixel 6 B of map (read once per camera: the frame is the fastest block index), 1 B/channel read through L1
// (neighbouring destinations read neighbouring sources), 1 B/channel written.  Thread = 4 adjacent destination pixels
// (one 128-bit map_xy load, one 64-bit map_frac load, one packed store per channel group).
#include "lgx_internal.cuh"

namespace lgx {
namespace {

template <int CN>
__global__ void __launch_bounds__(128) undistort_kernel(const uint8_t* __restrict__ src, size_t pitch, size_t fstride, int H, int W,
                                                        const short2* __restrict__ map_xy, const uint16_t* __restrict__ map_frac,
                                                        const int32_t* __restrict__ cam_index, uint8_t* __restrict__ dst) {
  // the frame is the fastest block index: CTAs that run together work on the same map tile (one DRAM read per
  // camera, L2 hits for every other frame) and each on its own frame's rows
  const int x0 = (blockIdx.z * 128 + threadIdx.x) * 4;
  const int y = blockIdx.y, f = blockIdx.x;
  if (x0 >= W) return;
  const int cam = cam_index ? cam_index[f] : 0;
  const size_t mrow = ((size_t)cam * H + y) * W + x0;
  const uint8_t* __restrict__ S = src + (size_t)f * fstride;
  uint8_t* __restrict__ D = dst + (((size_t)f * H + y) * W + x0) * CN;
  short2 xy[4];
  uint16_t fr[4];
  const int n = min(4, W - x0);
  if (n == 4 && (reinterpret_cast<uintptr_t>(map_xy + mrow) & 15) == 0 && (reinterpret_cast<uintptr_t>(map_frac + mrow) & 7) == 0) {
    const int4 q = *reinterpret_cast<const int4*>(map_xy + mrow);
    const uint2 r = *reinterpret_cast<const uint2*>(map_frac + mrow);
    xy[0] = *reinterpret_cast<const short2*>(&q.x); xy[1] = *reinterpret_cast<const short2*>(&q.y);
    xy[2] = *reinterpret_cast<const short2*>(&q.z); xy[3] = *reinterpret_cast<const short2*>(&q.w);
    fr[0] = (uint16_t)(r.x & 0xffffu); fr[1] = (uint16_t)(r.x >> 16); fr[2] = (uint16_t)(r.y & 0xffffu); fr[3] = (uint16_t)(r.y >> 16);
  } else {
    for (int i = 0; i < 4; ++i) {
      xy[i] = i < n ? map_xy[mrow + i] : make_short2(-2, -2);
      fr[i] = i < n ? map_frac[mrow + i] : (uint16_t)0;
    }
  }
  uint8_t out[4 * CN];
  const unsigned pitch32 = (unsigned)pitch;            // frames are < 4 GB: 32-bit offsets from the frame base
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int sx = xy[i].x, sy = xy[i].y;
    // all four weights carry the factor 32 of OpenCV's table, so (sum*32 + 2^14) >> 15 == (sum + 2^9) >> 10 with
    // sum = (s00*(32-fx) + s01*fx)*(32-fy) + (s10*(32-fx) + s11*fx)*fy  (exact integer identities)
    const int fx = fr[i] & 31, fy = (fr[i] >> 5) & 31;
    const int gx = 32 - fx, gy = 32 - fy;
    const uint8_t* p = S + ((unsigned)sy * pitch32 + (unsigned)sx * CN);
    if ((unsigned)sx < (unsigned)(W - 1) && (unsigned)sy < (unsigned)(H - 1)) {
      const uint8_t* q = p + pitch32;
#pragma unroll
      for (int c = 0; c < CN; ++c) {
        const int top = p[c] * gx + p[CN + c] * fx, bot = q[c] * gx + q[CN + c] * fx;
        out[i * CN + c] = (uint8_t)((top * gy + bot * fy + (1 << 9)) >> 10);
      }
    } else {
      const bool x0in = (unsigned)sx < (unsigned)W, x1in = (unsigned)(sx + 1) < (unsigned)W;
      const bool y0in = (unsigned)sy < (unsigned)H, y1in = (unsigned)(sy + 1) < (unsigned)H;
      const uint8_t* pb = S + (ptrdiff_t)sy * (ptrdiff_t)pitch + (ptrdiff_t)sx * CN;
#pragma unroll
      for (int c = 0; c < CN; ++c) {
        const int s00 = (x0in && y0in) ? pb[c] : 0, s01 = (x1in && y0in) ? pb[CN + c] : 0;
        const int s10 = (x0in && y1in) ? pb[pitch + c] : 0, s11 = (x1in && y1in) ? pb[pitch + CN + c] : 0;
        out[i * CN + c] = (uint8_t)(((s00 * gx + s01 * fx) * gy + (s10 * gx + s11 * fx) * fy + (1 << 9)) >> 10);
      }
    }
  }
  if (n == 4 && (reinterpret_cast<uintptr_t>(D) & 3) == 0) {
#pragma unroll
    for (int k = 0; k < CN; ++k)
      reinterpret_cast<uint32_t*>(D)[k] = (uint32_t)out[4 * k] | ((uint32_t)out[4 * k + 1] << 8) | ((uint32_t)out[4 * k + 2] << 16) |
                                          ((uint32_t)out[4 * k + 3] << 24);
  } else {
    for (int k = 0; k < n * CN; ++k) D[k] = out[k];
  }
}

}  // namespace

cudaError_t launch_undistort(const uint8_t* src, int channels, int batch, int H, int W, size_t pitch, size_t fstride,
                             const int16_t* map_xy, const uint16_t* map_frac, const int32_t* cam_index, uint8_t* dst,
                             cudaStream_t stream) {
  dim3 grid(batch, H, (W + 511) / 512);
  const short2* mxy = reinterpret_cast<const short2*>(map_xy);
  if (channels == 1)
    undistort_kernel<1><<<grid, 128, 0, stream>>>(src, pitch, fstride, H, W, mxy, map_frac, cam_index, dst);
  else
    undistort_kernel<3><<<grid, 128, 0, stream>>>(src, pitch, fstride, H, W, mxy, map_frac, cam_index, dst);
  return cudaGetLastError();
}

}  // namespace lgx
