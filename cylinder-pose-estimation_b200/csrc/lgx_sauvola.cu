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
//
// Two instantiations of the same arithmetic:
//   sauvola_tma_kernel  one warp per CTA = a 32-column strip of one frame.  Lane 0 streams the three planes with TMA
//                       (box 32 columns x 4 rows) into a 7-stage shared-memory ring, two stages ahead of the rows
//                       being summed; the ring is also the 15-row window of the column sums (5 live stages), so every
//                       plane value crosses the memory system once and the loads in flight do not cost registers.
//                       The kernel is HBM-bound (24 B/px read).
//   sauvola_kernel      one thread per column, direct global loads (fallback for planes that are not 16-byte aligned
//                       or narrower than a strip; LGX_OPT_SAUVOLA = 1 forces it for cross-checks).
#include "lgx_tma.cuh"

namespace lgx {
namespace {

constexpr int kThreads = 128;     // 256 / 512 threads, with or without block-wide pacing of the walk, measured slower

// kUnroll rows of loads are in flight per thread: 4 for launches that fill the GPU (more registers cost occupancy and
// DRAM locality there), 8 / 16 for launches of less than one wave, which are bound by the latency of the serial walk
// (2048 rows x one memory round trip per kUnroll rows).
template <int kUnroll, int kMinCtas>
__global__ void __launch_bounds__(kThreads, kMinCtas) sauvola_kernel(const SauvolaParams p) {
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

// ---- TMA ring instantiation --------------------------------------------------------------------------------------
constexpr int SV_RS = 4;                    // rows per stage
#ifndef LGX_SV_NST
#define LGX_SV_NST 7
#endif
constexpr int SV_NST = LGX_SV_NST;          // stages: rows y-7 .. y+7 of a 4-row group span 5 stages, SV_NST - 5 more are in flight
constexpr int SV_PLANE = SV_RS * 32 * 8;    // bytes of one plane tile of a stage
constexpr int SV_STAGE = 3 * SV_PLANE;      // rowsum(b) | rowsum(b*b) | b
constexpr int SV_SMEM = 128 + SV_NST * SV_STAGE + 128;

struct SvTmaParams {
  CUtensorMap tm_rb, tm_rq, tm_b;           // [frames][H][W] f64 (row pitch Wp), box 32 columns x SV_RS rows
  SauvolaParams sp;
};

__global__ void __launch_bounds__(32) sauvola_tma_kernel(const __grid_constant__ SvTmaParams p) {
  using namespace tma;
  extern __shared__ unsigned char sv_raw[];
  unsigned char* smem = sv_raw + ((128u - (smem_u32(sv_raw) & 127u)) & 127u);
  const uint32_t bar0 = smem_u32(smem);
  unsigned char* stages = smem + 128;
  const int lane = threadIdx.x;
  const int strip = blockIdx.x, frame = blockIdx.y;
  const int H = p.sp.H, W = p.sp.W, Wp = p.sp.Wp, WW = p.sp.WW;
  const int x = strip * 32 + lane;
  const bool active = x < W;
  const int nst = (H + SV_RS - 1) / SV_RS;
  uint8_t* __restrict__ obin = p.sp.binary ? p.sp.binary + (size_t)frame * H * W : nullptr;
  uint32_t* __restrict__ obits = p.sp.bits + (size_t)frame * H * WW;
  double* __restrict__ oT = p.sp.T ? p.sp.T + (size_t)frame * p.sp.plane_stride : nullptr;

  if (lane == 0) {
    for (int s = 0; s < SV_NST; ++s) mbar_init(bar0 + 8u * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();

  auto issue = [&](int st) {                // lane 0: rows 4*st .. 4*st+3 of the three planes (rows >= H arrive as zeros)
    const int slot = st % SV_NST;
    const uint32_t bar = bar0 + 8u * slot;
    const uint32_t dst = smem_u32(stages + slot * SV_STAGE);
    mbar_expect_tx(bar, SV_STAGE);
    tma_load_3d(&p.tm_rb, bar, dst, 32 * strip, st * SV_RS, frame);
    tma_load_3d(&p.tm_rq, bar, dst + SV_PLANE, 32 * strip, st * SV_RS, frame);
    tma_load_3d(&p.tm_b, bar, dst + 2 * SV_PLANE, 32 * strip, st * SV_RS, frame);
  };
  // this lane's rowsum(b) value of image row r (rowsum(b*b) at +SV_PLANE/8, b at +2*SV_PLANE/8 doubles)
  auto row_ptr = [&](int r) -> const double* {
    const int slot = (r >> 2) % SV_NST;
    return reinterpret_cast<const double*>(stages + slot * SV_STAGE + (r & 3) * 256) + lane;
  };
  constexpr int PQ = SV_PLANE / 8, PB = 2 * SV_PLANE / 8;

  int next_load = 0, next_wait = 0;
  for (; next_load < SV_NST - 2 && next_load < nst; ++next_load)
    if (lane == 0) issue(next_load);
  auto wait_upto = [&](int st_max) {
    for (; next_wait <= st_max; ++next_wait) mbar_wait(bar0 + 8u * (next_wait % SV_NST), (next_wait / SV_NST) & 1);
  };

  // SUM over the first 14 padded rows (padded row pr <-> image row clamp(pr-7)): rows 0..6 = stages 0, 1
  wait_upto(min(1, nst - 1));
  double sum_b = 0.0, sum_q = 0.0;
  for (int pr = 0; pr < 14; ++pr) {
    const double* q = row_ptr(min(max(pr - 7, 0), H - 1));
    sum_b = __dadd_rn(sum_b, q[0]);
    sum_q = __dadd_rn(sum_q, q[PQ]);
  }
  const double scale = 1.0 / 225;

  for (int j = 0; j < nst; ++j) {
    wait_upto(min(j + 2, nst - 1));         // rows up to 4j+10
    const int y0 = j * SV_RS;
    double nb[SV_RS], nq[SV_RS], ob[SV_RS], oq[SV_RS], bv[SV_RS];
#pragma unroll
    for (int u = 0; u < SV_RS; ++u) {
      const int y = min(y0 + u, H - 1);
      const double* qn = row_ptr(min(y + 7, H - 1));
      const double* qo = row_ptr(max(y - 7, 0));
      nb[u] = qn[0]; nq[u] = qn[PQ];
      ob[u] = qo[0]; oq[u] = qo[PQ];
      bv[u] = row_ptr(y)[PB];
    }
#pragma unroll
    for (int u = 0; u < SV_RS; ++u) {
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
      if (lane == 0) obits[(size_t)y * WW + strip] = word;
      if (active) {
        if (obin) obin[(size_t)y * W + x] = white ? 255 : 0;
        if (oT) oT[(size_t)y * Wp + x] = T;
      }
    }
    // stage j-2 is dead (stage j+1 reads rows >= 4j-3): refill its slot with stage j + SV_NST - 2
    __syncwarp();
    if (next_load < nst) {
      if (lane == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        issue(next_load);
      }
      ++next_load;
    }
  }
}

bool sauvola_tma_usable(const SauvolaParams& p) {
  auto a16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  return p.W >= 32 && p.H >= SV_RS && p.Wp % 2 == 0 && p.plane_stride % 2 == 0 && a16(p.b) && a16(p.rsb) && a16(p.rsb2) &&
         tma::encode_fn() != nullptr;
}

cudaError_t launch_sauvola_tma(const SauvolaParams& sp, int batch, cudaStream_t stream) {
  SvTmaParams p;
  p.sp = sp;
  const size_t prow = (size_t)sp.Wp * 8, pframe = sp.plane_stride * 8;
  bool ok = tma::encode3(&p.tm_rb, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, sp.rsb, sp.W, sp.H, batch, prow, pframe, 32, SV_RS, CU_TENSOR_MAP_SWIZZLE_NONE);
  ok = ok && tma::encode3(&p.tm_rq, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, sp.rsb2, sp.W, sp.H, batch, prow, pframe, 32, SV_RS, CU_TENSOR_MAP_SWIZZLE_NONE);
  ok = ok && tma::encode3(&p.tm_b, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, sp.b, sp.W, sp.H, batch, prow, pframe, 32, SV_RS, CU_TENSOR_MAP_SWIZZLE_NONE);
  if (!ok) return cudaErrorInvalidValue;
  static unsigned long long attr_done = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  if (!(attr_done >> (dev & 63) & 1ull)) {
    cudaError_t e = cudaFuncSetAttribute(sauvola_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SV_SMEM);
    if (e != cudaSuccess) return e;
    attr_done |= 1ull << (dev & 63);
  }
  dim3 grid((sp.W + 31) / 32, batch);
  sauvola_tma_kernel<<<grid, 32, SV_SMEM, stream>>>(p);
  return cudaGetLastError();
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

// bit plane -> dense u8 plane {0,255} (binary_img when the fused kernel produced only bits): thread = 4 pixels
__global__ void __launch_bounds__(256) unpack_bits_kernel(const uint32_t* __restrict__ bits, int H, int W, int WW, uint8_t* __restrict__ out) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;          // group of four pixels of a row
  const int y = blockIdx.y, frame = blockIdx.z;
  const int x = 4 * q;
  if (x >= W) return;
  const uint32_t wv = bits[((size_t)frame * H + y) * WW + (x >> 5)] >> (x & 31);
  uint8_t* o = out + ((size_t)frame * H + y) * W + x;
  const uint32_t packed = ((wv & 1u) ? 0xffu : 0u) | ((wv & 2u) ? 0xff00u : 0u) | ((wv & 4u) ? 0xff0000u : 0u) | ((wv & 8u) ? 0xff000000u : 0u);
  if (x + 3 < W && ((reinterpret_cast<uintptr_t>(o) & 3) == 0)) {
    *reinterpret_cast<uint32_t*>(o) = packed;
  } else {
    for (int e = 0; e < 4 && x + e < W; ++e) o[e] = (uint8_t)(packed >> (8 * e));
  }
}

}  // namespace

cudaError_t launch_unpack_bits(const uint32_t* bits, int batch, int H, int W, uint8_t* out, cudaStream_t stream) {
  dim3 grid(((W + 3) / 4 + 255) / 256, H, batch);
  unpack_bits_kernel<<<grid, 256, 0, stream>>>(bits, H, W, bits_pitch(W), out);
  return cudaGetLastError();
}

cudaError_t launch_sauvola(const SauvolaParams& p, int batch, int variant, cudaStream_t stream) {
  if (variant == 2 && sauvola_tma_usable(p)) return launch_sauvola_tma(p, batch, stream);
  dim3 grid((p.W + kThreads - 1) / kThreads, batch);
  const long long ctas = (long long)grid.x * grid.y;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (ctas <= 2LL * sms)          // 2 CTAs per SM at 230 registers
    sauvola_kernel<16, 2><<<grid, kThreads, 0, stream>>>(p);
  else if (ctas <= 3LL * sms)     // (measured: from ~4 CTAs per SM on, the 4-row instantiation is the faster one again)
    sauvola_kernel<8, 4><<<grid, kThreads, 0, stream>>>(p);
  else
    sauvola_kernel<4, 1><<<grid, kThreads, 0, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_pack_bits(const uint8_t* binary, int batch, int H, int W, uint32_t* bits, cudaStream_t stream) {
  dim3 grid((W + 127) / 128, H, batch);
  pack_bits_kernel<<<grid, 128, 0, stream>>>(binary, H, W, bits_pitch(W), bits);
  return cudaGetLastError();
}

}  // namespace lgx
