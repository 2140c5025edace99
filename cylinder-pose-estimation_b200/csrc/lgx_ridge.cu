// K0 "blur5" and K1 "ridge".
//
// K0: cv2.GaussianBlur(gray,(5,5),0) on u8/u16 — exact integer [1 4 6 4 1]^2, (acc+128)>>8, BORDER_REFLECT_101
//     (/root/reference/utils/util_cylinder.py:1790; CPU twin oracle/restate.py blur5).
// K1: img_as_float LUT -> 25-tap gaussian (rows first, then columns; scipy NI_Correlate1D symmetric order)
//     -> np.gradient x4 -> smaller Hessian eigenvalue b -> horizontal running sums of cv2.boxFilter
//     (RowSum) for b and b*b.  Replaces, bit for bit, util_cylinder.py:1793 (detect_ridges, :1734-1738) and
//     the row pass of the two cv2.boxFilter calls at :1755-1757.  Operation order: SURVEY.md App. A items
//     3-7, restated on the CPU in oracle/restate.py (gauss25, min_eigenvalue, row_sums15).
//
// K1 shape: one CTA owns a horizontal band of <= 60 image rows of one frame and sweeps it left to right in
// 32-column steps.  The sweep is what makes OpenCV's whole-row running sum (a serial chain from x = 0)
// fusable: each band row has one thread that carries its running sum in a register across the sweep; it
// also means the vertical gaussian is never recomputed for a column halo.  The kernel is FP64-pipe bound
// (~125 f64 instr / pixel, FMA contraction is not allowed), so everything else is kept off the issue
// slots: the blurred image comes from K0 as 32-bit words prefetched one step ahead into registers, all
// shared-memory planes are linear (compile-time offsets), and lanes are mapped to columns for column-wise
// phases and to rows (odd pitches) for row-wise phases so that every 64-bit access is conflict free.
#include "lgx_internal.cuh"

namespace lgx {

// gaussian taps: RidgeParams::w (kernel parameter = constant bank, per launch and therefore per handle).
// w[j], j = 0..11: weight of the pair (l-12+j, l+12-j); w[12]: centre

namespace {

__device__ __forceinline__ int reflect101(int p, int n) {
  // cv2 borderInterpolate(BORDER_REFLECT_101)
  if (n == 1) return 0;
  while (p < 0 || p >= n) p = (p < 0) ? -p : 2 * (n - 1) - p;
  return p;
}

// ---------------------------------------------------------------------------------------------------
// K0: tiled 5x5 blur.  128 x 32 output tile, 256 threads: thread = (column, 16-row half), vertical sliding
// window of horizontal 5-tap sums.
// ---------------------------------------------------------------------------------------------------
constexpr int BT_W = 128, BT_H = 32;

template <typename PIX>
__global__ void __launch_bounds__(256) blur5_kernel(const void* __restrict__ frames, size_t pitch_bytes, size_t fstride,
                                                    int H, int W, PIX* __restrict__ out_pad, int pad_pitch,
                                                    PIX* __restrict__ out_dense) {
  __shared__ PIX s_t[BT_H + 4][BT_W + 4 + 4];
  const int x0 = blockIdx.x * BT_W, y0 = blockIdx.y * BT_H, f = blockIdx.z;
  const PIX* __restrict__ src = reinterpret_cast<const PIX*>(reinterpret_cast<const unsigned char*>(frames) + (size_t)f * fstride);
  const size_t sp = pitch_bytes / sizeof(PIX);
  const int tid = threadIdx.x;
  const bool interior = x0 >= 2 && y0 >= 2 && x0 + BT_W + 2 <= W && y0 + BT_H + 2 <= H;
  for (int idx = tid; idx < (BT_H + 4) * (BT_W + 4); idx += 256) {
    int r = idx / (BT_W + 4), j = idx - r * (BT_W + 4);
    int y = y0 - 2 + r, x = x0 - 2 + j;
    PIX v = 0;
    if (interior) {
      v = src[(size_t)y * sp + x];
    } else if (y < H + 2 && x < W + 2) {
      v = src[(size_t)reflect101(y, H) * sp + reflect101(x, W)];
    }
    s_t[r][j] = v;
  }
  __syncthreads();
  const int c = tid & (BT_W - 1);
  const int rbase = (tid >> 7) * 16;
  const int x = x0 + c;
  int h0 = 0, h1 = 0, h2 = 0, h3 = 0, h4 = 0;
#pragma unroll
  for (int rr = 0; rr < 20; ++rr) {
    const PIX* row = &s_t[rbase + rr][c];
    int hs = (int)row[0] + (int)row[4] + 4 * ((int)row[1] + (int)row[3]) + 6 * (int)row[2];
    h0 = h1; h1 = h2; h2 = h3; h3 = h4; h4 = hs;
    if (rr >= 4) {
      const int y = y0 + rbase + rr - 4;
      if (y < H && x < W) {
        PIX bl = (PIX)((h0 + h4 + 4 * (h1 + h3) + 6 * h2 + 128) >> 8);
        if (out_pad) out_pad[((size_t)f * H + y) * pad_pitch + x] = bl;
        if (out_dense) out_dense[((size_t)f * H + y) * W + x] = bl;
      }
    }
  }
}

// u8 specialisation: 256 x 64 tile as 32-bit words in shared memory (row = bytes x0-16 .. x0+271, filled with
// aligned 128-bit loads wherever the 16 bytes lie inside the image row, byte-wise reflect-101 at the edges only),
// 128 threads, thread = 4 adjacent columns x 32 rows (36 horizontal passes for 32 outputs), two pixels per
// register in 16-bit lanes, byte pairs extracted with PRMT, packed 32-bit stores.  ~11 instructions per pixel.
constexpr int B8_W = 256, B8_H = 64, B8_ROWW = B8_W / 4 + 8;   // 72 words per tile row

__global__ void __launch_bounds__(128) blur5_u8_kernel(const uint8_t* __restrict__ frames, size_t pitch, size_t fstride,
                                                       int H, int W, uint8_t* __restrict__ out_pad, int pad_pitch,
                                                       uint8_t* __restrict__ out_dense) {
  __shared__ __align__(16) uint32_t s_w[B8_H + 4][B8_ROWW];
  const int x0 = blockIdx.x * B8_W, y0 = blockIdx.y * B8_H, f = blockIdx.z;
  const uint8_t* __restrict__ src = frames + (size_t)f * fstride;
  const int tid = threadIdx.x;
  const bool aligned = (pitch & 15) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0;
  constexpr int Q = B8_ROWW / 4;                       // 18 quads per tile row
  for (int qi = tid; qi < (B8_H + 4) * Q; qi += 128) {
    const int r = qi / Q, k = qi - r * Q;
    const int y = y0 - 2 + r, xb = x0 - 16 + 16 * k;   // first byte of the quad
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (y < H + 2) {
      const uint8_t* row = src + (size_t)reflect101(y, H) * pitch;
      if (aligned && xb >= 0 && xb + 16 <= W) {
        v = *reinterpret_cast<const uint4*>(row + xb);
      } else {
        // a quad that straddles an image edge (or an unaligned image): only the bytes x in [-2, W+2) are ever read
        uint32_t w4[4] = {0u, 0u, 0u, 0u};
        for (int e = max(0, -2 - xb); e < min(16, W + 2 - xb); ++e)
          w4[e >> 2] |= (uint32_t)row[reflect101(xb + e, W)] << (8 * (e & 3));
        v = make_uint4(w4[0], w4[1], w4[2], w4[3]);
      }
    }
    *reinterpret_cast<uint4*>(&s_w[r][4 * k]) = v;
  }
  __syncthreads();
  const int cg = tid & 63, r0 = (tid >> 6) * 32;
  const int x = x0 + 4 * cg;
  if (x >= W) return;
  // Two pixels per 32-bit register in 16-bit lanes: lane sums stay below 2^16 (16*255 horizontally, 16*4080+128
  // vertically).  A = columns (0,2), B = columns (1,3) of the thread's four.
  constexpr uint32_t M = 0x00ff00ffu;
  constexpr uint32_t ODD = 0x4341u;                    // PRMT selector: bytes (1, -, 3, -) of a word, zero between
  uint32_t hA[5], hB[5];
  const bool full = x + 3 < W;
  uint8_t* dp = out_pad ? out_pad + ((size_t)f * H + y0 + r0) * pad_pitch + x : nullptr;
  uint8_t* dd = out_dense ? out_dense + ((size_t)f * H + y0 + r0) * W + x : nullptr;
  const bool dd_word = (reinterpret_cast<uintptr_t>(dd) & 3) == 0 && (W & 3) == 0;
#pragma unroll
  for (int rr = 0; rr < 36; ++rr) {
    const uint32_t w0 = s_w[r0 + rr][cg + 3], w1 = s_w[r0 + rr][cg + 4], w2 = s_w[r0 + rr][cg + 5];
    const uint32_t lo = __funnelshift_r(w0, w1, 16);    // pixels x-2 .. x+1
    const uint32_t hi = __funnelshift_r(w1, w2, 16);    // pixels x+2 .. x+5
    const uint32_t y0_ = lo & M;                          // (p0,p2)
    const uint32_t y1_ = __byte_perm(lo, 0u, ODD);        // (p1,p3)
    const uint32_t y2_ = w1 & M;                          // (p2,p4)
    const uint32_t y3_ = __byte_perm(w1, 0u, ODD);        // (p3,p5)
    const uint32_t y4_ = hi & M;                          // (p4,p6)
    const uint32_t y5_ = __byte_perm(hi, 0u, ODD);        // (p5,p7)
    const uint32_t sA = y0_ + y4_ + ((y1_ + y3_) << 2) + 6u * y2_;
    const uint32_t sB = y1_ + y5_ + ((y2_ + y4_) << 2) + 6u * y3_;
#pragma unroll
    for (int i = 0; i < 4; ++i) { hA[i] = hA[i + 1]; hB[i] = hB[i + 1]; }
    hA[4] = sA; hB[4] = sB;
    if (rr >= 4) {
      const int y = y0 + r0 + rr - 4;
      if (y < H) {
        const uint32_t vA = __byte_perm(hA[0] + hA[4] + 0x00800080u + ((hA[1] + hA[3]) << 2) + 6u * hA[2], 0u, ODD);
        const uint32_t vB = __byte_perm(hB[0] + hB[4] + 0x00800080u + ((hB[1] + hB[3]) << 2) + 6u * hB[2], 0u, ODD);
        const uint32_t packed = __byte_perm(vA, vB, 0x6240u);   // bytes vA.0, vB.0, vA.2, vB.2
        if (full) {
          if (dp) *reinterpret_cast<uint32_t*>(dp) = packed;
          if (dd) {
            if (dd_word) *reinterpret_cast<uint32_t*>(dd) = packed;
            else { dd[0] = packed; dd[1] = packed >> 8; dd[2] = packed >> 16; dd[3] = packed >> 24; }
          }
        } else {
          for (int k = 0; k < 4 && x + k < W; ++k) {
            if (dp) dp[k] = (uint8_t)(packed >> (8 * k));
            if (dd) dd[k] = (uint8_t)(packed >> (8 * k));
          }
        }
      }
      if (dp) dp += pad_pitch;
      if (dd) dd += W;
    }
  }
}

// u16 specialisation: 128 x 64 tile, 256 threads, thread = 4 columns x 8 rows, two pixels per 32-bit word, 32-bit lane arithmetic
// (16 * 65535 horizontally, 256 * 65535 + 128 vertically both fit).
constexpr int B16_W = 128, B16_H = 64;
constexpr int B16_WORDS = B16_W / 2 + 2;   // tile columns x0-2 .. x0+129 as 66 words

__global__ void __launch_bounds__(256) blur5_u16_kernel(const uint16_t* __restrict__ frames, size_t pitch_px, size_t fstride_bytes,
                                                        int H, int W, uint16_t* __restrict__ out_pad, int pad_pitch,
                                                        uint16_t* __restrict__ out_dense) {
  __shared__ uint32_t s_w[B16_H + 4][B16_WORDS + 1];
  const int x0 = blockIdx.x * B16_W, y0 = blockIdx.y * B16_H, f = blockIdx.z;
  const uint16_t* __restrict__ src = reinterpret_cast<const uint16_t*>(reinterpret_cast<const unsigned char*>(frames) + (size_t)f * fstride_bytes);
  const int tid = threadIdx.x;
  const bool fast = x0 >= 2 && y0 >= 2 && x0 + B16_W + 2 <= W && y0 + B16_H + 2 <= H && (pitch_px & 1) == 0 &&
                    (reinterpret_cast<uintptr_t>(src) & 3) == 0;
  for (int wi = tid; wi < (B16_H + 4) * B16_WORDS; wi += 256) {
    const int r = wi / B16_WORDS, j = wi - r * B16_WORDS;
    const int y = y0 - 2 + r, xw = x0 - 2 + 2 * j;
    uint32_t v = 0;
    if (fast) {
      v = *reinterpret_cast<const uint32_t*>(src + (size_t)y * pitch_px + xw);
    } else if (y < H + 2) {
      const uint16_t* row = src + (size_t)reflect101(y, H) * pitch_px;
      if (xw >= -2 && xw < W + 2) v = row[reflect101(xw, W)];
      if (xw + 1 >= -2 && xw + 1 < W + 2) v |= (uint32_t)row[reflect101(xw + 1, W)] << 16;
    }
    s_w[r][j] = v;
  }
  __syncthreads();
  const int cg = tid & 31, r0 = (tid >> 5) * 8;
  const int x = x0 + 4 * cg;
  if (x >= W) return;
  uint32_t h[4][5];
#pragma unroll
  for (int rr = 0; rr < 12; ++rr) {
    uint32_t p[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t wv = s_w[r0 + rr][2 * cg + k];      // pixels x-2+2k, x-1+2k
      p[2 * k] = wv & 0xffffu;
      p[2 * k + 1] = wv >> 16;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t hs = p[k] + p[k + 4] + 4u * (p[k + 1] + p[k + 3]) + 6u * p[k + 2];
      h[k][0] = h[k][1]; h[k][1] = h[k][2]; h[k][2] = h[k][3]; h[k][3] = h[k][4]; h[k][4] = hs;
    }
    if (rr >= 4) {
      const int y = y0 + r0 + rr - 4;
      if (y < H) {
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) o[k] = (h[k][0] + h[k][4] + 4u * (h[k][1] + h[k][3]) + 6u * h[k][2] + 128u) >> 8;
        const size_t row = (size_t)f * H + y;
        if (x + 3 < W) {
          if (out_pad) *reinterpret_cast<uint2*>(out_pad + row * pad_pitch + x) = make_uint2(o[0] | (o[1] << 16), o[2] | (o[3] << 16));
          if (out_dense) {
            uint16_t* d = out_dense + row * W + x;
            if ((reinterpret_cast<uintptr_t>(d) & 7) == 0) *reinterpret_cast<uint2*>(d) = make_uint2(o[0] | (o[1] << 16), o[2] | (o[3] << 16));
            else { d[0] = o[0]; d[1] = o[1]; d[2] = o[2]; d[3] = o[3]; }
          }
        } else {
          for (int k = 0; k < 4 && x + k < W; ++k) {
            if (out_pad) out_pad[row * pad_pitch + x + k] = (uint16_t)o[k];
            if (out_dense) out_dense[row * W + x + k] = (uint16_t)o[k];
          }
        }
      }
    }
  }
}

// cv2.cvtColor(BGR2GRAY), 15-bit fixed point (util_cylinder.py:1789 for a true-colour input; identity for R=G=B)
template <typename PIX>
__global__ void bgr2gray_kernel(const PIX* __restrict__ bgr, size_t npix, PIX* __restrict__ gray) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x) {
    unsigned b = bgr[3 * i], g = bgr[3 * i + 1], r = bgr[3 * i + 2];
    gray[i] = (PIX)((3735u * b + 19235u * g + 9798u * r + 16384u) >> 15);
  }
}

// ---------------------------------------------------------------------------------------------------
// K1
// ---------------------------------------------------------------------------------------------------
// Geometry of a band for NW warps per CTA: GROWS = 8*NW rows of the gaussian image, BROWS = GROWS-4 rows of b.
// NW = 8: 64/60 rows, 256 threads, 2 CTAs/SM.  NW = 4: 32/28 rows, 128 threads, 4 CTAs/SM (more independent phase
// streams per SM at the price of 7 % more gaussian halo work).
constexpr int V_PITCH = 65;                        // vertical-pass plane: ring of two 32-column slots (x & 63)
constexpr int G_HIST = 6, G_PITCH = 39;            // gaussian plane:      6 history + 32 new columns
constexpr int B_HIST = 16, B_PITCH = 49;           // eigenvalue plane:   16 history + 32 new columns

template <int NW>
struct Geo {
  static constexpr int kThreads = 32 * NW;
  static constexpr int kG = 8 * NW;                // g rows
  static constexpr int kB = kG - 4;                // b rows
  static constexpr int kF = kG + 2 * kRadius;      // f rows
  static constexpr size_t kSmem = (size_t)(kF * kChunk + kG * V_PITCH + kG * G_PITCH + kB * B_PITCH + 256) * sizeof(double);
};

__device__ __forceinline__ double min_eig(double Hrr, double Hrc, double Hcc) {
  // (M00 + M11)/2 - sqrt(4*M01**2 + (M00 - M11)**2)/2     (skimage _image_orthogonal_matrix22_eigvals)
  double s = __dadd_rn(Hrr, Hcc);
  double d = __dsub_rn(Hrr, Hcc);
  double r = __dsqrt_rn(__dadd_rn(__dmul_rn(4.0, __dmul_rn(Hrc, Hrc)), __dmul_rn(d, d)));
  return __dsub_rn(__dmul_rn(s, 0.5), __dmul_rn(r, 0.5));
}

// np.gradient-of-np.gradient at (y, x) with every border rule, reading g from the band plane.
// gy0: image row of plane row 0; gx0: image column of plane column 0.
__device__ __noinline__ double b_generic(const double* __restrict__ s_g, int gy0, int gx0, int y, int x, int H, int W,
                                         int mixed) {
  auto G = [&](int yy, int xx) { return s_g[(yy - gy0) * G_PITCH + (xx - gx0)]; };
  auto sc = [](int i, int n) { return (i > 0 && i < n - 1) ? 0.5 : 1.0; };
  auto Dr = [&](int yy, int xx) {
    int yp = min(yy + 1, H - 1), ym = max(yy - 1, 0);
    return __dmul_rn(__dsub_rn(G(yp, xx), G(ym, xx)), sc(yy, H));
  };
  auto Dc = [&](int yy, int xx) {
    int xp = min(xx + 1, W - 1), xm = max(xx - 1, 0);
    return __dmul_rn(__dsub_rn(G(yy, xp), G(yy, xm)), sc(xx, W));
  };
  int yp = min(y + 1, H - 1), ym = max(y - 1, 0);
  int xp = min(x + 1, W - 1), xm = max(x - 1, 0);
  double sr = sc(y, H), scx = sc(x, W);
  double Hrr = __dmul_rn(__dsub_rn(Dr(yp, x), Dr(ym, x)), sr);
  double Hcc = __dmul_rn(__dsub_rn(Dc(y, xp), Dc(y, xm)), scx);
  double Hrc = mixed ? __dmul_rn(__dsub_rn(Dc(yp, x), Dc(ym, x)), sr)
                     : __dmul_rn(__dsub_rn(Dr(y, xp), Dr(y, xm)), scx);
  return min_eig(Hrr, Hrc, Hcc);
}

template <typename PIX>
struct Px;
template <>
struct Px<uint8_t> { static constexpr int kPerWord = 4; };
template <>
struct Px<uint16_t> { static constexpr int kPerWord = 2; };

template <typename PIX, int NW>
__global__ void __launch_bounds__(32 * NW, 16 / NW) ridge_kernel(const RidgeParams p) {
  using G = Geo<NW>;
  constexpr int T = G::kThreads, GR = G::kG, BR = G::kB, FR = G::kF;
  extern __shared__ __align__(16) unsigned char smem[];
  double* s_f = reinterpret_cast<double*>(smem);
  double* s_v = s_f + FR * kChunk;
  double* s_g = s_v + GR * V_PITCH;
  double* s_b = s_g + GR * G_PITCH;
  double* s_lut = s_b + BR * B_PITCH;

  constexpr int PPW = Px<PIX>::kPerWord;          // pixels per 32-bit word
  constexpr int WPR = kChunk / PPW;               // words per tile row
  constexpr int NWORDS = (FR * WPR + T - 1) / T;  // words per thread

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int H = p.H, W = p.W, Wp = p.Wp;
  const int band = blockIdx.x;
  const int frame = blockIdx.y;
  const int y0 = band * p.rows_per_band;           // first b row of the band
  const int nrows = min(p.rows_per_band, H - y0);  // b rows of the band
  const int yg0 = y0 - 2;                          // image row of g row 0
  const int yf0 = yg0 - kRadius;                   // image row of f row 0

  const uint32_t* __restrict__ blur =
      reinterpret_cast<const uint32_t*>(reinterpret_cast<const PIX*>(p.blur) + (size_t)frame * H * p.blur_pitch);
  const int blur_pitch_w = p.blur_pitch / PPW;     // words per row of the padded blurred plane
  double* __restrict__ out_b = p.b + (size_t)frame * p.plane_stride;
  double* __restrict__ out_g = p.g ? p.g + (size_t)frame * p.plane_stride : nullptr;

  for (int i = tid; i < GR * V_PITCH; i += T) s_v[i] = 0.0;  // columns x < 0 are zero padding
  if (sizeof(PIX) == 1)
    for (int i = tid; i < 256; i += T) s_lut[i] = p.lut[i];

  // f-tile word ownership: word q of this thread <-> tile row wi / WPR, word-in-row wi % WPR (step invariant)
  uint32_t pre[NWORDS];
  int f_off[NWORDS];        // word offset of (row, word-in-row) in the blurred plane, -1 if the row is outside the image
  int f_dst[NWORDS];        // f64 index in the f tile, -1 if this thread has no q-th word
#pragma unroll
  for (int q = 0; q < NWORDS; ++q) {
    const int wi = tid + q * T;
    const int r = wi / WPR, wq = wi - r * WPR;
    const int y = yf0 + r;
    f_dst[q] = (wi < FR * WPR) ? r * kChunk + wq * PPW : -1;
    f_off[q] = (wi < FR * WPR && y >= 0 && y < H) ? y * blur_pitch_w + wq : -1;
  }
  auto prefetch = [&](int x0) {
#pragma unroll
    for (int q = 0; q < NWORDS; ++q) {
      pre[q] = 0;
      const int wq_x = x0 + ((f_dst[q] & (kChunk - 1)));      // first pixel of the word
      if (f_off[q] >= 0 && wq_x < W) pre[q] = __ldg(blur + f_off[q] + x0 / PPW);
    }
  };
  // prefetched words -> LUT -> f tile (zero outside the image)
  auto fill_f = [&](int x0) {
    const bool full = x0 + kChunk <= W;      // no per-pixel column check needed
#pragma unroll
    for (int q = 0; q < NWORDS; ++q) {
      if (f_dst[q] >= 0) {
        double2* dst = reinterpret_cast<double2*>(s_f + f_dst[q]);   // 128-bit stores: conflict free
        const bool rowok = f_off[q] >= 0;
        const int xw = x0 + (f_dst[q] & (kChunk - 1));
        double fv[PPW];
#pragma unroll
        for (int e = 0; e < PPW; ++e) {
          const uint32_t v = (sizeof(PIX) == 1) ? ((pre[q] >> (8 * e)) & 0xffu) : ((pre[q] >> (16 * e)) & 0xffffu);
          fv[e] = 0.0;
          // u8: 256-entry table in shared memory; u16: skimage.img_as_float's own operation, v * (1/65535) (or the IEEE
          // division with LGX_OPT_FLOAT_DIV) instead of a 64 K-entry gather (32 sectors per warp instruction)
          if (rowok && (full || xw + e < W)) fv[e] = (sizeof(PIX) == 1) ? s_lut[v] : (p.float_div ? __ddiv_rn((double)v, 65535.0) : __dmul_rn((double)v, 1.0 / 65535.0));
        }
#pragma unroll
        for (int e = 0; e < PPW; e += 2) dst[e >> 1] = make_double2(fv[e], fv[e + 1]);
      }
    }
  };

  const int nchunks = (W > 8 ? (W - 8 + kChunk - 1) / kChunk : 0) + 1;
  prefetch(0);
  __syncthreads();     // LUT visible
  fill_f(0);
  if (nchunks > 1) prefetch(kChunk);

  double chain_b = 0.0, chain_q = 0.0;  // running row sums of b and b*b (chain warps, lane <-> band row)
  // optional phase clock (debug option LGX_OPT_RIDGE_PROF): cycles thread 0 spends between barriers
  long long tprof[5] = {0, 0, 0, 0, 0};
  long long tlast = p.prof ? clock64() : 0;
#define LGX_TICK(i)                                  \
  if (p.prof && tid == 0) {                          \
    long long tnow_ = clock64();                     \
    tprof[i] += tnow_ - tlast;                       \
    tlast = tnow_;                                   \
  }

  // row / segment of this thread in the row-wise phases (S3, S4): 32 rows per warp, 8-column segments
  constexpr int RW = GR / 32;                      // warps per 32-row slab... GR = 32 or 64
  const int rrow = (warp % RW) * 32 + lane;        // g row (S3) / b row (S4)
  const int rseg = warp / RW;                      // 0..3

  for (int k = 0; k < nchunks; ++k) {
    const int x0 = k * kChunk;
    const int slot = (k & 1) * 32;                 // ring slot written by this step
    __syncthreads();  // f tile of this step complete; previous step fully consumed
    LGX_TICK(4)

    // ---- S2: vertical 25-tap gaussian.  lane = column, warp = 8-row group of g rows.  Also shifts the b and g
    // planes (their last readers of the previous step are behind the barrier above).
    if (k > 0) {
      for (int idx = tid; idx < BR * B_HIST; idx += T) {
        int r = idx >> 4, j = idx & 15;
        s_b[r * B_PITCH + j] = s_b[r * B_PITCH + kChunk + j];
      }
      for (int idx = tid; idx < GR * G_HIST; idx += T) {
        int r = idx / G_HIST, j = idx - r * G_HIST;
        s_g[r * G_PITCH + j] = s_g[r * G_PITCH + kChunk + j];
      }
    }
    {
      const int c = lane;
      const int q0 = warp * 8;
      double in[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) in[i] = s_f[(q0 + i) * kChunk + c];
      double* vdst = s_v + q0 * V_PITCH + slot + c;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        double acc = __dmul_rn(in[q + 12], p.w[12]);
#pragma unroll
        for (int j = 0; j < 12; ++j)
          acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(in[q + j], in[q + 24 - j]), p.w[j]));
        vdst[q * V_PITCH] = acc;
      }
    }
    __syncthreads();
    LGX_TICK(0)

    // ---- S3: horizontal 25-tap gaussian.  lane = g row, 8-column segment per warp group.
    // ring column of x is x & 63; segment s reads x = x0-24+8s .. +32 and produces g columns x0-12+8s .. +8
    {
      const int xs = x0 - 12 + rseg * 8;
      if (xs + 8 > 0 && xs < W) {
        double in[32];
        const double* row = s_v + rrow * V_PITCH;
        const int c0 = (xs - kRadius) & 63;
#pragma unroll
        for (int i = 0; i < 32; ++i) in[i] = row[(c0 + i) & 63];
        double* grow = s_g + rrow * G_PITCH + G_HIST + rseg * 8;
        const int y = yg0 + rrow;
        const bool store_g = out_g && y >= y0 && y < y0 + nrows;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          double acc = __dmul_rn(in[q + 12], p.w[12]);
#pragma unroll
          for (int j = 0; j < 12; ++j)
            acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(in[q + j], in[q + 24 - j]), p.w[j]));
          grow[q] = acc;
          if (store_g && xs + q >= 0 && xs + q < W) out_g[(size_t)y * Wp + xs + q] = acc;
        }
      }
    }
    __syncthreads();
    LGX_TICK(1)

    // ---- S4: Hessian by nested np.gradient, min eigenvalue.  lane = b row, 8-column segment as two 4-column halves.
    // plane column j of s_g <-> x = x0 - 18 + j; segment s produces b columns x0-16+8s .. +8
    {
      const int rb = rrow;
      const int xb = x0 - 16 + rseg * 8;
      const int y = y0 + rb;
      if (rb < nrows && xb >= 0 && xb < W) {
        double* brow = s_b + rb * B_PITCH + B_HIST + rseg * 8;
        const bool interior = (y >= 2) && (y <= H - 3) && (xb >= 2) && (xb + 7 <= W - 3);
        if (interior) {
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const double* gr = s_g + (rb + 2) * G_PITCH + rseg * 8 + hh * 4;  // plane column of x-2 of the half
            double g0[8], gm1[6], gp1[6], gm2[4], gp2[4];
#pragma unroll
            for (int i = 0; i < 8; ++i) g0[i] = gr[i];
#pragma unroll
            for (int i = 0; i < 6; ++i) { gm1[i] = gr[-G_PITCH + 1 + i]; gp1[i] = gr[G_PITCH + 1 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i) { gm2[i] = gr[-2 * G_PITCH + 2 + i]; gp2[i] = gr[2 * G_PITCH + 2 + i]; }
            double gc0[6], gx[6];  // g_c(y, x-1+i); gx: g_r(y, x-1+i)
#pragma unroll
            for (int i = 0; i < 6; ++i) gc0[i] = __dmul_rn(__dsub_rn(g0[i + 2], g0[i]), 0.5);
            if (!p.mixed_from_cols) {
#pragma unroll
              for (int i = 0; i < 6; ++i) gx[i] = __dmul_rn(__dsub_rn(gp1[i], gm1[i]), 0.5);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              double grp = __dmul_rn(__dsub_rn(gp2[q], g0[q + 2]), 0.5);   // g_r(y+1, x)
              double grm = __dmul_rn(__dsub_rn(g0[q + 2], gm2[q]), 0.5);   // g_r(y-1, x)
              double Hrr = __dmul_rn(__dsub_rn(grp, grm), 0.5);
              double Hrc;
              if (!p.mixed_from_cols) {
                Hrc = __dmul_rn(__dsub_rn(gx[q + 2], gx[q]), 0.5);
              } else {
                double gcp = __dmul_rn(__dsub_rn(gp1[q + 2], gp1[q]), 0.5);  // g_c(y+1, x)
                double gcm = __dmul_rn(__dsub_rn(gm1[q + 2], gm1[q]), 0.5);  // g_c(y-1, x)
                Hrc = __dmul_rn(__dsub_rn(gcp, gcm), 0.5);
              }
              double Hcc = __dmul_rn(__dsub_rn(gc0[q + 2], gc0[q]), 0.5);
              brow[hh * 4 + q] = min_eig(Hrr, Hrc, Hcc);
            }
          }
        } else {
#pragma unroll 1
          for (int q = 0; q < 8; q += 2) {
            double v0 = (xb + q < W) ? b_generic(s_g, yg0, x0 - 18, y, xb + q, H, W, p.mixed_from_cols) : 0.0;
            double v1 = (xb + q + 1 < W) ? b_generic(s_g, yg0, x0 - 18, y, xb + q + 1, H, W, p.mixed_from_cols) : 0.0;
            brow[q] = v0;
            brow[q + 1] = v1;
          }
        }
      }
    }
    __syncthreads();
    LGX_TICK(2)

    // ---- S5: chain warps (the first RW): cv2 RowSum chains of b and b*b (lane = band row; 32 serial steps, both
    // chains in one lane so every b value is read once).  s_b column j <-> x = x0 - 32 + j.  The running sums are
    // staged in dead shared memory (columns 0..31 of the gaussian plane; the ring slot the vertical pass does not
    // use this step) so that every global store of this kernel is a coalesced row segment (a lane-per-row store
    // touches 32 sectors per instruction).  The other warps meanwhile write the new b columns out; then every
    // warp fills the f tile of the next step (its last reader, S2, is behind two barriers) from the words
    // prefetched one step earlier.
    double* stage_q = s_v + (32 - slot);
    if (warp < RW) {
      const int rb = warp * 32 + lane;
      if (rb < nrows) {
        const double* brow = s_b + rb * B_PITCH;
        double* ob = s_g + rb * G_PITCH;
        double* oq = stage_q + rb * V_PITCH;
        if (x0 >= 32 && x0 + 15 <= W - 1) {
          // interior: chain += b[c+7] - b[c-8] for c = x0-24 .. x0+7  (plane columns c+7 -> 15+i, c-8 -> i)
          // (loads are batched ahead of the staging stores: the compiler cannot prove the planes do not alias)
#pragma unroll
          for (int i0 = 0; i0 < 32; i0 += 16) {
            double db[16], dq[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const double a = brow[15 + i0 + e], b = brow[i0 + e];
              db[e] = __dsub_rn(a, b);
              dq[e] = __dsub_rn(__dmul_rn(a, a), __dmul_rn(b, b));
            }
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              chain_b = __dadd_rn(chain_b, db[e]);
              chain_q = __dadd_rn(chain_q, dq[e]);
              db[e] = chain_b;
              dq[e] = chain_q;
            }
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              ob[i0 + e] = db[e];
              oq[i0 + e] = dq[e];
            }
          }
        } else {
          auto B = [&](int x) { return brow[x - (x0 - 32)]; };
#pragma unroll 1
          for (int i = 0; i < 32; ++i) {
            const int c = x0 - 24 + i;
            if (c < 0 || c >= W) continue;
            if (c == 0) {
              double sb = 0.0, sqv = 0.0;
              for (int t = 0; t < 15; ++t) {
                const double v = B(min(max(t - 7, 0), W - 1));
                sb = __dadd_rn(sb, v);
                sqv = __dadd_rn(sqv, __dmul_rn(v, v));
              }
              chain_b = sb;
              chain_q = sqv;
            } else {
              const double a = B(min(c + 7, W - 1)), b = B(max(c - 8, 0));
              chain_b = __dadd_rn(chain_b, __dsub_rn(a, b));
              chain_q = __dadd_rn(chain_q, __dsub_rn(__dmul_rn(a, a), __dmul_rn(b, b)));
            }
            ob[i] = chain_b;
            oq[i] = chain_q;
          }
        }
      }
    } else {
      // new b columns x0-16 .. x0+15 (plane columns 16..47), one 256-byte row segment per warp instruction
      constexpr int CW = NW - RW;                    // copy warps
      constexpr int NT = (BR + CW - 1) / CW;
      const int x = x0 - 16 + lane;
      const bool xok = x >= 0 && x < W;
      double vb[NT];
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const int rb = (warp - RW) + CW * t;
        vb[t] = (rb < BR) ? s_b[rb * B_PITCH + B_HIST + lane] : 0.0;
      }
      unsigned off = (unsigned)((y0 + warp - RW) * Wp + x);   // 32-bit element offsets inside the frame's plane
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const int rb = (warp - RW) + CW * t;
        if (rb < nrows && xok) out_b[off] = vb[t];
        off += (unsigned)(CW * Wp);
      }
    }
    if (k + 1 < nchunks) {
      fill_f(x0 + kChunk);
      if (k + 2 < nchunks) prefetch(x0 + 2 * kChunk);
    }
    __syncthreads();
    LGX_TICK(3)

    // ---- S6: staged running sums -> global, coalesced (columns c = x0-24 .. x0+7)
    {
      double* __restrict__ orb = p.rsb + (size_t)frame * p.plane_stride;
      double* __restrict__ orq = p.rsb2 + (size_t)frame * p.plane_stride;
      // thread <-> (column i = lane, rows warp, warp+NW, ...): loads first, then stores
      constexpr int NT = (BR + NW - 1) / NW;
      const int c = x0 - 24 + lane;
      const bool cok = c >= 0 && c < W;
      double vb[NT], vq[NT];
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const int rb = warp + NW * t;
        if (rb < BR) {
          vb[t] = s_g[rb * G_PITCH + lane];
          vq[t] = stage_q[rb * V_PITCH + lane];
        }
      }
      unsigned off = (unsigned)((y0 + warp) * Wp + c);
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const int rb = warp + NW * t;
        if (rb < nrows && cok) {
          orb[off] = vb[t];
          orq[off] = vq[t];
        }
        off += (unsigned)(NW * Wp);
      }
    }
  }
  if (p.prof && tid == 0) {
    for (int i = 0; i < 5; ++i) atomicAdd(p.prof + i, (unsigned long long)tprof[i]);
    atomicAdd(p.prof + 5, 1ull);
  }
#undef LGX_TICK
}

}  // namespace

template <typename PIX, int NW>
static cudaError_t launch_ridge_t(const RidgeParams& p, int batch, cudaStream_t stream) {
  static unsigned long long attr_done = 0;      // per-device bit: the attribute belongs to the function on one device
  int dev = 0;
  cudaGetDevice(&dev);
  if (!(attr_done >> (dev & 63) & 1ull)) {
    cudaError_t e = cudaFuncSetAttribute(ridge_kernel<PIX, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Geo<NW>::kSmem);
    if (e != cudaSuccess) return e;
    attr_done |= 1ull << (dev & 63);
  }
  dim3 grid(p.bands, batch);
  ridge_kernel<PIX, NW><<<grid, Geo<NW>::kThreads, Geo<NW>::kSmem, stream>>>(p);
  return cudaGetLastError();
}

int ridge_band_rows(int nwarps) { return 8 * nwarps - 4; }

cudaError_t launch_ridge(const RidgeParams& p, int bits, int batch, int nwarps, cudaStream_t stream) {
  if (nwarps == 4) return bits == 8 ? launch_ridge_t<uint8_t, 4>(p, batch, stream) : launch_ridge_t<uint16_t, 4>(p, batch, stream);
  return bits == 8 ? launch_ridge_t<uint8_t, 8>(p, batch, stream) : launch_ridge_t<uint16_t, 8>(p, batch, stream);
}

cudaError_t launch_bgr2gray(const void* bgr, int bits, size_t npix, void* gray, cudaStream_t stream) {
  int blocks = (int)((npix + 255) / 256 < 148 * 16 ? (npix + 255) / 256 : 148 * 16);
  if (bits == 8)
    bgr2gray_kernel<uint8_t><<<blocks, 256, 0, stream>>>((const uint8_t*)bgr, npix, (uint8_t*)gray);
  else
    bgr2gray_kernel<uint16_t><<<blocks, 256, 0, stream>>>((const uint16_t*)bgr, npix, (uint16_t*)gray);
  return cudaGetLastError();
}

cudaError_t launch_blur5(const void* frames, int bits, int batch, int H, int W, size_t pitch, size_t fstride,
                         void* out_pad, int pad_pitch, void* out_dense, cudaStream_t stream) {
  dim3 grid((W + BT_W - 1) / BT_W, (H + BT_H - 1) / BT_H, batch);
  if (bits == 8)
    blur5_u8_kernel<<<dim3((W + B8_W - 1) / B8_W, (H + B8_H - 1) / B8_H, batch), 128, 0, stream>>>(
        (const uint8_t*)frames, pitch, fstride, H, W, (uint8_t*)out_pad, pad_pitch, (uint8_t*)out_dense);
  else if ((pitch & 1) == 0)
    blur5_u16_kernel<<<dim3((W + B16_W - 1) / B16_W, (H + B16_H - 1) / B16_H, batch), 256, 0, stream>>>(
        (const uint16_t*)frames, pitch / 2, fstride, H, W, (uint16_t*)out_pad, pad_pitch, (uint16_t*)out_dense);
  else
    blur5_kernel<uint16_t><<<grid, 256, 0, stream>>>(frames, pitch, fstride, H, W, (uint16_t*)out_pad, pad_pitch, (uint16_t*)out_dense);
  return cudaGetLastError();
}

}  // namespace lgx
