// K1 "ridge": fused  5x5 integer blur -> img_as_float LUT -> 25-tap gaussian (rows first, then
// columns; scipy NI_Correlate1D symmetric order) -> np.gradient x4 -> smaller Hessian eigenvalue b
// -> horizontal running sums of cv2.boxFilter (RowSum) for b and b*b.
//
// Replaces, bit for bit:  /root/reference/utils/util_cylinder.py:1789-1793 (blur + detect_ridges)
// and the row pass of the two cv2.boxFilter calls at :1755-1757.  Operation order: SURVEY.md App. A
// items 2-7, restated on the CPU in oracle/restate.py (blur5, gauss25, min_eigenvalue, row_sums15).
//
// Shape: one CTA owns a horizontal band of <= 60 image rows of one frame and sweeps it left to right
// in 32-column steps.  The sweep is what makes OpenCV's whole-row running sum (a serial chain from
// x = 0) fusable: each of the band's rows has one thread that carries its running sum in a register
// across the whole sweep.  It also means the vertical gaussian is never recomputed for a column
// halo.  All f64 work is FP64-pipe bound (~125 instr / pixel, no FMA allowed); shared memory holds
// five small planes (lanes are mapped to columns for column-wise phases and to rows, with odd
// pitches, for row-wise phases so that every 64-bit access is conflict free).
#include "lgx_internal.cuh"

namespace lgx {

__constant__ double c_w[13];  // c_w[j], j = 0..11: weight of the pair (l-12+j, l+12-j); c_w[12]: centre

cudaError_t upload_gauss_weights(const double* w13) {
  return cudaMemcpyToSymbol(c_w, w13, 13 * sizeof(double));
}

namespace {

constexpr int IN_ROWS = kGRows + 2 * kRadius + 4;  // 92
constexpr int IN_COLS = kChunk + 4;                // 36
constexpr int F_ROWS = kGRows + 2 * kRadius;       // 88
constexpr int V_PITCH = 65;                        // ring of 64 columns, odd pitch
constexpr int G_COLS = kChunk + 4;                 // 36: 4 history + 32 new
constexpr int G_PITCH = 37;
constexpr int B_COLS = kChunk + 16;                // 48: 16 history + 32 new
constexpr int B_PITCH = 49;

constexpr size_t SM_IN = IN_ROWS * IN_COLS * sizeof(uint16_t);   // 6624
constexpr size_t SM_F = F_ROWS * kChunk * sizeof(double);        // 22528
constexpr size_t SM_V = kGRows * V_PITCH * sizeof(double);       // 33280
constexpr size_t SM_G = kGRows * G_PITCH * sizeof(double);       // 18944
constexpr size_t SM_B = kBRows * B_PITCH * sizeof(double);       // 23520
constexpr size_t SM_LUT = 256 * sizeof(double);                  // 2048
constexpr size_t SM_TOTAL = SM_F + SM_V + SM_G + SM_B + SM_LUT + SM_IN;

__device__ __forceinline__ int reflect101(int p, int n) {
  // cv2 borderInterpolate(BORDER_REFLECT_101)
  if (n == 1) return 0;
  while (p < 0 || p >= n) p = (p < 0) ? -p : 2 * (n - 1) - p;
  return p;
}

__device__ __forceinline__ double min_eig(double Hrr, double Hrc, double Hcc) {
  // (M00 + M11)/2 - sqrt(4*M01**2 + (M00 - M11)**2)/2     (skimage _image_orthogonal_matrix22_eigvals)
  double s = __dadd_rn(Hrr, Hcc);
  double d = __dsub_rn(Hrr, Hcc);
  double r = __dsqrt_rn(__dadd_rn(__dmul_rn(4.0, __dmul_rn(Hrc, Hrc)), __dmul_rn(d, d)));
  return __dsub_rn(__dmul_rn(s, 0.5), __dmul_rn(r, 0.5));
}

// np.gradient-of-np.gradient at (y, x) with every border rule, reading g from the band buffer.
// gy0: image row of buffer row 0; gx0: image column of buffer column 0.
__device__ double b_generic(const double* __restrict__ s_g, int gy0, int gx0, int y, int x, int H, int W,
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
__global__ void __launch_bounds__(kRidgeThreads, 2) ridge_kernel(const RidgeParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  double* s_f = reinterpret_cast<double*>(smem);
  double* s_v = s_f + F_ROWS * kChunk;
  double* s_g = s_v + kGRows * V_PITCH;
  double* s_b = s_g + kGRows * G_PITCH;
  double* s_lut = s_b + kBRows * B_PITCH;
  uint16_t* s_in = reinterpret_cast<uint16_t*>(s_lut + 256);

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int H = p.H, W = p.W, Wp = p.Wp;
  const int band = blockIdx.x;
  const int frame = blockIdx.y;
  const int y0 = band * p.rows_per_band;          // first b row of the band
  const int nrows = min(p.rows_per_band, H - y0);  // b rows of the band
  const int yg0 = y0 - 2;                          // image row of g row 0
  const int yf0 = yg0 - kRadius;                   // image row of f row 0
  const int yi0 = yf0 - 2;                         // image row of input-tile row 0

  const PIX* __restrict__ src =
      reinterpret_cast<const PIX*>(reinterpret_cast<const unsigned char*>(p.frames) + (size_t)frame * p.frame_stride_bytes);
  const size_t src_pitch = p.pitch_bytes / sizeof(PIX);
  double* __restrict__ out_b = p.b + (size_t)frame * p.plane_stride;
  double* __restrict__ out_rs = ((warp >> 1) == 0 ? p.rsb : p.rsb2) + (size_t)frame * p.plane_stride;
  double* __restrict__ out_g = p.g ? p.g + (size_t)frame * p.plane_stride : nullptr;
  PIX* __restrict__ out_blur = p.blurred ? reinterpret_cast<PIX*>(p.blurred) + (size_t)frame * H * W : nullptr;

  for (int i = tid; i < kGRows * V_PITCH; i += kRidgeThreads) s_v[i] = 0.0;  // columns x < 0 are zero padding
  if (sizeof(PIX) == 1) s_lut[tid] = p.lut[tid];

  double chain = 0.0;  // running row sum (warps 0-1: b, warps 2-3: b*b), lane <-> band row
  const int nchunks = (W > 8 ? (W - 8 + kChunk - 1) / kChunk : 0) + 1;

  for (int k = 0; k < nchunks; ++k) {
    const int x0 = k * kChunk;
    __syncthreads();  // previous chunk fully consumed

    // ---- S0: input tile (rows yi0.., cols x0-2..), reflect-101 at the image border; history shifts
    for (int idx = tid; idx < IN_ROWS * IN_COLS; idx += kRidgeThreads) {
      int r = idx / IN_COLS, j = idx - r * IN_COLS;
      int y = yi0 + r, x = x0 - 2 + j;
      uint16_t v = 0;
      if (y >= -2 && y < H + 2 && x >= -2 && x < W + 2) {
        v = src[(size_t)reflect101(y, H) * src_pitch + reflect101(x, W)];
      }
      s_in[idx] = v;
    }
    if (k > 0) {
      {  // g history: last 4 columns -> first 4
        int r = tid >> 2, j = tid & 3;
        s_g[r * G_PITCH + j] = s_g[r * G_PITCH + kChunk + j];
      }
      for (int idx = tid; idx < kBRows * 16; idx += kRidgeThreads) {
        int r = idx >> 4, j = idx & 15;
        s_b[r * B_PITCH + j] = s_b[r * B_PITCH + kChunk + j];
      }
    }
    __syncthreads();

    // ---- S1: 5x5 integer blur (separable, sliding), LUT to f64.  lane = column, warp = 11-row group
    {
      const int c = lane;
      const int x = x0 + c;
      const int rbase = warp * 11;
      int h0 = 0, h1 = 0, h2 = 0, h3 = 0, h4 = 0;
#pragma unroll
      for (int rr = 0; rr < 15; ++rr) {
        const uint16_t* row = s_in + (rbase + rr) * IN_COLS + c;
        int hs = (int)row[0] + (int)row[4] + 4 * ((int)row[1] + (int)row[3]) + 6 * (int)row[2];
        h0 = h1; h1 = h2; h2 = h3; h3 = h4; h4 = hs;
        if (rr >= 4) {
          int r = rbase + rr - 4;
          int y = yf0 + r;
          int bl = (h0 + h4 + 4 * (h1 + h3) + 6 * h2 + 128) >> 8;
          bool valid = (y >= 0) && (y < H) && (x < W);
          double f = 0.0;
          if (valid) f = (sizeof(PIX) == 1) ? s_lut[bl] : __ldg(p.lut + bl);
          s_f[r * kChunk + c] = f;
          if (out_blur && valid && y >= y0 && y < y0 + nrows) out_blur[(size_t)y * W + x] = (PIX)bl;
        }
      }
    }
    __syncthreads();

    // ---- S2: vertical 25-tap gaussian.  lane = column, warp = 8-row group of g rows
    {
      const int c = lane;
      const int q0 = warp * 8;
      double in[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) in[i] = s_f[(q0 + i) * kChunk + c];
      const int col = (x0 + c) & 63;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        double acc = __dmul_rn(in[q + 12], c_w[12]);
#pragma unroll
        for (int j = 0; j < 12; ++j)
          acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(in[q + j], in[q + 24 - j]), c_w[j]));
        s_v[(q0 + q) * V_PITCH + col] = acc;
      }
    }
    __syncthreads();

    // ---- S3: horizontal 25-tap gaussian.  lane = g row (32 per warp), warp>>1 = 8-column segment
    {
      const int r = (warp & 1) * 32 + lane;
      const int xs = x0 - 14 + (warp >> 1) * 8;  // first g column of the segment
      if (xs + 8 > 0 && xs < W) {
        double in[32];
        const double* row = s_v + r * V_PITCH;
        const int c0 = (xs - kRadius) & 63;
#pragma unroll
        for (int i = 0; i < 32; ++i) in[i] = row[(c0 + i) & 63];
        double* grow = s_g + r * G_PITCH + 4 + (warp >> 1) * 8;
        const int y = yg0 + r;
        const bool store_g = out_g && y >= y0 && y < y0 + nrows;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          double acc = __dmul_rn(in[q + 12], c_w[12]);
#pragma unroll
          for (int j = 0; j < 12; ++j)
            acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(in[q + j], in[q + 24 - j]), c_w[j]));
          grow[q] = acc;
          if (store_g && xs + q >= 0 && xs + q < W) out_g[(size_t)y * Wp + xs + q] = acc;
        }
      }
    }
    __syncthreads();

    // ---- S4: Hessian by nested np.gradient, min eigenvalue.  lane = b row, warp>>1 = 8-column segment
    {
      const int rb = (warp & 1) * 32 + lane;
      const int seg = warp >> 1;
      const int xb = x0 - 16 + seg * 8;  // first b column of the segment
      const int y = y0 + rb;
      if (rb < nrows && xb >= 0 && xb < W) {
        double bv[8];
        const int gx0 = x0 - 18;  // image column of g buffer column 0
        const bool interior = (y >= 2) && (y <= H - 3) && (xb >= 2) && (xb + 7 <= W - 3);
        if (interior) {
          const double* gr = s_g + (rb + 2) * G_PITCH + seg * 8;  // buffer column of xb-2
          double g0[12], gm1[10], gp1[10], gm2[8], gp2[8];
#pragma unroll
          for (int i = 0; i < 12; ++i) g0[i] = gr[i];
#pragma unroll
          for (int i = 0; i < 10; ++i) { gm1[i] = gr[-G_PITCH + 1 + i]; gp1[i] = gr[G_PITCH + 1 + i]; }
#pragma unroll
          for (int i = 0; i < 8; ++i) { gm2[i] = gr[-2 * G_PITCH + 2 + i]; gp2[i] = gr[2 * G_PITCH + 2 + i]; }
          double gc0[10];  // g_c(y, xb-1+i)
#pragma unroll
          for (int i = 0; i < 10; ++i) gc0[i] = __dmul_rn(__dsub_rn(g0[i + 2], g0[i]), 0.5);
          if (!p.mixed_from_cols) {
            double gr0[10];  // g_r(y, xb-1+i)
#pragma unroll
            for (int i = 0; i < 10; ++i) gr0[i] = __dmul_rn(__dsub_rn(gp1[i], gm1[i]), 0.5);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              double grp = __dmul_rn(__dsub_rn(gp2[q], g0[q + 2]), 0.5);   // g_r(y+1, x)
              double grm = __dmul_rn(__dsub_rn(g0[q + 2], gm2[q]), 0.5);   // g_r(y-1, x)
              double Hrr = __dmul_rn(__dsub_rn(grp, grm), 0.5);
              double Hrc = __dmul_rn(__dsub_rn(gr0[q + 2], gr0[q]), 0.5);
              double Hcc = __dmul_rn(__dsub_rn(gc0[q + 2], gc0[q]), 0.5);
              bv[q] = min_eig(Hrr, Hrc, Hcc);
            }
          } else {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              double grp = __dmul_rn(__dsub_rn(gp2[q], g0[q + 2]), 0.5);
              double grm = __dmul_rn(__dsub_rn(g0[q + 2], gm2[q]), 0.5);
              double Hrr = __dmul_rn(__dsub_rn(grp, grm), 0.5);
              double gcp = __dmul_rn(__dsub_rn(gp1[q + 2], gp1[q]), 0.5);  // g_c(y+1, x)
              double gcm = __dmul_rn(__dsub_rn(gm1[q + 2], gm1[q]), 0.5);  // g_c(y-1, x)
              double Hrc = __dmul_rn(__dsub_rn(gcp, gcm), 0.5);
              double Hcc = __dmul_rn(__dsub_rn(gc0[q + 2], gc0[q]), 0.5);
              bv[q] = min_eig(Hrr, Hrc, Hcc);
            }
          }
        } else {
#pragma unroll 1
          for (int q = 0; q < 8; ++q)
            bv[q] = (xb + q < W) ? b_generic(s_g, yg0, gx0, y, xb + q, H, W, p.mixed_from_cols) : 0.0;
        }
        double* brow = s_b + rb * B_PITCH + 16 + seg * 8;
#pragma unroll
        for (int q = 0; q < 8; ++q) brow[q] = bv[q];
        double2* o = reinterpret_cast<double2*>(out_b + (size_t)y * Wp + xb);
#pragma unroll
        for (int q = 0; q < 4; ++q) o[q] = make_double2(bv[2 * q], bv[2 * q + 1]);
      }
    }
    __syncthreads();

    // ---- S5: cv2 RowSum chains.  warps 0-1: b, warps 2-3: b*b; lane = band row; 32 serial steps
    if (warp < 4) {
      const int rb = (warp & 1) * 32 + lane;
      if (rb < nrows) {
        const bool sq = (warp >> 1) != 0;
        const double* brow = s_b + rb * B_PITCH - (x0 - 32);  // brow[x] = b(y, x)
        double* orow = out_rs + (size_t)(y0 + rb) * Wp;
        auto B = [&](int x) {
          double v = brow[x];
          return sq ? __dmul_rn(v, v) : v;
        };
#pragma unroll 1
        for (int c4 = x0 - 24; c4 < x0 + 8; c4 += 4) {
          if (c4 < 0 || c4 >= W) continue;
          double o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            int c = c4 + e;
            if (c == 0) {
              double s = 0.0;
              for (int i = 0; i < 15; ++i) s = __dadd_rn(s, B(min(max(i - 7, 0), W - 1)));
              chain = s;
            } else if (c < W) {
              chain = __dadd_rn(chain, __dsub_rn(B(min(c + 7, W - 1)), B(max(c - 8, 0))));
            }
            o[e] = chain;
          }
          double2* dst = reinterpret_cast<double2*>(orow + c4);
          dst[0] = make_double2(o[0], o[1]);
          dst[1] = make_double2(o[2], o[3]);
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

// stand-alone 5x5 blur (parity entry point lgx_blur5); the fused kernel does not call it
template <typename PIX>
__global__ void blur5_kernel(const void* frames, size_t pitch_bytes, size_t fstride, int H, int W, PIX* out) {
  int x = blockIdx.x * blockDim.x + threadIdx.x;
  int y = blockIdx.y;
  int f = blockIdx.z;
  if (x >= W) return;
  const PIX* src = reinterpret_cast<const PIX*>(reinterpret_cast<const unsigned char*>(frames) + (size_t)f * fstride);
  size_t sp = pitch_bytes / sizeof(PIX);
  const int kk[5] = {1, 4, 6, 4, 1};
  int acc = 0;
  for (int i = 0; i < 5; ++i) {
    int yy = reflect101(y + i - 2, H);
    int hs = 0;
    for (int j = 0; j < 5; ++j) hs += kk[j] * (int)src[(size_t)yy * sp + reflect101(x + j - 2, W)];
    acc += kk[i] * hs;
  }
  out[((size_t)f * H + y) * W + x] = (PIX)((acc + 128) >> 8);
}

}  // namespace

cudaError_t launch_ridge(const RidgeParams& p, int bits, int batch, cudaStream_t stream) {
  static bool attr_done[2] = {false, false};
  dim3 grid(p.bands, batch);
  if (bits == 8) {
    if (!attr_done[0]) {
      cudaError_t e = cudaFuncSetAttribute(ridge_kernel<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_TOTAL);
      if (e != cudaSuccess) return e;
      attr_done[0] = true;
    }
    ridge_kernel<uint8_t><<<grid, kRidgeThreads, SM_TOTAL, stream>>>(p);
  } else {
    if (!attr_done[1]) {
      cudaError_t e = cudaFuncSetAttribute(ridge_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_TOTAL);
      if (e != cudaSuccess) return e;
      attr_done[1] = true;
    }
    ridge_kernel<uint16_t><<<grid, kRidgeThreads, SM_TOTAL, stream>>>(p);
  }
  return cudaGetLastError();
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
                         void* out, cudaStream_t stream) {
  dim3 grid((W + 127) / 128, H, batch);
  if (bits == 8)
    blur5_kernel<uint8_t><<<grid, 128, 0, stream>>>(frames, pitch, fstride, H, W, (uint8_t*)out);
  else
    blur5_kernel<uint16_t><<<grid, 128, 0, stream>>>(frames, pitch, fstride, H, W, (uint16_t*)out);
  return cudaGetLastError();
}

}  // namespace lgx
