// K1 "ridge", warp-specialised instantiation for large launches (the dominant kernel of the path).
//
// Same arithmetic, bit for bit, as ridge_kernel in lgx_ridge.cu (img_as_float -> 25-tap gaussian rows-then-columns
// in scipy's NI_Correlate1D order -> np.gradient x4 -> smaller Hessian eigenvalue -> cv2 RowSum chains of b and
// b*b; /root/reference/utils/util_cylinder.py:1734-1738, 1755-1757), but organised as
// a pipeline of warp roles instead of block-wide phases.  The kernel is bound by the FP64 pipe (no FMA
// contraction is allowed), and the phase version left that pipe idle half of the time: every thread re-loaded a
// 32-value window per 8 outputs (shared-memory bandwidth as loaded as the FP64 pipe) and the whole block met at
// six barriers per step.  Here:
//
//   one CTA per SM = one band of 128 gaussian rows (<= 124 rows of b) of one frame, swept left to right in
//   32-column steps by four roles of four warps (one warp of each role per SM sub-partition, so the four FP64
//   pipes carry the same load), plus two single-thread TMA warps:
//     L  (1 thread)  TMA-loads the blurred u8/u16 tile of a step (152 rows x 32 columns, zero fill outside the
//                    image = scipy's mode='constant') into a 3-stage ring;
//     V  (4 warps)   vertical 25-tap: lane = column, a thread slides down its warp's 32 rows with the window in
//                    registers (56 loads per 32 outputs);
//     H  (4 warps)   horizontal 25-tap: lane = row, the 24-value window lives in registers across the whole
//                    sweep (one load per output);
//     E  (4 warps)   Hessian / eigenvalue / RowSum chains: lane = row, sliding differences along x, the 16-value
//                    window of the running sums in registers (5 loads per pixel); results go to shared memory
//                    tiles in the TMA 64-byte swizzle as 128-bit stores;
//     S  (1 thread)  TMA-stores the b / rowsum(b) / rowsum(b*b) tiles (the hardware clips columns < 0, >= W and
//                    rows >= H).
//   Roles hand buffers over through mbarriers (full / empty pairs); there is no block-wide barrier in the sweep.
//
// Exact-scaling note (E): np.gradient's "/ 2.0" and the eigenvalue formula's "/ 2" are multiplications by a power
// of two, which commute with IEEE rounding (no intermediate of this path is subnormal: the smallest non-zero
// magnitudes are multiples of ~2^-76).  The differences are therefore formed unscaled and b is scaled once by
// 0.125; 4*Hrc^2 + d^2 is one DFMA with an exact product (4.0 * x).  tests/test_gpu_parity.py compares the planes
// with oracle/restate.py and with the phase kernel bit for bit.
#include "lgx_sqrt.cuh"
#include "lgx_tma.cuh"

namespace lgx {

namespace {

using namespace tma;

constexpr int WS_GR = 128;                 // gaussian rows per band
constexpr int WS_BR = 124;                 // b rows per band (max)
constexpr int WS_VBLK = 32 * 33;           // one V->H block: 32 rows x 32 columns, pitch 33
constexpr int WS_GT = 8;                   // tail columns kept in front of a gaussian slot
constexpr int WS_GP = WS_GT + 32 + 1;      // gaussian slot pitch (odd)
constexpr int WS_GSLOT = WS_GR * WS_GP;
constexpr int WS_OT = 8192;                // bytes of one staging tile plane: 128 rows x 64 B
constexpr int WS_THREADS = 12 * 32;           // 3 warps per SM sub-partition: up to 168 registers per thread
constexpr int WS_VR = 32 + 2 * kRadius;    // blurred rows one V warp needs for its 32 gaussian rows (56)

template <typename PIX>
struct WsCfg;
template <>
struct WsCfg<uint8_t> { static constexpr int NS = 3; };
template <>
struct WsCfg<uint16_t> { static constexpr int NS = 2; };

// shared-memory layout (byte offsets from a 1024-byte aligned base)
constexpr int OFF_BAR = 0;                                  // 40 mbarriers
constexpr int OFF_LUT = 512;                                // 256 doubles
template <typename PIX>                                     // 2 buffers x 3 planes x WS_OT (no LUT for u16)
__host__ __device__ constexpr int off_out() { return sizeof(PIX) == 1 ? 3072 : 1024; }
template <typename PIX>                                     // per V warp: NS stages of its 56-row blurred tile
__host__ __device__ constexpr int off_in() { return off_out<PIX>() + 6 * WS_OT; }
template <typename PIX>
__host__ __device__ constexpr int off_v() { return off_in<PIX>() + 4 * WsCfg<PIX>::NS * WS_VR * 32 * (int)sizeof(PIX); }
template <typename PIX>
__host__ __device__ constexpr int off_g() { return off_v<PIX>() + 8 * WS_VBLK * 8; }
template <typename PIX>
__host__ __device__ constexpr int ws_smem_bytes() { return off_g<PIX>() + 2 * WS_GSLOT * 8 + 1024; }

// barrier indices
constexpr int B_FULL_IN = 0;                                // [warp * 4 + stage]
constexpr int B_FULL_V = 16, B_EMPTY_V = 24;                // [warp * 2 + slot]
constexpr int B_FULL_G = 32, B_EMPTY_G = 34;                // [slot]

struct WsParams {
  CUtensorMap tm_in;              // blurred frames, box 32 x 56
  CUtensorMap tm_o[3];            // b, rowsum(b), rowsum(b*b): box 8 columns x 32 rows
  CUtensorMap tm_p[3];            // the same with the row count of the last, partial E warp of a band
  int H, W, Wp;
  int rows_per_band;
  int bands, nitems;              // bands per frame; bands * frames = work items of the launch
  size_t plane_stride;
  double* g;                      // nullable (debug)
  const double* lut;              // 256 entries (u8)
  unsigned long long* prof;       // nullable: [16], entries 8..15 = role cycle counters (see the end of the kernel)
  double w[13];                   // gaussian taps (RidgeParams::w): read from the kernel-parameter constant bank
};

__device__ __forceinline__ double tap25(const double* in, const double* __restrict__ w) {
  // scipy NI_Correlate1D, symmetric kernel: centre tap first, then the pairs from the far end inwards
  double acc = __dmul_rn(in[12], w[12]);
#pragma unroll
  for (int j = 0; j < 12; ++j) acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(in[j], in[24 - j]), w[j]));
  return acc;
}

// skimage.img_as_float: u8 through the handle's 256-entry table; u16 computed, DIV = false: v * RN(1/65535), what
// scikit-image 0.19 does (np.multiply(image, 1. / imax_in)); DIV = true (LGX_OPT_FLOAT_DIV): v / 65535.0 correctly
// rounded without the division: q0 = v * RN(1/65535), one exact-remainder FMA, one correction FMA (Markstein) -
// equal to the IEEE quotient for all 65536 inputs (tests/test_host_logic.py replays it in exact rational arithmetic).
template <typename PIX, bool DIV>
__device__ __forceinline__ double px_to_f(const double* s_lut, PIX v) {
  if constexpr (sizeof(PIX) == 1) {
    return s_lut[v];
  } else {
    constexpr double R = 1.0 / 65535.0;
    const double x = (double)v;
    const double q0 = __dmul_rn(x, R);
    if constexpr (!DIV) return q0;
    const double rem = __fma_rn(-q0, 65535.0, x);
    return __fma_rn(rem, R, q0);
  }
}

// np.gradient-of-np.gradient at (y, x) with every border rule, reading g from a slot (pitch WS_GP).
// gy0: image row of slot row 0; gx0: image column of slot column 0.  Same operations as b_generic in lgx_ridge.cu.
__device__ __noinline__ double b_generic_ws(const double* __restrict__ s_g, int gy0, int gx0, int y, int x, int H, int W,
                                            int mixed) {
  auto G = [&](int yy, int xx) { return s_g[(yy - gy0) * WS_GP + (xx - gx0)]; };
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
  double Hrc = mixed ? __dmul_rn(__dsub_rn(Dc(yp, x), Dc(ym, x)), sr) : __dmul_rn(__dsub_rn(Dr(y, xp), Dr(y, xm)), scx);
  double s = __dadd_rn(Hrr, Hcc);
  double d = __dsub_rn(Hrr, Hcc);
  double r = __dsqrt_rn(__dadd_rn(__dmul_rn(4.0, __dmul_rn(Hrc, Hrc)), __dmul_rn(d, d)));
  return __dsub_rn(__dmul_rn(s, 0.5), __dmul_rn(r, 0.5));
}

// address of the 16-byte chunk (pair of columns 2m, 2m+1) of row rb in a staging tile (TMA SWIZZLE_64B)
__device__ __forceinline__ double2* tile_chunk(unsigned char* tile, int rb, int m) {
  return reinterpret_cast<double2*>(tile + rb * 64 + ((m ^ ((rb >> 1) & 3)) << 4));
}

// ---- E role state and steps -------------------------------------------------------------------------------------
// (the 15 older values of b that the running sums subtract are read back from the two staging tiles: they hold this
// row's last 16 columns)
struct EState {
  double chain_b = 0.0, chain_q = 0.0;   // cv2 RowSum of b and b*b at column x - 8 on entry of a pixel
  double b0 = 0.0, blast = 0.0;   // b(0), b(min(x, W-1))
};

// per-lane row geometry of np.gradient-of-np.gradient along axis 0 (slot row offsets in doubles, power-of-two scales)
struct ERows {
  int o0, oU, oD, a1, a2, a3, a4;   // slot offsets (row * WS_GP) of rows y, yp, ym and of the four rows of Hrr
  double m1, m2, mB;                // 4*s0*s1, 4*s0*s2, 2*s0  (1, 1, 1 on interior rows)
  __device__ __forceinline__ void init(int y, int yg0, int H) {
    auto sc = [&](int i) { return (i > 0 && i < H - 1) ? 0.5 : 1.0; };
    const int yc = min(y, H - 1);
    const int yp = min(yc + 1, H - 1), ym = max(yc - 1, 0);
    const double s0 = sc(yc);
    o0 = (yc - yg0) * WS_GP; oU = (yp - yg0) * WS_GP; oD = (ym - yg0) * WS_GP;
    a1 = (min(yp + 1, H - 1) - yg0) * WS_GP; a2 = (max(yp - 1, 0) - yg0) * WS_GP;
    a3 = (min(ym + 1, H - 1) - yg0) * WS_GP; a4 = (max(ym - 1, 0) - yg0) * WS_GP;
    m1 = 4.0 * s0 * sc(yp); m2 = 4.0 * s0 * sc(ym); mB = 2.0 * s0;
  }
};

// 8 interior columns (2 <= x <= W-3, x >= 16) of one row: columns j = 8*qt .. 8*qt+7 of the step.
// GEN = false: rows 2 .. H-3 (a2 = a3 = o0, scales 1).  Column c of the slot <-> x = xs - 4 + c.
// tb: staging tile of this quarter (still holding columns xq-16 .. xq-9), tprev: the other one (xq-8 .. xq-1).
template <bool MIXED, bool GEN>
__device__ __forceinline__ void e_quarter_fast(EState& st, const ERows& er, const double* __restrict__ gs, int qt,
                                               unsigned char* tb, const unsigned char* tprev, int rb_lane, int k, int y,
                                               int yg0, int H, int W) {
  unsigned char* ts = tb + WS_OT;
  unsigned char* tq = tb + 2 * WS_OT;
  const int c0 = 8 * qt + 4;                        // slot column of the quarter's first pixel
  const double* row0 = gs + er.o0 + c0;
  const double* rowU = gs + er.oU + c0;
  const double* rowD = gs + er.oD + c0;
  const double* ra1 = gs + er.a1 + c0;
  const double* ra4 = gs + er.a4 + c0;
  const double* ra2 = gs + er.a2 + c0;
  const double* ra3 = gs + er.a3 + c0;
  // sliding state at the first pixel
  double o_0 = row0[0], o_p1 = row0[1];
  double gc_m1 = __dsub_rn(o_0, row0[-2]);          // 2*g_c(x-1) = g(x) - g(x-2)
  double gc_0 = __dsub_rn(o_p1, row0[-1]);          // 2*g_c(x)
  double u_m1, u_0, d_m1, d_0;                      // !MIXED: u = g(yp,.) - g(ym,.) at x-1, x;  MIXED: g(yp,.) / g(ym,.) at x-1, x
  if (!MIXED) {
    u_m1 = __dsub_rn(rowU[-1], rowD[-1]);
    u_0 = __dsub_rn(rowU[0], rowD[0]);
    d_m1 = d_0 = 0.0;
  } else {
    u_m1 = rowU[-1]; u_0 = rowU[0];
    d_m1 = rowD[-1]; d_0 = rowD[0];
  }
  // phase 1: the eight eigenvalues, branch free (independent dependency chains for the scheduler to interleave)
  double bv[8], S[8], X[8];
  unsigned worst = 0;                                // max of the square root's range-check words
  // b(x - 15) for the eight pixels = columns 1..7 of this tile's old content and column 0 of the other tile
  double old[8];
  {
    const double2 c0 = *tile_chunk(tb, rb_lane, 0), c1 = *tile_chunk(tb, rb_lane, 1);
    const double2 c2 = *tile_chunk(tb, rb_lane, 2), c3 = *tile_chunk(tb, rb_lane, 3);
    const double2 n0 = *tile_chunk(const_cast<unsigned char*>(tprev), rb_lane, 0);
    old[0] = c0.y; old[1] = c1.x; old[2] = c1.y; old[3] = c2.x; old[4] = c2.y; old[5] = c3.x; old[6] = c3.y; old[7] = n0.x;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const double gn = row0[i + 2];                   // g(y, x+2)
    const double gc_p1 = __dsub_rn(gn, o_0);         // 2*g_c(x+1)
    const double C = __dsub_rn(gc_p1, gc_m1);        // 4*Hcc
    const double up = rowU[i + 1], dn = rowD[i + 1];
    double B;
    if (!MIXED) {
      const double u_p1 = __dsub_rn(up, dn);         // g(yp, x+1) - g(ym, x+1)
      B = __dsub_rn(u_p1, u_m1);
      u_m1 = u_0; u_0 = u_p1;
    } else {
      B = __dsub_rn(__dsub_rn(up, u_m1), __dsub_rn(dn, d_m1));   // 2*g_c(yp, x) - 2*g_c(ym, x)
      u_m1 = u_0; u_0 = up; d_m1 = d_0; d_0 = dn;
    }
    double A;
    if (!GEN) {
      A = __dsub_rn(__dsub_rn(ra1[i], o_0), __dsub_rn(o_0, ra4[i]));
    } else {
      A = __dsub_rn(__dmul_rn(__dsub_rn(ra1[i], ra2[i]), er.m1), __dmul_rn(__dsub_rn(ra3[i], ra4[i]), er.m2));
      B = __dmul_rn(B, er.mB);
    }
    gc_m1 = gc_0; gc_0 = gc_p1;
    o_0 = o_p1; o_p1 = gn;
    // A, B, C = 4*Hrr, 4*Hrc, 4*Hcc:  b = ((A + C) - sqrt(4*B*B + (A - C)^2)) / 8   (header note)
    S[i] = __dadd_rn(A, C);
    const double D = __dsub_rn(A, C);
    X[i] = __fma_rn(4.0, __dmul_rn(B, B), __dmul_rn(D, D));
  }
  // The eight square roots stage by stage (lgx_sqrt.cuh: branch-free, exact for x in [2^-970, +max]); `worst` records
  // whether any radicand fell outside (in this kernel only x == 0, black areas), and phase 2 below then redoes the
  // quarter with the library square root.
  {
    double R[8];
    sqrt_inrange<8>(X, R, worst);
#pragma unroll
    for (int i = 0; i < 8; ++i) bv[i] = __dmul_rn(__dsub_rn(S[i], R[i]), 0.125);
  }
  if (worst >= kSqrtOutOfRange) {
    // phase 2 (black areas: a radicand is exactly 0, outside the branch-free square root's range): the per-pixel
    // formula with the library square root for the whole quarter
    const int y_in = min(y, H - 1);
#pragma unroll 1
    for (int i = 0; i < 8; ++i) {
      const double v = b_generic_ws(gs, yg0, 32 * k - 20, y_in, 32 * k - 16 + 8 * qt + i, H, W, MIXED);
#pragma unroll
      for (int t = 0; t < 8; ++t) if (t == i) bv[t] = v;
    }
  }
  // phase 3: cv2 RowSum chains s(c) = s(c-1) + (b(c+7) - b(c-8)), c = x - 7; the tile column holds s(x - 8)
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    double vs[2], vq[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int i = 2 * m + e;
      vs[e] = st.chain_b; vq[e] = st.chain_q;
      st.chain_b = __dadd_rn(st.chain_b, __dsub_rn(bv[i], old[i]));
      st.chain_q = __dadd_rn(st.chain_q, __dsub_rn(__dmul_rn(bv[i], bv[i]), __dmul_rn(old[i], old[i])));
    }
    *tile_chunk(tb, rb_lane, m) = make_double2(bv[2 * m], bv[2 * m + 1]);
    *tile_chunk(ts, rb_lane, m) = make_double2(vs[0], vs[1]);
    *tile_chunk(tq, rb_lane, m) = make_double2(vq[0], vq[1]);
  }
}

// 8 columns of one row with every border rule spelled out per pixel (first step, last steps of a sweep).
__device__ __noinline__ void e_quarter_edge(EState& st, const double* __restrict__ gs, int k, int qt, unsigned char* tb,
                                            const unsigned char* tprev, int rb_lane, int y, int yg0, int H, int W, int mixed) {
  unsigned char* ts = tb + WS_OT;
  unsigned char* tq = tb + 2 * WS_OT;
  const bool row_in = y < H;
  double w[24];                                      // w[9 + i] = b(x - 15) of pixel i, w[16 + i] = b of pixel i
  {
    const double2 c0 = *tile_chunk(tb, rb_lane, 0), c1 = *tile_chunk(tb, rb_lane, 1);
    const double2 c2 = *tile_chunk(tb, rb_lane, 2), c3 = *tile_chunk(tb, rb_lane, 3);
    const double2 n0 = *tile_chunk(const_cast<unsigned char*>(tprev), rb_lane, 0);
    w[1] = c0.y; w[2] = c1.x; w[3] = c1.y; w[4] = c2.x; w[5] = c2.y; w[6] = c3.x; w[7] = c3.y; w[8] = n0.x;
  }
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    double vb[2], vs[2], vq[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int i = 2 * m + e;
      const int x = 32 * k - 16 + 8 * qt + i;
      double bv = 0.0;
      if (x >= 0 && x < W) {
        bv = row_in ? b_generic_ws(gs, yg0, 32 * k - 20, y, x, H, W, mixed) : 0.0;
        st.blast = bv;
        if (x == 0) st.b0 = bv;
      } else if (x >= W) {
        bv = st.blast;                               // b(min(c+7, W-1))
      }
      const double oldw = w[i + 1];                  // b(x - 15) (not meaningful before column 16: replaced below)
      w[16 + i] = bv;
      vs[e] = st.chain_b; vq[e] = st.chain_q;        // column x - 8
      const int c = x - 7;
      if (i == 7 && c == 0) {                        // x = 7 is pixel 7 of quarter 2 of step 0
        // cv2 RowSum start: the 15 replicated-border terms accumulated left to right from 0.0 (W >= 64 here)
        double sb = 0.0, sq = 0.0;
#pragma unroll
        for (int t = 0; t < 15; ++t) {
          const int bi = t > 7 ? t - 7 : 0;          // b(clamp(t - 7))
          const double v = (bi == 0) ? st.b0 : w[16 + bi];
          sb = __dadd_rn(sb, v);
          sq = __dadd_rn(sq, __dmul_rn(v, v));
        }
        st.chain_b = sb;
        st.chain_q = sq;
      } else if (c > 0 && c < W) {
        const double old = (x - 15 <= 0) ? st.b0 : oldw;   // b(max(c - 8, 0))
        st.chain_b = __dadd_rn(st.chain_b, __dsub_rn(bv, old));
        st.chain_q = __dadd_rn(st.chain_q, __dsub_rn(__dmul_rn(bv, bv), __dmul_rn(old, old)));
      }
      vb[e] = bv;
    }
    *tile_chunk(tb, rb_lane, m) = make_double2(vb[0], vb[1]);
    *tile_chunk(ts, rb_lane, m) = make_double2(vs[0], vs[1]);
    *tile_chunk(tq, rb_lane, m) = make_double2(vq[0], vq[1]);
  }
}

// Outputs per loop iteration of the two 25-tap roles: the window shift costs 48 register moves per iteration, so a
// larger group trades instruction-cache footprint (the three roles' loops must fit 32 KB together) for issue slots.
#ifndef LGX_WS_GROUP
#define LGX_WS_GROUP 16
#endif
constexpr int WG = LGX_WS_GROUP;

template <typename PIX, bool MIXED, bool DIV>
__global__ void __launch_bounds__(WS_THREADS, 1) ridge_ws_kernel(const __grid_constant__ WsParams p) {
  constexpr int NS = WsCfg<PIX>::NS;
  constexpr int TILE_IN = WS_VR * 32 * (int)sizeof(PIX);    // one V warp, one stage
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t bar0 = smem_u32(smem + OFF_BAR);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  double* s_lut = reinterpret_cast<double*>(smem + OFF_LUT);
  unsigned char* s_out = smem + off_out<PIX>();
  unsigned char* s_in = smem + off_in<PIX>();
  double* s_v = reinterpret_cast<double*>(smem + off_v<PIX>());
  double* s_g = reinterpret_cast<double*>(smem + off_g<PIX>());

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int H = p.H, W = p.W;
  const int nsteps = (W + 23) / 32 + 1;
  // Persistent CTAs: CTA c sweeps the work items c, c + gridDim.x, ... (item = frame * bands + band) back to back;
  // the roles' hand-over protocol simply runs on over the global step counter kg = item_index * nsteps + k, so
  // the pipeline never drains between sweeps (and a launch on fewer CTAs than SMs leaves the other SMs to the
  // memory-bound kernels of another stream).
  const int my_items = ((int)blockIdx.x < p.nitems) ? (p.nitems - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int total = my_items * nsteps;
  auto item_of = [&](int j, int& frame, int& y0) {
    const int item = (int)blockIdx.x + j * (int)gridDim.x;
    frame = item / p.bands;
    y0 = (item - frame * p.bands) * p.rows_per_band;
  };

  if (tid == 0) {
    for (int s = 0; s < 16; ++s) mbar_init(BAR(B_FULL_IN + s), 1);
    for (int s = 0; s < 8; ++s) { mbar_init(BAR(B_FULL_V + s), 32); mbar_init(BAR(B_EMPTY_V + s), 32); }
    for (int s = 0; s < 2; ++s) { mbar_init(BAR(B_FULL_G + s), 128); mbar_init(BAR(B_EMPTY_G + s), 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (sizeof(PIX) == 1)
    for (int i = tid; i < 256; i += WS_THREADS) s_lut[i] = p.lut[i];
  for (int i = tid; i < 2 * WS_GSLOT; i += WS_THREADS) s_g[i] = 0.0;   // tails of step 0, rows never written
  __syncthreads();

  const int role = warp >> 2;
  const bool prof = p.prof != nullptr;
  long long wait_a = 0, wait_b = 0;                 // cycles this warp waited for its input / its output buffer
  const long long t_begin = prof ? clock64() : 0;
  if (role == 0) {
    // ------------------------------------------------------------------ V: vertical 25-tap, lane = column
    // Each V warp TMA-loads its own 56 blurred rows (lane 0 issues, NS - 1 steps ahead): no cross-warp hand-over.
    const int w = warp;
    unsigned char* my_in = s_in + w * NS * TILE_IN;
    auto issue_load = [&](int kgx) {                  // tile of global step kgx into its stage (lane 0 only)
      const int jx = kgx / nsteps, kx = kgx - jx * nsteps;
      int frame, y0;
      item_of(jx, frame, y0);
      const int sn = kgx % NS;
      mbar_expect_tx(BAR(B_FULL_IN + w * 4 + sn), TILE_IN);
      tma_load_3d(&p.tm_in, BAR(B_FULL_IN + w * 4 + sn), smem_u32(my_in + sn * TILE_IN), 32 * kx, y0 - 2 - kRadius + 32 * w, frame);
    };
    if (lane == 0)
      for (int kg = 0; kg < NS - 1 && kg < total; ++kg) issue_load(kg);
    for (int kg = 0; kg < total; ++kg) {
      const int stage = kg % NS, slot = kg & 1;
      // the stage read in step kg-1 is free (program order of this warp): refill it with the tile of step kg+NS-1
      __syncwarp();
      if (lane == 0 && kg + NS - 1 < total) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        issue_load(kg + NS - 1);
      }
      mbar_wait(BAR(B_FULL_IN + w * 4 + stage), (kg / NS) & 1, prof, wait_a);
      mbar_wait(BAR(B_EMPTY_V + w * 2 + slot), ((kg >> 1) & 1) ^ 1, prof, wait_b);
      const PIX* tile = reinterpret_cast<const PIX*>(my_in + stage * TILE_IN) + lane;
      double* vb = s_v + (w * 2 + slot) * WS_VBLK + lane;
      double in[24 + WG];
#pragma unroll
      for (int i = 0; i < 24; ++i) in[i] = px_to_f<PIX, DIV>(s_lut, tile[i * 32]);
      // one copy of the 8-output body (the three roles run different code at the same time: the whole sweep has to
      // fit the 32 KB instruction cache); the window shift is 24 register moves per 296 FP64 instructions
#pragma unroll 1
      for (int grp = 0; grp < 32 / WG; ++grp) {
        const PIX* tg = tile + (24 + WG * grp) * 32;
        double* vg = vb + WG * grp * 33;
#pragma unroll
        for (int i = 0; i < WG; ++i) in[24 + i] = px_to_f<PIX, DIV>(s_lut, tg[i * 32]);
#pragma unroll
        for (int q = 0; q < WG; ++q) vg[q * 33] = tap25(in + q, p.w);
#pragma unroll
        for (int i = 0; i < 24; ++i) in[i] = in[i + WG];
      }
      mbar_arrive(BAR(B_FULL_V + w * 2 + slot));
    }
  } else if (role == 1) {
    // ------------------------------------------------------------------ H: horizontal 25-tap, lane = row
    const int w = warp - 4;
    const int r = 32 * w + lane;
    double in[24 + WG];
    int k = 0, j = 0;
    for (int kg = 0; kg < total; ++kg) {
      const int slot = kg & 1;
      if (k == 0) {                                   // new sweep: v(x < 0) = 0
#pragma unroll
        for (int i = 0; i < 24; ++i) in[i] = 0.0;
      }
      mbar_wait(BAR(B_FULL_V + w * 2 + slot), (kg >> 1) & 1, prof, wait_a);
      const double* vrow = s_v + (w * 2 + slot) * WS_VBLK + lane * 33;
      double* grow = s_g + slot * WS_GSLOT + r * WS_GP + WS_GT;    // slot column c <-> x = 32k - 12 + c
#pragma unroll 1
      for (int grp = 0; grp < 32 / WG; ++grp) {
        const double* vg = vrow + WG * grp;
        double* gg = grow + WG * grp;
#pragma unroll
        for (int i = 0; i < WG; ++i) in[24 + i] = vg[i];
#pragma unroll
        for (int q = 0; q < WG; ++q) gg[q] = tap25(in + q, p.w);
#pragma unroll
        for (int i = 0; i < 24; ++i) in[i] = in[i + WG];
      }
      mbar_arrive(BAR(B_EMPTY_V + w * 2 + slot));
      mbar_arrive(BAR(B_FULL_G + slot));
      if (p.g) {        // debug plane (lgx_ridge with d_g): re-read this row of the slot
        int frame, y0;
        item_of(j, frame, y0);
        const int y = y0 - 2 + r;
        if (y >= y0 && y < min(y0 + p.rows_per_band, H)) {
          double* out_g = p.g + (size_t)frame * p.plane_stride;
#pragma unroll 1
          for (int c = 0; c < 32; ++c) {
            const int x = 32 * k - 12 + c;
            if (x >= 0 && x < W) out_g[(size_t)y * p.Wp + x] = grow[c];
          }
        }
      }
      if (kg + 1 < total) {
        // the last 8 columns are also the tail of the next slot; E must have finished step kg-1 in it
        mbar_wait(BAR(B_EMPTY_G + (slot ^ 1)), (((kg + 1) >> 1) & 1) ^ 1, prof, wait_b);
        double* gt = s_g + (slot ^ 1) * WS_GSLOT + r * WS_GP;
#pragma unroll
        for (int q = 0; q < 8; ++q) gt[q] = grow[24 + q];
      }
      if (++k == nsteps) { k = 0; ++j; }
    }
  } else if (role == 2) {
    // ------------------------------------------------------------------ E: Hessian, eigenvalue, RowSum chains
    const int w = warp - 8;
    const int rb_lane = 32 * w + lane;
    const int rb = min(rb_lane, WS_BR - 1);          // lanes 124..127 shadow row 123 (their tile rows are never stored)
    // rows of the band this warp stores itself (TMA box rows: 32, or the remainder for the band's last warp)
    const int my_rows = min(32, max(0, p.rows_per_band - 32 * w));
    ERows er;
    EState st;
    int frame = 0, y0 = 0, y = 0, yg0 = 0;
    bool gen_rows = false;
    int k = 0, j = 0;
    for (int kg = 0; kg < total; ++kg) {
      const int slot = kg & 1;
      if (k == 0) {                                   // new sweep
        item_of(j, frame, y0);
        y = y0 + rb;
        yg0 = y0 - 2;
        er.init(y, yg0, H);
        // rows 0, 1, H-2, H-1 follow np.gradient's one-sided rules: a warp that owns one of them (or rows below
        // the image) runs the variant with per-lane row offsets and scales, every other warp the plain interior one
        gen_rows = __any_sync(0xffffffffu, !(y >= 2 && y <= H - 3));
        st = EState();
      }
      mbar_wait(BAR(B_FULL_G + slot), (kg >> 1) & 1, prof, wait_a);
      const double* gs = s_g + slot * WS_GSLOT;       // slot column 0 <-> x = 32k - 20
      const int xs = 32 * k - 16;                     // b column of j = 0
      const bool fast = k >= 1 && xs + 31 <= W - 3;
#pragma unroll 1
      for (int qt = 0; qt < 4; ++qt) {
        const int buf = qt & 1;
        // this warp's rows of the staging buffer were handed to the TMA two quarters ago (bulk group of lane 0)
        if (lane == 0) {
          const long long t0 = prof ? clock64() : 0;
          asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          if (prof) wait_b += clock64() - t0;
        }
        __syncwarp();
        unsigned char* tb = s_out + buf * 3 * WS_OT;
        const unsigned char* tprev = s_out + (buf ^ 1) * 3 * WS_OT;
        if (fast) {
          if (!gen_rows) e_quarter_fast<MIXED, false>(st, er, gs, qt, tb, tprev, rb_lane, k, y, yg0, H, W);
          else e_quarter_fast<MIXED, true>(st, er, gs, qt, tb, tprev, rb_lane, k, y, yg0, H, W);
        } else {
          EState tmp = st;     // the out-of-line edge step takes the state by address; keep `st` itself in registers
          e_quarter_edge(tmp, gs, k, qt, tb, tprev, rb_lane, y, yg0, H, W, MIXED);
          st = tmp;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          // b columns x .. x+7; running sums columns x-8 .. x-1; rows 32w .. of the band (the hardware clips)
          const int x = xs + 8 * qt;
          if (my_rows > 0) {
            const CUtensorMap* tm = my_rows == 32 ? p.tm_o : p.tm_p;
            const uint32_t src = smem_u32(tb + 32 * w * 64);
            if (x + 8 > 0 && x < W) tma_store_3d(tm + 0, src, x, y0 + 32 * w, frame);
            if (x > 0 && x - 8 < W) {
              tma_store_3d(tm + 1, src + WS_OT, x - 8, y0 + 32 * w, frame);
              tma_store_3d(tm + 2, src + 2 * WS_OT, x - 8, y0 + 32 * w, frame);
            }
          }
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      mbar_arrive(BAR(B_EMPTY_G + slot));
      if (++k == nsteps) { k = 0; ++j; }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  if (prof && lane == 0 && role < 3) {
    // per role (V, H, E): [8 + 2*role] wait for input, [9 + 2*role] wait for the output buffer; [14] total warp
    // cycles of the three roles (12 warps per CTA), [15] CTAs
    atomicAdd(p.prof + 8 + 2 * role, (unsigned long long)wait_a);
    atomicAdd(p.prof + 9 + 2 * role, (unsigned long long)wait_b);
    atomicAdd(p.prof + 14, (unsigned long long)(clock64() - t_begin));
    if (warp == 0) atomicAdd(p.prof + 15, 1ull);
  }
}

template <typename PIX, bool MIXED, bool DIV>
cudaError_t launch_t(const WsParams& p, int ctas, cudaStream_t stream) {
  static unsigned long long attr_done = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  constexpr int smem = ws_smem_bytes<PIX>();
  if (!(attr_done >> (dev & 63) & 1ull)) {
    cudaError_t e = cudaFuncSetAttribute(ridge_ws_kernel<PIX, MIXED, DIV>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    attr_done |= 1ull << (dev & 63);
  }
  ridge_ws_kernel<PIX, MIXED, DIV><<<ctas, WS_THREADS, smem, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace

bool ridge_ws_usable(const RidgeParams& rp, int bits) {
  const size_t psz = (size_t)bits / 8;
  auto a16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  return rp.W >= 64 && rp.H >= 8 && a16(rp.blur) && a16(rp.b) && a16(rp.rsb) && a16(rp.rsb2) &&
         ((size_t)rp.blur_pitch * psz) % 16 == 0 && encode_fn() != nullptr;
}

int ridge_ws_band_rows() { return WS_BR; }

// rp.bands / rp.rows_per_band must have been computed with ridge_ws_band_rows().
cudaError_t launch_ridge_ws(const RidgeParams& rp, int bits, int batch, int max_ctas, cudaStream_t stream) {
  WsParams p;
  const size_t psz = (size_t)bits / 8;
  const size_t in_row = (size_t)rp.blur_pitch * psz;
  bool ok = encode3(&p.tm_in, bits == 8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_UINT16, (int)psz, rp.blur,
                    rp.W, rp.H, batch, in_row, in_row * rp.H, 32, WS_VR, CU_TENSOR_MAP_SWIZZLE_NONE);
  const size_t prow = (size_t)rp.Wp * 8, pframe = rp.plane_stride * 8;
  const int rem = rp.rows_per_band % 32 ? rp.rows_per_band % 32 : 32;
  const double* planes[3] = {rp.b, rp.rsb, rp.rsb2};
  for (int i = 0; i < 3; ++i) {
    ok = ok && encode3(&p.tm_o[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, planes[i], rp.W, rp.H, batch, prow, pframe, 8, 32,
                       CU_TENSOR_MAP_SWIZZLE_64B);
    ok = ok && encode3(&p.tm_p[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, planes[i], rp.W, rp.H, batch, prow, pframe, 8, rem,
                       CU_TENSOR_MAP_SWIZZLE_64B);
  }
  if (!ok) return cudaErrorInvalidValue;
  p.H = rp.H; p.W = rp.W; p.Wp = rp.Wp;
  p.rows_per_band = rp.rows_per_band;
  p.bands = rp.bands;
  p.nitems = rp.bands * batch;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (max_ctas > 0 && max_ctas < sms) sms = max_ctas;
  const int ctas = p.nitems < sms ? p.nitems : sms;
  p.plane_stride = rp.plane_stride;
  p.g = rp.g;
  p.lut = rp.lut;
  p.prof = rp.prof;
  for (int i = 0; i < 13; ++i) p.w[i] = rp.w[i];
  if (bits == 8)
    return rp.mixed_from_cols ? launch_t<uint8_t, true, false>(p, ctas, stream) : launch_t<uint8_t, false, false>(p, ctas, stream);
  if (rp.float_div)
    return rp.mixed_from_cols ? launch_t<uint16_t, true, true>(p, ctas, stream) : launch_t<uint16_t, false, true>(p, ctas, stream);
  return rp.mixed_from_cols ? launch_t<uint16_t, true, false>(p, ctas, stream) : launch_t<uint16_t, false, false>(p, ctas, stream);
}

}  // namespace lgx
