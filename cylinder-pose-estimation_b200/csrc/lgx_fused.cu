// K1F "ridge + sauvola, fused": img_as_float -> 25-tap gaussian (rows, columns) -> np.gradient x4 -> smaller Hessian
// eigenvalue b -> cv2.boxFilter(b), cv2.boxFilter(b*b) (RowSum and ColumnSum chains) -> Sauvola threshold -> compare,
// in ONE persistent kernel: the three f64 planes (b, rowsum b, rowsum b*b) that lgx_ridge_ws.cu writes and
// lgx_sauvola.cu reads back (24 B/px each way, 82 % of the path's DRAM traffic) never exist.
//
// Replaces, bit for bit, /root/reference/utils/util_cylinder.py:1734-1738 (detect_ridges), :1740-1766
// (sauvola_threshold_fast), :1798-1800 (compare / invert); CPU twins: oracle/restate.py (arithmetic) and
// oracle/fused_model.py (this kernel's lane schedule, rings and hand-over, checked on the CPU).
//
// Why it is not a plain fusion: the two running-sum chains of cv2.boxFilter are serial in orthogonal directions from
// the image edges (RowSum along x from column 0, ColumnSum along y from row 0) and floating-point addition does not
// re-associate, so a tile cannot start either chain in the middle.  Here:
//
//   * a CTA owns one band of 124 rows of one frame and sweeps it left to right (as lgx_ridge_ws.cu does); its four
//     SM sub-partitions each run one 32-row block of the band as a two-warp pipeline
//         VH_w  vertical 25-tap (lane = column, TMA-fed u8/u16 tile, window slides down in registers), then
//               horizontal 25-tap (lane = row, 24-value window lives in registers across the sweep) into a 64-column
//               ring of g (plus a 128-column ring of its last four rows for the block below);
//         EC_w  lane = row, lanes SKEWED by one column each (lane l is at column t - l in slot t): Hessian, eigenvalue
//               (branch-free interleaved square roots), RowSum chains in registers as before - and the ColumnSum
//               chain travels DOWN the lanes by shuffle, one lane per slot: lane l adds its row sum of column p to the
//               running sum lane l-1 produced for column p in the previous slot, subtracts the row sum of 14 rows up
//               (lane l-14, 14 slots ago, a 16-slot ring in shared memory), forms mean / variance / threshold and
//               compares with b of 7 rows up (lane l-7, 15 slots ago).  Compare bits are shifted into a per-lane
//               word and stored when 32 columns are complete;
//     block w runs 40 slots behind block w-1 (the chain crosses 32 lanes), and receives the row sums of the 14 rows
//     above it, b of the 7 rows above it and the running sums through a 128-column ring in global memory (L2
//     resident: 114 KB per CTA); band j+1 of a frame runs on the next CTA and receives the same from band j through
//     per-item arrays (296 B per column, the only DRAM traffic the fusion adds: ~25 MB per 2448x2048 frame instead
//     of 250 MB), gated by a progress counter per item.
//   * work items (frame, band) are dealt round-robin to the persistent CTAs, so the bands of a frame run on adjacent
//     CTAs as a wavefront; all CTAs are co-resident (cooperative launch), which the waits between bands need.
//
// Exactness notes: every f64 operation is individually rounded in the reference's order (-fmad=false; the explicit
// __fma_rn calls have an exact product).  Sauvola: mean * (1 + 0.5 * ((sd / 128) - 1)) is formed as
// m * (1 + fma(sd, 2^-8, -0.5)) - the power-of-two scalings commute with rounding (oracle/fused_model.py asserts
// the identity on every pixel it visits).
#include "lgx_sqrt.cuh"
#include "lgx_tma.cuh"

namespace lgx {
namespace {

using namespace tma;

constexpr int FZ_THREADS = 256;            // 8 warps: VH_0..3 (warps 0-3), EC_0..3 (warps 4-7); warp % 4 = sub-partition = block
constexpr int FZ_BR = 124;                 // rows of b per band
constexpr int FZ_VR = 32 + 2 * kRadius;    // blurred rows a VH warp needs for its 32 gaussian rows (56)
constexpr int FZ_NS = 1;                   // stages of the input ring of a VH warp (the next tile is requested right after the vertical
                                           // pass has read this one; the horizontal pass covers its latency)
constexpr int FZ_VB = 32 * 33;             // v block of a VH warp: 32 rows x 32 columns, pitch 33
constexpr int FZ_GP = 78;                  // pitch of a g ring row: 64 columns + the first 12 again + 2 (P - 1 odd: the skewed reads
                                           // of the EC lanes, address l * (P - 1) + t, hit 16 different bank pairs)
constexpr int FZ_GROWS = 36;               // g rows an EC warp reads: lane l needs rows l .. l+4 (rows 0..3 come from the block above)
constexpr int FZ_GW = FZ_GROWS * FZ_GP;
constexpr int FZ_HP = 142;                 // pitch of a halo ring row: 128 columns + 12 + 2
constexpr int FZ_HW = 4 * FZ_HP;
constexpr int FZ_LAG = 40;                 // slots block w runs behind block w-1: 32 lanes + one batch of 8
constexpr int FZ_RINGCOLS = 128;           // columns of the intra-CTA hand-over rings
constexpr int FZ_PUBLISH = 64;             // EC_3 publishes its progress to the next band every 64 slots (one gpu-scope fence each)
constexpr int WG2 = 16;                    // outputs per loop iteration of the two 25-tap phases

constexpr int WH2 = 8;                     // columns per hand-over group of the g ring (VH -> EC full / empty barriers)

// hand-over block of nc columns (everything in 16-byte units so that cp.async.cg can fetch it):
//   rsb[7][nc]  {rowsum b, rowsum b*b, b, -}   rows 0..6 of the 14 row-sum rows; b of the 7 rows above the block (b of column k in record k)
//   rs [7][nc]  {rowsum b, rowsum b*b}         rows 7..13
//   sum[nc]     {running column sum of b, of b*b} below the block's last row
__host__ __device__ constexpr size_t ho_bytes(size_t nc) { return nc * (7 * 32 + 7 * 16 + 16); }
__host__ __device__ constexpr size_t ho_off_rs(size_t nc, int r, size_t k) {     // byte offset of the row-sum pair of row r, column k
  return r < 7 ? ((size_t)r * nc + k) * 32 : nc * 7 * 32 + ((size_t)(r - 7) * nc + k) * 16;
}
__host__ __device__ constexpr size_t ho_off_sum(size_t nc, size_t k) { return nc * (7 * 32 + 7 * 16) + k * 16; }

// staging of an EC warp for the hand-over values of one batch (filled by cp.async): 14 lanes x 8 pixels x 32 B
// (lane pitch 272 B: 2-way instead of 14-way bank conflicts), then the 8 running-sum pairs of the first lane
constexpr int FZ_STG_LANE = 8 * 32 + 16;
constexpr int FZ_STG_SUM = 14 * FZ_STG_LANE;
constexpr int FZ_STG = FZ_STG_SUM + 8 * 16;

// shared-memory layout (byte offsets from a 1024-byte aligned base)
constexpr int FO_BAR = 0;                  // mbarriers
constexpr int FO_PROG = 960;               // 4 ints: slots completed by EC_w
constexpr int FO_LUT = 1024;               // 256 doubles
constexpr int FO_IN = 3072;
template <typename PIX>
__host__ __device__ constexpr int fz_tile_in() { return FZ_VR * 32 * (int)sizeof(PIX); }
template <typename PIX>
__host__ __device__ constexpr int fo_v() { return FO_IN + 4 * FZ_NS * fz_tile_in<PIX>(); }
template <typename PIX>
__host__ __device__ constexpr int fo_g() { return fo_v<PIX>() + 4 * FZ_VB * 8; }
template <typename PIX>
__host__ __device__ constexpr int fo_hl() { return fo_g<PIX>() + 4 * FZ_GW * 8; }
template <typename PIX>
__host__ __device__ constexpr int fo_bs() { return fo_hl<PIX>() + 3 * FZ_HW * 8; }
template <typename PIX>
__host__ __device__ constexpr int fo_rs() { return fo_bs<PIX>() + 4 * 16 * 32 * 8; }
template <typename PIX>
__host__ __device__ constexpr int fo_stg() { return fo_rs<PIX>() + 4 * 16 * 32 * 16; }
template <typename PIX>
__host__ __device__ constexpr int fz_smem_bytes() { return fo_stg<PIX>() + 4 * FZ_STG + 1024; }

// barrier indices
constexpr int FB_IN = 0;                   // [w]                             TMA tile landed
constexpr int FB_FULL_G = 4;               // [w * 8 + group]                 8 columns of the g ring written (32 arrivals)
constexpr int FB_EMPTY_G = 36;             // [w * 8 + group]                 ... read by every lane of EC_w (1 arrival)
constexpr int FB_FULL_H = 68;              // [(w-1) * 8 + group], w = 1..3   16 columns of the halo ring of EC_w written (4 arrivals)
constexpr int FB_EMPTY_H = 92;             // [(w-1) * 8 + group]             ... read by lanes 0..3 of EC_w (1 arrival)
constexpr int FB_COUNT = FB_EMPTY_H + 24;
static_assert(8 * FB_COUNT <= FO_PROG, "barriers overlap the progress words");

struct FzParams {
  CUtensorMap tm_in;              // blurred frames, box 32 x 56
  int H, W, WW;
  int bands, nitems, nsteps;      // bands per frame (over H + 7 rows), bands * frames, 32-column steps per sweep
  int batch, groups;              // frames; CTA c works on band c % bands of the frames c / bands + j * groups, j = 0, 1, ...
  uint32_t* bits;                 // [frames][H][WW]
  unsigned char* ho_items;        // [nitems] hand-over blocks of nc_item columns (band -> next band)
  unsigned char* ho_rings;        // [ctas][3] hand-over blocks of FZ_RINGCOLS columns (block -> next block of the CTA)
  int nc_item;
  int* prog;                      // [nitems] columns of the sweep whose hand-over EC_3 has published (zeroed before the launch)
  const double* lut;              // 256 entries (u8)
  double* dbg_b;                  // nullable: [frames][H][Wp]
  double* dbg_T;                  // nullable
  int Wp;
  size_t plane_stride;
  double w[13];                   // gaussian taps
};

__device__ __forceinline__ double tap25f(const double* in, const double* __restrict__ w) {
  // scipy NI_Correlate1D, symmetric kernel: centre tap first, then the pairs from the far end inwards
  double acc = __dmul_rn(in[12], w[12]);
#pragma unroll
  for (int j = 0; j < 12; ++j) acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(in[j], in[24 - j]), w[j]));
  return acc;
}

template <typename PIX, bool DIV>
__device__ __forceinline__ double px_to_ff(const double* s_lut, PIX v) {
  if constexpr (sizeof(PIX) == 1) {
    return s_lut[v];
  } else {
    constexpr double R = 1.0 / 65535.0;      // skimage 0.19 img_as_float: v * RN(1 / 65535)
    const double x = (double)v;
    const double q0 = __dmul_rn(x, R);
    if constexpr (!DIV) return q0;
    const double rem = __fma_rn(-q0, 65535.0, x);
    return __fma_rn(rem, R, q0);
  }
}

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_cta_shared(uint32_t a) {
  int v;
  asm volatile("ld.acquire.cta.shared::cta.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_cta_shared(uint32_t a, int v) {
  asm volatile("st.release.cta.shared::cta.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
// a wait longer than ~4 s is a protocol bug: trap instead of hanging the device
struct SpinGuard {
  unsigned long long t0 = 0;
  unsigned spins = 0;
  __device__ __forceinline__ void tick() {
    if ((++spins & 255u) == 0) {
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t0 == 0) t0 = t1;
      else if (t1 - t0 > 4000000000ull) __trap();
    }
    __nanosleep(64);
  }
};

// ---- the EC role --------------------------------------------------------------------------------------------------
// Per-lane geometry of the current sweep (item) - changes once per sweep and lane.
struct EcRow {
  int frame, y;                 // image row of the lane (>= H: virtual or idle)
  bool valid;                   // the row exists in this band (image rows and the 7 replicated rows below the image)
  bool real;                    // y < H
  bool emit;                    // y >= 7: the column role produces output row y - 7
  // np.gradient-of-np.gradient along axis 0 with the one-sided rules at rows 0, 1, H-2, H-1 (as ERows in lgx_ridge_ws.cu):
  // indices into the lane's five g rows (0..4 <-> image rows y-2..y+2) and power-of-two scales
  int i0, iU, iD, i1, i2, i3, i4;
  double m1, m2, mB;
  bool interior;                // 2 <= y <= H-3
  bool top;                     // block 0 of band 0: the first rows of the image (no block above)
  unsigned char* in_base;       // hand-over block this lane reads (block above) and writes (block below) in its current sweep
  unsigned char* out_base;
};

struct EcState {
  // row chain (cv2 RowSum of b and b*b at column x - 8 on entry of a pixel)
  double chain_b = 0.0, chain_q = 0.0, b0 = 0.0, blast = 0.0;
  // column chain: running sums below this lane for the column of the previous slot; `new` of the previous slot
  double sum_b = 0.0, sum_q = 0.0, newp_b = 0.0, newp_q = 0.0;
  unsigned word = 0;
};

// what a lane needs of the kernel's geometry
struct EcCtx {
  int H, W, SW, lane, w, first;       // first: first lane of the block that can hold a row (4 in block 0, else 0)
  uint32_t rowoff[5];                 // shared-memory byte address of the lane's g rows l .. l+4 (column position 0)
  uint32_t halo;                      // bit k: row k lives in the 128-column halo ring
  uint32_t bs, rs;                    // shared-memory byte addresses of this warp's Bs / Rs rings
};

__device__ __forceinline__ double lds_f64(uint32_t a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_f64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ double2 lds_f64x2(uint32_t a) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_f64x2(uint32_t a, double x, double y) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(x), "d"(y) : "memory");
}

// The lane's five g rows and which of them live in the 128-column halo ring, passed by value to the out-of-line helpers
// (a reference to the lane's context would force it into local memory).
struct GRows {
  uint32_t off[5];
  uint32_t halo;
};

// g(window row k, global column coordinate ug) for this lane
__device__ __forceinline__ double g_at(const GRows& g, int k, int ug) {
  const int m = (g.halo >> k) & 1u ? 127 : 63;
  return lds_f64(g.off[k] + 8u * (uint32_t)(ug & m));
}

// b at (y, x) with every border rule of np.gradient, for a lane whose current column x has the global coordinate ug
// (same operations as b_generic in lgx_ridge.cu)
__device__ __noinline__ double b_generic_f(GRows gr, int H, int W, int y, int x, int ug, bool mixed) {
  auto G = [&](int yy, int xx) { return g_at(gr, 2 + (yy - y), ug + (xx - x)); };
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

__device__ __forceinline__ void ec_set_row(EcRow& r, int H, int w, int lane, int frame, int y0, int rows_in_band) {
  const int local = 32 * w - 4 + lane;
  r.frame = frame;
  r.y = y0 + local;
  r.valid = local >= 0 && local < rows_in_band && r.y < H + 7;
  r.real = r.valid && r.y < H;
  r.emit = r.valid && r.y >= 7;
  const int y = min(r.y, H - 1);
  auto sc = [&](int i) { return (i > 0 && i < H - 1) ? 0.5 : 1.0; };
  const int yp = min(y + 1, H - 1), ym = max(y - 1, 0);
  const double s0 = sc(y);
  r.i0 = 2; r.iU = 2 + (yp - y); r.iD = 2 + (ym - y);
  r.i1 = 2 + (min(yp + 1, H - 1) - y); r.i2 = 2 + (max(yp - 1, 0) - y);
  r.i3 = 2 + (min(ym + 1, H - 1) - y); r.i4 = 2 + (max(ym - 1, 0) - y);
  r.m1 = 4.0 * s0 * sc(yp); r.m2 = 4.0 * s0 * sc(ym); r.mB = 2.0 * s0;
  r.interior = r.y >= 2 && r.y <= H - 3;
}

// Geometry and hand-over blocks of the item a lane enters with its next sweep (returned by value: see GRows).
__device__ __noinline__ EcRow ec_row_of_sweep(const FzParams& p, int w, int lane, int jl, int my_items, unsigned char* ring_in,
                                              unsigned char* ring_out) {
  int frame = 0, y0 = 0, rows = 0, band = 0, item = 0;
  if (jl < my_items) {
    const int grp = (int)blockIdx.x / p.bands;
    band = (int)blockIdx.x - grp * p.bands;
    frame = grp + jl * p.groups;
    item = frame * p.bands + band;
    y0 = band * FZ_BR;
    rows = FZ_BR;
  }
  EcRow row;
  ec_set_row(row, p.H, w, lane, frame, y0, rows);
  row.top = (band == 0 && w == 0);
  row.in_base = (w == 0) ? p.ho_items + (size_t)(band > 0 ? item - 1 : item) * ho_bytes((size_t)p.nc_item) : ring_in;
  row.out_base = (w == 3) ? p.ho_items + (size_t)item * ho_bytes((size_t)p.nc_item) : ring_out;
  return row;
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

#ifdef LGX_FZ_PROF
__device__ unsigned long long g_fz_prof[32];
#define FZ_CLK(var) const long long var = clock64()
#define FZ_ACC(slot, t1, t0) prof[slot] += (t1) - (t0)
#else
#define FZ_CLK(var)
#define FZ_ACC(slot, t1, t0)
#endif

// One batch: slots t0 .. t0+7 of an EC warp (t0 a multiple of 8).  GENERAL = false: every lane is in the interior
// columns of the same sweep (16 <= x, x + 7 <= W - 3), no lane changes its item, and the block has neither rows of the
// first 14 image rows nor replicated rows below the image; GENERAL = true: everything.
template <bool MIXED, bool GEN, bool GENERAL>
__device__ __forceinline__ void ec_batch(const FzParams& p, const EcCtx& c, EcRow& row, EcState& st, const int t0, int& ubase, int& jl,
                                         const int my_items, const uint32_t hin_mask, const uint32_t hout_mask, const int hin_nc,
                                         const int hout_nc, unsigned char* ring_in, unsigned char* ring_out, const uint32_t stg
#ifdef LGX_FZ_PROF
                                         , long long* prof
#endif
) {
  const int lane = c.lane;
  const int W = c.W, H = c.H;
  const int ul0 = t0 - lane;                         // global column coordinate of pixel 0
  const int rel = lane - c.first;                    // row of the hand-over block this lane reads (rs: rel < 14, b: rel < 7)
  FZ_CLK(tp0);
  if constexpr (!GENERAL) {
    // hand-over values of the block above for the eight pixels, fetched from L2 into the warp's staging while phases 1-2 run
    if (rel >= 0 && rel < 14) {
      const uint32_t dst = stg + (uint32_t)rel * FZ_STG_LANE;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t k = (uint32_t)(hin_mask == 0xffffffffu ? (ul0 + i - ubase + 4) : (ul0 + i - 8)) & hin_mask;
        const unsigned char* src = row.in_base + ho_off_rs((size_t)hin_nc, rel, k);
        cp_async16(dst + 32u * i, src);
        if (rel < 7) cp_async16(dst + 32u * i + 16u, src + 16);
      }
    }
    if (rel == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t k = (uint32_t)(hin_mask == 0xffffffffu ? (ul0 + i - ubase + 4) : (ul0 + i - 8)) & hin_mask;
        cp_async16(stg + FZ_STG_SUM + 16u * i, row.in_base + ho_off_sum((size_t)hin_nc, k));
      }
    }
  }
  GRows gr;
#pragma unroll
  for (int k = 0; k < 5; ++k) gr.off[k] = c.rowoff[k];
  gr.halo = c.halo;
  // ---- g window of the batch: columns ul0 - 2 .. ul0 + 9 of the lane's rows (contiguous: the rings repeat their first 12 columns)
  const uint32_t c63 = 8u * (uint32_t)((ul0 - 2) & 63), c127 = 8u * (uint32_t)((ul0 - 2) & 127);
  auto rowptr = [&](int k) { return c.rowoff[k] + (((c.halo >> k) & 1u) ? c127 : c63); };
  uint32_t r0, rU, rD, a1, a2, a3, a4;
  if constexpr (!GEN) {
    r0 = rowptr(2); rU = rowptr(3); rD = rowptr(1); a1 = rowptr(4); a4 = rowptr(0); a2 = a3 = r0;
  } else {
    // per-lane row choice (rows 0, 1, H-2, H-1): select among the lane's five rows
    uint32_t rp[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) rp[k] = rowptr(k);
    auto pick = [&](int i) { return i == 0 ? rp[0] : i == 1 ? rp[1] : i == 2 ? rp[2] : i == 3 ? rp[3] : rp[4]; };
    r0 = pick(row.i0); rU = pick(row.iU); rD = pick(row.iD);
    a1 = pick(row.i1); a2 = pick(row.i2); a3 = pick(row.i3); a4 = pick(row.i4);
  }
  // element e of a row pointer <-> column ul0 - 2 + e, i.e. pixel i's column x_i is element i + 2
  auto L = [&](uint32_t base, int e) { return lds_f64(base + 8u * (uint32_t)e); };

  // ---- phase 1: the eight eigenvalue ingredients, branch free ----------------------------------------------------
  double bv[8], S[8], X[8];
  {
    double o_0 = L(r0, 2), o_p1 = L(r0, 3);
    double gc_m1 = __dsub_rn(o_0, L(r0, 0));          // 2*g_c(x-1) = g(x) - g(x-2)
    double gc_0 = __dsub_rn(o_p1, L(r0, 1));          // 2*g_c(x)
    double u_m1, u_0, d_m1, d_0;
    if (!MIXED) {
      u_m1 = __dsub_rn(L(rU, 1), L(rD, 1));
      u_0 = __dsub_rn(L(rU, 2), L(rD, 2));
      d_m1 = d_0 = 0.0;
    } else {
      u_m1 = L(rU, 1); u_0 = L(rU, 2);
      d_m1 = L(rD, 1); d_0 = L(rD, 2);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const double gn = L(r0, i + 4);                  // g(y, x+2)
      const double gc_p1 = __dsub_rn(gn, o_0);         // 2*g_c(x+1)
      const double C = __dsub_rn(gc_p1, gc_m1);        // 4*Hcc
      const double up = L(rU, i + 3), dn = L(rD, i + 3);
      double B;
      if (!MIXED) {
        const double u_p1 = __dsub_rn(up, dn);
        B = __dsub_rn(u_p1, u_m1);
        u_m1 = u_0; u_0 = u_p1;
      } else {
        B = __dsub_rn(__dsub_rn(up, u_m1), __dsub_rn(dn, d_m1));   // 2*g_c(yp, x) - 2*g_c(ym, x)
        u_m1 = u_0; u_0 = up; d_m1 = d_0; d_0 = dn;
      }
      double A;
      if constexpr (!GEN) {
        A = __dsub_rn(__dsub_rn(L(a1, i + 2), o_0), __dsub_rn(o_0, L(a4, i + 2)));
      } else {
        A = __dsub_rn(__dmul_rn(__dsub_rn(L(a1, i + 2), L(a2, i + 2)), row.m1), __dmul_rn(__dsub_rn(L(a3, i + 2), L(a4, i + 2)), row.m2));
        B = __dmul_rn(B, row.mB);
      }
      gc_m1 = gc_0; gc_0 = gc_p1;
      o_0 = o_p1; o_p1 = gn;
      // A, B, C = 4*Hrr, 4*Hrc, 4*Hcc:  b = ((A + C) - sqrt(4*B*B + (A - C)^2)) / 8
      S[i] = __dadd_rn(A, C);
      const double D = __dsub_rn(A, C);
      X[i] = __fma_rn(4.0, __dmul_rn(B, B), __dmul_rn(D, D));
    }
  }
  // ring slots: pixel i lives in slot (t0 + i) & 15 = h + i;  b(x - 15) and b(y - 7, x - 8) of lane l-7: slot t - 15;
  // rs(y - 14, x - 8) of lane l-14: slot t - 14
  const uint32_t h = (uint32_t)(t0 & 8);
  const uint32_t bs_own = c.bs + 8u * (uint32_t)lane, bs_up = c.bs + 8u * (uint32_t)((lane + 25) & 31);   // lane - 7
  const uint32_t rs_own = c.rs + 16u * (uint32_t)lane, rs_up = c.rs + 16u * (uint32_t)((lane + 18) & 31);  // lane - 14
  auto slot15 = [&](int i) { return (uint32_t)((h + i + 1) & 15); };
  auto slot14 = [&](int i) { return (uint32_t)((h + i + 2) & 15); };
  double bold[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) bold[i] = lds_f64(bs_own + 256u * slot15(i));

  // ---- phase 2: the eight square roots ---------------------------------------------------------------------------
  unsigned worst = 0;
  {
    double R[8];
    sqrt_inrange<8>(X, R, worst);
#pragma unroll
    for (int i = 0; i < 8; ++i) bv[i] = __dmul_rn(__dsub_rn(S[i], R[i]), 0.125);
  }
  FZ_CLK(tp2);
  if constexpr (!GENERAL) cp_async_wait_all();
  FZ_CLK(tp2w);
  // ---- phase 3: per pixel, in slot order: row chain, hand-over, column chain -------------------------------------
  double m_[8], v_[8], bc_[8];
  unsigned emitmask = 0;
  int ub = ubase;
  const uint32_t stg_me = stg + (uint32_t)(rel & 15) * FZ_STG_LANE;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    __syncwarp();                                      // ring entries of earlier slots are visible; lanes are converged
    const int ul = ul0 + i;
    bool act;
    if constexpr (GENERAL) {
      if (ul - ub >= c.SW) {
        // the lane enters its next sweep: geometry and hand-over blocks of the new item, chains restart
        ub += c.SW;
        ++jl;
        row = ec_row_of_sweep(p, c.w, lane, jl, my_items, ring_in, ring_out);
        st.chain_b = st.chain_q = st.b0 = st.blast = 0.0;
        st.word = 0;
      }
      act = ul >= 0 && row.valid;
    } else {
      act = row.valid;
    }
    const int u = ul - ub;                             // column counter of the sweep: x = u - 12
    const int x = u - 12;
    const int pcol = x - 8;                            // column of the row sums and of the output
    const bool inimg = act && row.real && x >= 0 && x < W;
    // -- b of this pixel
    double b = bv[i];
    if constexpr (GENERAL) {
      if (inimg && (worst >= kSqrtOutOfRange || x < 2 || x > W - 3)) b = b_generic_f(gr, H, W, row.y, x, ul, MIXED);
      if (inimg) {
        st.blast = b;
        if (x == 0) st.b0 = b;
      } else {
        b = st.blast;                                  // right of the image: b(min(c + 7, W - 1)); elsewhere unused
      }
    } else {
      if (worst >= kSqrtOutOfRange) b = b_generic_f(gr, H, W, row.y, x, ul, MIXED);
    }
    if (p.dbg_b && inimg) p.dbg_b[(size_t)row.frame * p.plane_stride + (size_t)row.y * p.Wp + x] = b;
    // -- row chain: rs = s(x - 8); then s(x - 7)
    const double rs_b = st.chain_b, rs_q = st.chain_q;
    if constexpr (GENERAL) {
      const int cc = x - 7;
      if (act && row.real) {
        if (cc == 0) {
          // cv2 RowSum start: the 15 replicated-border terms accumulated left to right from 0.0: 8 x b(0), b(1) .. b(7)
          double sb = 0.0, sq = 0.0;
#pragma unroll 1
          for (int k = 0; k < 15; ++k) {
            const int bi = k > 7 ? k - 7 : 0;
            const double v = bi == 0 ? st.b0 : (bi == 7 ? b : lds_f64(bs_own + 256u * (uint32_t)((t0 + i - (7 - bi)) & 15)));
            sb = __dadd_rn(sb, v);
            sq = __dadd_rn(sq, __dmul_rn(v, v));
          }
          st.chain_b = sb; st.chain_q = sq;
        } else if (cc > 0 && cc < W) {
          const double old = (x - 15 <= 0) ? st.b0 : bold[i];
          st.chain_b = __dadd_rn(st.chain_b, __dsub_rn(b, old));
          st.chain_q = __dadd_rn(st.chain_q, __dsub_rn(__dmul_rn(b, b), __dmul_rn(old, old)));
        }
      }
    } else {
      st.chain_b = __dadd_rn(st.chain_b, __dsub_rn(b, bold[i]));
      st.chain_q = __dadd_rn(st.chain_q, __dsub_rn(__dmul_rn(b, b), __dmul_rn(bold[i], bold[i])));
    }
    // -- this pixel's ring entry of b (read 15 slots from now)
    sts_f64(bs_own + 256u * (uint32_t)((h + i) & 15), b);
    // -- column role at column pcol
    const bool pv = act && pcol >= 0 && pcol < W;
    // hand-over column index: per-item arrays are indexed by the sweep's column counter, the CTA's rings by the global one
    const uint32_t hc_in = (uint32_t)(hin_mask == 0xffffffffu ? (u + 4) : (ul - 8)) & hin_mask;
    const uint32_t hc_out = (uint32_t)(hout_mask == 0xffffffffu ? (u + 4) : (ul - 8)) & hout_mask;
    const bool has_in = !row.top;
    // new value of the column chain: the row sum of this row, or (replicated rows below the image) the one of the row above
    double nw_b = rs_b, nw_q = rs_q;
    if constexpr (GENERAL) {
      const double upb = __shfl_up_sync(0xffffffffu, st.newp_b, 1), upq = __shfl_up_sync(0xffffffffu, st.newp_q, 1);
      if (act && !row.real) {
        if (lane > c.first) { nw_b = upb; nw_q = upq; }
        else if (pv) {
          const double2 v = __ldcg(reinterpret_cast<const double2*>(row.in_base + ho_off_rs((size_t)hin_nc, 13, hc_in)));
          nw_b = v.x; nw_q = v.y;
        }
      }
      if (pv) { st.newp_b = nw_b; st.newp_q = nw_q; }
    }
    // rs(y - 14) and b(y - 7): from the rings of this warp, or from the hand-over of the block above
    double old_b, old_q, bcmp;
    {
      const double2 ro = lds_f64x2(rs_up + 512u * slot14(i));
      old_b = ro.x; old_q = ro.y;
      bcmp = lds_f64(bs_up + 256u * slot15(i));
      if constexpr (!GENERAL) {
        if (rel < 14) {
          const double2 v = lds_f64x2(stg_me + 32u * i);
          old_b = v.x; old_q = v.y;
          if (rel < 7) bcmp = lds_f64(stg_me + 32u * i + 16u);
        }
      } else {
        if (rel >= 0 && rel < 14 && has_in && pv) {
          const unsigned char* src = row.in_base + ho_off_rs((size_t)hin_nc, rel, hc_in);
          const double2 v = __ldcg(reinterpret_cast<const double2*>(src));
          old_b = v.x; old_q = v.y;
          if (rel < 7) bcmp = __ldcg(reinterpret_cast<const double*>(src + 16));
        }
        if (row.top && row.y < 14) {
          // first rows of the image: cv2 replicates row 0 above the image, so the "row 14 up" is row 0 (the block's first
          // lane had this column row.y slots ago); rows 0..6 only accumulate
          const double2 r0v = lds_f64x2(c.rs + 16u * (uint32_t)c.first + 512u * (uint32_t)((t0 + i - row.y) & 15));
          old_b = row.y >= 7 ? r0v.x : 0.0;
          old_q = row.y >= 7 ? r0v.y : 0.0;
        }
      }
    }
    // running sums of the lane above for this column (previous slot), or of the block above
    double sin_b = __shfl_up_sync(0xffffffffu, st.sum_b, 1), sin_q = __shfl_up_sync(0xffffffffu, st.sum_q, 1);
    if (lane == c.first) {
      if constexpr (!GENERAL) {
        const double2 v = lds_f64x2(stg + FZ_STG_SUM + 16u * i);
        sin_b = v.x; sin_q = v.y;
      } else {
        sin_b = 0.0; sin_q = 0.0;
        if (has_in && pv) {
          const double2 v = __ldcg(reinterpret_cast<const double2*>(row.in_base + ho_off_sum((size_t)hin_nc, hc_in)));
          sin_b = v.x; sin_q = v.y;
        }
      }
    }
    double s0b, s0q;
    if (GENERAL && row.top && row.y == 0) {
      // ColumnSum start: the first 14 padded rows are row 0 eight times, then rows 1..6
      s0b = 0.0; s0q = 0.0;
#pragma unroll 1
      for (int k = 0; k < 8; ++k) { s0b = __dadd_rn(s0b, nw_b); s0q = __dadd_rn(s0q, nw_q); }
      old_b = old_q = 0.0;
    } else {
      s0b = __dadd_rn(sin_b, nw_b);
      s0q = __dadd_rn(sin_q, nw_q);
    }
    if (pv) {
      st.sum_b = __dsub_rn(s0b, old_b);
      st.sum_q = __dsub_rn(s0q, old_q);
    }
    // ring entry of the row sums (read 14 slots from now by lane l + 14), hand-over for the block below
    sts_f64x2(rs_own + 512u * (uint32_t)((h + i) & 15), nw_b, nw_q);
    if (pv) {
      if (lane >= 18) __stcg(reinterpret_cast<double2*>(row.out_base + ho_off_rs((size_t)hout_nc, lane - 18, hc_out)), make_double2(nw_b, nw_q));
      if (lane == 31) __stcg(reinterpret_cast<double2*>(row.out_base + ho_off_sum((size_t)hout_nc, hc_out)), make_double2(st.sum_b, st.sum_q));
    }
    if (lane >= 25 && inimg) {
      const uint32_t hcb = (uint32_t)(hout_mask == 0xffffffffu ? (u + 12) : ul) & hout_mask;
      __stcg(reinterpret_cast<double*>(row.out_base + ho_off_rs((size_t)hout_nc, lane - 25, hcb) + 16), b);
    }
    // mean, variance (threshold and compare after the batched square roots)
    const double m = __dmul_rn(s0b, 1.0 / 225);
    const double msq = __dmul_rn(s0q, 1.0 / 225);
    double var = __dsub_rn(msq, __dmul_rn(m, m));
    if (var < 0) var = 0.0;
    m_[i] = m; v_[i] = var; bc_[i] = bcmp;
    if (pv && row.emit) emitmask |= 1u << i;
  }
  ubase = ub;
  FZ_CLK(tp3);
  // ---- phase 4: thresholds, compares, bit words ------------------------------------------------------------------
  {
    double sd[8];
    unsigned worst2 = 0;
    sqrt_inrange<8>(v_, sd, worst2);
    if (worst2 >= kSqrtOutOfRange) {
#pragma unroll
      for (int i = 0; i < 8; ++i) sd[i] = __dsqrt_rn(v_[i]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      // mean * (1 + 0.5 * ((sd / 128) - 1))
      const double T = __dmul_rn(m_[i], __dadd_rn(1.0, __fma_rn(sd[i], 0.00390625, -0.5)));
      const bool white = !(bc_[i] > T);               // binary = 255 - 255 * (b > T)
      if ((emitmask >> i) & 1u) {
        // (a batch in which the lane changes its sweep emits nothing: the last 8 columns of a sweep are past the image and
        // the first 8 are left of it, so the sweep base and the row of the end of the batch are right for every emitting pixel)
        const int pcol = (ul0 + i - ubase) - 20;
        st.word = (st.word >> 1) | (white ? 0x80000000u : 0u);
        const int yo = row.y - 7;
        if (p.dbg_T) p.dbg_T[(size_t)row.frame * p.plane_stride + (size_t)yo * p.Wp + pcol] = T;
        if ((pcol & 31) == 31 || pcol == W - 1) {
          p.bits[((size_t)row.frame * H + yo) * p.WW + (pcol >> 5)] = st.word >> (31 - (pcol & 31));
          st.word = 0;
        }
      }
    }
  }
  FZ_CLK(tp4);
  FZ_ACC(0, tp2, tp0); FZ_ACC(1, tp2w, tp2); FZ_ACC(2, tp3, tp2w); FZ_ACC(3, tp4, tp3);
}

template <typename PIX, bool MIXED, bool DIV>
__global__ void __launch_bounds__(FZ_THREADS, 1) ridge_fused_kernel(const __grid_constant__ FzParams p) {
  constexpr int TILE_IN = fz_tile_in<PIX>();
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = sbase + FO_BAR;
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  double* s_lut = reinterpret_cast<double*>(smem + FO_LUT);
  unsigned char* s_in = smem + FO_IN;
  double* s_v = reinterpret_cast<double*>(smem + fo_v<PIX>());
  double* s_g = reinterpret_cast<double*>(smem + fo_g<PIX>());
  double* s_hl = reinterpret_cast<double*>(smem + fo_hl<PIX>());

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int w = warp & 3;
  const int H = p.H, W = p.W;
  const int nsteps = p.nsteps;
  const int SW = 32 * nsteps;
  // CTA = (group, band): band `my_band` of the frames grp, grp + groups, ...  The band above runs on CTA blockIdx.x - 1
  // on the same frame at the same time, a fixed distance ahead: no CTA ever waits for more than that distance.
  const int grp = (int)blockIdx.x / p.bands;
  const int my_band = (int)blockIdx.x - grp * p.bands;
  const int my_items = grp < p.batch ? (p.batch - 1 - grp) / p.groups + 1 : 0;
  const int total = my_items * nsteps;                 // 32-column steps of this CTA
  const int total_slots = my_items * SW + 32;          // slots of an EC warp: the last lane is 31 columns behind
#ifdef LGX_FZ_PROF
  long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const long long t_begin = clock64();
#endif

  if (tid == 0) {
    for (int s = 0; s < 4; ++s) mbar_init(BAR(FB_IN + s), 1);
    for (int s = 0; s < 32; ++s) { mbar_init(BAR(FB_FULL_G + s), 32); mbar_init(BAR(FB_EMPTY_G + s), 1); }
    for (int s = 0; s < 24; ++s) { mbar_init(BAR(FB_FULL_H + s), 4); mbar_init(BAR(FB_EMPTY_H + s), 1); }
    for (int s = 0; s < 4; ++s) reinterpret_cast<volatile int*>(smem + FO_PROG)[s] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (sizeof(PIX) == 1)
    for (int i = tid; i < 256; i += FZ_THREADS) s_lut[i] = p.lut[i];
  // rings start as zeros (columns left of a sweep's first column, rows never written)
  for (int i = tid; i < 4 * FZ_GW + 3 * FZ_HW; i += FZ_THREADS) s_g[i] = 0.0;
  for (int i = tid; i < 4 * 16 * 32 * 3; i += FZ_THREADS) reinterpret_cast<double*>(smem + fo_bs<PIX>())[i] = 0.0;
  __syncthreads();

  if (warp < 4) {
    // ============================================================ VH_w: vertical then horizontal 25-tap of block w
    unsigned char* my_in = s_in + w * TILE_IN;
    double* vb = s_v + w * FZ_VB;
    double* G = s_g + w * FZ_GW;
    double* HL = (w < 3) ? s_hl + w * FZ_HW : nullptr;   // halo ring read by EC_{w+1}: this block's last four rows
    auto issue_load = [&](int kgx) {                     // tile of global step kgx (lane 0 only)
      const int jx = kgx / nsteps, kx = kgx - jx * nsteps;
      mbar_expect_tx(BAR(FB_IN + w), TILE_IN);
      // gaussian row r of the block <-> image row y0 - 2 + 32 w + r; it needs blurred rows -12 .. +12 around it
      tma_load_3d(&p.tm_in, BAR(FB_IN + w), smem_u32(my_in), 32 * kx, my_band * FZ_BR - 2 - kRadius + 32 * w, grp + jx * p.groups);
    };
    if (lane == 0 && total > 0) issue_load(0);
    double hin[24 + WG2];
    int k = 0;
    for (int kg = 0; kg < total; ++kg) {
      FZ_CLK(tv0);
      mbar_wait(BAR(FB_IN + w), kg & 1);
      FZ_CLK(tv1);
      // ---- vertical: lane = column, the window slides down the 32 rows
      {
        const PIX* tile = reinterpret_cast<const PIX*>(my_in) + lane;
        double in[24 + WG2];
#pragma unroll
        for (int i = 0; i < 24; ++i) in[i] = px_to_ff<PIX, DIV>(s_lut, tile[i * 32]);
#pragma unroll 1
        for (int grp2 = 0; grp2 < 32 / WG2; ++grp2) {
          const PIX* tg = tile + (24 + WG2 * grp2) * 32;
          double* vg = vb + WG2 * grp2 * 33 + lane;
#pragma unroll
          for (int i = 0; i < WG2; ++i) in[24 + i] = px_to_ff<PIX, DIV>(s_lut, tg[i * 32]);
#pragma unroll
          for (int q = 0; q < WG2; ++q) vg[q * 33] = tap25f(in + q, p.w);
#pragma unroll
          for (int i = 0; i < 24; ++i) in[i] = in[i + WG2];
        }
      }
      __syncwarp();                                      // the tile has been read, the v block is complete
      if (lane == 0 && kg + 1 < total) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        issue_load(kg + 1);                              // lands while the horizontal pass runs
      }
      FZ_CLK(tv2);
      // ---- horizontal: lane = row, the window lives in registers across the sweep; output column c of the step is
      // image column 32k - 12 + c = global column coordinate 32 kg + c
      if (k == 0) {
#pragma unroll
        for (int i = 0; i < 24; ++i) hin[i] = 0.0;       // new sweep: v(x < 0) = 0
      }
      const double* vrow = vb + lane * 33;
#pragma unroll 1
      for (int grp2 = 0; grp2 < 32 / WG2; ++grp2) {
        const int gg = 2 * kg + grp2;                    // global 16-column group (halo ring: 8 groups)
        if (HL && gg >= 8) {
          FZ_CLK(tw0);
          mbar_wait(BAR(FB_EMPTY_H + w * 8 + (gg & 7)), ((gg >> 3) & 1) ^ 1);
          FZ_CLK(tw1);
          FZ_ACC(2, tw1, tw0);
        }
        const double* vg = vrow + WG2 * grp2;
        double* hl_out = HL ? HL + (lane & 3) * FZ_HP + 16 * (gg & 7) : nullptr;
#pragma unroll
        for (int i = 0; i < WG2; ++i) hin[24 + i] = vg[i];
#pragma unroll
        for (int half = 0; half < WG2 / WH2; ++half) {
          const int g8 = 2 * gg + half;                  // global 8-column group (g ring: 8 groups)
          const int gi = g8 & 7;
          if (g8 >= 8) {
            FZ_CLK(tw0);
            mbar_wait(BAR(FB_EMPTY_G + w * 8 + gi), ((g8 >> 3) & 1) ^ 1);
            FZ_CLK(tw1);
            FZ_ACC(2, tw1, tw0);
          }
          double* gg_out = G + (4 + lane) * FZ_GP + WH2 * gi;
#pragma unroll
          for (int q = 0; q < WH2; ++q) {
            const double val = tap25f(hin + WH2 * half + q, p.w);
            gg_out[q] = val;
            if (gi == 0 || (gi == 1 && q < 4)) gg_out[64 + q] = val;   // the first 12 columns of the ring again behind its end
            if (hl_out && lane >= 28) {
              hl_out[WH2 * half + q] = val;
              if ((gg & 7) == 0 && WH2 * half + q < 12) hl_out[128 + WH2 * half + q] = val;
            }
          }
          mbar_arrive(BAR(FB_FULL_G + w * 8 + gi));
        }
#pragma unroll
        for (int i = 0; i < 24; ++i) hin[i] = hin[i + WG2];
        if (HL && lane >= 28) mbar_arrive(BAR(FB_FULL_H + w * 8 + (gg & 7)));
      }
      if (++k == nsteps) k = 0;
      FZ_CLK(tv3);
      FZ_ACC(0, tv1, tv0); FZ_ACC(1, tv2, tv1); FZ_ACC(3, tv3, tv2);
    }
#ifdef LGX_FZ_PROF
    if (lane == 0) {
      prof[3] -= prof[2];                                // horizontal pass without its waits
      for (int i = 0; i < 4; ++i) atomicAdd(&g_fz_prof[i], (unsigned long long)prof[i]);
      atomicAdd(&g_fz_prof[4], (unsigned long long)(clock64() - t_begin));
      atomicAdd(&g_fz_prof[5], 1ull);
    }
#endif
  } else {
    // ============================================================ EC_w
    EcCtx c;
    c.H = H; c.W = W; c.SW = SW; c.lane = lane; c.w = w; c.first = (w == 0) ? 4 : 0;
    c.halo = 0;
#pragma unroll
    for (int kk = 0; kk < 5; ++kk) {
      const int i = lane + kk;                          // window row
      if (i < 4 && w > 0) {
        c.rowoff[kk] = sbase + fo_hl<PIX>() + 8u * (uint32_t)((w - 1) * FZ_HW + i * FZ_HP);
        c.halo |= 1u << kk;
      } else {
        c.rowoff[kk] = sbase + fo_g<PIX>() + 8u * (uint32_t)(w * FZ_GW + max(i, 4) * FZ_GP);
      }
    }
    c.bs = sbase + fo_bs<PIX>() + (uint32_t)w * 16u * 32u * 8u;
    c.rs = sbase + fo_rs<PIX>() + (uint32_t)w * 16u * 32u * 16u;
    const uint32_t stg = sbase + fo_stg<PIX>() + (uint32_t)w * FZ_STG;
    const uint32_t prog_me = sbase + FO_PROG + 4u * (uint32_t)w;
    const uint32_t prog_up = sbase + FO_PROG + 4u * (uint32_t)(w - 1);
    const uint32_t prog_dn = sbase + FO_PROG + 4u * (uint32_t)(w + 1);
    // hand-over rings of this CTA: ring r (0..2) is written by block r and read by block r + 1
    unsigned char* ring_base = p.ho_rings + (size_t)blockIdx.x * 3 * ho_bytes(FZ_RINGCOLS);

    EcRow row;
    EcState st;
    unsigned char* ring_in = ring_base + (size_t)(w > 0 ? w - 1 : 0) * ho_bytes(FZ_RINGCOLS);
    unsigned char* ring_out = ring_base + (size_t)(w < 3 ? w : 0) * ho_bytes(FZ_RINGCOLS);
    row = ec_row_of_sweep(p, w, lane, my_items, my_items, ring_in, ring_out);   // not started: no valid row
    int ubase = -SW, jl = -1;                           // the lane's sweep base (global column coordinate of u = 0) and item index
    int g_ready = 0, g_freed = 0, h_ready = 0, h_freed = 0;
    int up_seen = 0, dn_seen = 0, band_seen = 0, band_j = -1;
    const int last_g8 = 4 * total - 1, last_g16 = 2 * total - 1;
    const int in_nc = (w == 0) ? p.nc_item : FZ_RINGCOLS, out_nc = (w == 3) ? p.nc_item : FZ_RINGCOLS;
    const uint32_t in_mask = (w == 0) ? 0xffffffffu : (uint32_t)(FZ_RINGCOLS - 1);
    const uint32_t out_mask = (w == 3) ? 0xffffffffu : (uint32_t)(FZ_RINGCOLS - 1);
    const int item_stride = p.groups * p.bands;         // distance of this CTA's consecutive items in the item arrays

    for (int t0 = 0; t0 < total_slots; t0 += 8) {
      // ---- the sweep of the block's most advanced lane decides what must be ready
      const int jlead = min(t0 / SW, my_items - 1);
      const int ts = t0 - jlead * SW;                    // its column counter at pixel 0 (lane 0; the first lane of block 0 is lane 4)
      const int item_l = (grp + jlead * p.groups) * p.bands + my_band;
      FZ_CLK(te0);
      {
        const int need_g = min((t0 + 9) >> 3, last_g8);
        while (g_ready <= need_g) { mbar_wait(BAR(FB_FULL_G + w * 8 + (g_ready & 7)), (g_ready >> 3) & 1); ++g_ready; }
        if (w > 0) {
          const int need_h = min((t0 + 9) >> 4, last_g16);
          while (h_ready <= need_h) { mbar_wait(BAR(FB_FULL_H + (w - 1) * 8 + (h_ready & 7)), (h_ready >> 3) & 1); ++h_ready; }
        }
      }
      FZ_CLK(te1);
      if (w > 0) {
        // block w-1 must be 40 slots ahead (its lane 31 has passed the columns of this batch)
        const int need = min(t0 + FZ_LAG, total_slots);
        if (up_seen < need) {
          SpinGuard sg;
          while ((up_seen = ld_acquire_cta_shared(prog_up)) < need) sg.tick();
        }
      } else if (my_band > 0 && t0 < my_items * SW) {
        // the band above (another CTA): its EC_3 must have published the columns of this batch
        if (band_j != jlead) { band_j = jlead; band_seen = 0; }
        const int need = min(ts + 8, SW);
        if (band_seen < need) {
          int v = 0;
          if (lane == 0) {
            SpinGuard sg;
            while ((v = ld_acquire_gpu(p.prog + item_l - 1)) < need) sg.tick();
          }
          band_seen = __shfl_sync(0xffffffffu, v, 0);
        }
      }
      FZ_CLK(te2);
      if (w < 3) {
        // back-pressure: the ring this block writes has 128 columns; the block below may be at most ~100 slots behind
        const int need = t0 - 96;
        if (dn_seen < need) {
          SpinGuard sg;
          while ((dn_seen = ld_acquire_cta_shared(prog_dn)) < need) sg.tick();
        }
      }
      FZ_CLK(te3);
      // ---- interior batch (every lane in the same sweep with 16 <= x and x + 7 <= W - 3, plain rows) or general
      const int y_blk = my_band * FZ_BR + 32 * w - 4;    // image row of lane 0
      const bool interior = t0 < my_items * SW && ts >= 59 && ts <= W + 2 && y_blk + c.first >= 14 && y_blk + 31 <= H - 3;
#ifdef LGX_FZ_PROF
#define FZ_PROF_ARG , prof + 4
#else
#define FZ_PROF_ARG
#endif
      if (interior) {
        ec_batch<MIXED, false, false>(p, c, row, st, t0, ubase, jl, my_items, in_mask, out_mask, in_nc, out_nc, ring_in, ring_out, stg FZ_PROF_ARG);
      } else {
        ec_batch<MIXED, true, true>(p, c, row, st, t0, ubase, jl, my_items, in_mask, out_mask, in_nc, out_nc, ring_in, ring_out, stg FZ_PROF_ARG);
      }
      __syncwarp();
      FZ_CLK(te4);
      // ---- release what this batch has finished with, publish progress
      if (lane == 0) {
        const int freeg = (t0 + 8 - 31 - 2) >> 3;        // groups below the column lane 31 needs next
        while (g_freed < freeg && g_freed <= last_g8) { mbar_arrive(BAR(FB_EMPTY_G + w * 8 + (g_freed & 7))); ++g_freed; }
        if (w > 0) {
          const int freeh = (t0 + 8 - 3 - 2) >> 4;
          while (h_freed < freeh && h_freed <= last_g16) { mbar_arrive(BAR(FB_EMPTY_H + (w - 1) * 8 + (h_freed & 7))); ++h_freed; }
        }
        st_release_cta_shared(prog_me, t0 + 8);
        if (w == 3 && (((t0 + 8) % FZ_PUBLISH) == 0 || t0 + 8 >= total_slots)) {
          // number of columns of its sweep that lane 31 has completed, for the item lane 31 is in; the item before it is complete
          const int done = t0 + 8 - 32;                  // lane 31 has completed the global column coordinates <= done
          if (done >= 0) {
            const int jj = min(done / SW, my_items - 1);
            const int itemj = (grp + jj * p.groups) * p.bands + my_band;
            if (jj > 0) st_release_gpu(p.prog + itemj - item_stride, SW);
            st_release_gpu(p.prog + itemj, min(done - jj * SW + 1, SW));
          }
        }
      }
      FZ_CLK(te5);
      FZ_ACC(0, te1, te0); FZ_ACC(1, te2, te1); FZ_ACC(2, te3, te2); FZ_ACC(3, te5, te4);
    }
#ifdef LGX_FZ_PROF
    if (lane == 0) {
      // [8..11] waits: g / up or band / down / release;  [12..15] phases 1-2 / cp.async wait / phase 3 / phase 4
      for (int i = 0; i < 8; ++i) atomicAdd(&g_fz_prof[8 + i], (unsigned long long)prof[i]);
      atomicAdd(&g_fz_prof[16], (unsigned long long)(clock64() - t_begin));
      atomicAdd(&g_fz_prof[17], 1ull);
      if (w == 0) atomicAdd(&g_fz_prof[18], (unsigned long long)prof[1]);
    }
#endif
  }
}

template <typename PIX, bool MIXED, bool DIV>
cudaError_t fz_launch_t(const FzParams& p, int ctas, cudaStream_t stream) {
  static unsigned long long attr_done = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  constexpr int smem = fz_smem_bytes<PIX>();
  auto kern = ridge_fused_kernel<PIX, MIXED, DIV>;
  if (!(attr_done >> (dev & 63) & 1ull)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    attr_done |= 1ull << (dev & 63);
  }
  void* args[] = {const_cast<FzParams*>(&p)};
  // cooperative: every CTA is resident, which the waits between bands (and between the blocks of a band) rely on
  return cudaLaunchCooperativeKernel((const void*)kern, dim3(ctas), dim3(FZ_THREADS), args, smem, stream);
}

}  // namespace

bool ridge_fused_usable(const RidgeParams& rp, int bits) {
  auto a16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  return rp.W >= 64 && rp.H >= 16 && a16(rp.blur) && ((size_t)rp.blur_pitch * (bits / 8)) % 16 == 0 && encode_fn() != nullptr;
}

size_t fused_item_bytes(int W) { return ho_bytes((size_t)((W + 64 + 7) & ~7)); }
size_t fused_ring_bytes(int ctas) { return (size_t)ctas * 3 * ho_bytes(FZ_RINGCOLS); }
int fused_bands(int H) { return (H + 7 + FZ_BR - 1) / FZ_BR; }

void fused_prof_read(unsigned long long* out32, bool reset) {
#ifdef LGX_FZ_PROF
  cudaMemcpyFromSymbol(out32, g_fz_prof, sizeof(unsigned long long) * 32);
  if (reset) {
    unsigned long long z[32] = {};
    cudaMemcpyToSymbol(g_fz_prof, z, sizeof(z));
  }
#else
  (void)reset;
  for (int i = 0; i < 32; ++i) out32[i] = 0;
#endif
}

cudaError_t launch_ridge_fused(const RidgeParams& rp, const FusedParams& fp, int bits, int batch, int max_ctas, cudaStream_t stream) {
  FzParams p;
  const size_t psz = (size_t)bits / 8;
  const size_t in_row = (size_t)rp.blur_pitch * psz;
  if (!encode3(&p.tm_in, bits == 8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_UINT16, (int)psz, rp.blur, rp.W, rp.H, batch,
               in_row, in_row * rp.H, 32, FZ_VR, CU_TENSOR_MAP_SWIZZLE_NONE))
    return cudaErrorInvalidValue;
  p.H = rp.H; p.W = rp.W; p.WW = bits_pitch(rp.W);
  p.bands = fused_bands(rp.H);
  p.nitems = p.bands * batch;
  p.nsteps = (rp.W + 59) / 32;     // a sweep is 32 * nsteps >= W + 28 columns: x = -12 .. W + 15 at least
  p.bits = fp.bits;
  p.ho_items = fp.ho_items;
  p.ho_rings = fp.ho_rings;
  p.nc_item = (rp.W + 64 + 7) & ~7;
  p.prog = fp.prog;
  p.lut = rp.lut;
  p.dbg_b = fp.dbg_b; p.dbg_T = fp.dbg_T;
  p.Wp = rp.Wp; p.plane_stride = rp.plane_stride;
  for (int i = 0; i < 13; ++i) p.w[i] = rp.w[i];
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (max_ctas > 0 && max_ctas < sms) sms = max_ctas;
  // one CTA per (group, band): every group works on one frame at a time, its bands as a wavefront on adjacent CTAs
  int groups = sms / p.bands;
  if (groups > batch) groups = batch;
  if (groups < 1) return cudaErrorInvalidValue;
  p.batch = batch; p.groups = groups;
  const int ctas = groups * p.bands;
  if (ctas > fp.ring_ctas) return cudaErrorInvalidValue;
  cudaError_t e = cudaMemsetAsync(p.prog, 0, (size_t)p.nitems * sizeof(int), stream);
  if (e != cudaSuccess) return e;
  if (bits == 8)
    return rp.mixed_from_cols ? fz_launch_t<uint8_t, true, false>(p, ctas, stream) : fz_launch_t<uint8_t, false, false>(p, ctas, stream);
  if (rp.float_div)
    return rp.mixed_from_cols ? fz_launch_t<uint16_t, true, true>(p, ctas, stream) : fz_launch_t<uint16_t, false, true>(p, ctas, stream);
  return rp.mixed_from_cols ? fz_launch_t<uint16_t, true, false>(p, ctas, stream) : fz_launch_t<uint16_t, false, false>(p, ctas, stream);
}

}  // namespace lgx
