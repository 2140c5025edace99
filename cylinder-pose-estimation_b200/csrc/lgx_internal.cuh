// lgx internal declarations shared by the kernel translation units.
// Compiled with -fmad=false: every f64 operation on this path must round exactly like the
// reference's NumPy / SciPy / OpenCV scalar code (no FMA contraction), see DESIGN.md.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "lgx.h"

namespace lgx {

// ---- ridge kernel geometry (DESIGN.md "K1"; band heights are per-instantiation, see lgx_ridge.cu Geo<>) ----
constexpr int kChunk = 32;        // columns advanced per sweep step
constexpr int kRadius = 12;       // int(4*3.0+0.5), scipy gaussian_filter truncate=4

// plane pitches
__host__ __device__ inline int plane_pitch(int w) { return (w + 7) & ~7; }
__host__ __device__ inline int bits_pitch(int w) { return (w + 31) >> 5; }

struct RidgeParams {
  const void* blur;               // [batch][H][blur_pitch] u8/u16: output of blur5 (padded pitch, multiple of 32 px)
  int blur_pitch;
  int H, W, Wp;
  int bands, rows_per_band;
  size_t plane_stride;            // H * Wp (f64 elements)
  double* b;
  double* rsb;
  double* rsb2;
  double* g;                      // nullable (debug)
  const double* lut;              // 256 entries (u8): v * (1/255), or v / 255.0 with LGX_OPT_FLOAT_DIV
  int mixed_from_cols;
  int float_div;                  // u16 conversion in the kernels: 0: v * (1/65535) (skimage 0.19), 1: v / 65535.0
  double w[13];                   // gaussian taps of the handle: w[j], j < 12: pair (l-12+j, l+12-j); w[12]: centre
  unsigned long long* prof;       // nullable: [6] phase cycle counters (S2,S3,S4,S5 own work, barrier-to-top wait, CTAs)
};

struct SauvolaParams {
  const double* b;
  const double* rsb;
  const double* rsb2;
  int H, W, Wp, WW;
  size_t plane_stride;
  uint8_t* binary;                // nullable, dense [batch][H][W]
  uint32_t* bits;                 // [batch][H][WW]
  double* T;                      // nullable (debug), plane layout
};

struct MorphParams {
  const uint32_t* bits;           // [batch][H][WW] binary as bits (1 = 255)
  int H, W, WW;
  uint32_t* hbits;                // nullable: the two masks as bit planes [batch][H][WW] (LGX_OPT_PACKED_MASKS)
  uint32_t* vbits;
  uint8_t* hmask;                 // nullable dense u8
  uint8_t* vmask;                 // nullable dense u8
  uint32_t* jbits;                // [batch][H][WW] joints = H & V
  int32_t* lab;                   // nullable: [batch][H*W] union-find parents, seeded at the word-run starts of jbits
  int32_t* active;                // nullable: [batch][H*WW] compacted list of the non-empty jbits words
  int32_t* nactive;               // [batch] (zeroed by the caller)
};

// fused ridge + sauvola kernel (lgx_fused.cu): outputs and hand-over scratch
struct FusedParams {
  uint32_t* bits;                 // [batch][H][WW] binary as bits (1 = 255)
  unsigned char* ho_items;        // [bands * batch] hand-over blocks between the bands of a frame (fused_item_bytes(W) each)
  unsigned char* ho_rings;        // [ring_ctas] hand-over rings between the four row blocks of a CTA (fused_ring_bytes(ctas))
  int ring_ctas;
  int* prog;                      // [bands * batch] progress counters (zeroed by the launcher)
  double* dbg_b;                  // nullable (parity tests): b plane, pitch Wp
  double* dbg_T;                  // nullable: Sauvola threshold plane
};

// joints (contour-equivalent) scratch, per chunk
struct JointsParams {
  const uint32_t* jbits;          // mask analysed in this pass (joints, or hole-filled joints)
  int H, W, WW;
  int pass;                       // 0: first pass on joints; 1: second pass on filled mask (flagged frames only)
  int32_t* lab;                   // [batch][H*W] sparse union-find parents (only run-start pixels are used)
  uint32_t* rootbits;             // [batch][H][WW]
  int32_t* rootpix;               // [batch][max_comp]
  unsigned long long* acc;        // [batch][max_comp][4]: (e4<<32 | a00), a10, a01, spare
  int32_t* ncomp;                 // [batch]
  uint32_t* flags;                // [batch]
  int max_comp;
  int32_t* holes;                 // [batch][kMaxHoles] ranks of components with holes
  int32_t* nholes;                // [batch]
  int32_t* nested;                // [batch][kMaxNested] ranks of components nested in a hole
  int32_t* nnested;               // [batch]
  int32_t* segcount;              // [batch][8] roots per rank segment
  const int32_t* active;          // nullable: [batch][H*WW] indices of the non-empty words of jbits (any order)
  const int32_t* nactive;         // [batch]
};

// strip-local first pass of the contour stage (lgx_joints_local.cu)
struct JointsLocalParams {
  const uint32_t* jbits;          // [batch][H][WW]
  int H, W, WW;
  int R, strips, capr;            // rows per strip (joints_local_rows(W)), strips per frame, runs per strip (joints_local_runs(W))
  int32_t* lab;                   // [batch][H*W] sparse global parents: boundary runs and the first pixels of boundary components only
  unsigned long long* rec;        // [batch][max_comp][4] records {a00 | e4 << 32, a10, a01, first pixel | flags << 32}, strip blocks
  int32_t* nrec;                  // [batch] records reserved (zeroed by the caller)
  int32_t* sbase;                 // [batch][strips] first record of a strip
  int32_t* scount;                // [batch][strips] records of a strip
  int32_t* sdead;                 // [batch][strips] records merged into another strip's component (zeroed by the caller)
  unsigned long long* acc;        // outputs, as JointsParams
  int32_t* rootpix;
  int32_t* ncomp;
  uint32_t* flags;
  int max_comp;
  int32_t* holes;
  int32_t* nholes;
};
int joints_local_rows(int W);     // < 2: too wide for the strip kernel
int joints_local_runs(int W);
cudaError_t launch_joints_local(const JointsLocalParams& p, int batch, cudaStream_t stream);   // local + border + merge + compact

constexpr int kMaxHoles = 64;
constexpr int kMaxNested = 256;

struct EmitParams {
  const unsigned long long* acc;
  const int32_t* rootpix;
  const int32_t* ncomp;
  uint32_t* flags;
  int max_comp;
  int32_t* centroids;             // [batch][max_cent][2]
  double* centroids_f;            // nullable
  int max_cent;
  int32_t* counts;
};

// launchers (each enqueues on `stream`, returns cudaGetLastError())
cudaError_t launch_ridge(const RidgeParams& p, int bits, int batch, int nwarps, cudaStream_t stream);
int ridge_band_rows(int nwarps);   // b rows a band of the ridge kernel produces (8*nwarps - 4)
cudaError_t launch_undistort(const uint8_t* src, int channels, int batch, int H, int W, size_t pitch, size_t fstride,
                             const int16_t* map_xy, const uint16_t* map_frac, const int32_t* cam_index, uint8_t* dst,
                             cudaStream_t stream);
cudaError_t launch_bgr2gray(const void* bgr, int bits, size_t npix, void* gray, cudaStream_t stream);
cudaError_t launch_blur5(const void* frames, int bits, int batch, int H, int W, size_t pitch, size_t fstride,
                         void* out_pad, int pad_pitch, void* out_dense, cudaStream_t stream);
__host__ __device__ inline int blur_pitch(int w) { return (w + 31) & ~31; }
// variant: 0 = column kernel (direct loads), 2 = TMA ring kernel when the planes allow it
cudaError_t launch_sauvola(const SauvolaParams& p, int batch, int variant, cudaStream_t stream);
cudaError_t launch_pack_bits(const uint8_t* binary, int batch, int H, int W, uint32_t* bits, cudaStream_t stream);
cudaError_t launch_morph(const MorphParams& p, int batch, cudaStream_t stream);
cudaError_t launch_joints_label(const JointsParams& p, int batch, bool seeded, cudaStream_t stream);   // [seed]+union+roots+rank+sums
cudaError_t launch_joints_holes(const JointsParams& p, int batch, bool listed, cudaStream_t stream);   // [list] + local fix + kill
cudaError_t launch_fill_holes(const uint32_t* jbits, uint32_t* filled, uint32_t* scratch, const uint32_t* flags,
                              int batch, int H, int W, cudaStream_t stream);
cudaError_t launch_emit(const EmitParams& p, int batch, cudaStream_t stream);
// warp-specialised ridge kernel for large launches (lgx_ridge_ws.cu): 124-row bands, one CTA per SM, TMA in/out
bool ridge_ws_usable(const RidgeParams& rp, int bits);   // W >= 64, 16-byte aligned planes, driver exports cuTensorMapEncodeTiled
int ridge_ws_band_rows();
// max_ctas: 0 = one persistent CTA per SM; smaller values leave SMs free for kernels of other streams
cudaError_t launch_ridge_ws(const RidgeParams& rp, int bits, int batch, int max_ctas, cudaStream_t stream);

// fused ridge + sauvola kernel for large launches (lgx_fused.cu): 124-row bands, one persistent CTA per SM, no f64 planes
bool ridge_fused_usable(const RidgeParams& rp, int bits);
size_t fused_item_bytes(int W);
size_t fused_ring_bytes(int ctas);
int fused_bands(int H);
void fused_prof_read(unsigned long long* out32, bool reset);
cudaError_t launch_ridge_fused(const RidgeParams& rp, const FusedParams& fp, int bits, int batch, int max_ctas, cudaStream_t stream);
// u8 planes from bit planes (binary_img when the fused kernel produced only bits)
cudaError_t launch_unpack_bits(const uint32_t* bits, int batch, int H, int W, uint8_t* out, cudaStream_t stream);

}  // namespace lgx
