/*
 * lgx — B200-native laser-grid point extractor: C ABI (drop-in boundary).
 *
 * The reference has no FFI of its own; the path is two plain Python functions
 * bound by module attribute:
 *
 *   load_and_preprocess_image(img)  -> (original, gray, blurred, binary)
 *        /root/reference/utils/util_cylinder.py:1769-1802  (= utils/util_plane.py:2459-2492)
 *   extract_joints(binary)          -> (horizontal_mask, vertical_mask, centroids)
 *        /root/reference/utils/util_cylinder.py:1805-1827  (= utils/util_plane.py:2494-2516)
 *   callers: python_grid_detection_cylinder.py:77,82 ; python_grid_detection_plane.py:84,89
 *
 * This header is what a binding for that path binds instead (ctypes stub in
 * INTEGRATION.md).  Plain pointers and sizes only; no torch / CUDA types.
 * `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *
 * All device buffers are caller-allocated; the handle owns only scratch.
 * Every function returns an lgx_status (0 = ok, negative = error) and never
 * throws.  A handle is bound to one device and is not thread-safe.
 * There is no CPU path: without a CUDA device lgx_create fails.
 */
#ifndef LGX_H_
#define LGX_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LGX_VERSION 100

typedef struct lgx_handle lgx_handle;

enum lgx_status {
  LGX_OK = 0,
  LGX_ERR_BAD_ARG = -1,      /* NULL pointer, bits not 8/16, w or h < 2, batch < 0, size > handle capacity */
  LGX_ERR_CUDA = -2,         /* a CUDA runtime call failed; lgx_last_cuda_error() has the text */
  LGX_ERR_NO_DEVICE = -3,    /* no CUDA device / device is not sm_100 */
  LGX_ERR_OOM = -4,          /* scratch allocation failed */
  LGX_ERR_CAPACITY = -5      /* a frame overflowed max_centroids / max_components (see frame flags) */
};

/* per-frame flag bits written to d_flags / returned by lgx_resolve */
#define LGX_FLAG_HOLES          1u  /* frame had contours with holes (handled; informational) */
#define LGX_FLAG_GENERIC_FILL   2u  /* whole-frame hole fill pass was needed (handled; informational) */
#define LGX_FLAG_COMP_OVERFLOW  4u  /* more connected components than max_components: result invalid */
#define LGX_FLAG_CENT_OVERFLOW  8u  /* more centroids than max_centroids: list truncated, count is true */

/* options for lgx_set_option */
#define LGX_OPT_MIXED_FROM_COLS 1   /* 1 (default): Hrc = d(g_c)/dr, what scikit-image 0.19.x (the reference's pinned version) forms for
                                       order='rc' ; 0: Hrc = d(g_r)/dc (scikit-image >= 0.20).  SURVEY.md §8c */
#define LGX_OPT_FLOAT_DIV       8   /* 0 (default): img_as_float = v * (1/imax), scikit-image 0.19 `_convert`; 1: v / imax */
#define LGX_OPT_RIDGE_PROF      3   /* 1: the ridge kernel accumulates per-phase cycle counters (lgx_get_ridge_prof; debug) */
#define LGX_OPT_RIDGE_WARPS     4   /* 16: warp-specialised TMA pipeline, 124-row bands, 1 CTA/SM; 8: 64-row bands, 2 CTAs/SM; 4: 32-row bands,
                                       4 CTAs/SM; 0 (default): chosen by launch size. Same results. */
#define LGX_OPT_RIDGE_SMS       5   /* persistent CTAs of the pipeline ridge kernel: 0 (default) = one per SM; N < SMs leaves SMs to other streams */
#define LGX_OPT_SAUVOLA         6   /* 0 (default): column kernel (direct loads); 2: TMA ring kernel (planes must be 16-byte aligned; same results) */
#define LGX_OPT_HOST_SPLIT_FIRST 7  /* 1 (default): lgx_frontend_host splits its first chunk 1/4 + 3/4 (shorter pipeline fill); 0: uniform chunks */
#define LGX_OPT_FUSED           9   /* 0 (default): stage 1 = blur5, ridge, sauvola kernels.  1: blur5 + ONE fused ridge/sauvola kernel (no f64
                                       planes; csrc/lgx_fused.cu) when the batch fills every group of CTAs (148 / bands frames), the
                                       three-kernel path below that; 2: fused whenever the geometry allows (width >= 64, height >= 16).
                                       Same results; experimental: parity-green but slower than the default. */
#define LGX_OPT_JOINTS_GLOBAL   10  /* 1: first pass of the contour stage as the whole-frame union-find of round 1 (csrc/lgx_joints.cu) instead of
                                       the strip-local labelling in shared memory (csrc/lgx_joints_local.cu).  Same results (cross-check). */
#define LGX_OPT_PACKED_MASKS    11  /* 1: the binary / hmask / vmask buffers of lgx_frontend and lgx_frontend_host hold BIT planes instead of u8
                                       planes: [batch][height][lgx_bits_pitch(width)] u32, bit i of word w = pixel 32 w + i, 1 = 255.  An eighth of
                                       the bytes over PCIe for callers that run the reference's stages 3-6 on the host (np.unpackbits, little). */
#define LGX_OPT_TIMING          2   /* 1: bracket each kernel group of lgx_frontend with CUDA events (lgx_get_stats) */

/* ---- lifetime -------------------------------------------------------------------------- */

/* Scratch bytes lgx_create will allocate for these capacities. */
size_t lgx_workspace_bytes(int max_w, int max_h, int chunk_frames, int max_components);

/* chunk_frames: frames processed per internal pass (scratch is sized for it; any batch size
 * is accepted by lgx_frontend and processed chunk by chunk).  max_components: connected
 * components per frame before compaction (0 = max_w*max_h/16). */
int lgx_create(int device, int max_w, int max_h, int chunk_frames, int max_components,
               lgx_handle** out);
int lgx_destroy(lgx_handle* h);
int lgx_set_option(lgx_handle* h, int option, int value);
/* Override the 25 gaussian taps (scipy _gaussian_kernel1d(3.0, 0, 12) as computed by the caller's own
 * numpy/scipy; must be symmetric).  Default: the values of numpy 2.3.5 / scipy 1.18.1. */
int lgx_set_gauss_weights(lgx_handle* h, const double* w25);
const char* lgx_strerror(int status);
const char* lgx_last_cuda_error(void);
int lgx_version(void);

/* ---- the hot path, device buffers ------------------------------------------------------ */

/* Stages 1+2 for `batch` frames resident in device memory.
 *   d_frames      [batch] frames of h rows, row pitch `pitch_bytes`, frame stride `frame_stride_bytes`;
 *                 bits = 8 (uint8) or 16 (uint16); single channel (gray).
 *   d_binary      [batch][h][w] u8 {0,255}      binary_img      (util_cylinder.py:1798-1800)   nullable
 *   d_hmask       [batch][h][w] u8 {0,255}      horizontal_mask (util_cylinder.py:1813)        nullable
 *   d_vmask       [batch][h][w] u8 {0,255}      vertical_mask   (util_cylinder.py:1814)        nullable
 *   d_blurred     [batch][h][w] u8/u16          blurred_img     (util_cylinder.py:1790)        nullable
 *   d_centroids   [batch][max_centroids][2] i32 (cX,cY) in the reference's list order (util_cylinder.py:1819-1825)
 *   d_centroids_f [batch][max_centroids][2] f64 (m10/m00, m01/m00) before int()                nullable
 *   d_counts      [batch] i32   len(centroids) per frame
 *   d_flags       [batch] u32   LGX_FLAG_* per frame
 * Work is enqueued on `stream`; nothing is synchronised.  */
int lgx_frontend(lgx_handle* h, const void* d_frames, int bits, int batch, int height, int width,
                 size_t pitch_bytes, size_t frame_stride_bytes,
                 uint8_t* d_binary, uint8_t* d_hmask, uint8_t* d_vmask, void* d_blurred,
                 int32_t* d_centroids, double* d_centroids_f, int max_centroids,
                 int32_t* d_counts, uint32_t* d_flags, void* stream);

/* ---- the hot path, host buffers (what a reference-side caller holds) -------------------- */

/* Same as lgx_frontend with HOST pointers (pageable or pinned): copies frames in, runs, copies
 * the requested outputs back and synchronises `stream`.  Output pointers may be NULL to skip
 * that copy (d_centroids/d_counts are required).  The batch is pipelined in chunks of `chunk_frames` through three
 * device slots (copy-in, compute and copy-out of consecutive chunks overlap when the host buffers are page-locked);
 * only the used part of each centroid list is copied back.  */
int lgx_frontend_host(lgx_handle* h, const void* frames, int bits, int batch, int height, int width,
                      uint8_t* binary, uint8_t* hmask, uint8_t* vmask, void* blurred,
                      int32_t* centroids, double* centroids_f, int max_centroids,
                      int32_t* counts, uint32_t* flags, void* stream);

/* ---- per-stage entry points (parity tests; also the stage-2-only call) ------------------ */

/* cv2.cvtColor(BGR2GRAY) for a true-colour input (util_cylinder.py:1789): interleaved [batch][h][w][3] -> [batch][h][w]. */
int lgx_bgr2gray(const void* d_bgr, int bits, int batch, int height, int width, void* d_gray, void* stream);

/* Input side: the per-frame part of cv2.undistort (utils/iotool.py:22-39, cv2.undistort at :38;
 * callers python_grid_detection_cylinder.py:36-41, iotool.py:62-65) = cv2.remap(INTER_LINEAR, BORDER_CONSTANT 0) of
 * 8-bit images with `channels` = 1 or 3 interleaved channels, through the fixed-point maps of
 * cv2.initUndistortRectifyMap(CV_16SC2): d_map_xy [ncam][h][w][2] int16, d_map_frac [ncam][h][w] uint16 (the host
 * computes them once per camera, cylinder-pose-estimation_b200/iotool.py).  d_cam_index [batch] selects the
 * camera of each frame (NULL = camera 0 for all, e.g. L/R of a stereo rig = 0/1).  d_dst is dense
 * [batch][h][w][channels]; it must not alias d_src.  Needs no handle (no scratch). */
int lgx_undistort(const uint8_t* d_src, int channels, int batch, int height, int width, size_t pitch_bytes,
                  size_t frame_stride_bytes, const int16_t* d_map_xy, const uint16_t* d_map_frac,
                  const int32_t* d_cam_index, uint8_t* d_dst, void* stream);

/* cv2.GaussianBlur((5,5),0) on u8/u16 (util_cylinder.py:1790). */
int lgx_blur5(lgx_handle* h, const void* d_frames, int bits, int batch, int height, int width,
              size_t pitch_bytes, size_t frame_stride_bytes, void* d_blurred, void* stream);

/* detect_ridges(blurred, 3.0)[1] fused with the 5x5 blur and the horizontal running sums of
 * cv2.boxFilter (util_cylinder.py:1734-1738, :1755-1757).  Planes are f64 [batch][h][lgx_plane_pitch(w)].
 * d_g (gaussian-filtered image) is a debug output, nullable. */
int lgx_ridge(lgx_handle* h, const void* d_frames, int bits, int batch, int height, int width,
              size_t pitch_bytes, size_t frame_stride_bytes,
              double* d_b, double* d_rowsum_b, double* d_rowsum_b2, double* d_g, void* stream);

/* Column running sums + Sauvola threshold + compare (util_cylinder.py:1757-1765, :1798-1800).
 * d_T (threshold plane) is a debug output, nullable.  d_bits: [batch][h][lgx_bits_pitch(w)] u32. */
int lgx_sauvola(lgx_handle* h, const double* d_b, const double* d_rowsum_b, const double* d_rowsum_b2,
                int batch, int height, int width, uint8_t* d_binary, uint32_t* d_bits, double* d_T,
                void* stream);

/* Stage 1 in one fused kernel (csrc/lgx_fused.cu): detect_ridges + sauvola_threshold_fast + compare
 * (util_cylinder.py:1734-1766, :1798-1800) without the f64 planes in between.  d_bits: [batch][h][lgx_bits_pitch(w)] u32
 * (required); d_binary dense u8 (nullable); d_b / d_T: debug planes for parity tests, f64 [batch][h][lgx_plane_pitch(w)],
 * nullable.  Needs width >= 64 and height >= 16 (LGX_ERR_BAD_ARG otherwise). */
int lgx_ridge_sauvola(lgx_handle* h, const void* d_frames, int bits, int batch, int height, int width,
                      size_t pitch_bytes, size_t frame_stride_bytes,
                      double* d_b, double* d_T, uint8_t* d_binary, uint32_t* d_bits, void* stream);

/* Name of the kernel the last processed chunk used for the ridge stage ("ridge_fused_kernel<uint8_t>",
 * "ridge_ws_kernel<uint8_t>", "ridge_kernel<uint8_t,4>", ...): the dominant kernel bench.py reports its roofline for. */
const char* lgx_last_ridge_kernel(lgx_handle* h);
/* First-pass kernel of the contour stage in the last chunk: "jl_local" (strip-local) or "jl_union" (whole frame). */
const char* lgx_last_joints_kernel(lgx_handle* h);

/* extract_joints on a device-resident u8 binary image (util_cylinder.py:1805-1827). */
int lgx_extract_joints(lgx_handle* h, const uint8_t* d_binary, int batch, int height, int width,
                       uint8_t* d_hmask, uint8_t* d_vmask,
                       int32_t* d_centroids, double* d_centroids_f, int max_centroids,
                       int32_t* d_counts, uint32_t* d_flags, void* stream);

/* The contour part of extract_joints alone, on any device-resident u8 mask (non-zero = set): cv2.findContours(mask,
 * RETR_EXTERNAL, CHAIN_APPROX_SIMPLE) + cv2.moments + int(m10/m00), int(m01/m00) in contour order, contours of zero area
 * dropped (util_cylinder.py:1817-1825).  Outputs as lgx_extract_joints. */
int lgx_contour_centroids(lgx_handle* h, const uint8_t* d_mask, int batch, int height, int width,
                          int32_t* d_centroids, double* d_centroids_f, int max_centroids,
                          int32_t* d_counts, uint32_t* d_flags, void* stream);

/* Contour statistics of the last chunk processed (debug / parity): per frame, for each reported
 * contour in the reference's order: first pixel raster index and the Green sums a00,a10,a01.
 * d_out: [n][4] i64 for frame `frame_in_chunk`; returns the number written via *n. */
int lgx_debug_contours(lgx_handle* h, int frame_in_chunk, int64_t* out_host, int capacity, int* n);

/* Accumulated since the last reset: ms5[0..4] = device time (CUDA events on the launch stream, needs
 * LGX_OPT_TIMING) of the blur5 | ridge | sauvola | open_hv | joints kernels, `chunks` = kernel groups timed,
 * `launches` = kernels launched by this handle.  Synchronises on the last recorded event. */
int lgx_get_stats(lgx_handle* h, double* ms5, long long* chunks, long long* launches, int reset);

/* Debug: 16 counters.  Phase kernel: out[0..3] = cycles thread 0 of every ridge CTA spent in phases S2,S3,S4,S5
 * (own work + wait at the closing barrier), out[4] = wait at the loop-top barrier, out[5] = CTAs.  Pipeline kernel:
 * out[8],[9] / [10],[11] / [12],[13] = cycles the V / H / E warps (lane 0 of each) waited for their input and for
 * their output buffer, out[14] = total cycles of those warps, out[15] = CTAs.  Needs LGX_OPT_RIDGE_PROF. */
int lgx_get_ridge_prof(lgx_handle* h, unsigned long long* out16, int reset);

/* Parity of the branch-free square root the ridge / Sauvola kernels use (csrc/lgx_sqrt.cuh) against the device's
 * IEEE sqrt.rn.f64, bit for bit, on `n` radicands generated on the device from `seed` (mode 0: random bit patterns of
 * every non-negative finite double; mode 1: magnitudes 2^-120 .. 2^8; mode 2: the n values of extra_host).
 * out4[0] = values that differ, out4[1] = values the sequence flags as out of its range [2^-970, max] (the kernels
 * then take the library square root), out4[2] = out-of-range values that were not flagged, out4[3] = bits of the
 * first differing radicand | 1<<63.  Runs on the current device's default stream and synchronises. */
int lgx_debug_sqrt(unsigned long long seed, unsigned long long n, int mode, const double* extra_host,
                   unsigned long long* out4);

/* Debug: 32 cycle counters of the fused ridge + sauvola kernel (all zero unless the library was built with
 * -DLGX_FZ_PROF; tools/fused_prof.py prints them): [0..3] VH warps: wait for the tile | vertical pass | wait for a free
 * ring group | horizontal pass, [4] total, [5] warps; [8..11] EC warps: wait for g | for the block / band above | for
 * the block below | release + publish, [12..15] phases 1-2 | wait for the prefetched hand-over | phase 3 | phase 4,
 * [16] total, [17] warps, [18] wait for the band above (EC_0 only). */
int lgx_debug_fused_prof(unsigned long long* out32, int reset);

int lgx_plane_pitch(int width);   /* f64 elements per row of the b / rowsum planes */
int lgx_bits_pitch(int width);    /* u32 words per row of bit planes */

/* ---- synthetic frames (bench / tests): base f32 image + per-frame noise on the device ---- */

/* out[f][y][x] = clip(rint((base[f % n_base][y][x] + sigma * N(0,1; seed0+f)) * scale)), scale=1 (8 bit) or 257 (16 bit). */
int lgx_render_noisy(const float* d_base, int n_base, int batch, int height, int width,
                     float sigma, uint64_t seed0, int bits, void* d_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LGX_H_ */
