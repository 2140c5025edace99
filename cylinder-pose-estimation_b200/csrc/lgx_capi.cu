// C ABI of lgx (include/lgx.h): handle, scratch planes, chunked orchestration of the kernels.
// Mirrors the reference's two stage functions (util_cylinder.py:1769-1802, :1805-1827); see lgx.h.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>
#include "lgx_internal.cuh"

using namespace lgx;

constexpr int kHostSlots = 3;   // device mirrors of lgx_frontend_host: copy-in runs one chunk ahead of the compute queue

struct lgx_handle {
  int device = 0;
  int max_w = 0, max_h = 0, chunk = 0, max_comp = 0;
  int mixed = 1;                // LGX_OPT_MIXED_FROM_COLS: 1 = Hrc = d(g_c)/dr (scikit-image 0.19.x, the reference's pinned version)
  int float_div = 0;            // LGX_OPT_FLOAT_DIV: 0 = img_as_float multiplies by RN(1/imax) (scikit-image 0.19), 1 = divides
  double gw[13];                // gaussian taps of this handle (lgx_set_gauss_weights), passed to the kernels as parameters
  int split_first = 1;          // LGX_OPT_HOST_SPLIT_FIRST: lgx_frontend_host splits its first chunk 1/4 + 3/4 (shorter pipeline fill)
  int sauvola_variant = 0;      // LGX_OPT_SAUVOLA: 0 = column kernel, 2 = TMA ring kernel when usable
  int ridge_sms = 0;            // LGX_OPT_RIDGE_SMS: persistent CTAs of the pipeline ridge kernel (0 = one per SM)
  int fused = 0;                // LGX_OPT_FUSED: 0 = blur / ridge / sauvola as three kernels (default), 1 = fused ridge + sauvola kernel when the
                                // batch fills every CTA group, 2 = whenever the geometry allows
  const char* last_ridge_kernel = "";   // name of the kernel the last chunk used for the ridge stage (lgx_last_ridge_kernel)
  unsigned char *ho_items = nullptr, *ho_rings = nullptr;   // hand-over scratch of the fused kernel
  int* fprog = nullptr;
  int ring_ctas = 0;
  int ridge_warps = 0;          // LGX_OPT_RIDGE_WARPS: 16 (warp-specialised, 124-row bands, 1 CTA/SM), 8 (64-row bands, 2 CTAs/SM),
                                // 4 (32-row bands, 4 CTAs/SM), 0 = by launch size
  // per-chunk scratch
  double *b = nullptr, *rsb = nullptr, *rsb2 = nullptr;
  uint32_t *bits = nullptr, *jbits = nullptr, *rootbits = nullptr, *filled = nullptr, *oscr = nullptr;
  int32_t *lab = nullptr, *rootpix = nullptr, *ncomp = nullptr;
  int32_t* holework = nullptr;   // per chunk frame: nholes, nnested counters + the two lists
  int32_t* stripwork = nullptr;  // strip-local contour pass: [chunk] records reserved, then [chunk][max_h] x 3: merged, first record, records per strip
  unsigned long long* rec = nullptr;   // [chunk][max_comp][4] component records of the strip-local pass
  int packed = 0;                // LGX_OPT_PACKED_MASKS: mask outputs as bit planes
  int joints_global = 0;         // LGX_OPT_JOINTS_GLOBAL: 1 = whole-frame union-find as the first pass (cross-check)
  const char* last_joints_kernel = "";
  int32_t* active = nullptr;     // [chunk][h*ww] compacted non-empty joints words + [chunk] counters at the end
  unsigned long long* acc = nullptr;
  double* lut8 = nullptr;       // 256 entries (the 16-bit conversion is computed in the kernels)
  uint16_t* blur = nullptr;     // [chunk][h][blur_pitch(w)] u8 or u16
  unsigned long long* prof = nullptr;   // 8 counters, allocated when LGX_OPT_RIDGE_PROF is set
  // device mirrors for lgx_frontend_host (lazily sized)
  unsigned char* host_dev = nullptr;
  size_t host_dev_bytes = 0;
  cudaStream_t s_in = nullptr, s_out[kHostSlots] = {};
  cudaEvent_t ev_in[kHostSlots] = {}, ev_done[kHostSlots] = {}, ev_small[kHostSlots] = {}, ev_out[kHostSlots] = {};
  cudaEvent_t ev_start = nullptr;
  // last chunk geometry (lgx_debug_contours)
  int last_h = 0, last_w = 0, last_n = 0;
  // optional per-kernel timing (LGX_OPT_TIMING): 5 events per chunk bracket ridge | sauvola | morph | joints
  int timing = 0;
  std::vector<cudaEvent_t> evs;
  size_t ev_used = 0;
  double ms_acc[5] = {0, 0, 0, 0, 0};
  long long chunks_timed = 0;
  long long launches = 0;
};

namespace {

thread_local char g_cuda_err[256] = "";

int fail_cuda(cudaError_t e, const char* what) {
  snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", what, cudaGetErrorString(e));
  return LGX_ERR_CUDA;
}

// Every entry point runs on the handle's device and leaves the caller's current device as it found it.
struct DeviceGuard {
  int prev = -1;
  cudaError_t err = cudaSuccess;
  explicit DeviceGuard(int dev) {
    err = cudaGetDevice(&prev);
    if (err == cudaSuccess && prev != dev) err = cudaSetDevice(dev);
    else if (err == cudaSuccess) prev = -1;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define LGX_ON_DEVICE(h)                                              \
  DeviceGuard guard_((h)->device);                                    \
  if (guard_.err != cudaSuccess) return fail_cuda(guard_.err, "cudaSetDevice")

#define LGX_CK(call)                                   \
  do {                                                 \
    cudaError_t e_ = (call);                           \
    if (e_ != cudaSuccess) return fail_cuda(e_, #call); \
  } while (0)

// numpy/scipy: phi = exp(-0.5/9 * x^2), x = -12..12, w = phi / phi.sum()  (scipy _gaussian_kernel1d).
// The 13 distinct values as produced by numpy 2.3.5 / scipy 1.18.1 (tests/test_host_logic.py compares them
// with the caller's own scipy); lgx_set_gauss_weights overrides them.
const double kGaussW[13] = {0x1.763a210dfb306p-15, 0x1.4fbe39149e277p-13, 0x1.0d8a5ad43c165p-11, 0x1.8345966f69518p-10,
                            0x1.f1e9915139406p-9,  0x1.1e6bccad344bap-7,  0x1.26defcaeb0202p-6,  0x1.0fa58939b528fp-5,
                            0x1.bfde9c12bec92p-5,  0x1.4a614d1afd337p-4,  0x1.b42a57d56c0bep-4,  0x1.01a25f86eb137p-3,
                            0x1.105a329f98197p-3};

struct Sizes {
  size_t plane, bitsz, lab, rootpix, acc, blur;
};
Sizes sizes_for(int w, int h, int chunk, int max_comp) {
  Sizes s;
  s.plane = (size_t)chunk * h * plane_pitch(w) * sizeof(double);
  s.bitsz = (size_t)chunk * h * bits_pitch(w) * sizeof(uint32_t);
  s.lab = (size_t)chunk * h * w * sizeof(int32_t);
  s.rootpix = (size_t)chunk * max_comp * sizeof(int32_t);
  s.acc = (size_t)chunk * max_comp * 4 * sizeof(unsigned long long);
  s.blur = (size_t)chunk * h * blur_pitch(w) * sizeof(uint16_t);
  return s;
}
int default_max_comp(int w, int h) {
  long long v = (long long)w * h / 16;
  return (int)(v < 1024 ? 1024 : v);
}

bool geometry_ok(const lgx_handle* h, int bits, int batch, int height, int width) {
  return h && (bits == 8 || bits == 16) && batch >= 0 && height >= 2 && width >= 2 && height <= h->max_h &&
         width <= h->max_w && (size_t)height * width <= (size_t)h->max_h * h->max_w;
}

int mark(lgx_handle* h, cudaStream_t st) {
  if (!h->timing) return LGX_OK;
  if (h->ev_used == h->evs.size()) {
    cudaEvent_t e;
    LGX_CK(cudaEventCreate(&e));
    h->evs.push_back(e);
  }
  LGX_CK(cudaEventRecord(h->evs[h->ev_used++], st));
  return LGX_OK;
}

// true when the morph kernel has to seed the whole-frame union-find and list the non-empty words (first pass of lgx_joints.cu)
bool joints_whole_frame(const lgx_handle* h, int W) { return h->joints_global || joints_local_rows(W) < 2; }

int run_joints(lgx_handle* h, int nb, int H, int W, int32_t* cent, double* centf, int max_cent, int32_t* counts,
               uint32_t* flags, bool seeded, cudaStream_t st) {
  JointsParams jp{};
  jp.jbits = h->jbits;
  jp.H = H; jp.W = W; jp.WW = bits_pitch(W);
  jp.pass = 0;
  jp.lab = h->lab; jp.rootbits = h->rootbits; jp.rootpix = h->rootpix; jp.acc = h->acc; jp.ncomp = h->ncomp;
  jp.flags = flags; jp.max_comp = h->max_comp;
  jp.nholes = h->holework;
  jp.nnested = h->holework + h->chunk;
  jp.holes = h->holework + 2 * h->chunk;
  jp.nested = jp.holes + (size_t)h->chunk * kMaxHoles;
  jp.segcount = jp.nested + (size_t)h->chunk * kMaxNested;
  jp.active = seeded ? h->active : nullptr;      // (the morph kernel lists the non-empty words when it seeds)
  jp.nactive = h->active + (size_t)h->chunk * H * bits_pitch(W);
  LGX_CK(cudaMemsetAsync(h->holework, 0, (size_t)2 * h->chunk * sizeof(int32_t), st));
  const int R = h->joints_global ? 0 : joints_local_rows(W);
  if (R >= 2) {
    // first pass strip by strip in shared memory (lgx_joints_local.cu)
    JointsLocalParams lp{};
    lp.jbits = h->jbits; lp.H = H; lp.W = W; lp.WW = bits_pitch(W);
    lp.R = R; lp.strips = (H + R - 1) / R; lp.capr = joints_local_runs(W);
    lp.lab = h->lab; lp.rec = h->rec;
    lp.nrec = h->stripwork; lp.sdead = h->stripwork + h->chunk;
    lp.sbase = lp.sdead + (size_t)h->chunk * lp.strips; lp.scount = lp.sbase + (size_t)h->chunk * lp.strips;
    lp.acc = h->acc; lp.rootpix = h->rootpix; lp.ncomp = h->ncomp; lp.flags = flags; lp.max_comp = h->max_comp;
    lp.holes = jp.holes; lp.nholes = jp.nholes;
    LGX_CK(cudaMemsetAsync(h->stripwork, 0, ((size_t)h->chunk + (size_t)h->chunk * lp.strips) * sizeof(int32_t), st));   // nrec, sdead
    LGX_CK(launch_joints_local(lp, nb, st));
    LGX_CK(launch_joints_holes(jp, nb, true, st));
    h->last_joints_kernel = "jl_local";
  } else {
    LGX_CK(cudaMemsetAsync(h->rootbits, 0, (size_t)nb * H * bits_pitch(W) * sizeof(uint32_t), st));   // only non-empty words are visited
    LGX_CK(launch_joints_label(jp, nb, seeded, st));
    LGX_CK(launch_joints_holes(jp, nb, false, st));
    h->last_joints_kernel = "jl_union";
  }
  LGX_CK(launch_fill_holes(h->jbits, h->filled, h->oscr, flags, nb, H, W, st));
  jp.pass = 1;
  jp.jbits = h->filled;
  LGX_CK(launch_joints_label(jp, nb, false, st));
  EmitParams ep{};
  ep.acc = h->acc; ep.rootpix = h->rootpix; ep.ncomp = h->ncomp; ep.flags = flags; ep.max_comp = h->max_comp;
  ep.centroids = cent; ep.centroids_f = centf; ep.max_cent = max_cent; ep.counts = counts;
  LGX_CK(launch_emit(ep, nb, st));
  h->launches += R >= 2 ? 14 : 16;   // (local, border, merge, compact | union, roots, rank x2, sums, hole list) + hole fix/kill + fill
                                     // + (seed, union, roots, rank x2, sums) + emit
  h->last_h = H; h->last_w = W; h->last_n = nb;
  return LGX_OK;
}

// skimage.img_as_float(uint8) as a table: v * RN(1/255) (scikit-image 0.19: np.multiply(image, 1. / imax_in)), or the
// division v / 255.0 with LGX_OPT_FLOAT_DIV.  The 16-bit conversion is computed inside the kernels.
int upload_lut8(lgx_handle* h) {
  double lut[256];
  for (int v = 0; v < 256; ++v) lut[v] = h->float_div ? (double)v / 255.0 : (double)v * (1.0 / 255.0);
  LGX_CK(cudaMemcpy(h->lut8, lut, sizeof(lut), cudaMemcpyHostToDevice));
  return LGX_OK;
}

}  // namespace

extern "C" {

int lgx_version(void) { return LGX_VERSION; }
int lgx_plane_pitch(int width) { return plane_pitch(width); }
int lgx_bits_pitch(int width) { return bits_pitch(width); }
const char* lgx_last_cuda_error(void) { return g_cuda_err; }

const char* lgx_strerror(int status) {
  switch (status) {
    case LGX_OK: return "ok";
    case LGX_ERR_BAD_ARG: return "bad argument";
    case LGX_ERR_CUDA: return "CUDA error";
    case LGX_ERR_NO_DEVICE: return "no usable CUDA device (sm_100 required); lgx has no CPU path";
    case LGX_ERR_OOM: return "out of device memory";
    case LGX_ERR_CAPACITY: return "capacity exceeded";
    default: return "unknown status";
  }
}

size_t lgx_workspace_bytes(int max_w, int max_h, int chunk_frames, int max_components) {
  if (max_w < 2 || max_h < 2 || chunk_frames < 1) return 0;
  if (max_components <= 0) max_components = default_max_comp(max_w, max_h);
  Sizes s = sizes_for(max_w, max_h, chunk_frames, max_components);
  return 3 * s.plane + 6 * s.bitsz + s.lab + s.rootpix + 2 * s.acc + (size_t)chunk_frames * (1 + 3 * (size_t)max_h) * 4 + s.blur + (size_t)chunk_frames * 4 + 256 * sizeof(double) +
         (size_t)chunk_frames * fused_bands(max_h) * (fused_item_bytes(max_w) + sizeof(int)) + fused_ring_bytes(148);
}

int lgx_create(int device, int max_w, int max_h, int chunk_frames, int max_components, lgx_handle** out) {
  if (!out || max_w < 2 || max_h < 2 || chunk_frames < 1) return LGX_ERR_BAD_ARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) {
    cudaGetLastError();
    return LGX_ERR_NO_DEVICE;
  }
  cudaDeviceProp prop;
  LGX_CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return LGX_ERR_NO_DEVICE;   // built for sm_100a only
  DeviceGuard guard(device);
  if (guard.err != cudaSuccess) return fail_cuda(guard.err, "cudaSetDevice");
  lgx_handle* h = new (std::nothrow) lgx_handle();
  if (!h) return LGX_ERR_OOM;
  h->device = device; h->max_w = max_w; h->max_h = max_h; h->chunk = chunk_frames;
  h->max_comp = max_components > 0 ? max_components : default_max_comp(max_w, max_h);
  for (int i = 0; i < 13; ++i) h->gw[i] = kGaussW[i];
  Sizes s = sizes_for(max_w, max_h, chunk_frames, h->max_comp);
  bool ok = true;
  auto alloc = [&](void** p, size_t n) { if (ok && cudaMalloc(p, n) != cudaSuccess) { ok = false; cudaGetLastError(); } };
  alloc((void**)&h->b, s.plane); alloc((void**)&h->rsb, s.plane); alloc((void**)&h->rsb2, s.plane);
  alloc((void**)&h->bits, s.bitsz); alloc((void**)&h->jbits, s.bitsz); alloc((void**)&h->rootbits, s.bitsz);
  alloc((void**)&h->filled, s.bitsz); alloc((void**)&h->oscr, s.bitsz);
  alloc((void**)&h->lab, s.lab); alloc((void**)&h->rootpix, s.rootpix); alloc((void**)&h->acc, s.acc);
  alloc((void**)&h->ncomp, (size_t)chunk_frames * sizeof(int32_t));
  alloc((void**)&h->holework, (size_t)chunk_frames * (2 + kMaxHoles + kMaxNested + 8) * sizeof(int32_t));
  alloc((void**)&h->active, s.bitsz + (size_t)chunk_frames * sizeof(int32_t));
  alloc((void**)&h->stripwork, (size_t)chunk_frames * (1 + 3 * (size_t)max_h) * sizeof(int32_t));
  alloc((void**)&h->rec, s.acc);
  alloc((void**)&h->blur, s.blur);
  alloc((void**)&h->lut8, 256 * sizeof(double));
  {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    h->ring_ctas = sms;
    const size_t items = (size_t)chunk_frames * fused_bands(max_h);
    alloc((void**)&h->ho_items, items * fused_item_bytes(max_w));
    alloc((void**)&h->ho_rings, fused_ring_bytes(sms));
    alloc((void**)&h->fprog, items * sizeof(int));
  }
  if (!ok) { lgx_destroy(h); return LGX_ERR_OOM; }
  const int rc = upload_lut8(h);
  if (rc != LGX_OK) { lgx_destroy(h); return rc; }
  *out = h;
  return LGX_OK;
}

int lgx_destroy(lgx_handle* h) {
  if (!h) return LGX_OK;
  DeviceGuard guard(h->device);
  void* ptrs[] = {h->b, h->rsb, h->rsb2, h->bits, h->jbits, h->rootbits, h->filled, h->oscr, h->lab, h->rootpix,
                  h->acc, h->ncomp, h->lut8, h->host_dev, h->ho_items, h->ho_rings, h->fprog, h->blur, h->prof, h->holework, h->active, h->stripwork, h->rec};
  for (void* p : ptrs) if (p) cudaFree(p);
  for (cudaEvent_t e : h->evs) cudaEventDestroy(e);
  if (h->s_in) {
    cudaStreamDestroy(h->s_in);
    cudaEventDestroy(h->ev_start);
    for (int s = 0; s < kHostSlots; ++s) {
      cudaStreamDestroy(h->s_out[s]);
      cudaEventDestroy(h->ev_in[s]); cudaEventDestroy(h->ev_done[s]); cudaEventDestroy(h->ev_small[s]); cudaEventDestroy(h->ev_out[s]);
    }
  }
  delete h;
  return LGX_OK;
}

int lgx_set_option(lgx_handle* h, int option, int value) {
  if (!h) return LGX_ERR_BAD_ARG;
  if (option == LGX_OPT_MIXED_FROM_COLS) { h->mixed = value ? 1 : 0; return LGX_OK; }
  if (option == LGX_OPT_FLOAT_DIV) {
    LGX_ON_DEVICE(h);
    h->float_div = value ? 1 : 0;
    LGX_CK(cudaDeviceSynchronize());     // kernels in flight still read the old table
    return upload_lut8(h);
  }
  if (option == LGX_OPT_HOST_SPLIT_FIRST) { h->split_first = value ? 1 : 0; return LGX_OK; }
  if (option == LGX_OPT_PACKED_MASKS) { h->packed = value ? 1 : 0; return LGX_OK; }
  if (option == LGX_OPT_JOINTS_GLOBAL) { h->joints_global = value ? 1 : 0; return LGX_OK; }
  if (option == LGX_OPT_SAUVOLA) { h->sauvola_variant = value == 2 ? 2 : 0; return LGX_OK; }
  if (option == LGX_OPT_TIMING) { h->timing = value ? 1 : 0; return LGX_OK; }
  if (option == LGX_OPT_FUSED) {
    if (value < 0 || value > 2) return LGX_ERR_BAD_ARG;
    h->fused = value;
    return LGX_OK;
  }
  if (option == LGX_OPT_RIDGE_WARPS) {
    if (value != 0 && value != 4 && value != 8 && value != 16) return LGX_ERR_BAD_ARG;
    h->ridge_warps = value;
    return LGX_OK;
  }
  if (option == LGX_OPT_RIDGE_SMS) {
    if (value < 0) return LGX_ERR_BAD_ARG;
    h->ridge_sms = value;
    return LGX_OK;
  }
  if (option == LGX_OPT_RIDGE_PROF) {
    LGX_ON_DEVICE(h);
    if (value && !h->prof) {
      LGX_CK(cudaMalloc((void**)&h->prof, 16 * sizeof(unsigned long long)));
      LGX_CK(cudaMemset(h->prof, 0, 16 * sizeof(unsigned long long)));
    } else if (!value && h->prof) {
      cudaFree(h->prof);
      h->prof = nullptr;
    }
    return LGX_OK;
  }
  return LGX_ERR_BAD_ARG;
}

int lgx_set_gauss_weights(lgx_handle* h, const double* w25) {
  if (!h || !w25) return LGX_ERR_BAD_ARG;
  for (int i = 0; i < 12; ++i) if (w25[i] != w25[24 - i]) return LGX_ERR_BAD_ARG;
  for (int i = 0; i < 13; ++i) h->gw[i] = w25[i];     // per handle: the kernels take the taps as parameters
  return LGX_OK;
}

int lgx_bgr2gray(const void* d_bgr, int bits, int batch, int height, int width, void* d_gray, void* stream) {
  if (!d_bgr || !d_gray || (bits != 8 && bits != 16) || batch < 0 || height < 1 || width < 1) return LGX_ERR_BAD_ARG;
  if (batch == 0) return LGX_OK;
  LGX_CK(launch_bgr2gray(d_bgr, bits, (size_t)batch * height * width, d_gray, (cudaStream_t)stream));
  return LGX_OK;
}

int lgx_undistort(const uint8_t* d_src, int channels, int batch, int height, int width, size_t pitch_bytes,
                  size_t frame_stride_bytes, const int16_t* d_map_xy, const uint16_t* d_map_frac,
                  const int32_t* d_cam_index, uint8_t* d_dst, void* stream) {
  if (!d_src || !d_map_xy || !d_map_frac || !d_dst || (channels != 1 && channels != 3) || batch < 0 || height < 1 || width < 1 ||
      height > 32767 || width > 32767)
    return LGX_ERR_BAD_ARG;
  if (pitch_bytes < (size_t)width * channels || frame_stride_bytes < pitch_bytes * (size_t)height) return LGX_ERR_BAD_ARG;
  if (pitch_bytes * (size_t)height > 0xffffffffull) return LGX_ERR_BAD_ARG;   // the kernel uses 32-bit offsets inside a frame
  if (batch == 0) return LGX_OK;
  LGX_CK(launch_undistort(d_src, channels, batch, height, width, pitch_bytes, frame_stride_bytes, d_map_xy, d_map_frac,
                          d_cam_index, d_dst, (cudaStream_t)stream));
  return LGX_OK;
}

int lgx_blur5(lgx_handle* h, const void* d_frames, int bits, int batch, int height, int width, size_t pitch_bytes,
              size_t frame_stride_bytes, void* d_blurred, void* stream) {
  if (!geometry_ok(h, bits, batch, height, width) || !d_frames || !d_blurred) return LGX_ERR_BAD_ARG;
  if (batch == 0) return LGX_OK;
  LGX_ON_DEVICE(h);
  LGX_CK(launch_blur5(d_frames, bits, batch, height, width, pitch_bytes, frame_stride_bytes, nullptr, 0, d_blurred, (cudaStream_t)stream));
  return LGX_OK;
}

static int ridge_chunk(lgx_handle* h, const void* d_frames, int bits, int nb, int H, int W, size_t pitch, size_t fstride,
                       double* b, double* rsb, double* rsb2, double* g, void* blurred, cudaStream_t st, bool timed = false) {
  LGX_CK(launch_blur5(d_frames, bits, nb, H, W, pitch, fstride, h->blur, blur_pitch(W), blurred, st));
  if (timed) { int rc = mark(h, st); if (rc) return rc; }
  RidgeParams rp{};
  rp.blur = h->blur; rp.blur_pitch = blur_pitch(W);
  rp.H = H; rp.W = W; rp.Wp = plane_pitch(W);
  rp.plane_stride = (size_t)H * rp.Wp;
  rp.b = b; rp.rsb = rsb; rp.rsb2 = rsb2; rp.g = g;
  rp.lut = h->lut8;
  rp.mixed_from_cols = h->mixed;
  rp.float_div = h->float_div;
  for (int i = 0; i < 13; ++i) rp.w[i] = h->gw[i];
  rp.prof = h->prof;
  // Large launches: the warp-specialised kernel (one 124-row band per SM).  Small launches (single frames) fill
  // the SMs better with the phase kernel's 4-warp CTAs; all instantiations give bit-identical planes.
  int nwarps = h->ridge_warps;
  const int ws_bands = (H + ridge_ws_band_rows() - 1) / ridge_ws_band_rows();
  if ((nwarps == 16 || (nwarps == 0 && ws_bands * nb >= 148)) && ridge_ws_usable(rp, bits)) {
    rp.bands = ws_bands;
    rp.rows_per_band = (H + rp.bands - 1) / rp.bands;
    LGX_CK(launch_ridge_ws(rp, bits, nb, h->ridge_sms, st));
    h->last_ridge_kernel = bits == 8 ? "ridge_ws_kernel<uint8_t>" : "ridge_ws_kernel<uint16_t>";
    return LGX_OK;
  }
  if (nwarps == 16) nwarps = 8;
  if (nwarps == 0) nwarps = ((H + 59) / 60) * nb < 2 * 148 ? 4 : 8;
  const int brows = ridge_band_rows(nwarps);
  rp.bands = (H + brows - 1) / brows;
  rp.rows_per_band = (H + rp.bands - 1) / rp.bands;
  LGX_CK(launch_ridge(rp, bits, nb, nwarps, st));
  h->last_ridge_kernel = bits == 8 ? (nwarps == 4 ? "ridge_kernel<uint8_t,4>" : "ridge_kernel<uint8_t,8>")
                                   : (nwarps == 4 ? "ridge_kernel<uint16_t,4>" : "ridge_kernel<uint16_t,8>");
  return LGX_OK;
}

// Stage 1 of one chunk with the fused kernel: blur5 -> ridge + sauvola (bit plane in h->bits).  Returns 1 if the fused
// kernel does not apply to this launch (the caller then runs the two-kernel path), 0 on success, < 0 on error.
static int fused_chunk(lgx_handle* h, const void* d_frames, int bits, int nb, int H, int W, size_t pitch, size_t fstride,
                       void* blurred, uint32_t* d_bits, double* dbg_b, double* dbg_T, cudaStream_t st, bool timed) {
  if (!h->fused) return 1;
  RidgeParams rp{};
  rp.blur = h->blur; rp.blur_pitch = blur_pitch(W);
  rp.H = H; rp.W = W; rp.Wp = plane_pitch(W);
  rp.plane_stride = (size_t)H * rp.Wp;
  rp.lut = h->lut8;
  rp.mixed_from_cols = h->mixed;
  rp.float_div = h->float_div;
  for (int i = 0; i < 13; ++i) rp.w[i] = h->gw[i];
  if (!ridge_fused_usable(rp, bits)) return 1;
  const int fbands = fused_bands(H);
  if (fbands > h->ring_ctas) return 1;
  if (h->fused == 1 && nb < h->ring_ctas / fbands) return 1;   // fewer frames than CTA groups: the phase kernels fill the SMs better
  LGX_CK(launch_blur5(d_frames, bits, nb, H, W, pitch, fstride, h->blur, blur_pitch(W), blurred, st));
  if (timed) { int rc = mark(h, st); if (rc) return rc; }
  FusedParams fp{};
  fp.bits = d_bits;
  fp.ho_items = h->ho_items; fp.ho_rings = h->ho_rings; fp.ring_ctas = h->ring_ctas; fp.prog = h->fprog;
  fp.dbg_b = dbg_b; fp.dbg_T = dbg_T;
  LGX_CK(launch_ridge_fused(rp, fp, bits, nb, h->ridge_sms, st));
  h->last_ridge_kernel = bits == 8 ? "ridge_fused_kernel<uint8_t>" : "ridge_fused_kernel<uint16_t>";
  return LGX_OK;
}

int lgx_ridge_sauvola(lgx_handle* h, const void* d_frames, int bits, int batch, int height, int width, size_t pitch_bytes,
                      size_t frame_stride_bytes, double* d_b, double* d_T, uint8_t* d_binary, uint32_t* d_bits, void* stream) {
  if (!geometry_ok(h, bits, batch, height, width) || !d_frames || !d_bits) return LGX_ERR_BAD_ARG;
  LGX_ON_DEVICE(h);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t ps = (size_t)height * plane_pitch(width), bs = (size_t)height * bits_pitch(width);
  const int saved = h->fused;
  h->fused = 2;
  int rc = LGX_OK;
  for (int c0 = 0; c0 < batch && rc == LGX_OK; c0 += h->chunk) {
    const int nb = batch - c0 < h->chunk ? batch - c0 : h->chunk;
    rc = fused_chunk(h, (const unsigned char*)d_frames + (size_t)c0 * frame_stride_bytes, bits, nb, height, width, pitch_bytes,
                     frame_stride_bytes, nullptr, d_bits + c0 * bs, d_b ? d_b + c0 * ps : nullptr, d_T ? d_T + c0 * ps : nullptr, st, false);
    if (rc == 1) rc = LGX_ERR_BAD_ARG;                      // geometry the fused kernel does not take (width < 64, height < 16)
    if (rc == LGX_OK && d_binary)
      rc = launch_unpack_bits(d_bits + c0 * bs, nb, height, width, d_binary + (size_t)c0 * height * width, st) == cudaSuccess ? LGX_OK : LGX_ERR_CUDA;
  }
  h->fused = saved;
  return rc;
}

const char* lgx_last_ridge_kernel(lgx_handle* h) { return h ? h->last_ridge_kernel : ""; }
const char* lgx_last_joints_kernel(lgx_handle* h) { return h ? h->last_joints_kernel : ""; }

int lgx_ridge(lgx_handle* h, const void* d_frames, int bits, int batch, int height, int width, size_t pitch_bytes,
              size_t frame_stride_bytes, double* d_b, double* d_rowsum_b, double* d_rowsum_b2, double* d_g, void* stream) {
  if (!geometry_ok(h, bits, batch, height, width) || !d_frames || !d_b || !d_rowsum_b || !d_rowsum_b2) return LGX_ERR_BAD_ARG;
  LGX_ON_DEVICE(h);
  const size_t ps = (size_t)height * plane_pitch(width);
  for (int c0 = 0; c0 < batch; c0 += h->chunk) {
    const int nb = batch - c0 < h->chunk ? batch - c0 : h->chunk;
    int rc = ridge_chunk(h, (const unsigned char*)d_frames + (size_t)c0 * frame_stride_bytes, bits, nb, height, width,
                         pitch_bytes, frame_stride_bytes, d_b + c0 * ps, d_rowsum_b + c0 * ps, d_rowsum_b2 + c0 * ps,
                         d_g ? d_g + c0 * ps : nullptr, nullptr, (cudaStream_t)stream);
    if (rc) return rc;
  }
  return LGX_OK;
}

int lgx_sauvola(lgx_handle* h, const double* d_b, const double* d_rowsum_b, const double* d_rowsum_b2, int batch,
                int height, int width, uint8_t* d_binary, uint32_t* d_bits, double* d_T, void* stream) {
  if (!geometry_ok(h, 8, batch, height, width) || !d_b || !d_rowsum_b || !d_rowsum_b2 || !d_bits) return LGX_ERR_BAD_ARG;
  if (batch == 0) return LGX_OK;
  LGX_ON_DEVICE(h);
  SauvolaParams sp{};
  sp.b = d_b; sp.rsb = d_rowsum_b; sp.rsb2 = d_rowsum_b2;
  sp.H = height; sp.W = width; sp.Wp = plane_pitch(width); sp.WW = bits_pitch(width);
  sp.plane_stride = (size_t)height * sp.Wp;
  sp.binary = d_binary; sp.bits = d_bits; sp.T = d_T;
  LGX_CK(launch_sauvola(sp, batch, h->sauvola_variant, (cudaStream_t)stream));
  return LGX_OK;
}

int lgx_frontend(lgx_handle* h, const void* d_frames, int bits, int batch, int height, int width, size_t pitch_bytes,
                 size_t frame_stride_bytes, uint8_t* d_binary, uint8_t* d_hmask, uint8_t* d_vmask, void* d_blurred,
                 int32_t* d_centroids, double* d_centroids_f, int max_centroids, int32_t* d_counts, uint32_t* d_flags,
                 void* stream) {
  if (!geometry_ok(h, bits, batch, height, width) || !d_frames || !d_centroids || !d_counts || !d_flags || max_centroids < 1)
    return LGX_ERR_BAD_ARG;
  if (pitch_bytes < (size_t)width * (bits / 8) || frame_stride_bytes < pitch_bytes * (size_t)height) return LGX_ERR_BAD_ARG;
  LGX_ON_DEVICE(h);
  cudaStream_t st = (cudaStream_t)stream;
  const int H = height, W = width;
  const size_t npix = (size_t)H * W;
  const size_t mplane = h->packed ? (size_t)H * bits_pitch(W) * sizeof(uint32_t) : npix;   // bytes of one mask plane in the caller's buffers
  const int pixb = bits / 8;
  for (int c0 = 0; c0 < batch; c0 += h->chunk) {
    const int nb = batch - c0 < h->chunk ? batch - c0 : h->chunk;
    LGX_CK(cudaMemsetAsync(d_flags + c0, 0, (size_t)nb * sizeof(uint32_t), st));
    const unsigned char* fr = (const unsigned char*)d_frames + (size_t)c0 * frame_stride_bytes;
    void* bl = d_blurred ? (unsigned char*)d_blurred + (size_t)c0 * npix * pixb : nullptr;
    int rc = mark(h, st);
    if (rc) return rc;
    SauvolaParams sp{};
    sp.b = h->b; sp.rsb = h->rsb; sp.rsb2 = h->rsb2;
    sp.H = H; sp.W = W; sp.Wp = plane_pitch(W); sp.WW = bits_pitch(W);
    sp.plane_stride = (size_t)H * sp.Wp;
    sp.binary = (d_binary && !h->packed) ? d_binary + (size_t)c0 * npix : nullptr;
    sp.bits = h->bits; sp.T = nullptr;
    rc = fused_chunk(h, fr, bits, nb, H, W, pitch_bytes, frame_stride_bytes, bl, h->bits, nullptr, nullptr, st, true);
    if (rc < 0) return rc;
    if (rc == LGX_OK) {
      // fused: the bit plane is the kernel's only output; the dense binary_img (if asked for) is unpacked from it
      if ((rc = mark(h, st))) return rc;
      if (sp.binary) LGX_CK(launch_unpack_bits(h->bits, nb, H, W, sp.binary, st));
      if ((rc = mark(h, st))) return rc;
      h->launches -= 1;   // (no sauvola kernel; the unpack kernel replaces it when binary_img is requested)
      if (sp.binary) h->launches += 1;
    } else {
      rc = ridge_chunk(h, fr, bits, nb, H, W, pitch_bytes, frame_stride_bytes, h->b, h->rsb, h->rsb2, nullptr, bl, st, true);
      if (rc) return rc;
      if ((rc = mark(h, st))) return rc;
      LGX_CK(launch_sauvola(sp, nb, h->sauvola_variant, st));
      if ((rc = mark(h, st))) return rc;
    }
    MorphParams mp{};
    mp.bits = h->bits; mp.H = H; mp.W = W; mp.WW = sp.WW;
    if (h->packed) {
      if (d_binary) LGX_CK(cudaMemcpyAsync(d_binary + (size_t)c0 * mplane, h->bits, (size_t)nb * mplane, cudaMemcpyDeviceToDevice, st));
      mp.hbits = d_hmask ? reinterpret_cast<uint32_t*>(d_hmask + (size_t)c0 * mplane) : nullptr;
      mp.vbits = d_vmask ? reinterpret_cast<uint32_t*>(d_vmask + (size_t)c0 * mplane) : nullptr;
    } else {
      mp.hmask = d_hmask ? d_hmask + (size_t)c0 * npix : nullptr;
      mp.vmask = d_vmask ? d_vmask + (size_t)c0 * npix : nullptr;
    }
    mp.jbits = h->jbits;
    if (joints_whole_frame(h, W)) {     // the whole-frame union-find wants its seeds and the list of non-empty words
      mp.lab = h->lab;
      mp.active = h->active;
      mp.nactive = h->active + (size_t)h->chunk * H * bits_pitch(W);
      LGX_CK(cudaMemsetAsync(mp.nactive, 0, (size_t)nb * sizeof(int32_t), st));
    }
    LGX_CK(launch_morph(mp, nb, st));
    if ((rc = mark(h, st))) return rc;
    rc = run_joints(h, nb, H, W, d_centroids + (size_t)c0 * max_centroids * 2,
                    d_centroids_f ? d_centroids_f + (size_t)c0 * max_centroids * 2 : nullptr, max_centroids,
                    d_counts + c0, d_flags + c0, true, st);
    if (rc) return rc;
    if ((rc = mark(h, st))) return rc;
    h->launches += 4;   // blur5, ridge, sauvola, morph
    if (h->timing) h->chunks_timed++;
  }
  return LGX_OK;
}

int lgx_get_stats(lgx_handle* h, double* ms5, long long* chunks, long long* launches, int reset) {
  if (!h) return LGX_ERR_BAD_ARG;
  LGX_ON_DEVICE(h);
  if (h->ev_used) {
    LGX_CK(cudaEventSynchronize(h->evs[h->ev_used - 1]));
    for (size_t i = 0; i + 5 < h->ev_used && i + 5 < h->evs.size(); i += 6)
      for (int j = 0; j < 5; ++j) {
        float ms = 0;
        LGX_CK(cudaEventElapsedTime(&ms, h->evs[i + j], h->evs[i + j + 1]));
        h->ms_acc[j] += ms;
      }
    h->ev_used = 0;
  }
  if (ms5) for (int j = 0; j < 5; ++j) ms5[j] = h->ms_acc[j];
  if (chunks) *chunks = h->chunks_timed;
  if (launches) *launches = h->launches;
  if (reset) {
    for (int j = 0; j < 5; ++j) h->ms_acc[j] = 0;
    h->chunks_timed = 0;
    h->launches = 0;
  }
  return LGX_OK;
}

int lgx_get_ridge_prof(lgx_handle* h, unsigned long long* out16, int reset) {
  if (!h || !out16 || !h->prof) return LGX_ERR_BAD_ARG;
  LGX_ON_DEVICE(h);
  LGX_CK(cudaDeviceSynchronize());
  LGX_CK(cudaMemcpy(out16, h->prof, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  if (reset) LGX_CK(cudaMemset(h->prof, 0, 16 * sizeof(unsigned long long)));
  return LGX_OK;
}

int lgx_contour_centroids(lgx_handle* h, const uint8_t* d_mask, int batch, int height, int width, int32_t* d_centroids,
                          double* d_centroids_f, int max_centroids, int32_t* d_counts, uint32_t* d_flags, void* stream) {
  if (!geometry_ok(h, 8, batch, height, width) || !d_mask || !d_centroids || !d_counts || !d_flags || max_centroids < 1)
    return LGX_ERR_BAD_ARG;
  LGX_ON_DEVICE(h);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t npix = (size_t)height * width;
  for (int c0 = 0; c0 < batch; c0 += h->chunk) {
    const int nb = batch - c0 < h->chunk ? batch - c0 : h->chunk;
    LGX_CK(cudaMemsetAsync(d_flags + c0, 0, (size_t)nb * sizeof(uint32_t), st));
    LGX_CK(launch_pack_bits(d_mask + (size_t)c0 * npix, nb, height, width, h->jbits, st));
    int rc = run_joints(h, nb, height, width, d_centroids + (size_t)c0 * max_centroids * 2,
                        d_centroids_f ? d_centroids_f + (size_t)c0 * max_centroids * 2 : nullptr, max_centroids,
                        d_counts + c0, d_flags + c0, false, st);
    if (rc) return rc;
  }
  return LGX_OK;
}

int lgx_extract_joints(lgx_handle* h, const uint8_t* d_binary, int batch, int height, int width, uint8_t* d_hmask,
                       uint8_t* d_vmask, int32_t* d_centroids, double* d_centroids_f, int max_centroids,
                       int32_t* d_counts, uint32_t* d_flags, void* stream) {
  if (!geometry_ok(h, 8, batch, height, width) || !d_binary || !d_centroids || !d_counts || !d_flags || max_centroids < 1)
    return LGX_ERR_BAD_ARG;
  LGX_ON_DEVICE(h);
  cudaStream_t st = (cudaStream_t)stream;
  const int H = height, W = width;
  const size_t npix = (size_t)H * W;
  for (int c0 = 0; c0 < batch; c0 += h->chunk) {
    const int nb = batch - c0 < h->chunk ? batch - c0 : h->chunk;
    LGX_CK(cudaMemsetAsync(d_flags + c0, 0, (size_t)nb * sizeof(uint32_t), st));
    LGX_CK(launch_pack_bits(d_binary + (size_t)c0 * npix, nb, H, W, h->bits, st));
    MorphParams mp{};
    mp.bits = h->bits; mp.H = H; mp.W = W; mp.WW = bits_pitch(W);
    mp.hmask = d_hmask ? d_hmask + (size_t)c0 * npix : nullptr;
    mp.vmask = d_vmask ? d_vmask + (size_t)c0 * npix : nullptr;
    mp.jbits = h->jbits;
    if (joints_whole_frame(h, W)) {     // the whole-frame union-find wants its seeds and the list of non-empty words
      mp.lab = h->lab;
      mp.active = h->active;
      mp.nactive = h->active + (size_t)h->chunk * H * bits_pitch(W);
      LGX_CK(cudaMemsetAsync(mp.nactive, 0, (size_t)nb * sizeof(int32_t), st));
    }
    LGX_CK(launch_morph(mp, nb, st));
    int rc = run_joints(h, nb, H, W, d_centroids + (size_t)c0 * max_centroids * 2,
                        d_centroids_f ? d_centroids_f + (size_t)c0 * max_centroids * 2 : nullptr, max_centroids,
                        d_counts + c0, d_flags + c0, true, st);
    if (rc) return rc;
  }
  return LGX_OK;
}

// Host-buffer entry point.  Three device mirrors ("slots") rotate: the copy-in of chunk c+1 is queued on an internal
// stream before the host waits for the counts of chunk c-1, so the H2D engine never idles behind the host; chunk c
// computes on `stream`; the outputs of chunk c-1 leave on a third stream (fully overlapped when the caller's buffers
// are pinned; pageable buffers work but serialise inside the driver).
int lgx_frontend_host(lgx_handle* h, const void* frames, int bits, int batch, int height, int width, uint8_t* binary,
                      uint8_t* hmask, uint8_t* vmask, void* blurred, int32_t* centroids, double* centroids_f,
                      int max_centroids, int32_t* counts, uint32_t* flags, void* stream) {
  if (!geometry_ok(h, bits, batch, height, width) || !frames || !centroids || !counts || max_centroids < 1) return LGX_ERR_BAD_ARG;
  if (batch == 0) return LGX_OK;
  cudaStream_t st = (cudaStream_t)stream;
  LGX_ON_DEVICE(h);
  if (!h->s_in) {
    LGX_CK(cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking));
    LGX_CK(cudaEventCreateWithFlags(&h->ev_start, cudaEventDisableTiming));
    for (int s = 0; s < kHostSlots; ++s) {
      LGX_CK(cudaStreamCreateWithFlags(&h->s_out[s], cudaStreamNonBlocking));
      LGX_CK(cudaEventCreateWithFlags(&h->ev_in[s], cudaEventDisableTiming));
      LGX_CK(cudaEventCreateWithFlags(&h->ev_done[s], cudaEventDisableTiming));
      LGX_CK(cudaEventCreateWithFlags(&h->ev_small[s], cudaEventDisableTiming));
      LGX_CK(cudaEventCreateWithFlags(&h->ev_out[s], cudaEventDisableTiming));
    }
  }
  const size_t npix = (size_t)height * width;
  const size_t mplane = h->packed ? (size_t)height * bits_pitch(width) * sizeof(uint32_t) : npix;   // bytes of one mask plane
  const int pixb = bits / 8;
  const int nbmax = batch < h->chunk ? batch : h->chunk;
  // Chunk schedule: the first chunk is split 1/4 + 3/4 so that the pipeline fill (the one copy-in that overlaps
  // nothing) is a quarter as long; every other chunk is full size (a geometric ramp 8, 12, 18, 27 ... was measured
  // slower: short chunks pay more for their small launches than the fill saves).
  std::vector<int> cstart;
  const int split = (h->split_first && h->chunk >= 16 && batch > h->chunk) ? h->chunk / 4 : 0;
  for (int c0 = 0; c0 < batch; c0 += h->chunk) {
    cstart.push_back(c0);
    if (c0 == 0 && split) cstart.push_back(split);
  }
  cstart.push_back(batch);
  const int nchunks = (int)cstart.size() - 1;
  const int nslots = nchunks < kHostSlots ? nchunks : kHostSlots;
  auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
  // device mirrors for one chunk
  const size_t o_in = 0;
  const size_t o_bin = o_in + up(nbmax * npix * pixb);
  const size_t o_h = o_bin + up(binary ? nbmax * mplane : 0);
  const size_t o_v = o_h + up(hmask ? nbmax * mplane : 0);
  const size_t o_bl = o_v + up(vmask ? nbmax * mplane : 0);
  const size_t o_c = o_bl + up(blurred ? nbmax * npix * pixb : 0);
  const size_t o_cf = o_c + up((size_t)nbmax * max_centroids * 2 * sizeof(int32_t));
  const size_t o_n = o_cf + up(centroids_f ? (size_t)nbmax * max_centroids * 2 * sizeof(double) : 0);
  const size_t o_fl = o_n + up((size_t)nbmax * sizeof(int32_t));
  const size_t slot_bytes = o_fl + up((size_t)nbmax * sizeof(uint32_t));
  if (nslots * slot_bytes > h->host_dev_bytes) {
    LGX_CK(cudaDeviceSynchronize());
    if (h->host_dev) cudaFree(h->host_dev);
    h->host_dev = nullptr; h->host_dev_bytes = 0;
    if (cudaMalloc((void**)&h->host_dev, nslots * slot_bytes) != cudaSuccess) { cudaGetLastError(); return LGX_ERR_OOM; }
    h->host_dev_bytes = nslots * slot_bytes;
  }
  std::vector<uint32_t> flag_tmp;
  if (!flags) { flag_tmp.resize(batch); flags = flag_tmp.data(); }
  // order the internal copy-in stream after whatever the caller already queued on `stream`
  LGX_CK(cudaEventRecord(h->ev_start, st));
  LGX_CK(cudaStreamWaitEvent(h->s_in, h->ev_start, 0));

  auto chunk_nb = [&](int c) { return cstart[c + 1] - cstart[c]; };

  auto copy_in = [&](int c) -> int {       // chunk c -> its slot, as soon as the compute of chunk c-nslots has read it
    const int s = c % nslots, c0 = cstart[c], nb = chunk_nb(c);
    unsigned char* d = h->host_dev + (size_t)s * slot_bytes;
    if (c >= nslots) LGX_CK(cudaStreamWaitEvent(h->s_in, h->ev_done[s], 0));
    LGX_CK(cudaMemcpyAsync(d + o_in, (const unsigned char*)frames + (size_t)c0 * npix * pixb, (size_t)nb * npix * pixb,
                           cudaMemcpyHostToDevice, h->s_in));
    LGX_CK(cudaEventRecord(h->ev_in[s], h->s_in));
    return LGX_OK;
  };

  auto compute = [&](int c) -> int {       // chunk c on the caller's stream, then its fixed-size outputs on s_out
    const int s = c % nslots, c0 = cstart[c], nb = chunk_nb(c);
    unsigned char* d = h->host_dev + (size_t)s * slot_bytes;
    LGX_CK(cudaStreamWaitEvent(st, h->ev_in[s], 0));
    if (c >= nslots) LGX_CK(cudaStreamWaitEvent(st, h->ev_out[s], 0));     // outputs of chunk c-nslots have left the slot
    int rc = lgx_frontend(h, d + o_in, bits, nb, height, width, (size_t)width * pixb, npix * pixb,
                          binary ? d + o_bin : nullptr, hmask ? d + o_h : nullptr, vmask ? d + o_v : nullptr,
                          blurred ? d + o_bl : nullptr, (int32_t*)(d + o_c), centroids_f ? (double*)(d + o_cf) : nullptr,
                          max_centroids, (int32_t*)(d + o_n), (uint32_t*)(d + o_fl), st);
    if (rc) return rc;
    LGX_CK(cudaEventRecord(h->ev_done[s], st));
    cudaStream_t so = h->s_out[s];
    LGX_CK(cudaStreamWaitEvent(so, h->ev_done[s], 0));
    LGX_CK(cudaMemcpyAsync(counts + c0, d + o_n, (size_t)nb * sizeof(int32_t), cudaMemcpyDeviceToHost, so));
    LGX_CK(cudaMemcpyAsync(flags + c0, d + o_fl, (size_t)nb * sizeof(uint32_t), cudaMemcpyDeviceToHost, so));
    LGX_CK(cudaEventRecord(h->ev_small[s], so));
    if (binary) LGX_CK(cudaMemcpyAsync(binary + (size_t)c0 * mplane, d + o_bin, (size_t)nb * mplane, cudaMemcpyDeviceToHost, so));
    if (hmask) LGX_CK(cudaMemcpyAsync(hmask + (size_t)c0 * mplane, d + o_h, (size_t)nb * mplane, cudaMemcpyDeviceToHost, so));
    if (vmask) LGX_CK(cudaMemcpyAsync(vmask + (size_t)c0 * mplane, d + o_v, (size_t)nb * mplane, cudaMemcpyDeviceToHost, so));
    if (blurred) LGX_CK(cudaMemcpyAsync((unsigned char*)blurred + (size_t)c0 * npix * pixb, d + o_bl, (size_t)nb * npix * pixb, cudaMemcpyDeviceToHost, so));
    return LGX_OK;
  };

  auto finalize = [&](int c) -> int {      // chunk c: counts are on the host -> copy the used part of each list
    const int s = c % nslots, c0 = cstart[c], nb = chunk_nb(c);
    unsigned char* d = h->host_dev + (size_t)s * slot_bytes;
    LGX_CK(cudaEventSynchronize(h->ev_small[s]));
    for (int f = 0; f < nb; ++f) {
      int n = counts[c0 + f] < max_centroids ? counts[c0 + f] : max_centroids;
      if (n <= 0) continue;
      LGX_CK(cudaMemcpyAsync(centroids + ((size_t)(c0 + f) * max_centroids) * 2, d + o_c + (size_t)f * max_centroids * 2 * sizeof(int32_t),
                             (size_t)n * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, h->s_out[s]));
      if (centroids_f)
        LGX_CK(cudaMemcpyAsync(centroids_f + ((size_t)(c0 + f) * max_centroids) * 2, d + o_cf + (size_t)f * max_centroids * 2 * sizeof(double),
                               (size_t)n * 2 * sizeof(double), cudaMemcpyDeviceToHost, h->s_out[s]));
    }
    LGX_CK(cudaEventRecord(h->ev_out[s], h->s_out[s]));
    return LGX_OK;
  };

  int rc = copy_in(0);
  if (rc) return rc;
  for (int c = 0; c < nchunks; ++c) {
    // chunk c+1's slot was last used by chunk c+1-nslots <= c-2 (nslots = 3), whose compute is already queued
    if (c + 1 < nchunks && (rc = copy_in(c + 1))) return rc;
    if ((rc = compute(c))) return rc;
    if (c >= 1 && (rc = finalize(c - 1))) return rc;
  }
  if ((rc = finalize(nchunks - 1))) return rc;
  for (int s = 0; s < nslots; ++s) LGX_CK(cudaStreamSynchronize(h->s_out[s]));
  LGX_CK(cudaStreamSynchronize(st));
  return LGX_OK;
}

int lgx_debug_contours(lgx_handle* h, int frame_in_chunk, int64_t* out_host, int capacity, int* n_out) {
  if (!h || !out_host || !n_out || frame_in_chunk < 0 || frame_in_chunk >= h->last_n) return LGX_ERR_BAD_ARG;
  LGX_ON_DEVICE(h);
  LGX_CK(cudaDeviceSynchronize());
  int32_t n = 0;
  LGX_CK(cudaMemcpy(&n, h->ncomp + frame_in_chunk, sizeof(int32_t), cudaMemcpyDeviceToHost));
  std::vector<int32_t> rp(n);
  std::vector<unsigned long long> acc((size_t)n * 4);
  if (n) {
    LGX_CK(cudaMemcpy(rp.data(), h->rootpix + (size_t)frame_in_chunk * h->max_comp, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost));
    LGX_CK(cudaMemcpy(acc.data(), h->acc + (size_t)frame_in_chunk * h->max_comp * 4, (size_t)n * 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  }
  int m = n < capacity ? n : capacity;
  for (int i = 0; i < m; ++i) {
    int k = n - 1 - i;
    out_host[4 * i + 0] = rp[k];
    out_host[4 * i + 1] = (int64_t)(acc[(size_t)k * 4] & 0xffffffffull);
    out_host[4 * i + 2] = (int64_t)acc[(size_t)k * 4 + 1];
    out_host[4 * i + 3] = (int64_t)acc[(size_t)k * 4 + 2];
  }
  *n_out = n;
  return LGX_OK;
}

}  // extern "C"
