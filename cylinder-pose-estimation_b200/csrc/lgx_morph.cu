// K3 "open_hv": cv2.morphologyEx(MORPH_OPEN) with a 20x1 and a 1x20 rectangle, and their AND,
// on bit-packed rows (32 pixels per word).
//
// Replaces, bit for bit: /root/reference/utils/util_cylinder.py:1810-1816.  Pixel rule (SURVEY.md App. A.9,
// CPU twin oracle/restate.py open_line): erode = AND over offsets [-10,+9] with outside = white,
// dilate = OR over the SAME offsets with outside = black (so the result is the true opening shifted by
// +1 px, and border-touching runs survive when shorter).  Outputs the two masks as u8 planes (what
// the reference's later stages read) and the joints mask as a bit plane for the contour stage.
#include "lgx_internal.cuh"

namespace lgx {
namespace {

constexpr int TR = 32;            // output rows per CTA
constexpr int TWW = 32;           // output words per CTA (1024 px)
constexpr int IN_R = TR + 38;     // input rows y0-20 .. y0+TR+17
constexpr int IN_W = TWW + 2;     // input words w0-1 .. w0+TWW
constexpr int ER_ROWS = TR + 19;  // vertically eroded rows y0-10 .. y0+TR+8

// bit i of the result = bit (32 + i + s) of the 96-bit string p:c:n (p lowest), s in [-31, 31]
__device__ __forceinline__ uint32_t win(uint32_t p, uint32_t c, uint32_t n, int s) {
  return s >= 0 ? __funnelshift_r(c, n, s) : __funnelshift_r(p, c, 32 + s);
}

__device__ __forceinline__ uint32_t and20(uint32_t p, uint32_t c, uint32_t n) {
  uint32_t r = 0xffffffffu;
#pragma unroll
  for (int d = -10; d <= 9; ++d) r &= win(p, c, n, d);
  return r;
}
__device__ __forceinline__ uint32_t or20(uint32_t p, uint32_t c, uint32_t n) {
  uint32_t r = 0;
#pragma unroll
  for (int d = -10; d <= 9; ++d) r |= win(p, c, n, d);
  return r;
}

__device__ __forceinline__ uint32_t valid_mask(int w, int W) {
  int rem = W - w * 32;
  return rem >= 32 ? 0xffffffffu : (rem <= 0 ? 0u : ((1u << rem) - 1u));
}

// 4 mask bits -> 4 bytes of 0x00 / 0xff
__device__ __forceinline__ uint32_t spread4(uint32_t n) { return ((n * 0x00204081u) & 0x01010101u) * 0xffu; }

__device__ __forceinline__ void store_mask_row(uint8_t* __restrict__ dst, uint32_t bits, int x0, int W) {
  if (x0 + 32 <= W && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    uint4 a, b;
    a.x = spread4(bits & 15u); a.y = spread4((bits >> 4) & 15u); a.z = spread4((bits >> 8) & 15u); a.w = spread4((bits >> 12) & 15u);
    b.x = spread4((bits >> 16) & 15u); b.y = spread4((bits >> 20) & 15u); b.z = spread4((bits >> 24) & 15u); b.w = spread4(bits >> 28);
    reinterpret_cast<uint4*>(dst)[0] = a;
    reinterpret_cast<uint4*>(dst)[1] = b;
  } else {
    for (int i = 0; i < 32 && x0 + i < W; ++i) dst[i] = ((bits >> i) & 1u) ? 255 : 0;
  }
}

__global__ void __launch_bounds__(256) morph_kernel(const MorphParams p) {
  __shared__ uint32_t s_in[IN_R][IN_W + 1];
  __shared__ uint32_t s_er[ER_ROWS][TWW + 1];
  const int H = p.H, W = p.W, WW = p.WW;
  const int frame = blockIdx.z;
  const int w0 = blockIdx.x * TWW;
  const int y0 = blockIdx.y * TR;
  const uint32_t* __restrict__ in = p.bits + (size_t)frame * H * WW;
  const int tid = threadIdx.y * 32 + threadIdx.x;

  // input rows y0-20 .. y0+TR+17, words w0-1 .. w0+TWW; outside the image = white (erode border rule)
  for (int idx = tid; idx < IN_R * IN_W; idx += 256) {
    int r = idx / IN_W, j = idx - r * IN_W;
    int y = y0 - 20 + r, w = w0 - 1 + j;
    uint32_t v = 0xffffffffu;
    if (y >= 0 && y < H && w >= 0 && w < WW) v = in[(size_t)y * WW + w] | ~valid_mask(w, W);
    s_in[r][j] = v;
  }
  __syncthreads();

  const int lw = threadIdx.x;           // word within tile
  const int w = w0 + lw;
  const uint32_t vm = (w < WW) ? valid_mask(w, W) : 0u;

  // vertical erode: er row r <-> y' = y0-10+r; in[y'+d] is input row r+10+d, d in [-10, 9]
  for (int r = threadIdx.y; r < ER_ROWS; r += 8) {
    uint32_t e = 0xffffffffu;
#pragma unroll
    for (int d = 0; d < 20; ++d) e &= s_in[r + d][lw + 1];
    int y = y0 - 10 + r;
    s_er[r][lw] = (y >= 0 && y < H) ? (e & vm) : 0u;   // dilate border rule: outside = black
  }
  __syncthreads();

  for (int r = threadIdx.y; r < TR; r += 8) {
    const int y = y0 + r;
    if (y >= H || w >= WW) continue;
    // vertical dilate: out[y] = OR_{d=-10..9} er[y+d]; er row of y+d is r+10+d
    uint32_t v = 0;
#pragma unroll
    for (int d = 0; d < 20; ++d) v |= s_er[r + d][lw];
    v &= vm;
    // horizontal open on row y (input row r+20).  Only bits >= 22 of the previous word's erosion and
    // bits <= 8 of the next word's are consumed, and those never depend on words w-2 / w+2.
    const uint32_t* row = &s_in[r + 20][lw];   // row[0] = word w-1, row[1] = w, row[2] = w+1
    const uint32_t pw = row[0], cw = row[1], nw = row[2];
    const uint32_t e_c = and20(pw, cw, nw) & vm;
    const uint32_t e_p = (w >= 1) ? and20(0xffffffffu, pw, cw) : 0u;
    const uint32_t e_n = (w + 1 < WW) ? (and20(cw, nw, 0xffffffffu) & valid_mask(w + 1, W)) : 0u;
    const uint32_t hbits = or20(e_p, e_c, e_n) & vm;
    const size_t row_o = (size_t)frame * H + y;
    const uint32_t j = hbits & v;
    p.jbits[row_o * WW + w] = j;
    if (p.active && j) {   // compacted list of non-empty words (warp-aggregated append; order is irrelevant)
      const unsigned am = __activemask();
      const int leader = __ffs(am) - 1;
      int base = 0;
      if ((int)threadIdx.x == leader) base = atomicAdd(&p.nactive[frame], __popc(am));
      base = __shfl_sync(am, base, leader);
      p.active[(size_t)frame * H * WW + base + __popc(am & ((1u << threadIdx.x) - 1u))] = y * WW + w;
    }
    if (p.lab && j) {   // union-find seed of the contour stage: every word-run start is its own parent
      uint32_t starts = j & ~(j << 1);
      const int base = y * W + w * 32;
      int32_t* L = p.lab + (size_t)frame * H * W + base;
      while (starts) {
        const int s = __ffs(starts) - 1;
        starts &= starts - 1;
        L[s] = base + s;
      }
    }
    if (p.hmask) store_mask_row(p.hmask + row_o * W + w * 32, hbits, w * 32, W);
    if (p.vmask) store_mask_row(p.vmask + row_o * W + w * 32, v, w * 32, W);
  }
}

}  // namespace

cudaError_t launch_morph(const MorphParams& p, int batch, cudaStream_t stream) {
  dim3 grid((p.WW + TWW - 1) / TWW, (p.H + TR - 1) / TR, batch);
  dim3 block(32, 8);
  morph_kernel<<<grid, block, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace lgx
