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

// 20-wide AND / OR of a 64-bit string by doubling: r[k] = op(x[k] .. x[k+19]) (bits shifted in from above are the
// identity of the other operation's border rule, see callers)
__device__ __forceinline__ uint64_t and20_64(uint64_t x) {
  const uint64_t c2 = x & (x >> 1);
  const uint64_t c4 = c2 & (c2 >> 2);
  const uint64_t c8 = c4 & (c4 >> 4);
  const uint64_t c16 = c8 & (c8 >> 8);
  return c16 & (c4 >> 16);
}
__device__ __forceinline__ uint64_t or20_64(uint64_t x) {
  const uint64_t c2 = x | (x >> 1);
  const uint64_t c4 = c2 | (c2 >> 2);
  const uint64_t c8 = c4 | (c4 >> 4);
  const uint64_t c16 = c8 | (c8 >> 8);
  return c16 | (c4 >> 16);
}

// horizontal open of word `cw` given its neighbours (S = pw:cw:nw, bit 0 of S = bit 0 of pw, pixel x = 32(w-1)+j).
// erode: er[j] = AND S[j-10 .. j+9]; dilate: out[i] = OR er[i-10 .. i+9] for i = 32..63, er outside the image = 0.
__device__ __forceinline__ uint32_t open_h_word(uint32_t pw, uint32_t cw, uint32_t nw, int w, int W) {
  const uint64_t lo = ((uint64_t)cw << 32) | pw;            // S[0..63]
  const uint64_t hi = ((uint64_t)nw << 32) | cw;            // S[32..95]
  const uint64_t XA = (lo >> 12) | (hi << 20);              // S[12..75]
  const uint64_t ea = and20_64(XA);                         // ea[k] = er[22+k], valid k = 0..44
  const uint64_t eb = and20_64(hi);                         // eb[k] = er[42+k], valid k = 0..44
  uint64_t Y = (ea & ((1ull << 45) - 1ull)) | ((eb >> 25) << 45);   // Y[k] = er[22+k], k = 0..50 (+ junk above, masked)
  // er[22+k] is pixel x = 32w - 10 + k: black outside [0, W)
  int vlo = 10 - 32 * w;                                    // first valid k
  int vhi = W - 32 * w + 10;                                // one past the last valid k
  uint64_t mask = (1ull << 51) - 1ull;
  if (vlo > 0) mask &= ~((1ull << vlo) - 1ull);
  if (vhi < 51) mask &= (vhi <= 0) ? 0ull : ((1ull << vhi) - 1ull);
  Y &= mask;
  return (uint32_t)or20_64(Y);                              // out[32+m] = OR Y[m .. m+19], m = 0..31
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

#ifndef LGX_MORPH_MINB
#define LGX_MORPH_MINB 1
#endif
__global__ void __launch_bounds__(256, LGX_MORPH_MINB) morph_kernel(const MorphParams p) {
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
  const int ty = threadIdx.y;
  const int w = w0 + lw;
  const uint32_t vm = (w < WW) ? valid_mask(w, W) : 0u;

  // vertical erode by doubling: thread (lw, ty) owns eroded rows e0 .. e0+6 (er row r <-> y' = y0-10+r; it is the AND
  // of input rows r .. r+19)
  {
    const int e0 = ty * 7;
    if (e0 < ER_ROWS) {
      uint32_t x[26];
#pragma unroll
      for (int i = 0; i < 26; ++i) x[i] = (e0 + i < IN_R) ? s_in[e0 + i][lw + 1] : 0xffffffffu;
      uint32_t c2[25], c4[23], c8[15];
#pragma unroll
      for (int i = 0; i < 25; ++i) c2[i] = x[i] & x[i + 1];
#pragma unroll
      for (int i = 0; i < 23; ++i) c4[i] = c2[i] & c2[i + 2];
#pragma unroll
      for (int i = 0; i < 15; ++i) c8[i] = c4[i] & c4[i + 4];
#pragma unroll
      for (int i = 0; i < 7; ++i) {
        const int r = e0 + i;
        if (r < ER_ROWS) {
          const uint32_t e = c8[i] & c8[i + 8] & c4[i + 16];
          const int y = y0 - 10 + r;
          s_er[r][lw] = (y >= 0 && y < H) ? (e & vm) : 0u;   // dilate border rule: outside = black
        }
      }
    }
  }
  __syncthreads();

  // vertical dilate by doubling + horizontal open: thread (lw, ty) owns output rows 4*ty .. 4*ty+3
  {
    const int r0 = ty * 4;
    uint32_t e[23];
#pragma unroll
    for (int i = 0; i < 23; ++i) e[i] = (r0 + i < ER_ROWS) ? s_er[r0 + i][lw] : 0u;
    uint32_t d2[22], d4[20], d8[12];
#pragma unroll
    for (int i = 0; i < 22; ++i) d2[i] = e[i] | e[i + 1];
#pragma unroll
    for (int i = 0; i < 20; ++i) d4[i] = d2[i] | d2[i + 2];
#pragma unroll
    for (int i = 0; i < 12; ++i) d8[i] = d4[i] | d4[i + 4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = r0 + i;
      const int y = y0 + r;
      if (y >= H || w >= WW) continue;
      const uint32_t v = (d8[i] | d8[i + 8] | d4[i + 16]) & vm;      // out[y] = OR er rows r .. r+19
      const uint32_t* row = &s_in[r + 20][lw];                        // row[0] = word w-1, row[1] = w, row[2] = w+1
      const uint32_t hbits = open_h_word(row[0], row[1], row[2], w, W) & vm;
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
      if (p.hbits) p.hbits[row_o * WW + w] = hbits;
      if (p.vbits) p.vbits[row_o * WW + w] = v;
      if (p.hmask) store_mask_row(p.hmask + row_o * W + w * 32, hbits, w * 32, W);
      if (p.vmask) store_mask_row(p.vmask + row_o * W + w * 32, v, w * 32, W);
    }
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
