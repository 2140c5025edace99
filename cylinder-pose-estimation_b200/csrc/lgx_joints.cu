// K4 "joints": the data-parallel, tracing-free equivalent of
//     cv2.findContours(joints, RETR_EXTERNAL, CHAIN_APPROX_SIMPLE) + cv2.moments(contour) + int(m10/m00), int(m01/m00)
// i.e. /root/reference/utils/util_cylinder.py:1817-1825, in the reference's list order.
//
// Equivalence used (CPU twin oracle/restate.py contour_sums, checked against cv2 by
// tests/test_oracle.py::test_quad_sums_match_findcontours_moments): one reported contour == one 8-connected component of the hole-filled mask;
// contour order == descending raster index of the component's first pixel; the contour's Green sums
// (a00,a10,a01) == sums of per-2x2-quad integer terms over the filled component.
//
// Work is done on bit-packed rows.  Union-find elements are "word-runs" (maximal runs of set bits inside one
// 32-bit word), identified by the raster index of their first pixel, so the parent array is a sparse
// [H*W] int32 map touched only at run starts, and a component's root is its first raster pixel for free.
//   seed -> union (left word + three words above) -> flatten/roots -> rank (deterministic scan, ascending
//   raster order) -> per-run quad sums (integer atomics, order independent) -> emit (descending order scan).
// Holes (rare: ~1e-4 of components) are detected per component from the Euler number of its quads
// (E = (Q1 - Q3 - 2 QD)/4 = 1 - holes); frames that have one get a whole-frame background flood
// (band-parallel bit sweeps) and a second labelling pass on the filled mask.
#include "lgx_internal.cuh"
#include "lgx_joints.cuh"

namespace lgx {
namespace {

constexpr int kWordThreads = 256;

// word handled by this thread: dense indexing, or (first pass, when the morph kernel built it) the compacted list of
// non-empty words, which keeps warps full of similar work instead of a few active lanes with divergent loops
__device__ __forceinline__ int word_index(const JointsParams& p, int frame) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int NW = p.H * p.WW;
  if (p.active && p.pass == 0) {
    if (i >= min(p.nactive[frame], NW)) return -1;
    return p.active[(size_t)frame * NW + i];
  }
  return i < NW ? i : -1;
}

__device__ __forceinline__ bool frame_active(const JointsParams& p, int frame) {
  return p.pass == 0 || (p.flags[frame] & LGX_FLAG_GENERIC_FILL);
}

// ---- seed: every word-run start is its own parent -------------------------------------------
__device__ __forceinline__ void jl_seed_word(const JointsParams& p, const int frame, const int idx) {
  const uint32_t cur = p.jbits[(size_t)frame * p.H * p.WW + idx];
  if (!cur) return;
  const int y = idx / p.WW, w = idx - y * p.WW;
  int32_t* L = p.lab + (size_t)frame * p.H * p.W;
  uint32_t starts = cur & ~(cur << 1);
  const int base = y * p.W + w * 32;
  while (starts) {
    int s = __ffs(starts) - 1;
    starts &= starts - 1;
    L[base + s] = base + s;
  }
}

__global__ void __launch_bounds__(kWordThreads) jl_seed(const JointsParams p) {
  const int frame = blockIdx.y;
  if (!frame_active(p, frame)) return;
  if (p.pass == 0) {
    const int idx = word_index(p, frame);
    if (idx >= 0) jl_seed_word(p, frame, idx);
  } else {
    // second pass (frames that needed the whole-frame hole fill): small grid, grid-stride over all words
    for (int idx = blockIdx.x * kWordThreads + threadIdx.x; idx < p.H * p.WW; idx += gridDim.x * kWordThreads)
      jl_seed_word(p, frame, idx);
  }
}

// ---- union: link each word-run to the run left of it (across the word boundary) and to the runs it
// touches in the row above (8-connectivity: columns s-1 .. e+1) -----------------------------------
__device__ __forceinline__ void jl_union_word(const JointsParams& p, const int frame, const int idx) {
  const int H = p.H, W = p.W, WW = p.WW;
  const uint32_t* __restrict__ jb = p.jbits + (size_t)frame * H * WW;
  const uint32_t cur = jb[idx];
  if (!cur) return;
  const int y = idx / WW, w = idx - y * WW;
  int32_t* L = p.lab + (size_t)frame * H * W;
  const uint32_t left = (w > 0) ? jb[idx - 1] : 0u;
  uint32_t up_p = 0, up_c = 0, up_n = 0;
  if (y > 0) {
    const uint32_t* up = jb + (size_t)(y - 1) * WW;
    up_c = up[w];
    if (w > 0) up_p = up[w - 1];
    if (w + 1 < WW) up_n = up[w + 1];
  }
  const uint64_t U = (uint64_t)(up_p >> 31) | ((uint64_t)up_c << 1) | ((uint64_t)(up_n & 1u) << 33);
  uint32_t m = cur;
  while (m) {
    const int s = __ffs(m) - 1;
    const int len = run_len32(m, s);
    const int e = s + len - 1;
    const int id = y * W + w * 32 + s;
    if (s == 0 && (left >> 31)) {
      int st = run_start32(left, 31);
      uf_union(L, id, y * W + (w - 1) * 32 + st);
    }
    // window bits s .. e+2  <->  pixels x = 32w+s-1 .. 32w+e+1 of the row above
    uint64_t mask = ((1ull << (e + 3)) - 1ull) & ~((1ull << s) - 1ull);
    uint64_t mm = U & mask;
    while (mm) {
      const int i = __ffsll((long long)mm) - 1;
      uint32_t word;
      int bb, wu;
      if (i == 0) { word = up_p; bb = 31; wu = w - 1; }
      else if (i <= 32) { word = up_c; bb = i - 1; wu = w; }
      else { word = up_n; bb = 0; wu = w + 1; }
      const int st = run_start32(word, bb);
      uf_union(L, id, (y - 1) * W + wu * 32 + st);
      const uint64_t t2 = ~(U >> i);
      const int len2 = __ffsll((long long)t2) - 1;      // U < 2^34, so t2 != 0
      mm &= ~(((1ull << len2) - 1ull) << i);
    }
    m &= ~((len == 32 ? 0xffffffffu : ((1u << len) - 1u)) << s);
  }
}

__global__ void __launch_bounds__(kWordThreads) jl_union(const JointsParams p) {
  const int frame = blockIdx.y;
  if (!frame_active(p, frame)) return;
  if (p.pass == 0) {
    const int idx = word_index(p, frame);
    if (idx >= 0) jl_union_word(p, frame, idx);
  } else {
    // second pass (frames that needed the whole-frame hole fill): small grid, grid-stride over all words
    for (int idx = blockIdx.x * kWordThreads + threadIdx.x; idx < p.H * p.WW; idx += gridDim.x * kWordThreads)
      jl_union_word(p, frame, idx);
  }
}

// ---- flatten + root bits ---------------------------------------------------------------------
__device__ __forceinline__ void jl_roots_word(const JointsParams& p, const int frame, const int idx) {
  const int H = p.H, W = p.W, WW = p.WW;
  const uint32_t cur = p.jbits[(size_t)frame * H * WW + idx];
  uint32_t roots = 0;
  if (cur) {
    const int y = idx / WW, w = idx - y * WW;
    int32_t* L = p.lab + (size_t)frame * H * W;
    uint32_t starts = cur & ~(cur << 1);
    const int base = y * W + w * 32;
    while (starts) {
      int s = __ffs(starts) - 1;
      starts &= starts - 1;
      int r = uf_find(L, base + s);
      if (r == base + s) roots |= 1u << s;
      else L[base + s] = r;
    }
  }
  p.rootbits[(size_t)frame * H * WW + idx] = roots;
}

__global__ void __launch_bounds__(kWordThreads) jl_roots(const JointsParams p) {
  const int frame = blockIdx.y;
  if (!frame_active(p, frame)) return;
  if (p.pass == 0) {
    const int idx = word_index(p, frame);
    if (idx >= 0) jl_roots_word(p, frame, idx);
  } else {
    // second pass (frames that needed the whole-frame hole fill): small grid, grid-stride over all words
    for (int idx = blockIdx.x * kWordThreads + threadIdx.x; idx < p.H * p.WW; idx += gridDim.x * kWordThreads)
      jl_roots_word(p, frame, idx);
  }
}

// ---- rank: ascending raster order of roots (deterministic).  kRankSegs CTAs per frame, each owning a contiguous
// segment of the root-bit plane; warps read 32 consecutive words per step (coalesced) and scan with shuffles.
// jl_rank_count publishes the number of roots per segment; jl_rank_assign turns them into ranks.
constexpr int kRankSegs = 8;

__device__ __forceinline__ int block_exclusive_scan_1024(int v, int* s_warp, int* s_total) {
  const int tid = threadIdx.x;
  int incl = v;
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, incl, o);
    if ((tid & 31) >= o) incl += t;
  }
  if ((tid & 31) == 31) s_warp[tid >> 5] = incl;
  __syncthreads();
  if (tid < 32) {
    int w = s_warp[tid], in2 = w;
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, in2, o);
      if (tid >= o) in2 += t;
    }
    s_warp[tid] = in2 - w;
    if (tid == 31) *s_total = in2;
  }
  __syncthreads();
  return s_warp[tid >> 5] + incl - v;
}

// words [lo, hi) of this warp inside segment `seg`
__device__ __forceinline__ void rank_warp_range(int NW, int seg, int warp, int& lo, int& hi) {
  const int segw = ((NW + kRankSegs - 1) / kRankSegs + 31) & ~31;
  const int wlen = ((segw + 31) / 32 + 31) & ~31;          // words per warp, multiple of 32
  const int s0 = min(seg * segw, NW), s1 = min(s0 + segw, NW);
  lo = min(s0 + warp * wlen, s1);
  hi = min(lo + wlen, s1);
}

__global__ void __launch_bounds__(1024) jl_rank_count(const JointsParams p) {
  const int frame = blockIdx.y, seg = blockIdx.x;
  if (!frame_active(p, frame)) return;
  __shared__ int s_warp[32];
  __shared__ int s_total;
  const int NW = p.H * p.WW;
  const uint32_t* __restrict__ rb = p.rootbits + (size_t)frame * NW;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int lo, hi;
  rank_warp_range(NW, seg, warp, lo, hi);
  int cnt = 0;
#pragma unroll 4
  for (int i = lo + lane; i < hi; i += 32) cnt += __popc(rb[i]);
  block_exclusive_scan_1024(cnt, s_warp, &s_total);
  if (threadIdx.x == 0) p.segcount[frame * kRankSegs + seg] = s_total;
}

__global__ void __launch_bounds__(1024) jl_rank_assign(const JointsParams p) {
  const int frame = blockIdx.y, seg = blockIdx.x;
  if (!frame_active(p, frame)) return;
  __shared__ int s_warp[32];
  __shared__ int s_total;
  const int H = p.H, W = p.W, WW = p.WW;
  const int NW = H * WW;
  const uint32_t* __restrict__ rb = p.rootbits + (size_t)frame * NW;
  int32_t* L = p.lab + (size_t)frame * H * W;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int base = 0, total = 0;
  for (int s = 0; s < kRankSegs; ++s) {
    const int c = p.segcount[frame * kRankSegs + s];
    if (s < seg) base += c;
    total += c;
  }
  int lo, hi;
  rank_warp_range(NW, seg, warp, lo, hi);
  // per-warp totals -> exclusive offsets of the warps inside the segment
  int cnt = 0;
#pragma unroll 4
  for (int i = lo + lane; i < hi; i += 32) cnt += __popc(rb[i]);
  for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);   // warp total in every lane
  int wbase = block_exclusive_scan_1024(lane == 0 ? cnt : 0, s_warp, &s_total);
  wbase = __shfl_sync(0xffffffffu, wbase, 0);
  int rank0 = base + wbase;                       // rank of the first root of this warp
  int32_t* rootpix = p.rootpix + (size_t)frame * p.max_comp;
  unsigned long long* acc = p.acc + (size_t)frame * p.max_comp * 4;
  for (int i0 = lo; i0 < hi; i0 += 32) {
    const int i = i0 + lane;
    uint32_t bits = (i < hi) ? rb[i] : 0u;
    const int c = __popc(bits);
    int incl = c;
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    int rank = rank0 + incl - c;
    rank0 += __shfl_sync(0xffffffffu, incl, 31);
    if (bits) {
      const int y = i / WW, w = i - y * WW;
      const int pix0 = y * W + w * 32;
      while (bits) {
        const int s = __ffs(bits) - 1;
        bits &= bits - 1;
        L[pix0 + s] = ~rank;
        if (rank < p.max_comp) {
          rootpix[rank] = pix0 + s;
          acc[(size_t)rank * 4] = 0ull; acc[(size_t)rank * 4 + 1] = 0ull; acc[(size_t)rank * 4 + 2] = 0ull; acc[(size_t)rank * 4 + 3] = 0ull;
        }
        ++rank;
      }
    }
  }
  if (seg == 0 && tid == 0) {
    p.ncomp[frame] = min(total, p.max_comp);
    if (total > p.max_comp) atomicOr(&p.flags[frame], LGX_FLAG_COMP_OVERFLOW);
  }
}

// ---- per-run quad sums --------------------------------------------------------------------------
// Every 2x2 quad of pixel centres (including quads overlapping the 1-px zero pad) with k set pixels adds to
// its component: k==4: a00 += 2, a10 += 6x+3, a01 += 6y+3;  k==3: a00 += 1, a10 += sum of the three x,
// a01 += sum of the three y  (x,y = quad's top-left pixel).  Quads are owned by the word-run of their
// top-left pixel, else of their top-right pixel, else (Euler count only) of their single bottom pixel.
__device__ __forceinline__ void jl_sums_word(const JointsParams& p, const int frame, const int idx) {
  const int H = p.H, W = p.W, WW = p.WW;
  const uint32_t* __restrict__ jb = p.jbits + (size_t)frame * H * WW;
  const uint32_t cur = jb[idx];
  if (!cur) return;
  const int y = idx / WW, w = idx - y * WW;
  const int32_t* L = p.lab + (size_t)frame * H * W;
  unsigned long long* acc = p.acc + (size_t)frame * p.max_comp * 4;
  const uint64_t A = window34(jb + (size_t)y * WW, w, WW);
  const uint64_t Bn = (y + 1 < H) ? window34(jb + (size_t)(y + 1) * WW, w, WW) : 0ull;
  const uint64_t Up = (y > 0) ? window34(jb + (size_t)(y - 1) * WW, w, WW) : 0ull;
  const uint64_t tl = A, tr = A >> 1, bl = Bn, br = Bn >> 1;
  const uint64_t k4 = tl & tr & bl & br;
  const uint64_t k3 = (tl & tr & (bl ^ br)) | (bl & br & (tl ^ tr));
  const uint64_t k1 = ((tl ^ tr) & ~bl & ~br) | ((bl ^ br) & ~tl & ~tr);
  const uint64_t kd = (tl & br & ~tr & ~bl) | (tr & bl & ~tl & ~br);
  // quads of the row above whose top is empty and whose bottom (this row) has exactly one set pixel
  const uint64_t kb = (A ^ (A >> 1)) & ~Up & ~(Up >> 1);
  const int xbase = w * 32 - 1;   // x of window bit 0
  uint32_t m = cur;
  while (m) {
    const int s = __ffs(m) - 1;
    const int len = run_len32(m, s);
    const int e = s + len - 1;
    m &= ~((len == 32 ? 0xffffffffu : ((1u << len) - 1u)) << s);
    int r = L[y * W + w * 32 + s];
    if (r >= 0) r = L[r];
    const int rank = ~r;
    if (rank >= p.max_comp) continue;
    // owned quads: window indices s+1 .. e+1 (top-left in the run) and s if the pixel left of the run is unset
    uint64_t own = ((1ull << (e + 2)) - 1ull) & ~((1ull << (s + 1)) - 1ull);
    if (!((A >> s) & 1ull)) own |= 1ull << s;
    // bottom-anchored Euler quads: index s (bottom-right = run start) and e+1 (bottom-left = run end)
    const uint64_t ownb = (1ull << s) | (1ull << (e + 1));
    const uint64_t q4 = k4 & own, q3 = k3 & own;
    const int n4 = __popcll(q4), n3 = __popcll(q3);
    const int e4 = __popcll(k1 & own) + __popcll(kb & ownb) - n3 - 2 * __popcll(kd & own);
    const long long sx4 = (long long)n4 * xbase + sum_bit_index(q4);
    const long long sx3 = (long long)n3 * xbase + sum_bit_index(q3);
    const long long a00 = 2 * n4 + n3;
    const long long a10 = 6 * sx4 + 3 * n4 + 3 * sx3 + __popcll(q3 & tr) + __popcll(q3 & br);
    const long long a01 = (long long)n4 * (6 * y + 3) + 3ll * y * n3 + __popcll(q3 & bl) + __popcll(q3 & br);
    const unsigned long long packed = (unsigned long long)a00 + ((unsigned long long)(long long)e4 << 32);
    unsigned long long* a = acc + (size_t)rank * 4;
    if (packed) atomicAdd(&a[0], packed);
    if (a10) atomicAdd(&a[1], (unsigned long long)a10);
    if (a01) atomicAdd(&a[2], (unsigned long long)a01);
  }
}

__global__ void __launch_bounds__(kWordThreads) jl_sums(const JointsParams p) {
  const int frame = blockIdx.y;
  if (!frame_active(p, frame)) return;
  if (p.pass == 0) {
    const int idx = word_index(p, frame);
    if (idx >= 0) jl_sums_word(p, frame, idx);
  } else {
    // second pass (frames that needed the whole-frame hole fill): small grid, grid-stride over all words
    for (int idx = blockIdx.x * kWordThreads + threadIdx.x; idx < p.H * p.WW; idx += gridDim.x * kWordThreads)
      jl_sums_word(p, frame, idx);
  }
}

// ---- components with holes (Euler number != 1): per-frame list --------------------------------------
__global__ void __launch_bounds__(256) jl_hole_list(const JointsParams p) {
  const int frame = blockIdx.y;
  const int n = p.ncomp[frame];
  for (int k = blockIdx.x * 256 + threadIdx.x; k < n; k += gridDim.x * 256) {
    const unsigned long long v = p.acc[((size_t)frame * p.max_comp + k) * 4];
    if ((int)(v >> 32) != 4) {
      const int slot = atomicAdd(&p.nholes[frame], 1);
      if (slot < kMaxHoles) p.holes[(size_t)frame * kMaxHoles + slot] = k;
      else atomicOr(&p.flags[frame], LGX_FLAG_GENERIC_FILL);
      atomicOr(&p.flags[frame], LGX_FLAG_HOLES);
    }
  }
}

// ---- local hole fix: one warp per listed component -------------------------------------------------
// Works on a 128 x 64 pixel window around the component's first pixel (4 words x 64 rows, two rows per lane):
//   X = the component (8-connected flood from its first pixel through the joints mask),
//   O = what the window's outer ring reaches through ~X with 4-connectivity,
//   filled component = ~O.  Its quad sums replace the component's accumulators (what findContours' outer
//   contour encloses); every other component inside ~O is nested in a hole, which RETR_EXTERNAL does not
//   report: it is queued and zeroed by jl_hole_kill.  A component that reaches the window's ring is left to
//   the whole-frame flood (LGX_FLAG_GENERIC_FILL).
struct Row128 {
  uint32_t w[4];
};
__device__ __forceinline__ Row128 r_or(Row128 a, Row128 b) { Row128 r; for (int i = 0; i < 4; ++i) r.w[i] = a.w[i] | b.w[i]; return r; }
__device__ __forceinline__ Row128 r_and(Row128 a, Row128 b) { Row128 r; for (int i = 0; i < 4; ++i) r.w[i] = a.w[i] & b.w[i]; return r; }
__device__ __forceinline__ Row128 r_not(Row128 a) { Row128 r; for (int i = 0; i < 4; ++i) r.w[i] = ~a.w[i]; return r; }
__device__ __forceinline__ bool r_ne(Row128 a, Row128 b) { return (a.w[0] ^ b.w[0]) | (a.w[1] ^ b.w[1]) | (a.w[2] ^ b.w[2]) | (a.w[3] ^ b.w[3]); }
__device__ __forceinline__ bool r_any(Row128 a) { return a.w[0] | a.w[1] | a.w[2] | a.w[3]; }
__device__ __forceinline__ Row128 r_up1(Row128 a) {   // bit i <- bit i-1 (towards higher x)
  Row128 r;
  r.w[0] = a.w[0] << 1;
  for (int i = 1; i < 4; ++i) r.w[i] = (a.w[i] << 1) | (a.w[i - 1] >> 31);
  return r;
}
__device__ __forceinline__ Row128 r_dn1(Row128 a) {   // bit i <- bit i+1
  Row128 r;
  for (int i = 0; i < 3; ++i) r.w[i] = (a.w[i] >> 1) | (a.w[i + 1] << 31);
  r.w[3] = a.w[3] >> 1;
  return r;
}
__device__ __forceinline__ Row128 r_dil(Row128 a) { return r_or(a, r_or(r_up1(a), r_dn1(a))); }
// all pixels of the runs of `bg` that contain a seed (seed must be a subset of bg)
__device__ __forceinline__ Row128 r_hfill(Row128 bg, Row128 seed) {
  Row128 f;
  uint32_t cin = 0;
  for (int i = 0; i < 4; ++i) {
    const uint32_t b = bg.w[i], s = seed.w[i] | (cin & b);
    f.w[i] = (((b + s) ^ b) & b) | s;
    cin = f.w[i] >> 31;
  }
  cin = 0;
  for (int i = 3; i >= 0; --i) {
    const uint32_t b = __brev(bg.w[i]), s = __brev(f.w[i]) | (cin & b);
    const uint32_t g = (((b + s) ^ b) & b) | s;
    cin = g >> 31;
    f.w[i] = __brev(g);
  }
  return f;
}
__device__ __forceinline__ Row128 r_shfl(Row128 a, int src_lane) {
  Row128 r;
  for (int i = 0; i < 4; ++i) r.w[i] = __shfl_sync(0xffffffffu, a.w[i], src_lane);
  return r;
}

__global__ void __launch_bounds__(32) jl_hole_fix(const JointsParams p) {
  const int frame = blockIdx.y;
  const int nh = min(p.nholes[frame], kMaxHoles);
  if ((int)blockIdx.x >= nh || (p.flags[frame] & LGX_FLAG_GENERIC_FILL)) return;
  const int H = p.H, W = p.W, WW = p.WW;
  const int lane = threadIdx.x;
  const int rank = p.holes[(size_t)frame * kMaxHoles + blockIdx.x];
  const int pix = p.rootpix[(size_t)frame * p.max_comp + rank];
  const int yr = pix / W, xr = pix - yr * W;
  const int wr0 = (xr >> 5) - 1;          // first word of the window
  const int ry0 = yr - 1;                 // first row of the window
  const uint32_t* __restrict__ jb = p.jbits + (size_t)frame * H * WW;
  const Row128 zero = {{0u, 0u, 0u, 0u}};
  const Row128 ones = {{~0u, ~0u, ~0u, ~0u}};
  auto load_row = [&](int r) {
    Row128 v = zero;
    const int y = ry0 + r;
    if (y >= 0 && y < H)
      for (int i = 0; i < 4; ++i) {
        const int w = wr0 + i;
        if (w >= 0 && w < WW) v.w[i] = jb[(size_t)y * WW + w];
      }
    return v;
  };
  // rows 2*lane and 2*lane+1
  const Row128 J0 = load_row(2 * lane), J1 = load_row(2 * lane + 1);
  Row128 ring0 = zero, ring1 = zero;      // the window's outer ring
  ring0.w[0] = 1u; ring0.w[3] = 0x80000000u; ring1 = ring0;
  if (lane == 0) ring0 = ones;
  if (lane == 31) ring1 = ones;
  // X: 8-connected flood from the first pixel (window row 1, bit (xr & 31) of word 1)
  Row128 X0 = zero, X1 = zero;
  if (lane == 0) X1.w[1] = 1u << (xr & 31);
  for (;;) {
    const Row128 up0 = r_shfl(X1, (lane + 31) & 31), dn1 = r_shfl(X0, (lane + 1) & 31);
    Row128 n0 = r_or(r_dil(X1), X0), n1 = r_or(r_dil(X0), X1);
    if (lane > 0) n0 = r_or(n0, r_dil(up0));
    if (lane < 31) n1 = r_or(n1, r_dil(dn1));
    const Row128 y0 = r_hfill(J0, r_and(n0, J0)), y1 = r_hfill(J1, r_and(n1, J1));
    const bool ch = r_ne(y0, X0) || r_ne(y1, X1);
    X0 = y0; X1 = y1;
    if (!__any_sync(0xffffffffu, ch)) break;
  }
  if (__any_sync(0xffffffffu, r_any(r_and(X0, ring0)) || r_any(r_and(X1, ring1)))) {
    if (lane == 0) atomicOr(&p.flags[frame], LGX_FLAG_GENERIC_FILL);
    return;
  }
  // O: 4-connected flood of ~X from the ring
  const Row128 B0 = r_not(X0), B1 = r_not(X1);
  Row128 O0 = ring0, O1 = ring1;
  for (;;) {
    const Row128 up0 = r_shfl(O1, (lane + 31) & 31), dn1 = r_shfl(O0, (lane + 1) & 31);
    Row128 n0 = r_or(O1, O0), n1 = r_or(O0, O1);
    if (lane > 0) n0 = r_or(n0, up0);
    if (lane < 31) n1 = r_or(n1, dn1);
    const Row128 y0 = r_hfill(B0, r_and(n0, B0)), y1 = r_hfill(B1, r_and(n1, B1));
    const bool ch = r_ne(y0, O0) || r_ne(y1, O1);
    O0 = y0; O1 = y1;
    if (!__any_sync(0xffffffffu, ch)) break;
  }
  const Row128 M0 = r_not(O0), M1 = r_not(O1);   // filled component (never touches the ring)
  // quad sums over M: pairs (row 2l, 2l+1) and (row 2l+1, 2l+2)
  const Row128 Mn = r_shfl(M0, (lane + 1) & 31);  // row 2l+2
  long long a00 = 0, a10 = 0, a01 = 0;
  auto quad_pair = [&](Row128 A, Row128 Bv, int ytop) {
    const Row128 tr = r_dn1(A), br = r_dn1(Bv);
    for (int i = 0; i < 4; ++i) {
      const uint32_t tl = A.w[i], t2 = tr.w[i], bl = Bv.w[i], b2 = br.w[i];
      uint32_t k4 = tl & t2 & bl & b2;
      uint32_t k3 = (tl & t2 & (bl ^ b2)) | (bl & b2 & (tl ^ t2));
      const int xb = (wr0 + i) * 32;
      const int n4 = __popc(k4), n3 = __popc(k3);
      a00 += 2 * n4 + n3;
      a01 += (long long)n4 * (6 * ytop + 3) + 3ll * ytop * n3 + __popc(k3 & bl) + __popc(k3 & b2);
      a10 += 3ll * n4 + __popc(k3 & t2) + __popc(k3 & b2);
      while (k4) { a10 += 6ll * (xb + __ffs(k4) - 1); k4 &= k4 - 1; }
      while (k3) { a10 += 3ll * (xb + __ffs(k3) - 1); k3 &= k3 - 1; }
    }
  };
  quad_pair(M0, M1, ry0 + 2 * lane);
  if (lane < 31) quad_pair(M1, Mn, ry0 + 2 * lane + 1);
  for (int o = 16; o; o >>= 1) {
    a00 += __shfl_xor_sync(0xffffffffu, a00, o);
    a10 += __shfl_xor_sync(0xffffffffu, a10, o);
    a01 += __shfl_xor_sync(0xffffffffu, a01, o);
  }
  unsigned long long* acc = p.acc + ((size_t)frame * p.max_comp + rank) * 4;
  if (lane == 0) {
    acc[0] = (unsigned long long)a00 | (4ull << 32);
    acc[1] = (unsigned long long)a10;
    acc[2] = (unsigned long long)a01;
  }
  // components nested in the holes: every word-run start of the joints mask inside M \ X that is the first pixel of a
  // component (binary search in rootpix, which is ascending)
  const int32_t* rpx = p.rootpix + (size_t)frame * p.max_comp;
  const int ncomp = p.ncomp[frame];
  auto nested_row = [&](Row128 Jr, Row128 Mr, Row128 Xr, int r) {
    const int y = ry0 + r;
    for (int i = 0; i < 4; ++i) {
      const uint32_t cur = Jr.w[i];
      uint32_t starts = cur & ~(cur << 1) & Mr.w[i] & ~Xr.w[i];
      while (starts) {
        const int s = __ffs(starts) - 1;
        starts &= starts - 1;
        const int pixn = y * W + (wr0 + i) * 32 + s;
        int lo = 0, hi = ncomp - 1;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (rpx[mid] < pixn) lo = mid + 1;
          else hi = mid;
        }
        if (ncomp == 0 || rpx[lo] != pixn) continue;
        const int slot = atomicAdd(&p.nnested[frame], 1);
        if (slot < kMaxNested) p.nested[(size_t)frame * kMaxNested + slot] = lo;
        else atomicOr(&p.flags[frame], LGX_FLAG_GENERIC_FILL);
      }
    }
  };
  nested_row(J0, M0, X0, 2 * lane);
  nested_row(J1, M1, X1, 2 * lane + 1);
}

// nested components are not reported: a00 = 0 makes emit drop them
__global__ void __launch_bounds__(256) jl_hole_kill(const JointsParams p) {
  const int frame = blockIdx.x;
  if (p.flags[frame] & LGX_FLAG_GENERIC_FILL) return;
  const int n = min(p.nnested[frame], kMaxNested);
  for (int i = threadIdx.x; i < n; i += 256) {
    const int rank = p.nested[(size_t)frame * kMaxNested + i];
    if (rank >= 0 && rank < p.max_comp) p.acc[((size_t)frame * p.max_comp + rank) * 4] = 4ull << 32;
  }
}

// ---- whole-frame hole fill: flood the background from the image border (4-connectivity) -----------
// One CTA per flagged frame; each warp owns bands of 32 rows and sweeps them down and up (Gauss-Seidel
// inside a band, Jacobi across bands) until no bit changes.  A row is filled horizontally with the
// carry trick: ((bg + seed) ^ bg) & bg | seed smears seeds upward through runs of ones, and word-level
// carries are resolved with the same trick on ballot masks.
constexpr int kMaxSegs = 8;   // up to 8*32 words = 8192 px wide

struct RowFill {
  int nseg, lane, WW;
  // fills `o` (seeds, subset of bg) along the row in both directions
  __device__ void run(const uint32_t* bg, uint32_t* o) const {
    // towards higher x
    uint32_t cin = 0;
    for (int sg = 0; sg < nseg; ++sg) {
      const uint32_t b = bg[sg], s = o[sg];
      uint32_t f = (((b + s) ^ b) & b) | s;
      const uint32_t G = __ballot_sync(0xffffffffu, f >> 31);
      const uint32_t P = __ballot_sync(0xffffffffu, b == 0xffffffffu);
      const uint32_t Aa = P | G;
      const uint64_t sum = (uint64_t)Aa + (uint64_t)G + cin;
      const uint32_t C = ((uint32_t)sum) ^ Aa ^ G;      // carry into each lane
      cin = (uint32_t)(sum >> 32);
      if ((C >> lane) & 1u) {
        const uint32_t s2 = s | (b & 1u);
        f = (((b + s2) ^ b) & b) | s2;
      }
      o[sg] = f;
    }
    // towards lower x (bit-reversed words, reversed lane order)
    cin = 0;
    for (int sg = nseg - 1; sg >= 0; --sg) {
      const uint32_t b = __brev(bg[sg]), s = __brev(o[sg]);
      uint32_t f = (((b + s) ^ b) & b) | s;
      const uint32_t G = __brev(__ballot_sync(0xffffffffu, f >> 31));
      const uint32_t P = __brev(__ballot_sync(0xffffffffu, b == 0xffffffffu));
      const uint32_t Aa = P | G;
      const uint64_t sum = (uint64_t)Aa + (uint64_t)G + cin;
      const uint32_t C = __brev(((uint32_t)sum) ^ Aa ^ G);
      cin = (uint32_t)(sum >> 32);
      if ((C >> lane) & 1u) {
        const uint32_t s2 = s | (b & 1u);
        f = (((b + s2) ^ b) & b) | s2;
      }
      o[sg] = __brev(f);
    }
  }
};

__global__ void __launch_bounds__(1024) fill_holes_kernel(const uint32_t* __restrict__ jbits_all, uint32_t* __restrict__ filled_all,
                                                          uint32_t* __restrict__ scratch_all, const uint32_t* __restrict__ flags,
                                                          int H, int W, int WW) {
  const int frame = blockIdx.x;
  if (!(flags[frame] & LGX_FLAG_GENERIC_FILL)) return;
  const uint32_t* jb = jbits_all + (size_t)frame * H * WW;
  uint32_t* O = scratch_all + (size_t)frame * H * WW;
  uint32_t* F = filled_all + (size_t)frame * H * WW;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nseg = (WW + 31) >> 5;
  RowFill rf{nseg, lane, WW};
  const int nbands = (H + 31) >> 5;
  auto valid = [&](int w) -> uint32_t {
    if (w >= WW) return 0u;
    int rem = W - w * 32;
    return rem >= 32 ? 0xffffffffu : ((1u << rem) - 1u);
  };
  auto load_bg = [&](int y, uint32_t* bg) {
    for (int sg = 0; sg < nseg; ++sg) {
      int w = sg * 32 + lane;
      bg[sg] = (w < WW) ? (~jb[(size_t)y * WW + w] & valid(w)) : 0u;
    }
  };
  auto edge = [&](int y, const uint32_t* bg, uint32_t* sd) {
    for (int sg = 0; sg < nseg; ++sg) {
      int w = sg * 32 + lane;
      uint32_t e = 0;
      if (y == 0 || y == H - 1) e = 0xffffffffu;
      if (w == 0) e |= 1u;
      if (w == WW - 1) e |= 1u << ((W - 1) & 31);
      sd[sg] = e & bg[sg];
    }
  };
  uint32_t bg[kMaxSegs], o[kMaxSegs], prev[kMaxSegs];
  // initial state: every row filled from its border seeds
  for (int y = warp; y < H; y += 32) {
    load_bg(y, bg);
    edge(y, bg, o);
    rf.run(bg, o);
    for (int sg = 0; sg < nseg; ++sg) {
      int w = sg * 32 + lane;
      if (w < WW) __stcg(&O[(size_t)y * WW + w], o[sg]);
    }
  }
  __syncthreads();
  for (;;) {
    bool changed = false;
    for (int band = warp; band < nbands; band += 32) {
      const int ya = band * 32, yb = min(ya + 32, H);
      for (int dir = 0; dir < 2; ++dir) {
        const int ystart = dir ? yb - 1 : ya, yend = dir ? ya - 1 : yb, step = dir ? -1 : 1;
        const int yn0 = ystart - step;   // neighbour row outside the band in sweep direction
        for (int sg = 0; sg < nseg; ++sg) {
          int w = sg * 32 + lane;
          prev[sg] = (yn0 >= 0 && yn0 < H && w < WW) ? __ldcg(&O[(size_t)yn0 * WW + w]) : 0u;
        }
        for (int y = ystart; y != yend; y += step) {
          load_bg(y, bg);
          bool any = false;
          for (int sg = 0; sg < nseg; ++sg) {
            int w = sg * 32 + lane;
            uint32_t cur = (w < WW) ? __ldcg(&O[(size_t)y * WW + w]) : 0u;
            uint32_t nv = cur | (prev[sg] & bg[sg]);
            any |= (nv != cur);
            o[sg] = nv;
          }
          if (__any_sync(0xffffffffu, any)) {
            rf.run(bg, o);
            for (int sg = 0; sg < nseg; ++sg) {
              int w = sg * 32 + lane;
              if (w < WW) __stcg(&O[(size_t)y * WW + w], o[sg]);
            }
            changed = true;
          }
          for (int sg = 0; sg < nseg; ++sg) prev[sg] = o[sg];
        }
      }
    }
    if (!__syncthreads_or(changed ? 1 : 0)) break;
  }
  // filled = joints | (background not reached from the border)
  for (int i = threadIdx.x; i < H * WW; i += 1024) {
    int w = i % WW;
    F[i] = jb[i] | (~__ldcg(&O[i]) & valid(w));
  }
}

// ---- emit: centroids in the reference's list order (descending first-pixel raster index) --------------
// kEmitSegs CTAs per frame: CTA `seg` owns positions j = seg*L .. (seg+1)*L-1 of the descending list (component
// k = n-1-j).  It first counts the components with a non-zero area in front of its range (one word per component),
// which is its write offset, then computes and compacts its own.  One CTA per frame needs n/1024 dependent rounds
// (67 us per launch whatever the batch); eight need an eighth of that.
constexpr int kEmitSegs = 8;

__global__ void __launch_bounds__(1024) emit_kernel(const EmitParams p) {
  const int seg = blockIdx.x, frame = blockIdx.y;
  __shared__ int s_w[32];
  __shared__ int s_run;
  const int n = p.ncomp[frame];
  const unsigned long long* __restrict__ acc = p.acc + (size_t)frame * p.max_comp * 4;
  int32_t* __restrict__ out = p.centroids + (size_t)frame * p.max_cent * 2;
  double* __restrict__ outf = p.centroids_f ? p.centroids_f + (size_t)frame * p.max_cent * 2 : nullptr;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int L = ((n + kEmitSegs * 1024 - 1) / (kEmitSegs * 1024)) * 1024;     // positions per segment, multiple of 1024
  const int j0 = min(seg * L, n), j1 = min(j0 + L, n);
  // valid components in front of this segment
  int before = 0;
  for (int j = tid; j < j0; j += 1024) before += (acc[(size_t)(n - 1 - j) * 4] & 0xffffffffull) != 0ull;
  for (int o = 16; o; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
  if (lane == 0) s_w[warp] = before;
  __syncthreads();
  if (tid == 0) {
    int t = 0;
    for (int i = 0; i < 32; ++i) t += s_w[i];
    s_run = t;
  }
  __syncthreads();
  for (int base = j0; base < j1; base += 1024) {
    const int j = base + tid;
    const int k = n - 1 - j;
    unsigned long long a00 = 0, a10 = 0, a01 = 0;
    if (j < j1) {
      a00 = acc[(size_t)k * 4] & 0xffffffffull;
      a10 = acc[(size_t)k * 4 + 1];
      a01 = acc[(size_t)k * 4 + 2];
    }
    const bool valid = (j < j1) && (a00 != 0);
    const unsigned bal = __ballot_sync(0xffffffffu, valid);
    __syncthreads();                       // s_w of the previous round (or of the prefix count) has been read
    if (lane == 0) s_w[warp] = __popc(bal);
    __syncthreads();
    int off = s_run;
    for (int i = 0; i < warp; ++i) off += s_w[i];
    const int pos = off + __popc(bal & ((1u << lane) - 1u));
    if (valid && pos < p.max_cent) {
      // cv2.moments: m00 = a00*0.5, m10 = a10*(1/6), m01 = a01*(1/6); centroid = int(m10/m00), int(m01/m00)
      const double m00 = __dmul_rn((double)a00, 0.5);
      const double fx = __ddiv_rn(__dmul_rn((double)a10, 1.0 / 6), m00);
      const double fy = __ddiv_rn(__dmul_rn((double)a01, 1.0 / 6), m00);
      out[2 * pos] = __double2int_rz(fx);
      out[2 * pos + 1] = __double2int_rz(fy);
      if (outf) { outf[2 * pos] = fx; outf[2 * pos + 1] = fy; }
    }
    __syncthreads();
    if (tid == 0) {
      int t = 0;
      for (int i = 0; i < 32; ++i) t += s_w[i];
      s_run += t;
    }
    __syncthreads();
  }
  if (seg == kEmitSegs - 1 && tid == 0) {  // the last segment ends at n: its running count is the frame's total
    p.counts[frame] = s_run;
    if (s_run > p.max_cent) atomicOr(&p.flags[frame], LGX_FLAG_CENT_OVERFLOW);
  }
}

}  // namespace

cudaError_t launch_joints_label(const JointsParams& p, int batch, bool seeded, cudaStream_t stream) {
  const int NW = p.H * p.WW;
  // first pass: one thread per (non-empty) word; second pass: 64 CTAs per frame that leave at once unless flagged
  dim3 gw(p.pass == 0 ? (NW + kWordThreads - 1) / kWordThreads : 64, batch);
  if (!seeded) jl_seed<<<gw, kWordThreads, 0, stream>>>(p);
  jl_union<<<gw, kWordThreads, 0, stream>>>(p);
  jl_roots<<<gw, kWordThreads, 0, stream>>>(p);
  jl_rank_count<<<dim3(kRankSegs, batch), 1024, 0, stream>>>(p);
  jl_rank_assign<<<dim3(kRankSegs, batch), 1024, 0, stream>>>(p);
  jl_sums<<<gw, kWordThreads, 0, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_joints_holes(const JointsParams& p, int batch, bool listed, cudaStream_t stream) {
  if (!listed) jl_hole_list<<<dim3(32, batch), 256, 0, stream>>>(p);
  jl_hole_fix<<<dim3(kMaxHoles, batch), 32, 0, stream>>>(p);
  jl_hole_kill<<<batch, 256, 0, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_fill_holes(const uint32_t* jbits, uint32_t* filled, uint32_t* scratch, const uint32_t* flags,
                              int batch, int H, int W, cudaStream_t stream) {
  if (bits_pitch(W) > kMaxSegs * 32) return cudaErrorInvalidValue;
  fill_holes_kernel<<<batch, 1024, 0, stream>>>(jbits, filled, scratch, flags, H, W, bits_pitch(W));
  return cudaGetLastError();
}

cudaError_t launch_emit(const EmitParams& p, int batch, cudaStream_t stream) {
  emit_kernel<<<dim3(kEmitSegs, batch), 1024, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace lgx
