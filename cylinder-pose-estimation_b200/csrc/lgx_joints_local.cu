// K4 first pass, strip-local: connected components of the joints bit plane labelled in SHARED memory.
//
// Same contract as lgx_joints.cu (the contour-equivalent of cv2.findContours(RETR_EXTERNAL) + cv2.moments,
// /root/reference/utils/util_cylinder.py:1817-1825; oracle/restate.py contour_sums): after this pass
//   rootpix[frame][k]  first raster pixel of component k, ascending in k        acc[frame][k] = {a00 | e4 << 32, a10, a01, 0}
//   ncomp[frame]       number of components                                     holes[frame]  components whose Euler number is not 1
// but without the sparse [H*W] parent map of the whole-frame union-find: a laser-grid joint is a blob of ~5x5 pixels, so
// almost every component lives inside one strip of rows and never needs global memory.
//
//   jl_local    one CTA per strip of R full-width rows (R from the shared-memory budget: 32 rows at 2448 px, 21 at 4096 px):
//               word-runs (maximal runs of set bits inside a 32-bit word) are the union-find elements, identified by
//               (word index in the strip) * 16 + (start bit / 2) - two runs of a word cannot start in the same bit pair -
//               so the parent array is dense per word and a component's root is its first raster pixel in the strip.
//               union (left word, three words above) -> flatten -> rank the roots in raster order (block scan) -> per-run
//               quad sums by shared-memory atomics -> one 32-byte record per component {a00 | e4 << 32, a10, a01, first pixel}
//               in a block reserved with one atomicAdd per strip.  Components with a run that touches a set pixel of the
//               neighbouring strip are flagged; only their runs on the boundary rows and their roots enter the global map.
//   jl_border   the boundary rows: 8-connectivity links between the last row of a strip and the first row of the next one,
//               as a union-find over the flagged components' first pixels (global map, a few thousand entries per frame).
//   jl_merge    every flagged record whose global root is another component adds its sums to that component's record
//               (found by binary search among the root strip's records, which are sorted by first pixel) and dies.
//   jl_compact  per strip: prefix of the live records of the strips above = rank of its first record; copies the live
//               records to acc / rootpix in raster order and lists the components with holes.
#include <cstdlib>

#include "lgx_joints.cuh"

namespace lgx {
namespace {

#ifndef LGX_JL_THREADS
#define LGX_JL_THREADS 512
#endif
constexpr int kLocalThreads = LGX_JL_THREADS;
constexpr int kLocalCap = 512;             // components per strip whose sums are accumulated in shared memory (the rest: global atomics)
constexpr uint32_t kRootBit = 0x80000000u;
constexpr unsigned long long kRecBorder = 1ull << 32, kRecDead = 1ull << 33;
constexpr uint32_t kSkipFlags = LGX_FLAG_GENERIC_FILL | LGX_FLAG_COMP_OVERFLOW;   // frames the border / merge / compact kernels leave alone

struct LocalSmem {
  uint32_t* w;        // [(R + 2)][WW]   strip rows with one halo row above and below
  uint32_t* root;     // [R][WW]         start bits of the runs that are roots
  uint16_t* rbase;    // [R][WW]         id of the word's first run (runs are numbered in raster order)
  uint32_t* par;      // [capr]          parents (run ids); roots end as kRootBit | rank
  uint32_t* runs;     // [capr]          word index << 5 | start bit, raster order
  uint32_t* acc;      // [kLocalCap][6]  a00, e4, a10 lo / hi, a01 lo / hi
  uint32_t* rpix;     // [kLocalCap]     first pixel of the component | kRootBit when it touches a neighbouring strip
  int* scan;          // [34]
};

__device__ __forceinline__ uint32_t sfind(uint32_t* P, uint32_t a) {
  uint32_t q = ((volatile uint32_t*)P)[a];
  while (q != a) {
    a = q;
    q = ((volatile uint32_t*)P)[a];
  }
  return a;
}

__device__ __forceinline__ void sunion(uint32_t* P, uint32_t a, uint32_t b) {
  bool done;
  do {
    a = sfind(P, a);
    b = sfind(P, b);
    if (a < b) {
      const uint32_t old = atomicMin(&P[b], a);
      done = (old == b);
      b = old;
    } else if (b < a) {
      const uint32_t old = atomicMin(&P[a], b);
      done = (old == a);
      a = old;
    } else {
      done = true;
    }
  } while (!done);
}

__device__ __forceinline__ int block_scan_excl(int v, int* s_warp, int* s_total) {
  const int tid = threadIdx.x;
  int incl = v;
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if ((tid & 31) >= o) incl += t;
  }
  if ((tid & 31) == 31) s_warp[tid >> 5] = incl;
  __syncthreads();
  if (tid < 32) {
    const int wv = (tid < (int)(blockDim.x >> 5)) ? s_warp[tid] : 0;
    int in2 = wv;
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, in2, o);
      if (tid >= o) in2 += t;
    }
    s_warp[tid] = in2 - wv;
    if (tid == 31) *s_total = in2;
  }
  __syncthreads();
  return s_warp[tid >> 5] + incl - v;
}

// id of the run of `word` (word index idx) that starts at bit st
__device__ __forceinline__ uint32_t run_id(const uint16_t* rbase, int idx, uint32_t word, int st) {
  return (uint32_t)rbase[idx] + (uint32_t)__popc(word & ~(word << 1) & ((1u << st) - 1u));
}

// sum of the indices of the set bits of m, without a loop over the bits (every lane the same instruction count)
__device__ __forceinline__ int sum_bit_index_flat(uint64_t m) {
  return __popcll(m & 0xaaaaaaaaaaaaaaaaull) + 2 * __popcll(m & 0xccccccccccccccccull) + 4 * __popcll(m & 0xf0f0f0f0f0f0f0f0ull) +
         8 * __popcll(m & 0xff00ff00ff00ff00ull) + 16 * __popcll(m & 0xffff0000ffff0000ull) + 32 * __popcll(m & 0xffffffff00000000ull);
}

__global__ void __launch_bounds__(kLocalThreads) jl_local(const JointsLocalParams p) {
  extern __shared__ unsigned char smem_raw[];
  const int H = p.H, W = p.W, WW = p.WW, R = p.R;
  const int frame = blockIdx.y, strip = blockIdx.x;
  const int y0 = strip * R;
  const int rows = min(R, H - y0);
  const int NW = rows * WW;
  LocalSmem s;
  s.w = reinterpret_cast<uint32_t*>(smem_raw);
  s.root = s.w + (size_t)(R + 2) * WW;
  s.runs = s.root + (size_t)R * WW;
  s.par = s.runs + p.capr;
  s.acc = s.par + p.capr;
  s.rpix = s.acc + 6 * kLocalCap;
  s.scan = reinterpret_cast<int*>(s.rpix + kLocalCap);
  s.rbase = reinterpret_cast<uint16_t*>(s.scan + 34);
  const int tid = threadIdx.x;
  const uint32_t* __restrict__ jb = p.jbits + (size_t)frame * H * WW;

  // ---- rows y0 - 1 .. y0 + rows (outside the image: empty)
  for (int i = tid; i < (rows + 2) * WW; i += kLocalThreads) {
    const int r = i / WW, w = i - r * WW;
    const int y = y0 - 1 + r;
    s.w[i] = (y >= 0 && y < H) ? jb[(size_t)y * WW + w] : 0u;
  }
  __syncthreads();
  const uint32_t* sw = s.w + WW;            // row 0 of the strip
  // thread t owns the words [t * wpt, (t + 1) * wpt) wherever raster order matters
  const int wpt = (NW + kLocalThreads - 1) / kLocalThreads;
  const int i0 = min(tid * wpt, NW), i1 = min(i0 + wpt, NW);
  // ---- the runs in raster order (word index << 5 | start bit), every run its own parent.  From here on a thread works on
  // a run, not on a word: lanes stay converged whatever the words look like.
  int nr = 0;
  for (int i = i0; i < i1; ++i) {
    const uint32_t cur = sw[i];
    nr += __popc(cur & ~(cur << 1));
  }
  int rpos = block_scan_excl(nr, s.scan, s.scan + 32);
  const int nruns = s.scan[32];
  if (nruns > p.capr || nruns == 0) {
    // more runs than the arrays hold (dense noise): the whole-frame flood + relabel pass redoes the frame
    if (tid == 0) {
      if (nruns) atomicOr(&p.flags[frame], LGX_FLAG_GENERIC_FILL);
      p.sbase[(size_t)frame * p.strips + strip] = 0;
      p.scount[(size_t)frame * p.strips + strip] = 0;
    }
    return;
  }
  for (int i = i0; i < i1; ++i) {
    const uint32_t cur = sw[i];
    s.rbase[i] = (uint16_t)rpos;
    s.root[i] = 0u;
    uint32_t starts = cur & ~(cur << 1);
    while (starts) {
      const int st = __ffs(starts) - 1;
      starts &= starts - 1;
      s.runs[rpos] = ((uint32_t)i << 5) | (uint32_t)st;
      s.par[rpos] = (uint32_t)rpos;
      ++rpos;
    }
  }
  __syncthreads();
  // ---- union: the run left of the word boundary, the runs touched in the row above (columns s-1 .. e+1), inside the strip
  for (int r = tid; r < nruns; r += kLocalThreads) {
    const uint32_t rd = s.runs[r];
    const int idx = (int)(rd >> 5), st = (int)(rd & 31u);
    const uint32_t cur = sw[idx];
    const int e = st + run_len32(cur, st) - 1;
    // (the word left of the first word of a row is the last word of the row above: its bit 31 is beyond the image or a pixel of
    // another row, so test the column)
    if (st == 0 && idx % WW != 0) {
      const uint32_t left = sw[idx - 1];
      if (left >> 31) sunion(s.par, (uint32_t)r, run_id(s.rbase, idx - 1, left, run_start32(left, 31)));
    }
    if (idx < WW) continue;                 // first row of the strip
    const int w = idx % WW;
    const uint32_t up_c = sw[idx - WW], up_p = w > 0 ? sw[idx - WW - 1] : 0u, up_n = w + 1 < WW ? sw[idx - WW + 1] : 0u;
    const uint64_t U = (uint64_t)(up_p >> 31) | ((uint64_t)up_c << 1) | ((uint64_t)(up_n & 1u) << 33);
    uint64_t mm = U & ((1ull << (e + 3)) - 1ull) & ~((1ull << st) - 1ull);
    while (mm) {
      const int i = __ffsll((long long)mm) - 1;
      uint32_t word;
      int bb, wi;
      if (i == 0) { word = up_p; bb = 31; wi = idx - WW - 1; }
      else if (i <= 32) { word = up_c; bb = i - 1; wi = idx - WW; }
      else { word = up_n; bb = 0; wi = idx - WW + 1; }
      sunion(s.par, (uint32_t)r, run_id(s.rbase, wi, word, run_start32(word, bb)));
      const uint64_t t2 = ~(U >> i);
      const int len2 = __ffsll((long long)t2) - 1;
      mm &= ~(((1ull << len2) - 1ull) << i);
    }
  }
  __syncthreads();
  // ---- flatten, root bits
  for (int r = tid; r < nruns; r += kLocalThreads) {
    const uint32_t root = sfind(s.par, (uint32_t)r);
    if (root == (uint32_t)r) {
      const uint32_t rd = s.runs[r];
      atomicOr(&s.root[rd >> 5], 1u << (rd & 31u));
    } else {
      s.par[r] = root;
    }
  }
  __syncthreads();
  // ---- rank of the roots in raster order
  int cnt = 0;
  for (int i = i0; i < i1; ++i) cnt += __popc(s.root[i]);
  int rank = block_scan_excl(cnt, s.scan, s.scan + 32);
  const int count = s.scan[32];
  if (tid == 0) {
    // a block of records for this strip (any order between strips: jl_compact walks the strips by their base)
    int base = atomicAdd(&p.nrec[frame], count);
    if (base + count > p.max_comp) {
      atomicOr(&p.flags[frame], LGX_FLAG_COMP_OVERFLOW);
      base = -1;
    }
    s.scan[33] = base;
    p.sbase[(size_t)frame * p.strips + strip] = base < 0 ? 0 : base;
    p.scount[(size_t)frame * p.strips + strip] = base < 0 ? 0 : count;
  }
  for (int i = tid; i < 6 * min(count, kLocalCap); i += kLocalThreads) s.acc[i] = 0u;
  __syncthreads();
  const int base = s.scan[33];
  if (base < 0) return;
  unsigned long long* rec = p.rec + ((size_t)frame * p.max_comp + base) * 4;
  if (cnt) {
    int ly = i0 / WW, w = i0 - ly * WW;
    for (int i = i0; i < i1; ++i) {
      uint32_t roots = s.root[i];
      while (roots) {
        const int st = __ffs(roots) - 1;
        roots &= roots - 1;
        const uint32_t pix = (uint32_t)((y0 + ly) * W + w * 32 + st);
        s.par[run_id(s.rbase, i, sw[i], st)] = kRootBit | (uint32_t)rank;
        if (rank < kLocalCap) s.rpix[rank] = pix;
        else {
          unsigned long long* r4 = rec + (size_t)rank * 4;
          r4[0] = 0ull; r4[1] = 0ull; r4[2] = 0ull; r4[3] = pix;
        }
        ++rank;
      }
      if (++w == WW) { w = 0; ++ly; }
    }
  }
  __syncthreads();
  // ---- per-run quad sums (ownership rules of lgx_joints.cu jl_sums_word), contact with the neighbouring strips
  int32_t* L = p.lab + (size_t)frame * H * W;
  for (int r = tid; r < nruns; r += kLocalThreads) {
    const uint32_t rd = s.runs[r];
    const int idx = (int)(rd >> 5), st = (int)(rd & 31u);
    const int ly = idx / WW, w = idx - ly * WW;
    const uint32_t cur = sw[idx];
    const int e = st + run_len32(cur, st) - 1;
    const int y = y0 + ly;
    const uint64_t A = window34(sw + (size_t)ly * WW, w, WW);
    const uint64_t Bn = window34(sw + (size_t)(ly + 1) * WW, w, WW);      // (halo row below the strip; empty below the image)
    const uint64_t Up = window34(sw + (size_t)(ly - 1) * WW, w, WW);
    uint64_t own = ((1ull << (e + 2)) - 1ull) & ~((1ull << (st + 1)) - 1ull);
    if (!((A >> st) & 1ull)) own |= 1ull << st;
    const uint64_t ownb = (1ull << st) | (1ull << (e + 1));
    const uint64_t tl = A, tr = A >> 1, bl = Bn, br = Bn >> 1;
    const uint64_t q4 = tl & tr & bl & br & own;
    const uint64_t q3 = ((tl & tr & (bl ^ br)) | (bl & br & (tl ^ tr))) & own;
    const uint64_t k1 = ((tl ^ tr) & ~bl & ~br) | ((bl ^ br) & ~tl & ~tr);
    const uint64_t kd = (tl & br & ~tr & ~bl) | (tr & bl & ~tl & ~br);
    const uint64_t kb = (A ^ (A >> 1)) & ~Up & ~(Up >> 1);
    const int xbase = w * 32 - 1;
    uint32_t pr = s.par[r];
    if (!(pr & kRootBit)) pr = s.par[pr];
    const int rk = (int)(pr & ~kRootBit);
    const int n4 = __popcll(q4), n3 = __popcll(q3);
    const int e4 = __popcll(k1 & own) + __popcll(kb & ownb) - n3 - 2 * __popcll(kd & own);
    const long long sx4 = (long long)n4 * xbase + sum_bit_index_flat(q4);
    const long long sx3 = (long long)n3 * xbase + sum_bit_index_flat(q3);
    const int a00 = 2 * n4 + n3;
    const unsigned long long a10 = (unsigned long long)(6 * sx4 + 3 * n4 + 3 * sx3 + __popcll(q3 & tr) + __popcll(q3 & br));
    const unsigned long long a01 = (unsigned long long)((long long)n4 * (6 * y + 3) + 3ll * y * n3 + __popcll(q3 & bl) + __popcll(q3 & br));
    if (rk < kLocalCap) {
      // {a00, e4, a10 lo, a10 hi, a01 lo, a01 hi}: native 32-bit shared-memory atomics, carries by hand
      uint32_t* a = s.acc + 6 * rk;
      if (a00) atomicAdd(&a[0], (uint32_t)a00);
      if (e4) atomicAdd(&a[1], (uint32_t)e4);
      if (a10) {
        const uint32_t lo = (uint32_t)a10, old = atomicAdd(&a[2], lo);
        const uint32_t hi = (uint32_t)(a10 >> 32) + ((uint32_t)(old + lo) < old ? 1u : 0u);
        if (hi) atomicAdd(&a[3], hi);
      }
      if (a01) {
        const uint32_t lo = (uint32_t)a01, old = atomicAdd(&a[4], lo);
        const uint32_t hi = (uint32_t)(a01 >> 32) + ((uint32_t)(old + lo) < old ? 1u : 0u);
        if (hi) atomicAdd(&a[5], hi);
      }
    } else {
      unsigned long long* a = rec + (size_t)rk * 4;
      const unsigned long long packed = (unsigned long long)a00 + ((unsigned long long)(long long)e4 << 32);
      if (packed) atomicAdd(&a[0], packed);
      if (a10) atomicAdd(&a[1], a10);
      if (a01) atomicAdd(&a[2], a01);
    }
    // 8-connected contact across the strip boundary: window bits st .. e + 2 of the row above / below
    const uint64_t span = ((1ull << (e + 3)) - 1ull) & ~((1ull << st) - 1ull);
    if ((ly == 0 && (Up & span)) || (ly == rows - 1 && (Bn & span))) {
      uint32_t rp;
      if (rk < kLocalCap) rp = atomicOr(&s.rpix[rk], kRootBit) & ~kRootBit;
      else rp = (uint32_t)(atomicOr(&rec[(size_t)rk * 4 + 3], kRecBorder) & 0xffffffffull);
      L[y * W + w * 32 + st] = (int32_t)rp;      // boundary run -> its component's first pixel (read by jl_border)
      L[rp] = (int32_t)rp;                       // ... which starts as its own global root
    }
  }
  __syncthreads();
  for (int i = tid; i < min(count, kLocalCap); i += kLocalThreads) {
    const uint32_t rp = s.rpix[i];
    const uint32_t* a = s.acc + 6 * i;
    ulonglong2* r2 = reinterpret_cast<ulonglong2*>(rec + (size_t)i * 4);
    r2[0] = make_ulonglong2((unsigned long long)a[0] | ((unsigned long long)a[1] << 32), (unsigned long long)a[2] | ((unsigned long long)a[3] << 32));
    r2[1] = make_ulonglong2((unsigned long long)a[4] | ((unsigned long long)a[5] << 32),
                            (unsigned long long)(rp & ~kRootBit) | ((rp & kRootBit) ? kRecBorder : 0ull));
  }
}

// ---- links across strip boundaries ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) jl_border(const JointsLocalParams p) {
  const int H = p.H, W = p.W, WW = p.WW;
  const int frame = blockIdx.z;
  const int y = ((int)blockIdx.y + 1) * p.R;        // first row of a strip
  const int w = blockIdx.x * 256 + threadIdx.x;
  if (y >= H || w >= WW || (p.flags[frame] & kSkipFlags)) return;   // (a strip gave up: the whole-frame pass redoes the frame, or capacity)
  const uint32_t* __restrict__ jb = p.jbits + (size_t)frame * H * WW;
  const uint32_t cur = jb[(size_t)y * WW + w];
  if (!cur) return;
  const uint32_t* up = jb + (size_t)(y - 1) * WW;
  const uint32_t up_c = up[w], up_p = w > 0 ? up[w - 1] : 0u, up_n = w + 1 < WW ? up[w + 1] : 0u;
  const uint64_t U = (uint64_t)(up_p >> 31) | ((uint64_t)up_c << 1) | ((uint64_t)(up_n & 1u) << 33);
  int32_t* L = p.lab + (size_t)frame * H * W;
  uint32_t m = cur;
  while (m) {
    const int st = __ffs(m) - 1;
    const int len = run_len32(m, st);
    const int e = st + len - 1;
    m &= ~((len == 32 ? 0xffffffffu : ((1u << len) - 1u)) << st);
    uint64_t mm = U & ((1ull << (e + 3)) - 1ull) & ~((1ull << st) - 1ull);
    if (!mm) continue;
    const int a = L[y * W + w * 32 + st];
    while (mm) {
      const int i = __ffsll((long long)mm) - 1;
      uint32_t word;
      int bb, wu;
      if (i == 0) { word = up_p; bb = 31; wu = w - 1; }
      else if (i <= 32) { word = up_c; bb = i - 1; wu = w; }
      else { word = up_n; bb = 0; wu = w + 1; }
      uf_union(L, a, L[(y - 1) * W + wu * 32 + run_start32(word, bb)]);
      const uint64_t t2 = ~(U >> i);
      const int len2 = __ffsll((long long)t2) - 1;
      mm &= ~(((1ull << len2) - 1ull) << i);
    }
  }
}

// ---- flagged components whose global root is another component: sums move to the root's record -----------------------------
__global__ void __launch_bounds__(256) jl_merge(const JointsLocalParams p) {
  const int frame = blockIdx.y;
  if (p.flags[frame] & kSkipFlags) return;
  const int n = min(p.nrec[frame], p.max_comp);
  unsigned long long* rec = p.rec + (size_t)frame * p.max_comp * 4;
  int32_t* L = p.lab + (size_t)frame * p.H * p.W;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
    const unsigned long long w3 = rec[(size_t)i * 4 + 3];
    if (!(w3 & kRecBorder)) continue;
    const int pix = (int)(w3 & 0xffffffffull);
    const int r = uf_find(L, pix);
    if (r == pix) continue;
    // the root's record: binary search among the records of its strip (sorted by first pixel)
    const int rs = (r / p.W) / p.R;
    int lo = p.sbase[(size_t)frame * p.strips + rs], hi = lo + p.scount[(size_t)frame * p.strips + rs] - 1;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if ((int)(rec[(size_t)mid * 4 + 3] & 0xffffffffull) < r) lo = mid + 1;
      else hi = mid;
    }
    unsigned long long* dst = rec + (size_t)lo * 4;
    atomicAdd(&dst[0], rec[(size_t)i * 4]);
    atomicAdd(&dst[1], rec[(size_t)i * 4 + 1]);
    atomicAdd(&dst[2], rec[(size_t)i * 4 + 2]);
    rec[(size_t)i * 4 + 3] = w3 | kRecDead;
    atomicAdd(&p.sdead[(size_t)frame * p.strips + (pix / p.W) / p.R], 1);
  }
}

// ---- live records -> acc / rootpix in raster order; components with holes ---------------------------------------------------
__global__ void __launch_bounds__(256) jl_compact(const JointsLocalParams p) {
  const int frame = blockIdx.y, strip = blockIdx.x;
  if (p.flags[frame] & kSkipFlags) {
    // more components than max_components: nothing is reported for the frame (the flag says so); a frame that goes to the
    // whole-frame pass gets its count there
    if ((p.flags[frame] & LGX_FLAG_COMP_OVERFLOW) && strip == 0 && threadIdx.x == 0) p.ncomp[frame] = 0;
    return;
  }
  __shared__ int s_warp[8];
  __shared__ int s_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int32_t* cnt = p.scount + (size_t)frame * p.strips;
  const int32_t* dead = p.sdead + (size_t)frame * p.strips;
  // live records of the strips above, and of the frame
  int before = 0, total = 0;
  for (int sidx = tid; sidx < p.strips; sidx += 256) {
    const int live = cnt[sidx] - dead[sidx];
    total += live;
    if (sidx < strip) before += live;
  }
  for (int o = 16; o; o >>= 1) {
    before += __shfl_xor_sync(0xffffffffu, before, o);
    total += __shfl_xor_sync(0xffffffffu, total, o);
  }
  if (lane == 0) { s_warp[warp] = before; }
  __syncthreads();
  if (tid == 0) {
    int t = 0;
    for (int i = 0; i < 8; ++i) t += s_warp[i];
    s_base = t;
  }
  __syncthreads();
  if (lane == 0) s_warp[warp] = total;
  __syncthreads();
  if (tid == 0) {
    int t = 0;
    for (int i = 0; i < 8; ++i) t += s_warp[i];
    if (strip == 0) {
      p.ncomp[frame] = min(t, p.max_comp);
      if (t > p.max_comp) atomicOr(&p.flags[frame], LGX_FLAG_COMP_OVERFLOW);
    }
  }
  __syncthreads();
  int run = s_base;
  const int b0 = p.sbase[(size_t)frame * p.strips + strip], n = cnt[strip];
  const unsigned long long* rec = p.rec + ((size_t)frame * p.max_comp + b0) * 4;
  unsigned long long* acc = p.acc + (size_t)frame * p.max_comp * 4;
  int32_t* rootpix = p.rootpix + (size_t)frame * p.max_comp;
  for (int i0 = 0; i0 < n; i0 += 256) {
    const int i = i0 + tid;
    ulonglong2 lo = make_ulonglong2(0, 0), hi = make_ulonglong2(0, kRecDead);
    if (i < n) {
      lo = reinterpret_cast<const ulonglong2*>(rec + (size_t)i * 4)[0];
      hi = reinterpret_cast<const ulonglong2*>(rec + (size_t)i * 4)[1];
    }
    const bool live = !(hi.y & kRecDead);
    const unsigned bal = __ballot_sync(0xffffffffu, live);
    __syncthreads();
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    int off = run;
    for (int k = 0; k < warp; ++k) off += s_warp[k];
    int tot = 0;
    for (int k = 0; k < 8; ++k) tot += s_warp[k];
    run += tot;
    const int rank = off + __popc(bal & ((1u << lane) - 1u));
    if (live && rank < p.max_comp) {
      ulonglong2* a2 = reinterpret_cast<ulonglong2*>(acc + (size_t)rank * 4);
      a2[0] = lo;
      a2[1] = make_ulonglong2(hi.x, 0ull);
      rootpix[rank] = (int32_t)(hi.y & 0xffffffffull);
      if ((int)(lo.x >> 32) != 4) {           // Euler number != 1: the component has holes (lgx_joints.cu jl_hole_fix)
        const int slot = atomicAdd(&p.nholes[frame], 1);
        if (slot < kMaxHoles) p.holes[(size_t)frame * kMaxHoles + slot] = rank;
        else atomicOr(&p.flags[frame], LGX_FLAG_GENERIC_FILL);
        atomicOr(&p.flags[frame], LGX_FLAG_HOLES);
      }
    }
  }
}

size_t local_smem_bytes(int R, int WW, int capr) {
  return (size_t)(R + 2) * WW * 4 + (size_t)R * WW * (4 + 2) + (size_t)capr * 8 + (size_t)kLocalCap * (24 + 4) + 34 * 4 + 16;
}

}  // namespace

// Strip geometry: about kStripWords words per strip keeps a CTA near 67 KB of shared memory (three CTAs per SM); the run arrays
// hold two runs per word (a laser-grid joints plane has ~0.8; more: the frame goes to the whole-frame pass).
static int strip_words() {
  static int v = 0;
  if (!v) {
    const char* e = getenv("LGX_JL_WORDS");      // tuning experiments only
    v = e ? atoi(e) : 2048;
    if (v < 256) v = 256;
  }
  return v;
}

int joints_local_rows(int W) {
  const int WW = bits_pitch(W);
  int r = strip_words() / WW;
  if (r > 32) r = 32;
  return r;                      // < 2: the image is too wide for this kernel (the whole-frame pass is used instead)
}

int joints_local_runs(int W) { return 2 * joints_local_rows(W) * bits_pitch(W); }

cudaError_t launch_joints_local(const JointsLocalParams& p, int batch, cudaStream_t stream) {
  static unsigned long long attr_done = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  if (!(attr_done >> (dev & 63) & 1ull)) {
    cudaError_t e = cudaFuncSetAttribute(jl_local, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e != cudaSuccess) return e;
    attr_done |= 1ull << (dev & 63);
  }
  jl_local<<<dim3(p.strips, batch), kLocalThreads, local_smem_bytes(p.R, p.WW, p.capr), stream>>>(p);
  if (p.strips > 1) jl_border<<<dim3((p.WW + 255) / 256, p.strips - 1, batch), 256, 0, stream>>>(p);
  if (p.strips > 1) jl_merge<<<dim3(16, batch), 256, 0, stream>>>(p);
  jl_compact<<<dim3(p.strips, batch), 256, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace lgx
