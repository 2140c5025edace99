// Bit-run helpers shared by the contour-equivalent kernels (lgx_joints.cu: whole-frame union-find, used for the rare
// second pass on hole-filled masks; lgx_joints_local.cu: strip-local labelling in shared memory, the default first pass).
#pragma once
#include "lgx_internal.cuh"

namespace lgx {
namespace {

__device__ __forceinline__ int uf_find(int32_t* L, int p) {
  int q = ((volatile int32_t*)L)[p];
  while (q != p) {
    p = q;
    q = ((volatile int32_t*)L)[p];
  }
  return p;
}

__device__ __forceinline__ void uf_union(int32_t* L, int a, int b) {
  bool done;
  do {
    a = uf_find(L, a);
    b = uf_find(L, b);
    if (a < b) {
      int old = atomicMin(&L[b], a);
      done = (old == b);
      b = old;
    } else if (b < a) {
      int old = atomicMin(&L[a], b);
      done = (old == a);
      a = old;
    } else {
      done = true;
    }
  } while (!done);
}

// length of the run of ones starting at bit s of m (bit s must be set)
__device__ __forceinline__ int run_len32(uint32_t m, int s) {
  uint32_t t = ~(m >> s);
  return t ? (__ffs(t) - 1) : 32;   // (m>>s) has zeros shifted in, so t != 0 unless s == 0 and m is all ones
}

// start bit of the word-run that contains bit b of word (bit b must be set)
__device__ __forceinline__ int run_start32(uint32_t word, int b) {
  uint32_t below = word << (31 - b);          // bit b -> bit 31
  int lead = __clz(~below);                   // leading ones (>= 1); 32 if below is all ones
  return b - (lead - 1);
}

// 34-bit window of a row around word w: bit i <-> pixel x = 32*w - 1 + i
__device__ __forceinline__ uint64_t window34(const uint32_t* __restrict__ row, int w, int WW) {
  uint64_t c = row[w];
  uint64_t p = (w > 0) ? (row[w - 1] >> 31) : 0u;
  uint64_t n = (w + 1 < WW) ? (row[w + 1] & 1u) : 0u;
  return p | (c << 1) | (n << 33);
}

__device__ __forceinline__ int sum_bit_index(uint64_t m) {
  int s = 0;
  while (m) {
    s += __ffsll((long long)m) - 1;
    m &= m - 1;
  }
  return s;
}

}  // namespace
}  // namespace lgx
