// mbarrier / TMA (cp.async.bulk.tensor) helpers and the host-side tensor-map encoder shared by the TMA kernels
// (lgx_ridge_ws.cu, lgx_sauvola.cu).  sm_100a only.
#pragma once
#include <cuda.h>

#include "lgx_internal.cuh"

namespace lgx {
namespace tma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(0x989680u)      // suspend-time hint: the thread sleeps until the phase completes
      : "memory");
  return ok != 0;
}
#ifndef LGX_MBAR_SLEEP_NS
#define LGX_MBAR_SLEEP_NS 0
#endif
// A role that waits longer than ~4 s is a protocol bug: trap instead of hanging the device.
static __device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity) {
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  unsigned spins = 0;
  while (!mbar_try(bar, parity)) {
#if LGX_MBAR_SLEEP_NS > 0
    __nanosleep(LGX_MBAR_SLEEP_NS);       // give the issue slots of a waiting role to the roles that have work
#endif
    if ((++spins & 63u) == 0) {
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 4000000000ull) __trap();
    }
  }
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (!mbar_try(bar, parity)) mbar_wait_slow(bar, parity);
}
// the same with the waiting time added to `acc` (debug option LGX_OPT_RIDGE_PROF)
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, bool prof, long long& acc) {
  const long long t0 = prof ? clock64() : 0;    // (the first try may already suspend the thread)
  if (!mbar_try(bar, parity)) mbar_wait_slow(bar, parity);
  if (prof) acc += clock64() - t0;
}

__device__ __forceinline__ void tma_load_3d(const CUtensorMap* tm, uint32_t bar, uint32_t dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* tm, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(tm)),
               "r"(src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

// ---- host side: cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda) ---------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

inline bool encode3(CUtensorMap* tm, CUtensorMapDataType dt, int esize, const void* base, int W, int H, int nb, size_t row_bytes,
             size_t frame_bytes, int box_w, int box_h, CUtensorMapSwizzle sw) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)nb};
  cuuint64_t strides[2] = {(cuuint64_t)row_bytes, (cuuint64_t)frame_bytes};
  cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1u};
  cuuint32_t estr[3] = {1u, 1u, 1u};
  (void)esize;
  return fn(tm, dt, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace tma
}  // namespace lgx
