// miso_b200/csrc/tile_mem.cuh -- shared-memory plumbing of the chain kernels (sm_100a):
// TMA bulk copy + mbarrier, explicit shared/global loads, warp shuffles of doubles.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace misob200 {

__device__ __forceinline__ double shfl_d(double v, int src) {
  return __shfl_sync(0xffffffffu, v, src);
}

// ---- TMA bulk copy global -> shared, completion on an mbarrier ---------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t) __cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// ---- explicit address-space loads ------------------------------------------------
// The tile normally sits in shared memory (32-bit shared addresses, LDS); genes
// whose tile does not fit in a slot are streamed from global/L2 instead.
template <bool SMEM> struct TileMem;
template <> struct TileMem<true> {
  using addr_t = uint32_t;
  static __device__ __forceinline__ addr_t base(const void *p) { return smem_u32(p); }
  static __device__ __forceinline__ uint32_t ld(addr_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
  }
  // 16 bytes from an 8-byte aligned address (two 8-byte loads)
  static __device__ __forceinline__ uint4 ld4(addr_t a) {
    uint4 v;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2+8];" : "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
  }
};
template <> struct TileMem<false> {
  using addr_t = const unsigned char *;
  static __device__ __forceinline__ addr_t base(const void *p) { return static_cast<addr_t>(p); }
  static __device__ __forceinline__ uint32_t ld(addr_t a) {
    uint32_t v;
    asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(a));
    return v;
  }
  static __device__ __forceinline__ uint4 ld4(addr_t a) {
    uint4 v;
    asm volatile("ld.global.nc.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(a));
    asm volatile("ld.global.nc.v2.u32 {%0,%1}, [%2+8];" : "=r"(v.z), "=r"(v.w) : "l"(a));
    return v;
  }
};
__device__ __forceinline__ double lds_f64(uint32_t a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
  return v;
}

// ---- shared math bodies ------------------------------------------------------------
// exp, log and the IEEE division expand to 30-60 instructions each.  Inlined at every
// call site of the scalar part they push the hot loop past the 32 KB instruction cache
// (ncu: a quarter of all warp stalls were "no instruction"); one out-of-line body per
// function keeps the whole iteration resident.
__device__ __noinline__ double d_exp(double x) { return exp(x); }
__device__ __noinline__ double d_log(double x) { return log(x); }
__device__ __noinline__ double d_div(double a, double b) { return a / b; }

}  // namespace misob200
