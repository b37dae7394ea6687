// PTX wrappers shared by the TMA-fed kernels (k_pyramid, k_find_points): mbarrier + 2-D tiled
// cp.async.bulk.tensor loads (SASS: UTMALDG / SYNCS), and the host-side tensor-map encoder.
#ifndef CSB_TMA_UTIL_H
#define CSB_TMA_UTIL_H

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

// Encodes a 2-D fp32 tiled tensor map (host).  dim0 = fastest dimension (elements), row stride in bytes
// (multiple of 16), box = elements per TMA load; out-of-bounds elements of a box are zero-filled.
// Returns 0 on success.  Uses the driver entry point, so the library does not link libcuda.
int csb_tmap_2d_f32(CUtensorMap *out, const float *base, uint64_t dim0, uint64_t dim1, uint64_t stride_bytes, uint32_t box0,
                    uint32_t box1);

#ifdef __CUDACC__
namespace tma {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
// one 2-D tiled TMA load: the map's box at element coordinates (x, y) -> shared memory, completion on `bar`
__device__ __forceinline__ void load_2d(void *dst_smem, const CUtensorMap *map, int x, int y, uint64_t *bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar))
               : "memory");
}
// orders generic-proxy accesses (global or shared) before subsequent async-proxy (TMA) accesses
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

}  // namespace tma
#endif

#endif
