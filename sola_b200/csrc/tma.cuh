// mbarrier + TMA (bulk async copy) helpers shared by the TMA-staged kernels (pair_iou.cu, jf_fused.cu).  sm_90+ PTX, built for sm_100a.
#pragma once
#include "common.cuh"

namespace sola {

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// Wait for the phase with the given parity.  A lost transaction must never hang the GPU (a hung box is a dead box), but a slow
// NVLink peer or a time-sliced context must not kill a healthy kernel either: the bound is wall time (20 s), checked every 64 K polls.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  const unsigned a = smem_u32(bar);
  unsigned done = 0;
  unsigned long long t0 = 0;
  for (unsigned spin = 0; !done; ++spin) {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(a), "r"(parity) : "memory");
    if (!done && (spin & 0xffffu) == 0xffffu) {
      const unsigned long long now = global_timer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 20000000000ull) __trap();
    }
  }
}

// rank-2 tiled tensor-map load (SASS UTMALDG.2D): box -> smem, completion counted in bytes on `bar`
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tensor_map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
                   "r"(smem_u32(smem_dst)), "l"(tensor_map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}

// 1-D bulk copy global -> shared (SASS UBLKCP): 16-byte aligned addresses, size a multiple of 16
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gmem_src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                   "r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

}  // namespace sola
