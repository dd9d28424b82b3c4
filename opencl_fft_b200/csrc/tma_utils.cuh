// tma_utils.cuh -- inline-PTX helpers for TMA-fed pipelines: 1-D bulk global->shared copies (cp.async.bulk, SASS
// UBLKCP) whose completion is counted on shared-memory mbarriers. Used by the partitioned-convolution MAC
// (pconv_kernels.cuh) and by the one-SM FFT (fft_sm.cuh), whose two input rounds are staged by the TMA engine.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace b2f {
namespace tma {
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t mbar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "TMA_WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra TMA_WAIT_DONE;\n"
      "bra TMA_WAIT_LOOP;\n"
      "TMA_WAIT_DONE:\n"
      "}\n" ::"r"(mbar),
      "r"(parity)
      : "memory");
}
// address of the same shared-memory variable in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_rank(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// arrive on an mbarrier of another CTA of the cluster; release at cluster scope: this thread's earlier writes --
// and, through a preceding bar.sync, its CTA's -- are visible to whoever observes the phase completing
__device__ __forceinline__ void mbar_arrive_cluster_release(uint32_t remote_mbar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote_mbar) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(mbar)
               : "memory");
}
// ask the TMA engine to bring `bytes` (multiple of 16) of global memory into L2: no registers, no shared memory, no
// completion to wait for (SASS UBLKPF)
__device__ __forceinline__ void prefetch_l2(const void *src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
}  // namespace tma

}  // namespace b2f
