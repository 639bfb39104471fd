// mbarrier / bulk-copy PTX helpers shared by the tcgen05 convolution and the smem-staged streaming kernels.
#pragma once
#include <stdint.h>
#include <stdio.h>

namespace basi {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must trap (and surface as a CUDA error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 8000000000LL) {   // ~4 s at 2 GHz
      printf("basi: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar,
             parity);
      __trap();
    }
  }
}
// 1-D bulk copy global -> shared, completion counted on an mbarrier (addresses and size multiples of 16 bytes)
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// one lane of a fully converged warp (elect.sync).  Code that issues bulk / tensor copies or tcgen05 instructions runs
// warp-uniformly and predicates only the instruction on the elected lane: the compiler then keeps the addresses in
// uniform registers instead of wrapping every uniform-datapath instruction of a single-lane branch in an
// ELECT / R2UR / BRA.U.ANY loop.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
  return pred != 0;
}
// Grid-wide barrier for kernels whose CTAs are all co-resident (cooperative launch): called by ONE thread per CTA,
// bracketed by CTA-level barriers.  The counter only ever grows (the target is the next multiple of gridDim.x), so it
// can be reused by the next barrier of the same launch; it must be zero at the first launch of a step.
__device__ __forceinline__ void grid_barrier_thread0(unsigned int* ctr) {
  __threadfence();
  const unsigned int old = atomicAdd(ctr, 1u);
  const unsigned int target = (old / gridDim.x + 1u) * gridDim.x;
  const long long t0 = clock64();
  for (;;) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    if ((int)(v - target) >= 0) break;
    if (clock64() - t0 > 8000000000LL) {
      printf("basi: grid barrier timed out (block %d, counter %u, target %u)\n", blockIdx.x, v, target);
      __trap();
    }
  }
  __threadfence();
}

__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

}  // namespace basi
