// Device helpers shared by the fused tcgen05 kernels (fused_em.cu, fused_readout.cu).
#pragma once

#include <cuda_fp16.h>
#include <stdint.h>

#include "tc05.cuh"

namespace swem {

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Wait until all T tiles of this unit have arrived at `counter`.  Bounded: a protocol bug aborts
// the kernel with an error instead of hanging the GPU.
__device__ __forceinline__ bool wait_counter(const unsigned* counter, unsigned target) {
  for (unsigned i = 0; i < (1u << 24); ++i) {
    if (ld_acquire_u32(counter) >= target) return true;
    __nanosleep(32);
  }
  return false;
}

// Block-wide wait on an mbarrier phase: ONE thread polls, everybody else parks on the hardware barrier.
// (256 threads polling try_wait starve the single issuing thread of the very mbarrier / TMA / MMA slots
//  it needs -- measured 5x slower MMA issue loops.)  A time-out latches `abort_flag` (uniform after the barrier).
#define SWEM_CTA_WAIT(bar, parity, abort_flag)                 \
  do {                                                         \
    if (threadIdx.x == 0) {                                    \
      if (!tc05::mbar_wait((bar), (parity))) (abort_flag) = 1; \
    }                                                          \
    __syncthreads();                                           \
  } while (0)

__device__ __forceinline__ long long global_ns() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// 2^x on the SFU (MUFU.EX2): relative error ~2^-22, flushes results below the normal range to 0
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void split_half(float v, __half& hi, __half& lo) {
  hi = __float2half_rn(v);
  lo = __float2half_rn(v - __half2float(hi));
}

// ---- thread-block clusters: rank, barrier, distributed shared memory ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t map_to_peer(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_f4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void st_cluster_f2(uint32_t addr, float a, float b) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float a) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(a) : "memory");
}


// ---- named barriers (ids 1..15; 0 is __syncthreads) -------------------------------------------
__device__ __forceinline__ void bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// ---- mbarriers across a cluster: arrive on a peer CTA's barrier (release at cluster scope: the st.shared::cluster
// stores issued before it are visible to the waiter), wait with acquire at cluster scope -----------------------------
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster_scope(uint64_t* bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cta.b64 _, [%0];" ::"r"(tc05::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(tc05::smem_u32(bar)), "r"(parity)
      : "memory");
  return ok;
}
__device__ __forceinline__ bool mbar_wait_cluster(uint64_t* bar, uint32_t parity, uint32_t max_spins = (1u << 24)) {
  if (mbar_try_wait_cluster(bar, parity)) return true;
#pragma unroll 1
  for (uint32_t i = 0; i < max_spins; ++i)
    if (mbar_try_wait_cluster(bar, parity)) return true;
  return false;
}
// one lane polls, the warp parks on the shuffle (polling with every thread starves the thread that issues the MMAs)
__device__ __forceinline__ bool warp_wait(uint64_t* bar, uint32_t parity) {
  int ok = 1;
  if ((threadIdx.x & 31) == 0) ok = tc05::mbar_wait(bar, parity) ? 1 : 0;
  return __shfl_sync(0xffffffffu, ok, 0) != 0;
}
__device__ __forceinline__ void red_release_add(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

}  // namespace swem
