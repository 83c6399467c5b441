// Communication latency / bandwidth probe for the cross-tile all-reduce of the EM kernel (sm_100a):
//   1. L2 path: red / bulk-reduce + completion + fence + flag, poll, read back (what em_res_kernel does per iteration)
//   2. distributed shared memory inside a cluster: st.async latency (ping-pong), st.async and bulk-copy bandwidth,
//      all-to-one reduction of 33 KB partials from CS - 1 peers (cluster sizes 2 / 4 / 8 / 16)
// Build: make -C tools comm_probe;  run: tools/comm_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ long long clk() { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)); return t; }
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t r) { uint32_t o; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(r)); return o; }
__device__ __forceinline__ void csync() { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void st_async_u4(uint32_t addr, uint4 v, uint32_t bar) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(bar) : "memory");
}
__device__ __forceinline__ void st_async_u2(uint32_t addr, uint32_t a, uint32_t b, uint32_t bar) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1, %2}, [%3];" ::"r"(addr), "r"(a), "r"(b), "r"(bar) : "memory");
}
__device__ __forceinline__ unsigned ld_acq(const unsigned* p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }

// ---- 1. L2 path -------------------------------------------------------------------------------------------------------
// grid = G CTAs of 512 threads; every CTA reduce-adds a 32 KB partial into acc[group], arrives on a counter, waits for the
// `per_group` CTAs of its group, reads the totals back.  CTA 0 records clock64 stamps.  mode 0: per-lane red; 1: bulk reduce
__global__ void __launch_bounds__(512, 1) l2_allreduce(float* acc, unsigned* counters, int per_group, int iters, int mode, long long* out, float* sink) {
  extern __shared__ __align__(128) uint8_t sm[];
  float* stage = reinterpret_cast<float*>(sm);
  const int tid = threadIdx.x, g = blockIdx.x / per_group;
  for (int i = tid; i < 8192; i += 512) stage[i] = 1.0f;
  __syncthreads();
  float s = 0.f;
  for (int it = 0; it < iters; ++it) {
    float* a = acc + ((size_t)it * (gridDim.x / per_group) + g) * 8192;
    unsigned* c = counters + it * (gridDim.x / per_group) + g;
    long long t0 = clk();
    if (mode == 0) {
      for (int j = 0; j < 16; ++j) atomicAdd(a + j * 512 + tid, stage[j * 512 + tid]);
      __syncthreads();
    } else {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      if (tid == 0) {
        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(a), "r"(smem_u32(stage)), "r"(32768) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      }
    }
    long long t1 = clk();
    if (tid == 0) __threadfence();
    long long t2 = clk();
    if (tid == 0) {
      atomicAdd(c, 1u);
      while (ld_acq(c) < (unsigned)per_group) {}
    }
    long long t3 = clk();
    __syncthreads();
    float v = 0.f;
    for (int j = 0; j < 16; ++j) v += __ldcg(a + j * 512 + tid);
    s += v;
    if (v == 123.f) s += 1.f;
    long long t4 = clk();
    if (blockIdx.x == 0 && tid == 0) { out[it * 4 + 0] = t1 - t0; out[it * 4 + 1] = t2 - t1; out[it * 4 + 2] = t3 - t2; out[it * 4 + 3] = t4 - t3; }
    __syncthreads();
  }
  if (s == 1234567.f) sink[0] = s;
}

// ---- 2. DSMEM -----------------------------------------------------------------------------------------------------------
// cluster of CS CTAs, 512 threads.  Test A: ping-pong of 8 bytes between rank 0 and rank 1 (latency).  Test B: every CTA sends
// `bytes` to rank (r + 1) % CS with st.async v4 (bandwidth, all links busy).  Test C: all-to-one -- every rank r != 0 sends
// bytes / CS ... i.e. reduce-scatter pattern: every CTA sends a slice of `bytes / CS` to every other CTA.
__global__ void __launch_bounds__(512, 1) dsmem_probe(int bytes, long long* out) {
  extern __shared__ __align__(128) uint8_t sm[];
  __shared__ uint64_t bar[4];
  const int tid = threadIdx.x;
  uint32_t CS;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(CS));
  const uint32_t r = cluster_rank();
  uint8_t* rx = sm;                    // receive buffer: up to 64 KB
  if (tid == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  csync();
  // ---- A: ping-pong latency (ranks 0, 1), 64 round trips
  long long tA = 0;
  if (r < 2 && tid == 0) {
    const uint32_t peer_rx = mapa(smem_u32(rx), r ^ 1), peer_bar = mapa(smem_u32(&bar[0]), r ^ 1);
    long long t0 = clk();
    for (int i = 0; i < 64; ++i) {
      if (r == 0) {
        mbar_expect(&bar[0], 8);
        st_async_u2(peer_rx, i, i, peer_bar);
        mbar_wait(&bar[0], i & 1);
      } else {
        mbar_expect(&bar[0], 8);
        mbar_wait(&bar[0], i & 1);
        st_async_u2(peer_rx, i, i, peer_bar);
      }
    }
    tA = (clk() - t0) / 64;
  }
  csync();
  // ---- B: ring send of `bytes` with st.async v4
  if (tid == 0) mbar_expect(&bar[1], bytes);
  csync();
  long long t0 = clk();
  {
    const uint32_t dst = mapa(smem_u32(rx), (r + 1) % CS), dbar = mapa(smem_u32(&bar[1]), (r + 1) % CS);
    for (int o = tid * 16; o < bytes; o += 512 * 16) st_async_u4(dst + o, make_uint4(o, o, o, o), dbar);
  }
  long long t1 = clk();
  if (tid == 0) mbar_wait(&bar[1], 0);
  __syncthreads();
  long long t2 = clk();
  csync();
  // ---- C: reduce-scatter pattern: slice of bytes / CS to every other rank (st.async v4)
  const int slice = bytes / CS;
  if (tid == 0) mbar_expect(&bar[2], slice * (CS - 1));
  csync();
  long long t3 = clk();
  for (uint32_t k = 1; k < CS; ++k) {
    const uint32_t to = (r + k) % CS;
    const uint32_t dst = mapa(smem_u32(rx), to) + ((r + CS - to) % CS - 1) * slice, dbar = mapa(smem_u32(&bar[2]), to);
    for (int o = tid * 16; o < slice; o += 512 * 16) st_async_u4(dst + o, make_uint4(o, o, o, o), dbar);
  }
  long long t4 = clk();
  if (tid == 0) mbar_wait(&bar[2], 0);
  __syncthreads();
  long long t5 = clk();
  csync();
  // ---- D: ring send with one bulk copy (cp.async.bulk.shared::cluster.shared::cta)
  if (tid == 0) mbar_expect(&bar[3], bytes);
  csync();
  long long t6 = clk();
  if (tid == 0) {
    const uint32_t dst = mapa(smem_u32(rx), (r + 1) % CS), dbar = mapa(smem_u32(&bar[3]), (r + 1) % CS);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "r"(smem_u32(sm + 65536)), "r"(bytes), "r"(dbar) : "memory");
    mbar_wait(&bar[3], 0);
  }
  __syncthreads();
  long long t7 = clk();
  long long tE0 = clk();
  csync();
  long long tE1 = clk();
  if (blockIdx.x == 0 && tid == 0) { out[0] = tA; out[1] = t1 - t0; out[2] = t2 - t0; out[3] = t4 - t3; out[4] = t5 - t3; out[5] = t7 - t6; out[6] = tE1 - tE0; }
  csync();
}

int main() {
  int dev = 0; cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, dev));
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  printf("%s, %d SMs, clock attr %.0f MHz (clock64 ticks at the current SM clock)\n", prop.name, prop.multiProcessorCount, khz / 1e3);
  float *acc, *sink; unsigned* cnt; long long* out;
  const int iters = 8;
  CK(cudaMalloc(&acc, (size_t)iters * 148 * 8192 * 4)); CK(cudaMalloc(&cnt, iters * 148 * 4)); CK(cudaMalloc(&out, 4096)); CK(cudaMalloc(&sink, 64));
  CK(cudaFuncSetAttribute(l2_allreduce, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
  long long h[64];
  for (int mode = 0; mode < 2; ++mode)
    for (int cfg = 0; cfg < 3; ++cfg) {
      const int grid = cfg == 0 ? 1 : cfg == 1 ? 13 : 130, per_group = cfg == 0 ? 1 : 13;
      for (int rep = 0; rep < 2; ++rep) {
        CK(cudaMemset(acc, 0, (size_t)iters * 148 * 8192 * 4)); CK(cudaMemset(cnt, 0, iters * 148 * 4));
        l2_allreduce<<<grid, 512, 32768>>>(acc, cnt, per_group, iters, mode, out, sink);
        CK(cudaDeviceSynchronize());
      }
      CK(cudaMemcpy(h, out, iters * 4 * 8, cudaMemcpyDeviceToHost));
      printf("L2 all-reduce 32 KB, %s, grid %3d (groups of %2d): cycles per iteration [reduce+complete | fence | arrive+wait | read-back]\n", mode ? "bulk reduce" : "per-lane red", grid, per_group);
      for (int it = 0; it < iters; ++it) printf("   it%d: %6lld %6lld %6lld %6lld\n", it, h[it * 4], h[it * 4 + 1], h[it * 4 + 2], h[it * 4 + 3]);
    }
  CK(cudaFuncSetAttribute(dsmem_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072 + 1024));
  CK(cudaFuncSetAttribute(dsmem_probe, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  for (int cs = 2; cs <= 16; cs *= 2) {
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(512, 1, 1); cfg.dynamicSmemBytes = 131072 + 1024;
    cudaLaunchAttribute attr[1]; attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cfg.gridDim = dim3(cs, 1, 1);
    int ncl = 0; cudaError_t e = cudaOccupancyMaxActiveClusters(&ncl, dsmem_probe, &cfg);
    printf("cluster size %2d: max active clusters %d (%s)\n", cs, ncl, cudaGetErrorString(e));
    if (e != cudaSuccess || ncl == 0) { cudaGetLastError(); continue; }
    for (int full = 0; full < 2; ++full) {
      cfg.gridDim = dim3(full ? cs * ncl : cs, 1, 1);
      for (int bytes = 4096; bytes <= 32768; bytes *= 8) {
        for (int rep = 0; rep < 2; ++rep) { CK(cudaLaunchKernelEx(&cfg, dsmem_probe, bytes, out)); CK(cudaDeviceSynchronize()); }
        CK(cudaMemcpy(h, out, 7 * 8, cudaMemcpyDeviceToHost));
        printf("   %s, %5d B: ping-pong round trip %lld cyc | ring st.async issue %lld / done %lld cyc (%.1f B/cyc) | scatter to %d peers issue %lld / done %lld cyc (%.1f B/cyc out) | ring bulk copy %lld cyc (%.1f B/cyc) | cluster barrier %lld cyc\n",
               full ? "chip full" : "1 cluster", bytes, h[0], h[1], h[2], (double)bytes / h[2], cs - 1, h[3], h[4], (double)(bytes / cs * (cs - 1)) / h[4], h[5], (double)bytes / h[5], h[6]);
      }
    }
  }
  return 0;
}
