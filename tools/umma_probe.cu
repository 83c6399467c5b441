// Standalone sm_100a probe: checks our tcgen05 shared-memory descriptor / operand-layout / TMEM
// conventions (swem_b200/csrc/tc05.cuh) against a CPU GEMM, one hypothesis per invocation so that
// a faulting variant cannot poison the others.   Build: make -C tools   Run: ./umma_probe <test#>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../swem_b200/csrc/tc05.cuh"

using namespace tc05;

struct Cfg {
  int M, N, K;
  int kind;      // 0 f16, 1 tf32
  int a_tmem;    // A operand read from TMEM (written with tcgen05.st, two halfs per column)
  int a_major, b_major;
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo;
  uint32_t a_kstep, b_kstep;  // bytes added to the start address per MMA (TMEM columns when a_tmem)
  int swap;                   // swap LBO/SBO fields in the descriptors (alternative reading of the spec)
  uint32_t a_bytes, b_bytes;
  int repeat;                 // issue the whole K loop this many times (timing)
  int swz;                    // 1: both operands K-major in the 128-byte-swizzled layout (rows of 64 halfs), SBO = 1024
};

__global__ void __launch_bounds__(128) probe(const uint8_t* a_img, const uint8_t* b_img, const __half* a_rowmajor,
                                             float* D, Cfg c, int* status, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* sa = smem;
  uint8_t* sb = smem + ((c.a_bytes + 1023) / 1024) * 1024;
  for (uint32_t i = tid * 16; i < c.a_bytes; i += 128 * 16) *(uint4*)(sa + i) = *(const uint4*)(a_img + i);
  for (uint32_t i = tid * 16; i < c.b_bytes; i += 128 * 16) *(uint4*)(sb + i) = *(const uint4*)(b_img + i);
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const uint32_t a_col0 = 256;
  if (c.a_tmem) {  // row tid of A, K halfs -> K/2 packed columns starting at column 256
    for (int k0 = 0; k0 < c.K; k0 += 16) {
      uint32_t r[8];
      for (int j = 0; j < 8; ++j) {
        __half2 h = __halves2half2(a_rowmajor[tid * c.K + k0 + 2 * j], a_rowmajor[tid * c.K + k0 + 2 * j + 1]);
        r[j] = *reinterpret_cast<uint32_t*>(&h);
      }
      tmem_st8(tmem_addr(tmem, warp * 32, a_col0 + k0 / 2), r);
    }
    tmem_st_wait();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
  }
  long long t0 = clock64();
  if (tid == 0) {
    const uint32_t fmt = c.kind == 0 ? kFmtF16 : kFmtTF32;
    const uint32_t idesc = make_idesc(c.M, c.N, fmt, fmt, c.a_major, c.b_major);
    const int kper = c.kind == 0 ? 16 : 8;
    for (int rep = 0; rep < c.repeat; ++rep)
      for (int k = 0; k < c.K / kper; ++k) {
        const uint64_t ad = c.swz ? make_sdesc_sw128(smem_u32(sa) + k * c.a_kstep, c.a_sbo)
                            : c.swap ? make_sdesc(smem_u32(sa) + k * c.a_kstep, c.a_sbo, c.a_lbo)
                                     : make_sdesc(smem_u32(sa) + k * c.a_kstep, c.a_lbo, c.a_sbo);
        const uint64_t bd = c.swz ? make_sdesc_sw128(smem_u32(sb) + k * c.b_kstep, c.b_sbo)
                            : c.swap ? make_sdesc(smem_u32(sb) + k * c.b_kstep, c.b_sbo, c.b_lbo)
                                     : make_sdesc(smem_u32(sb) + k * c.b_kstep, c.b_lbo, c.b_sbo);
        const uint32_t acc = (k > 0 || rep > 0) ? 1u : 0u;
        if (c.a_tmem) mma_f16_ts(tmem, tmem_addr(tmem, 0, a_col0 + k * c.a_kstep), bd, idesc, acc);
        else if (c.kind == 0) mma_f16_ss(tmem, ad, bd, idesc, acc);
        else mma_tf32_ss(tmem, ad, bd, idesc, acc);
      }
    mma_commit(&bar);
  }
  const bool ok = mbar_wait(&bar, 0, 1u << 22);
  long long t1 = clock64();
  tc_fence_after_sync();
  if (tid == 0) {
    *status = ok ? 0 : 1;
    *cycles = t1 - t0;
  }
  if (ok) {
    for (int c0 = 0; c0 < c.N; c0 += 16) {
      uint32_t r[16];
      tmem_ld16(tmem_addr(tmem, warp * 32, c0), r);
      tmem_ld_wait();
      for (int j = 0; j < 16; ++j) D[(warp * 32 + lane) * c.N + c0 + j] = __uint_as_float(r[j]);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---- throughput: 64 statically unrolled MMAs into one accumulator, SS vs TS (A from TMEM) -----
template <int N, bool TS, bool ALT>
__global__ void __launch_bounds__(128) tput(long long* cycles, int* status) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (16384 + N * 128 * 2) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // 1.0h
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  if (TS) {
    uint32_t r[8];
    for (int j = 0; j < 8; ++j) r[j] = 0x3c003c00u;
    for (int c = 0; c < 64; c += 8) tmem_st8(tmem_addr(tmem, warp * 32, 256 + c), r);
    tmem_st_wait();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
  }
  long long t0 = 0, t1 = 0;
  if (warp == 0) {
    if ((tid & 31) == 0) {
      const uint32_t idesc = make_idesc(128, N, kFmtF16, kFmtF16, kMajorK, kMajorK);
      const uint32_t sa = smem_u32(smem), sb = smem_u32(smem) + 16384;
      t0 = clock64();
#pragma unroll
      for (int k = 0; k < 64; ++k) {
        const uint64_t bd = make_sdesc(sb + (k & 7) * 256, 128, 1024);
        const uint32_t d = ALT ? tmem + (k & 1) * 128 : tmem;       // ALT: alternate between two accumulators
        if (TS) mma_f16_ts(d, tmem_addr(tmem, 0, 256 + (k & 7) * 8), bd, idesc, k > 1);
        else mma_f16_ss(d, make_sdesc(sa + (k & 7) * 256, 128, 1024), bd, idesc, k > 1);
      }
      mma_commit(&bar);
      t1 = clock64();
    }
    __syncwarp();
  }
  const bool ok = mbar_wait(&bar, 0, 1u << 22);
  const long long t2 = clock64();
  if (tid == 0) { cycles[0] = t1 - t0; cycles[1] = t2 - t0; *status = ok ? 0 : 1; }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int N, bool TS, bool ALT>
static int run_tput(const char* name) {
  long long* dc; int* ds;
  cudaMalloc(&dc, 16); cudaMalloc(&ds, 4);
  const size_t smem = 16384 + N * 128 * 2 + 1024;
  cudaFuncSetAttribute(tput<N, TS, ALT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  tput<N, TS, ALT><<<1, 128, smem>>>(dc, ds);
  cudaError_t e = cudaDeviceSynchronize();
  long long c[2]; int st;
  cudaMemcpy(c, dc, 16, cudaMemcpyDeviceToHost); cudaMemcpy(&st, ds, 4, cudaMemcpyDeviceToHost);
  printf("[tput] %s: %s issue %lld cyc, issue+complete %lld cyc for 64 MMAs (%.1f cyc/MMA)%s\n", name,
         e == cudaSuccess ? "ok" : cudaGetErrorString(e), c[0], c[1], c[1] / 64.0, st ? " TIMEOUT" : "");
  return 0;
}

// ---- host ------------------------------------------------------------------------------------
static int esz(int kind) { return kind == 0 ? 2 : 4; }

// byte offset of element (r, k) in the canonical no-swizzle layout
static uint32_t off(int major, int kind, int r, int k, uint32_t lbo, uint32_t sbo) {
  const int es = esz(kind), E = 16 / es;
  if (major == 0) return (r % 8) * 16 + (r / 8) * sbo + (k / E) * lbo + (k % E) * es;
  return (r % E) * es + (k % 8) * 16 + (r / E) * sbo + (k / 8) * lbo;
}

// byte offset of element (r, k) of a K-major f16 operand in the 128-byte-swizzled layout (K <= 64: one atom wide)
static uint32_t off_sw128(int r, int k) {
  const int chunk = (k * 2) / 16, i = r % 8;
  return (r / 8) * 1024 + i * 128 + ((chunk ^ i) * 16) + (k * 2) % 16;
}

static void put(std::vector<uint8_t>& img, uint32_t o, int kind, float v) {
  if (kind == 0) {
    __half h = __float2half(v);
    memcpy(&img[o], &h, 2);
  } else {
    memcpy(&img[o], &v, 4);
  }
}

int main(int argc, char** argv) {
  const int t = argc > 1 ? atoi(argv[1]) : 0;
  if (t == 20) return run_tput<128, false, false>("SS M128 N128 K16");
  if (t == 21) return run_tput<128, true, false>("TS M128 N128 K16");
  if (t == 22) return run_tput<256, false, false>("SS M128 N256 K16");
  if (t == 23) return run_tput<256, true, false>("TS M128 N256 K16");
  if (t == 24) return run_tput<64, false, false>("SS M128 N64 K16");
  if (t == 25) return run_tput<128, true, true>("TS M128 N128 K16, two accumulators alternating");
  if (t == 26) return run_tput<128, false, true>("SS M128 N128 K16, two accumulators alternating");
  Cfg c{};
  c.repeat = 1;
  const char* name = "";
  auto kk_p1 = [&](int M, int N, int K, int kind) {  // both K-major, k-groups adjacent (LBO=128)
    c.M = M; c.N = N; c.K = K; c.kind = kind; c.a_major = 0; c.b_major = 0;
    const int E = 16 / esz(kind);
    c.a_lbo = 128; c.a_sbo = (K / E) * 128; c.b_lbo = 128; c.b_sbo = (K / E) * 128;
    c.a_kstep = 256; c.b_kstep = 256;
  };
  switch (t) {
    case 0: name = "f16 K/K M128 N128 K64 LBO=k-stride"; kk_p1(128, 128, 64, 0); break;
    case 1: name = "f16 K/K M128 N128 K64 LBO/SBO swapped in descriptor"; kk_p1(128, 128, 64, 0); c.swap = 1; break;
    case 2: name = "f16 K/K M128 N128 K64 row-groups adjacent (SBO=128)"; kk_p1(128, 128, 64, 0);
      c.a_sbo = 128; c.a_lbo = (128 / 8) * 128; c.b_sbo = 128; c.b_lbo = (128 / 8) * 128;
      c.a_kstep = 2 * c.a_lbo; c.b_kstep = 2 * c.b_lbo; break;
    case 3: name = "f16 A MN-major (SBO=128, LBO=2048) B K-major M128 N256 K64"; kk_p1(128, 256, 64, 0);
      c.a_major = 1; c.a_sbo = 128; c.a_lbo = (128 / 8) * 128; c.a_kstep = 2 * c.a_lbo; break;
    case 4: name = "f16 A MN-major swapped"; kk_p1(128, 256, 64, 0);
      c.a_major = 1; c.a_sbo = 128; c.a_lbo = (128 / 8) * 128; c.a_kstep = 2 * c.a_lbo; c.swap = 1; break;
    case 5: name = "f16 K/K M128 N80 K64"; kk_p1(128, 80, 64, 0); break;
    case 6: name = "f16 K/K M64 N64 K64 (lane mapping)"; kk_p1(64, 64, 64, 0); break;
    case 7: name = "tf32 K/K M128 N64 K32"; kk_p1(128, 64, 32, 1); break;
    case 8: name = "f16 A from TMEM (packed pairs), B K-major M128 N64 K64"; kk_p1(128, 64, 64, 0);
      c.a_tmem = 1; c.a_kstep = 8; break;
    case 9: name = "timing f16 K/K M128 N256 K192 x8"; kk_p1(128, 256, 192, 0); c.repeat = 8; break;
    case 10: name = "timing f16 K/K M128 N128 K192 x8"; kk_p1(128, 128, 192, 0); c.repeat = 8; break;
    case 11: name = "timing f16 K/K M128 N80 K128 x8"; kk_p1(128, 80, 128, 0); c.repeat = 8; break;
    case 12: name = "tf32 A MN-major (SBO=128,LBO=4096) B K-major M128 N64 K32"; kk_p1(128, 64, 32, 1);
      c.a_major = 1; c.a_sbo = 128; c.a_lbo = (128 / 4) * 128; c.a_kstep = c.a_lbo; break;
    case 30: name = "f16 K/K 128B-swizzled M128 N256 K64 (k step +32 B)"; kk_p1(128, 256, 64, 0); c.swz = 1;
      c.a_sbo = 1024; c.b_sbo = 1024; c.a_kstep = 32; c.b_kstep = 32; break;
    case 31: name = "f16 K/K 128B-swizzled M128 N128 K64"; kk_p1(128, 128, 64, 0); c.swz = 1;
      c.a_sbo = 1024; c.b_sbo = 1024; c.a_kstep = 32; c.b_kstep = 32; break;
    case 32: name = "timing f16 K/K 128B-swizzled M128 N256 K64 x24"; kk_p1(128, 256, 64, 0); c.swz = 1; c.repeat = 24;
      c.a_sbo = 1024; c.b_sbo = 1024; c.a_kstep = 32; c.b_kstep = 32; break;
    case 33: name = "timing f16 K/K 128B-swizzled M128 N128 K64 x24"; kk_p1(128, 128, 64, 0); c.swz = 1; c.repeat = 24;
      c.a_sbo = 1024; c.b_sbo = 1024; c.a_kstep = 32; c.b_kstep = 32; break;
    default: printf("no such test\n"); return 2;
  }
  const int Mrows = 128;  // image always sized for 128 rows
  auto span = [&](int major, int rows, uint32_t lbo, uint32_t sbo) {
    uint32_t mx = 0;
    for (int r = 0; r < rows; ++r)
      for (int k = 0; k < c.K; ++k) mx = std::max(mx, off(major, c.kind, r, k, lbo, sbo));
    return ((mx + 16 + 15) / 16) * 16;
  };
  c.a_bytes = c.swz ? Mrows * 128 : span(c.a_major, Mrows, c.a_lbo, c.a_sbo);
  c.b_bytes = c.swz ? c.N * 128 : span(c.b_major, c.N, c.b_lbo, c.b_sbo);
  std::vector<uint8_t> a_img(c.a_bytes, 0), b_img(c.b_bytes, 0);
  std::vector<float> A(Mrows * c.K), B(c.N * c.K);
  std::vector<__half> a_rm(Mrows * c.K);
  srand(1234 + t);
  for (int r = 0; r < Mrows; ++r)
    for (int k = 0; k < c.K; ++k) {
      float v = (r < c.M) ? (float)(rand() % 5 - 2) : 0.f;
      A[r * c.K + k] = v;
      a_rm[r * c.K + k] = __float2half(v);
      put(a_img, c.swz ? off_sw128(r, k) : off(c.a_major, c.kind, r, k, c.a_lbo, c.a_sbo), c.kind, v);
    }
  for (int n = 0; n < c.N; ++n)
    for (int k = 0; k < c.K; ++k) {
      float v = (float)(rand() % 5 - 2);
      B[n * c.K + k] = v;
      put(b_img, c.swz ? off_sw128(n, k) : off(c.b_major, c.kind, n, k, c.b_lbo, c.b_sbo), c.kind, v);
    }
  uint8_t *da, *db;
  __half* darm;
  float* dD;
  int* dstat;
  long long* dcyc;
  cudaMalloc(&da, c.a_bytes); cudaMalloc(&db, c.b_bytes); cudaMalloc(&darm, a_rm.size() * 2);
  cudaMalloc(&dD, 128 * c.N * 4); cudaMalloc(&dstat, 4); cudaMalloc(&dcyc, 8);
  cudaMemcpy(da, a_img.data(), c.a_bytes, cudaMemcpyHostToDevice);
  cudaMemcpy(db, b_img.data(), c.b_bytes, cudaMemcpyHostToDevice);
  cudaMemcpy(darm, a_rm.data(), a_rm.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0xff, 128 * c.N * 4);
  const size_t smem = ((c.a_bytes + 1023) / 1024) * 1024 + c.b_bytes + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe<<<1, 128, smem>>>(da, db, darm, dD, c, dstat, dcyc);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("[test %d] %s: CUDA ERROR %s\n", t, name, cudaGetErrorString(e));
    return 1;
  }
  int stat;
  long long cyc;
  std::vector<float> D(128 * c.N);
  cudaMemcpy(&stat, dstat, 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  if (stat) {
    printf("[test %d] %s: TIMEOUT waiting for MMA commit\n", t, name);
    return 1;
  }
  // expected
  std::vector<float> R(c.M * c.N);
  for (int m = 0; m < c.M; ++m)
    for (int n = 0; n < c.N; ++n) {
      float s = 0;
      for (int k = 0; k < c.K; ++k) s += A[m * c.K + k] * B[n * c.K + k];
      R[m * c.N + n] = s * c.repeat;
    }
  int bad = 0;
  double maxerr = 0;
  if (c.M == 128) {
    for (int i = 0; i < c.M * c.N; ++i) {
      double d = fabs((double)D[i] - R[i]);
      maxerr = d > maxerr ? d : maxerr;
      bad += d > 1e-3;
    }
    printf("[test %d] %s: %s  mismatches=%d/%d maxerr=%g cycles=%lld\n", t, name, bad ? "FAIL" : "PASS", bad,
           c.M * c.N, maxerr, cyc);
  } else {
    // find for every expected row which TMEM lane holds it
    printf("[test %d] %s: row->lane map:", t, name);
    for (int m = 0; m < c.M; ++m) {
      int found = -1;
      for (int l = 0; l < 128 && found < 0; ++l) {
        bool eq = true;
        for (int n = 0; n < c.N && eq; ++n) eq = fabs(D[l * c.N + n] - R[m * c.N + n]) < 1e-3;
        if (eq) found = l;
      }
      if (m % 16 == 0) printf(" [%d]->%d", m, found);
      bad += found < 0;
    }
    printf("  unmatched rows=%d cycles=%lld\n", bad, cyc);
  }
  return bad ? 1 : 0;
}
