// Sequential weighted EM for sm_100a, "V-resident" kernel: the BASELINE shape family Ck = 64, L = 64 / 128 bases per side,
// Cv = 512 (every other covered shape runs em_pair_kernel, fused_em.cu).  Reference semantics: methods/SWEM/modules.py
// :129-168 (swem), :112-120 (E-step), :122-127 (M-step), :93-110 (W-step), :164-165 (nu).
//
// One launch per memorize call, one CTA of 16 warps per (unit u = (b, n), 128-pixel tile, side), the two sides of a tile
// forming a 2-CTA cluster.  What changed against em_pair_kernel (profiles/r1_phases_em_pair.txt -> r2_phases_em_res.txt):
//
//   * nu = Z^T V^T keeps the three split products (z_hi v_hi + z_hi v_lo + z_lo v_hi): a single-pass fp16 product leaves nu at
//     4e-4 of the reference -- inside the 1e-2 feature bar, but measured to double the mask disagreement of free-running
//     sequences and to push four mask tests over the 99.9 % gate (profiles/r2_single_pass_nu.txt) -- so precision is not traded.
//     The operand images of V (fp16 hi/lo, 8 x 32 KB per tile) are converted once per pair -- each CTA its channel half, in two
//     rounds of one warp step per warp placed where they do not delay latency-critical traffic -- into an L2-resident scratch
//     and streamed back through a 3-stage bulk-copy ring; the first three stages are filled as soon as the images exist.
//     (Converting straight into the ring with the warps as producers was measured at 12.7 us for the nu GEMM against 4.8 us:
//     one warp step of loads in flight per warp cannot cover the L2 latency; profiles/r2_em_res_phases.txt.)
//   * 16 warps: the softmax epilogue holds 32 logits per thread (4 threads per pixel, exps against the thread's own max and
//     one rescale -- no max exchange before the exps); the pixel statistics of the two sides meet through a remote mbarrier
//     (128 arrivals) instead of a full cluster barrier.
//   * M-step partials: all 16 warps reduce-add their 16 columns straight from TMEM; the prior term zita_ * kappa_ is added
//     once, by the CTA of tile 0, so the finalize is kappa = total / zita_total with no prior reloads, done by 4 threads per
//     row (16 channels each, norm through two shuffles).
//   * the kappa all-reduce of the last iteration completes under the nu GEMMs; the nu partial of the CTA's own side drains
//     while the tensor core is still busy with the peer side.
//
// Cross-tile sums still go through L2-resident accumulators with per-(unit, iteration, side) arrival counters (bounded
// spin; every CTA of a launch is co-resident, see the host side).
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc05.cuh"
#include "fused_common.cuh"

namespace swem {

using namespace tc05;

namespace emr {
constexpr int kTP = 128;      // pixels per CTA
constexpr int kL = 128;       // basis rows per CTA (L = 64: the upper half is zero padding)
constexpr int kCk = 64;
constexpr int kCv = 512;
constexpr int kDH = 256;      // value channels per CTA in the nu phase
constexpr int kThreads = 512;
constexpr float kKScale = 256.f;
constexpr float kZScale = 16384.f;

// ---- shared memory map (bytes).  Operand layouts as in fused_em.cu:
// XH : [c 0..79][p]  byte = (c%8)*16 + (c/8)*2048 + (p/8)*128 + (p%8)*2   (row 64 = ones -> zita column, 65.. = 0)
// XL : [c 0..63][p]
// KH/KL : [l][c] K-major: byte = (l%8)*16 + (l/8)*128 + (c/8)*2048 + (c%8)*2
// Z / ZL : [l][p] MN-major A: byte = (l%8)*2 + (p%8)*16 + (l/8)*2048 + (p/8)*128   (ZL aliases KH/KL)
// VS : ring of 3 V operand images [d 0..255][p 0..31] K-major B: byte = (d%8)*16 + (d/8)*128 + (p/8)*4096 + (p%8)*2, hi plane (16 KB) then lo
constexpr uint32_t kOffXH = 0;
constexpr uint32_t kOffXL = kOffXH + 10 * 2048;
constexpr uint32_t kOffKH = kOffXL + 8 * 2048;
constexpr uint32_t kOffKL = kOffKH + 8 * 2048;
constexpr uint32_t kOffZ = kOffKL + 8 * 2048;
constexpr uint32_t kOffZL = kOffKH;
constexpr uint32_t kOffVS = kOffZ + 16 * 2048;
constexpr int kStages = 3;
constexpr uint32_t kImgBytes = 32768;          // one V operand image: [256 d][32 px] fp16, hi plane then lo plane
constexpr uint32_t kVPlane = 16384;
constexpr int kImages = 8;                     // per tile: 2 channel halves x 4 pixel quarters
constexpr uint32_t kOffMisc = kOffVS + kStages * kImgBytes;
static_assert(kOffMisc >= 5 * 32768, "nu drain staging: five 32 KB buffers below the Misc block");

struct Misc {
  float inv_nx[kTP];
  float mask[kTP];
  float hmax[4][kTP];
  float hsum[4][kTP];
  float hew[4][kTP];
  float2 mbox[2][kTP];          // [iteration parity][pixel] = (max of the PEER side's logits, its W-step exp sum), written by the peer
  float rz[kL];
  float zp[kL];
  __align__(16) float zrow[kL];  // staged zita column of the M-step partial
  uint64_t bar_mma;
  uint64_t bar_nu[2];
  uint64_t bar_w;
  uint64_t bar_full[kStages];
  uint64_t bar_empty[kStages];
  uint64_t bar_vready;          // the peer's images are complete (remote arrival)
  uint32_t tmem_base;
  int abort_flag;
};
constexpr uint32_t kSmemBytes = kOffMisc + sizeof(Misc) + 128;
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");

constexpr uint32_t kColE = 0;      // [128 px][128]   logits of this side
                                   // [128 l][80]     M-step sums at column 128
                                   // nu: [128 l][512 d] at columns 0..511 (after the last M-step partial has been read out)
}  // namespace emr

struct EmResParams {
  const float* x;
  const float* v;
  const float* masks;
  const float* kappa_prior;
  const float* nu_prior;
  const float* zita_prior;
  float* kappa;
  float* nu;
  float* zita;
  float* z_last;
  uint8_t* vblob;        // [U][T][8][32 KB] scratch: operand images of V
  uint8_t* img_k;        // optional: the readout's khat / nu operand images of the OUTPUT bases (SwemEmArgs.image_workspace), bank
  uint8_t* img_v;        // img_bank of a memory of img_Lt / L banks; layouts of readout_prep_kernel (fused_readout.cu)
  int img_bank, img_Lt;
  float* acc_k;          // [U][n_iters][2][65][128]   (zeroed by the kernel itself, see "accumulators" below)
  float* acc_nu;         // [U][2][512][128]
  unsigned* counters;    // library-owned, zero between launches: [U][n_iters][2] M-step arrivals, then [U] accumulators zeroed,
                         // [U] nu drained, [U] departed (the last CTA of a unit to leave clears the unit's words again)
  int* status;
  long long* prof;
  int N, HW, T, n_iters, u0, L;
  int dbg;               // measurement switches (SWEM_EM_DBG): 4 = gpu-scope fence between the completed bulk reduction and the arrival
  int U;                 // units of the whole call (the nu counters follow the U * n_iters * 2 M-step counters)
  float c1s;             // log2(e) / (tau * kKScale)
};

// labelled time stamps of CTA 0: prof[0] = -count, prof[1 + k] = (ns << 8) | label
#define EMR_STAMP(id)                                                                                               \
  do {                                                                                                              \
    if (p.prof != nullptr && blockIdx.x == 0 && tid == 0 && n_stamp < 120) p.prof[1 + n_stamp++] = (global_ns() << 8) | (id); \
  } while (0)

__device__ __forceinline__ void st_cluster_u4(uint32_t addr, const uint4& v) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// remote store + transaction-count arrival on the destination CTA's mbarrier (no fence on either side: the data is visible
// to whoever observes the phase completion)
__device__ __forceinline__ void st_async_f2(uint32_t addr, float a, float b, uint32_t bar) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];" ::"r"(addr), "f"(a), "f"(b), "r"(bar) : "memory");
}
__device__ __forceinline__ void st_async_u4(uint32_t addr, const uint4& v, uint32_t bar) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(addr), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w), "r"(bar)
               : "memory");
}
__device__ __forceinline__ bool warp_wait(uint64_t* bar, uint32_t parity, int) {
  int ok = 1;
  if ((threadIdx.x & 31) == 0) ok = tc05::mbar_wait(bar, parity) ? 1 : 0;
  return __shfl_sync(0xffffffffu, ok, 0) != 0;
}
// cross-tile wait: one thread polls the arrival counter back to back (bounded)
__device__ __forceinline__ bool wait_counter_fast(const unsigned* counter, unsigned target) {
#pragma unroll 1
  for (unsigned i = 0; i < (1u << 22); ++i)
    if (ld_acquire_u32(counter) >= target) return true;
  return false;
}
__device__ __forceinline__ bool warp_wait_cluster(uint64_t* bar, uint32_t parity) {
  int ok = 1;
  if ((threadIdx.x & 31) == 0) ok = mbar_wait_cluster(bar, parity) ? 1 : 0;
  return __shfl_sync(0xffffffffu, ok, 0) != 0;
}

template <bool VPM>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(emr::kThreads, 1) em_res_kernel(const EmResParams p) {
  using namespace emr;
  extern __shared__ __align__(1024) uint8_t smem[];
  Misc& ms = *reinterpret_cast<Misc*>(smem + kOffMisc);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rank = (int)cluster_ctarank();
  const int sd = rank;                                  // side of the E / M steps (0 = background, 1 = foreground); value-channel half of the nu phase
  const int pair = blockIdx.x >> 1;
  const int tile = pair % p.T;
  const int u = p.u0 + pair / p.T;
  const int b = u / p.N;
  const int p0 = tile * kTP;
  const int HW = p.HW, I = p.n_iters, L = p.L;
  const uint32_t sbase = smem_u32(smem);
  int n_stamp = 0;
  EMR_STAMP(0);

  // thread roles
  const int q = warp & 3, cb = warp >> 2;               // TMEM lane quadrant, column block
  const int px = q * 32 + lane;                         // epilogue: pixel (32 logits columns [32 cb, +32)); reduce-add / finalize: row l (16 columns [16 cb, +16))
  const int gs = u * 2 + sd;

  // ---- every latency-critical global load of the set-up is issued first (prior kappa / zita, the pixel norms' channels, the X
  // tile, the mask): ~1 us of L2 / HBM latency that then runs under the TMEM allocation, the barrier set-up and the clearing of
  // the accumulators instead of after them (profiles/r2_phases.txt: set-up 5.5 -> 3.x us)
  const bool valid_row = px < L;
  float kap0[16], nrm[16];
  float4 xa[2], xb[2];
  const float zita_prior_f = valid_row ? __ldg(p.zita_prior + (size_t)gs * L + px) : 0.f;
  {
    const float* kprior = p.kappa_prior + ((size_t)gs * kCk + cb * 16) * L + (valid_row ? px : 0);
#pragma unroll
    for (int c = 0; c < 16; ++c) kap0[c] = valid_row ? __ldg(kprior + (size_t)c * L) : 0.f;
  }
  const int npq = tid & 127, ncq = tid >> 7;             // pixel norms: 4 threads per pixel, 16 channels each
  {
    const int pp = p0 + npq;
    const float* xp = p.x + ((size_t)b * kCk + ncq * 16) * HW + pp;
#pragma unroll
    for (int c = 0; c < 16; ++c) nrm[c] = pp < HW ? __ldg(xp + (size_t)c * HW) : 0.f;
  }
  const float mask_px = (ncq == 0 && p0 + npq < HW) ? __ldg(p.masks + (size_t)gs * HW + p0 + npq) : 0.f;
  const bool x_vec = (HW & 3) == 0;
  {
    const float* xrow = p.x + ((size_t)b * kCk + (tid >> 3)) * HW;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int pp0 = p0 + ((tid & 7) * 2 + j) * 8;
      if (x_vec && pp0 + 7 < HW) {
        xa[j] = __ldg(reinterpret_cast<const float4*>(xrow + pp0));
        xb[j] = __ldg(reinterpret_cast<const float4*>(xrow + pp0) + 1);
      } else {
        float t[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) t[e] = (pp0 + e < HW) ? __ldg(xrow + pp0 + e) : 0.f;
        xa[j] = make_float4(t[0], t[1], t[2], t[3]);
        xb[j] = make_float4(t[4], t[5], t[6], t[7]);
      }
    }
  }

  EMR_STAMP(35);                                        // set-up loads issued
  if (warp == 0) tmem_alloc(&ms.tmem_base, 512);
  EMR_STAMP(36);                                        // TMEM allocated
  if (tid == 0) {
    mbar_init(&ms.bar_mma, 1);
    mbar_init(&ms.bar_nu[0], 1);
    mbar_init(&ms.bar_nu[1], 1);
    mbar_init(&ms.bar_w, 1);                           // + transaction bytes: the peer's st.async stores
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&ms.bar_full[i], 1);
      mbar_init(&ms.bar_empty[i], 1);
    }
    mbar_init(&ms.bar_vready, 1);
    ms.abort_flag = 0;
    fence_mbar_init();
  }
  cluster_arrive();                                     // (waited for below: the peer is running and its barriers exist before any remote access)

  // ---- accumulators: every CTA clears its slice of the unit's L2 accumulators (the rows of acc_k it would finalize for every
  // iteration of its side, the value channels of acc_nu it finalizes at the end) with bulk stores from a zeroed shared-memory
  // buffer, waits for their completion at the end of the set-up and arrives on the unit's "zeroed" counter; nobody reduce-adds
  // before all 2 T CTAs of the unit have arrived.  No memset launch in front of the kernel.
  const int rows_per = (kCk + 1 + p.T - 1) / p.T;       // acc_k rows [c][128] per tile
  const int dper = (kCv + p.T - 1) / p.T;               // acc_nu value channels per tile
  const int d0 = tile * dper, d1 = min(kCv, d0 + dper);
  {
    constexpr uint32_t kZeroBytes = 16384;
    for (int i = tid; i < (int)(kZeroBytes / 16); i += kThreads) reinterpret_cast<uint4*>(smem + kOffZ)[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async_smem();
    __syncthreads();
    EMR_STAMP(37);                                      // zero buffer ready (first CTA-wide barrier)
    if (tid == 0) {
      auto zero_range = [&](float* dst, uint32_t bytes) {
        for (uint32_t o = 0; o < bytes; o += kZeroBytes) {
          const uint32_t n = min(kZeroBytes, bytes - o);
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(reinterpret_cast<uint8_t*>(dst) + o),
                       "r"(sbase + kOffZ), "r"(n)
                       : "memory");
        }
      };
      const int r0 = tile * rows_per, r1 = min(kCk + 1, r0 + rows_per);
      if (r1 > r0)
        for (int it = 0; it < I; ++it)
          zero_range(p.acc_k + ((size_t)((u * I + it) * 2 + sd)) * ((kCk + 1) * kL) + (size_t)r0 * kL, (uint32_t)(r1 - r0) * kL * 4);
      if (d1 > d0) zero_range(p.acc_nu + ((size_t)gs * kCv + d0) * kL, (uint32_t)(d1 - d0) * kL * 4);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      EMR_STAMP(30);                                    // clearing issued
    }
  }

  // ---- V -> fp16 hi/lo operand images in the L2-resident scratch.  Image k = (channel half k / 4, pixel quarter k % 4) takes 8 warp
  // steps (w8 = 0 .. 7); this CTA converts its channel half, images 4 rank .. 4 rank + 3, in two rounds of 16 warp steps.  Load phase (8 x 16 bytes per lane in flight) and
  // convert / store phase are separate calls so that the loads overlap whatever lies between them.
  float4 vf[8];
  auto v_load = [&](int k, int w8) {
    const int h = k >> 2, qd = k & 3;
    if constexpr (VPM) {                                // v [U][HW][512]: 8 pixels x 128 channels per step, the lane owns 4 channels
      const int g = w8 >> 1, d = (w8 & 1) * 128 + lane * 4;
      const float* src = p.v + ((size_t)u * HW) * kCv + h * 256 + d;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int pp = p0 + qd * 32 + g * 8 + e;
        vf[e] = pp < HW ? __ldg(reinterpret_cast<const float4*>(src + (size_t)pp * kCv)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    } else {                                            // v [U][512][HW]: 32 channels x 32 pixels per step, the lane owns 4 x (1 channel, 8 pixels)
      const int pp0 = p0 + qd * 32 + (lane & 3) * 8;
      const bool vec = ((HW & 3) == 0) && (pp0 + 7 < HW);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int d = w8 * 32 + j * 8 + (lane >> 2);
        const float* src = p.v + ((size_t)u * kCv + h * 256 + d) * HW + pp0;
        if (vec) {
          vf[2 * j] = __ldg(reinterpret_cast<const float4*>(src));
          vf[2 * j + 1] = __ldg(reinterpret_cast<const float4*>(src) + 1);
        } else {
          float t[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) t[e] = (pp0 + e < HW) ? __ldg(src + e) : 0.f;
          vf[2 * j] = make_float4(t[0], t[1], t[2], t[3]);
          vf[2 * j + 1] = make_float4(t[4], t[5], t[6], t[7]);
        }
      }
    }
  };
  uint8_t* const vimg = p.vblob + ((size_t)u * p.T + tile) * kImages * kImgBytes;
  auto v_store = [&](int k, int w8) {
    uint8_t* img = vimg + (size_t)k * kImgBytes;
    if constexpr (VPM) {
      const int g = w8 >> 1, d = (w8 & 1) * 128 + lane * 4;
      uint8_t* dst = img + d * 16 + g * 4096;           // (d % 8) * 16 + (d / 8) * 128 = d * 16
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        __align__(16) __half hi[8];
        __align__(16) __half lo[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) split_half(c == 0 ? vf[e].x : c == 1 ? vf[e].y : c == 2 ? vf[e].z : vf[e].w, hi[e], lo[e]);
        *reinterpret_cast<uint4*>(dst + c * 16) = *reinterpret_cast<uint4*>(hi);
        *reinterpret_cast<uint4*>(dst + kVPlane + c * 16) = *reinterpret_cast<uint4*>(lo);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int d = w8 * 32 + j * 8 + (lane >> 2);
        const float t[8] = {vf[2 * j].x, vf[2 * j].y, vf[2 * j].z, vf[2 * j].w, vf[2 * j + 1].x, vf[2 * j + 1].y, vf[2 * j + 1].z, vf[2 * j + 1].w};
        __align__(16) __half hi[8];
        __align__(16) __half lo[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) split_half(t[e], hi[e], lo[e]);
        uint8_t* dst = img + d * 16 + (lane & 3) * 4096;
        *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<uint4*>(hi);
        *reinterpret_cast<uint4*>(dst + kVPlane) = *reinterpret_cast<uint4*>(lo);
      }
    }
  };
  auto v_publish = [&]() {                              // generic-proxy global stores -> visible to the bulk copies (async proxy) of the pair
    __threadfence();
    asm volatile("fence.proxy.async;" ::: "memory");
  };
  auto load_image = [&](int seq) {                      // ring stage seq % 3 <- the seq-th image in consumption order (own channel half first)
    const int k = (((seq >> 2) ^ rank) << 2) + (seq & 3);
    mbar_expect_tx(&ms.bar_full[seq % kStages], kImgBytes);
    bulk_g2s(smem + kOffVS + (seq % kStages) * kImgBytes, vimg + (size_t)k * kImgBytes, kImgBytes, &ms.bar_full[seq % kStages]);
  };
  const int vround_img = rank * 4 + (warp >> 3), vround_w8 = warp & 7;   // round r converts image vround_img + 2 r
  // Round 0 is loaded at the end of the set-up and stored after the first logits GEMM, round 1 is loaded under the first M-step GEMM
  // and stored while the first cross-tile reduction is in flight -- never ahead of the X / prior loads (they delay the first logits
  // GEMM by 3 us) nor next to a cross-tile reduction (they delay its completion by 1 us; profiles/r2_em_res_phases.txt).
  // ---- khat = l2norm(kappa) * 256 -> fp16 hi/lo K-major rows (reference :115).  Thread <-> (row px, channels [16 cb, +16)):
  // global accesses are coalesced over the rows, the squared norm meets in shared memory across the 4 warps of a lane quadrant
  auto stage_khat = [&](const float (&kap)[16]) {
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < 16; ++c) ss = fmaf(kap[c], kap[c], ss);
    ms.hsum[cb][px] = ss;
    bar_sync(9 + q, 128);
    ss = (ms.hsum[0][px] + ms.hsum[1][px]) + (ms.hsum[2][px] + ms.hsum[3][px]);
    const float sc = kKScale / (sqrtf(ss) + kEpsNorm);
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      __align__(16) __half hi[8];
      __align__(16) __half lo[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) split_half(kap[g * 8 + e] * sc, hi[e], lo[e]);
      const uint32_t off = (px % 8) * 16 + (px / 8) * 128 + (cb * 2 + g) * 2048;
      *reinterpret_cast<uint4*>(smem + kOffKH + off) = *reinterpret_cast<uint4*>(hi);
      *reinterpret_cast<uint4*>(smem + kOffKL + off) = *reinterpret_cast<uint4*>(lo);
    }
  };
  // prior term of the M-step, added once per (unit, side) by the CTA of tile 0: zita_ * kappa_ (and zita_ itself), scaled like
  // the tensor-core sums (same thread mapping as the reduce-add)
  float pri[16], pri_z = 0.f;
  {
    const float zp = (tile == 0) ? zita_prior_f * kZScale : 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) pri[j] = zp * kap0[j];
    pri_z = zp;
    stage_khat(kap0);
  }
  EMR_STAMP(31);                                       // prior khat staged
  // pixel norms (4 threads per pixel, 16 channels each) + this side's mask
  {
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < 16; ++c) ss = fmaf(nrm[c], nrm[c], ss);
    ms.hew[ncq][npq] = ss;
    if (ncq == 0) ms.mask[npq] = mask_px;
  }
  // X tile -> fp16 hi/lo chunks: thread -> (channel c = tid / 8, 2 pixel groups of 8)
  {
    const int c = tid >> 3;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int pg = (tid & 7) * 2 + j;
      const float t[8] = {xa[j].x, xa[j].y, xa[j].z, xa[j].w, xb[j].x, xb[j].y, xb[j].z, xb[j].w};
      __align__(16) __half hi[8];
      __align__(16) __half lo[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) split_half(t[e], hi[e], lo[e]);
      const uint32_t off = (c % 8) * 16 + (c / 8) * 2048 + pg * 128;
      *reinterpret_cast<uint4*>(smem + kOffXH + off) = *reinterpret_cast<uint4*>(hi);
      *reinterpret_cast<uint4*>(smem + kOffXL + off) = *reinterpret_cast<uint4*>(lo);
    }
  }
  if (tid < 256) {                    // augmented rows 64..79 of XH: row 64 = 1 (-> zita), rest 0
    const int r = kCk + (tid >> 4), pg = tid & 15;
    const __half one = __float2half_rn(r == kCk ? 1.f : 0.f);
    __align__(16) __half vals[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) vals[e] = one;
    *reinterpret_cast<uint4*>(smem + kOffXH + (r % 8) * 16 + (r / 8) * 2048 + pg * 128) = *reinterpret_cast<uint4*>(vals);
  }
  EMR_STAMP(32);                                       // X staged
  v_load(vround_img, vround_w8);
  unsigned* const cnt_m = p.counters + (size_t)u * I * 2 + sd;            // + 2 it
  unsigned* const cnt_zero = p.counters + (size_t)p.U * I * 2 + u;
  unsigned* const cnt_nu = cnt_zero + p.U;
  unsigned* const cnt_dep = cnt_nu + p.U;
  if (tid == 0) {
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");             // this CTA's slices are zero at L2
    atomicAdd(cnt_zero, 1u);                                              // (waited for under the first logits GEMM)
    EMR_STAMP(34);                                                        // clearing complete + arrived
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (tid < kTP) ms.inv_nx[tid] = 1.f / (sqrtf(ms.hew[0][tid] + ms.hew[1][tid] + ms.hew[2][tid] + ms.hew[3][tid]) + kEpsNorm);
  cluster_wait();
  const uint32_t tmem = ms.tmem_base;
  uint32_t ph_mma = 0, ph_w = 0;
  bool failed = false;
  EMR_STAMP(1);                      // set-up done

  const uint32_t idesc_e = make_idesc(128, 128, kFmtF16, kFmtF16, kMajorMN, kMajorK);
  const uint32_t idesc_mhi = make_idesc(128, kCk + 16, kFmtF16, kFmtF16, kMajorMN, kMajorK);
  const uint32_t idesc_mlo = make_idesc(128, kCk, kFmtF16, kFmtF16, kMajorMN, kMajorK);
  const uint32_t idesc_nu = make_idesc(128, 256, kFmtF16, kFmtF16, kMajorMN, kMajorK);
  const uint32_t peer = (uint32_t)(rank ^ 1);
  const uint32_t peer_mbox = map_to_peer(smem_u32(&ms.mbox[0][0]), peer);
  const uint32_t peer_bar_w = map_to_peer(smem_u32(&ms.bar_w), peer);
  const uint32_t peer_bar_vready = map_to_peer(smem_u32(&ms.bar_vready), peer);
  constexpr uint32_t col_m = 128;
  const bool mma_thread = (tid == 32);                  // (thread 0 runs the cross-tile barriers, which block in fences)
  constexpr float kInvZ = 1.f / kZScale;

  // finalize from the completed totals of iteration `it`: kappa = total / zita_total (the prior is part of the totals)
  auto finalize = [&](const float* acc, bool last) {
    const float zt = __ldcg(acc + kCk * kL + px);
    float kap[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) kap[c] = __ldcg(acc + (cb * 16 + c) * kL + px);
    const float rz = valid_row ? 1.f / zt : 0.f;
#pragma unroll
    for (int c = 0; c < 16; ++c) kap[c] *= rz;
    EMR_STAMP(23);                                      // totals loaded
    if (last) {
      if (cb == 0) {
        ms.rz[px] = valid_row ? kZScale * rz : 0.f;     // 1 / zita in true units
        ms.zp[px] = zita_prior_f;
      }
      if (tile == 0 && valid_row) {
        if (cb == 0) p.zita[(size_t)gs * L + px] = zt * kInvZ;
        float* kout = p.kappa + ((size_t)gs * kCk + cb * 16) * L + px;
#pragma unroll
        for (int c = 0; c < 16; ++c) kout[(size_t)c * L] = kap[c];
      }
      if (tile == 0 && p.img_k != nullptr) {
        // the readout's khat operand of these bases: l2norm(kappa) * 256 as fp16 hi/lo, K-major rows j = bank L + l of the
        // [Lt rows][Ck] block of (u, side) (readout_prep_kernel's layout), straight from the registers that hold kappa
        float ss = 0.f;
#pragma unroll
        for (int c = 0; c < 16; ++c) ss = fmaf(kap[c], kap[c], ss);
        ms.hsum[cb][px] = ss;
        bar_sync(9 + q, 128);
        ss = (ms.hsum[0][px] + ms.hsum[1][px]) + (ms.hsum[2][px] + ms.hsum[3][px]);
        if (valid_row) {
          const float sc = kKScale / (sqrtf(ss) + kEpsNorm);
          const int R = p.img_Lt;                         // (Lt <= 256: one block per side)
          const uint32_t plane = (uint32_t)R * kCk * 2, lbo = (uint32_t)R * 16;
          const int j = p.img_bank * L + px;
          uint8_t* base = p.img_k + (size_t)gs * 2 * plane + (j % 8) * 16 + (j / 8) * 128;
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            __align__(16) __half hi[8];
            __align__(16) __half lo[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) split_half(kap[g * 8 + e] * sc, hi[e], lo[e]);
            *reinterpret_cast<uint4*>(base + (cb * 2 + g) * lbo) = *reinterpret_cast<uint4*>(hi);
            *reinterpret_cast<uint4*>(base + plane + (cb * 2 + g) * lbo) = *reinterpret_cast<uint4*>(lo);
          }
        }
      }
    } else {
      stage_khat(kap);
    }
  };

  // column block leader (warp 4 cb, lane 0): reduce-add the block's 8 KB of the staged M-step partial (block 0: and the zita
  // row), wait for the reduction to complete and arrive.  Release: every byte goes through the bulk-copy engine and the arrival
  // is issued only after `wait_group 0` has reported the reductions performed at L2, the point of coherence of the polling
  // CTAs (ld.acquire.gpu + ld.global.cg); a gpu-scope fence in between costs a further L2 round trip (0.7 us, measured:
  // profiles/r2_em_res_fence_ab.txt) and can be switched on for comparison with SWEM_EM_DBG=4.
  auto reduce_issue = [&](float* acc, uint32_t stage_off) {
    bar_sync(5 + cb, 128);                              // the 4 warps of this column block have staged their rows
    if (q == 0 && lane == 0) {
      asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(acc + cb * 16 * kL),
                   "r"(sbase + stage_off + (uint32_t)(cb * 16 * kL * 4)), "r"(16 * kL * 4)
                   : "memory");
      if (cb == 0)
        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(acc + kCk * kL),
                     "r"(smem_u32(&ms.zrow[0])), "r"(kL * 4)
                     : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      EMR_STAMP(20);                                    // block staged (barrier) + bulk reduce issued
    }
  };
  auto reduce_arrive = [&](unsigned* counter) {
    if (q == 0 && lane == 0) {
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      EMR_STAMP(21);                                    // bulk reduce complete
      if (p.dbg & 4) asm volatile("fence.acq_rel.gpu;" ::: "memory");
      atomicAdd(counter, 1u);
      EMR_STAMP(22);                                    // arrived
    }
    __syncwarp();
  };

  float4 nu_pre[3];
  for (int it = 0; it < I; ++it) {
    const bool last = (it == I - 1);
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();

    // ---- (1) logits of this side: a[p, l] = x_p . khat_l (hi/lo split, 3 products) ----------------------------------
    if (mma_thread) {
#pragma unroll
      for (int term = 0; term < 3; ++term) {
        const uint32_t xa = sbase + (term == 2 ? kOffXL : kOffXH);
        const uint32_t kb = sbase + (term == 1 ? kOffKL : kOffKH);
#pragma unroll
        for (int kk = 0; kk < kCk / 16; ++kk) {
          const uint64_t ad = make_sdesc(xa + kk * 2 * 2048, /*lbo*/ 2048, /*sbo*/ 128);
          const uint64_t bd = make_sdesc(kb + kk * 2 * 2048, /*lbo*/ 2048, /*sbo*/ 128);
          mma_f16_ss(tmem + kColE, ad, bd, idesc_e, (term | kk) ? 1u : 0u);
        }
      }
      mma_commit(&ms.bar_mma);
    }
    if (it == 0 && tid == 0) {       // every CTA of the unit has cleared its accumulator slices (polled under the GEMM; the CTAs
      if (!wait_counter_fast(cnt_zero, 2u * (unsigned)p.T)) ms.abort_flag = 1;   // of a launch start together: no wait in practice)
    }
    SWEM_CTA_WAIT(&ms.bar_mma, ph_mma, ms.abort_flag);
    ph_mma ^= 1;
    tc_fence_after_sync();
    EMR_STAMP(2);                    // logits GEMM done
    if (it == 0) {
      v_store(vround_img, vround_w8);
      if (I == 1) {                                     // (I = 1: round 1 cannot hide anywhere)
        v_load(vround_img + 2, vround_w8);
        v_store(vround_img + 2, vround_w8);
        v_publish();
      }
    }

    // ---- (2) epilogue: thread <-> (pixel px, columns [32 cb, +32) of this side's bases) -------------------------------
    {
      const bool do_w = it > 0;
      const bool active = cb * 32 < L;                  // L = 64: the upper half of the columns is padding
      if (tid == 0) {                                   // arm the receive barriers of this iteration
        if (do_w) mbar_expect_tx(&ms.bar_w, kTP * sizeof(float2));
      }
      float a[32];
      {
        uint32_t r[32];
        tmem_ld32(tmem_addr(tmem, q * 32, kColE + cb * 32), r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) a[j] = __uint_as_float(r[j]);
      }
      // exps against the thread's own max; the blocks of a pixel are rescaled onto the side max afterwards:
      // exp(t - M) = exp(t - m) * exp(m - M).  W-step (reference :93-110): same logits times 1 / ||x_p||.
      const float cw = ms.inv_nx[px] * p.c1s;
      float mloc = -3.0e38f, se = 0.f, ew = 0.f, fix_e = 1.f;
      if (active) {
#pragma unroll
        for (int j = 0; j < 32; ++j) mloc = fmaxf(mloc, a[j]);
        // one FFMA + one MUFU per exponential, four independent partial sums
        // (the rounding of the offset -m c, up to 2^-24 of an exponent of several hundred, is recovered exactly by a second
        //  FMA and applied to the sums / the final scale as the factor 2^residual)
        float s4[4] = {0.f, 0.f, 0.f, 0.f};
        if (do_w) {
          const float bw = -mloc * cw;
#pragma unroll
          for (int j = 0; j < 32; ++j) s4[j & 3] += fast_exp2(fmaf(a[j], cw, bw));
          ew = ((s4[0] + s4[1]) + (s4[2] + s4[3])) * fast_exp2(fmaf(-mloc, cw, -bw));
          s4[0] = s4[1] = s4[2] = s4[3] = 0.f;
        }
        const float be = -mloc * p.c1s;
        fix_e = fast_exp2(fmaf(-mloc, p.c1s, -be));
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          a[j] = fast_exp2(fmaf(a[j], p.c1s, be));
          s4[j & 3] += a[j];
        }
        se = ((s4[0] + s4[1]) + (s4[2] + s4[3])) * fix_e;
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) a[j] = 0.f;
      }
      ms.hmax[cb][px] = mloc;
      ms.hsum[cb][px] = se;
      ms.hew[cb][px] = ew;
      bar_sync(1 + q, 128);                             // the 4 warps that share this lane quadrant
      float m_side = fmaxf(fmaxf(ms.hmax[0][px], ms.hmax[1][px]), fmaxf(ms.hmax[2][px], ms.hmax[3][px]));
      float S = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) S += ms.hsum[k][px] * fast_exp2((ms.hmax[k][px] - m_side) * p.c1s);
      float w = ms.mask[px];
      if (do_w) {
        float EW = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) EW += ms.hew[k][px] * fast_exp2((ms.hmax[k][px] - m_side) * cw);
        const int par = it & 1;
        if (cb == 0) st_async_f2(peer_mbox + (uint32_t)((par * kTP + px) * sizeof(float2)), m_side, EW, peer_bar_w);
        EMR_STAMP(25);                                  // own statistics sent
        if (!warp_wait(&ms.bar_w, ph_w, 0)) ms.abort_flag = 1;
        EMR_STAMP(26);                                  // peer statistics received
        ph_w ^= 1;
        const float2 o = ms.mbox[par][px];              // the peer side's (max, W-step sum)
        const float gm = fmaxf(m_side, o.x);
        const float e_own = EW * fast_exp2((m_side - gm) * cw), e_peer = o.y * fast_exp2((o.x - gm) * cw);
        w *= 1.f - e_own / (e_own + e_peer);
      }
      const float scale = active ? (w / S) * fast_exp2((mloc - m_side) * p.c1s) * fix_e : 0.f;
      const float zs = scale * kZScale;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        __align__(16) __half hi[8];
        __align__(16) __half lo[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) split_half(a[g * 8 + k] * zs, hi[k], lo[k]);
        const uint32_t off = (px % 8) * 16 + (px / 8) * 128 + (cb * 4 + g) * 2048;
        *reinterpret_cast<uint4*>(smem + kOffZ + off) = *reinterpret_cast<uint4*>(hi);
        *reinterpret_cast<uint4*>(smem + kOffZL + off) = *reinterpret_cast<uint4*>(lo);
      }
      if (p.z_last != nullptr && last && p0 + px < HW && active) {
        float4* dst = reinterpret_cast<float4*>(p.z_last + ((size_t)gs * HW + p0 + px) * L + cb * 32);
#pragma unroll
        for (int g = 0; g < 8; ++g)
          dst[g] = make_float4(a[g * 4] * scale, a[g * 4 + 1] * scale, a[g * 4 + 2] * scale, a[g * 4 + 3] * scale);
      }
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    EMR_STAMP(3);                    // epilogue done
    if (it == (I == 1 ? 0 : 1)) {    // this CTA's images are complete and published: fill the ring, tell the peer
      if (tid == 64) {
        asm volatile("fence.proxy.async;" ::: "memory");
        for (int seq = 0; seq < kStages; ++seq) load_image(seq);
      }
      if (tid == 0) mbar_arrive_remote(peer_bar_vready);
    }

    // ---- (3) M-step GEMM: [sum_p z x | sum_p z] for this side's 128 bases (3 products) ----------------------------------
    if (mma_thread) {
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const uint64_t ad = make_sdesc(sbase + kOffZ + kk * 2 * 128, /*lbo*/ 128, /*sbo*/ 2048);
        const uint64_t al = make_sdesc(sbase + kOffZL + kk * 2 * 128, /*lbo*/ 128, /*sbo*/ 2048);
        const uint64_t bh = make_sdesc(sbase + kOffXH + kk * 2 * 128, /*lbo*/ 128, /*sbo*/ 2048);
        const uint64_t bl = make_sdesc(sbase + kOffXL + kk * 2 * 128, /*lbo*/ 128, /*sbo*/ 2048);
        mma_f16_ss(tmem + col_m, ad, bh, idesc_mhi, kk ? 1u : 0u);   // z_hi x_hi (+ zita column)
        mma_f16_ss(tmem + col_m, ad, bl, idesc_mlo, 1u);             // z_hi x_lo
        mma_f16_ss(tmem + col_m, al, bh, idesc_mhi, 1u);             // z_lo x_hi
      }
      mma_commit(&ms.bar_mma);
    }
    if (it == 0 && I > 1) v_load(vround_img + 2, vround_w8);
    SWEM_CTA_WAIT(&ms.bar_mma, ph_mma, ms.abort_flag);
    ph_mma ^= 1;
    tc_fence_after_sync();
    EMR_STAMP(4);                    // M GEMM done

    // ---- (4) partial of this tile -> fp32 reductions straight from TMEM into the L2-resident accumulator [c][l] ----------
    float* acc = p.acc_k + ((size_t)((u * I + it) * 2 + sd)) * ((kCk + 1) * kL);
    // (staged in the khat / z_lo region -- dead between the M-step GEMM and the finalize -- and reduce-added by the bulk-copy
    //  engine in four 8 KB pieces, one per column block: far fewer L2 atomic transactions than per-lane reductions)
    unsigned* counter = cnt_m + 2 * it;
    {
      uint32_t r[16];
      tmem_ld16(tmem_addr(tmem, q * 32, col_m + cb * 16), r);
      uint32_t rz16[16];
      if (cb == 0) tmem_ld16(tmem_addr(tmem, q * 32, col_m + kCk), rz16);
      tmem_ld_wait();
      float* ns = reinterpret_cast<float*>(smem + (last ? kOffXH : kOffKH));   // [64 c][128 l] fp32 (last iteration: X is dead, z_lo is not)
#pragma unroll
      for (int j = 0; j < 16; ++j) ns[(cb * 16 + j) * kL + px] = __uint_as_float(r[j]) + pri[j];
      if (cb == 0) ms.zrow[px] = __uint_as_float(rz16[0]) + pri_z;
      fence_proxy_async_smem();
    }
    if (!last) {
      tc_fence_before_sync();
      EMR_STAMP(5);                                     // partial staged
      reduce_issue(acc, kOffKH);
      reduce_arrive(counter);
      EMR_STAMP(6);                                     // reduce-added + arrived
      if (it == 0) {
        v_store(vround_img + 2, vround_w8);
        v_publish();
      }
      if (tid == 0) {
        if (!wait_counter_fast(counter, 4u * (unsigned)p.T)) ms.abort_flag = 1;
        EMR_STAMP(7);                                   // all tiles arrived
      }
      __syncthreads();
      EMR_STAMP(24);                                    // CTA released
      if (!ms.abort_flag) finalize(acc, false);
      EMR_STAMP(8);                                     // finalize done (before the loop-top barrier)
    } else {
      // ---- last iteration: nu = Z^T V^T for this side, one pass over the 8 operand images through the ring; the kappa
      // all-reduce completes under it and the first channel half drains while the second is still being multiplied ----------
      tc_fence_before_sync();
      __syncthreads();                                  // every warp has read its M-step columns: TMEM is free for nu
      tc_fence_after_sync();
      EMR_STAMP(5);
      reduce_issue(acc, kOffXH);
      reduce_arrive(counter);
      EMR_STAMP(6);
      // One thread of warp 1 issues the MMAs and refills the ring (the stage of image seq - 1 with image seq + 2, after the MMAs of
      // seq - 1 have retired: those of seq are queued behind them, so the tensor core does not wait).  The other 15 warps drain
      // the first channel half to complete -- the CTA's own -- while the second is still being multiplied.
      auto drain_block = [&](int dcol, int blk, float* ns) {   // this warp's lane quadrant x the 16 channels of column block `blk`
        uint32_t r[16];
        tmem_ld16(tmem_addr(tmem, q * 32, dcol + blk * 16), r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) ns[(blk * 16 + j) * kL + px] = __uint_as_float(r[j]);
      };
      auto drain_issue = [&](int dcol, float* ns) {            // leader of column block cb: its 8 KB -> acc_nu
        if (q == 0 && lane == 0) {
          float* dst = p.acc_nu + ((size_t)gs * kCv + dcol + cb * 16) * kL;
          asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst),
                       "r"(smem_u32(ns + cb * 16 * kL)), "r"(16 * kL * 4)
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      };
      if (warp == 1) {
        if (lane == 0) {
          bool peer_ready = false;
#pragma unroll 1
          for (int seq = 0; seq < kImages; ++seq) {
            const int st = seq % kStages;
            if (!mbar_wait(&ms.bar_full[st], (seq / kStages) & 1)) ms.abort_flag = 1;
            tc_fence_after_sync();
            const int h = (seq >> 2) ^ rank, qd = seq & 3;
            const uint32_t vb = sbase + kOffVS + st * kImgBytes;
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
              const uint64_t ad = make_sdesc(sbase + kOffZ + (qd * 2 + kk) * 2 * 128, /*lbo*/ 128, /*sbo*/ 2048);
              const uint64_t al = make_sdesc(sbase + kOffZL + (qd * 2 + kk) * 2 * 128, /*lbo*/ 128, /*sbo*/ 2048);
              const uint64_t bh = make_sdesc(vb + kk * 2 * 4096, /*lbo*/ 4096, /*sbo*/ 128);
              const uint64_t bl = make_sdesc(vb + kVPlane + kk * 2 * 4096, /*lbo*/ 4096, /*sbo*/ 128);
              mma_f16_ss(tmem + h * 256, ad, bh, idesc_nu, (qd | kk) ? 1u : 0u);   // z_hi v_hi
              mma_f16_ss(tmem + h * 256, ad, bl, idesc_nu, 1u);                    // z_hi v_lo
              mma_f16_ss(tmem + h * 256, al, bh, idesc_nu, 1u);                    // z_lo v_hi
            }
            mma_commit(&ms.bar_empty[st]);
            if (qd == 3) mma_commit(&ms.bar_nu[seq >> 2]);   // this channel half is complete
            if (seq >= 1 && seq + 2 < kImages) {
              if (!mbar_wait(&ms.bar_empty[(seq - 1) % kStages], ((seq - 1) / kStages) & 1)) ms.abort_flag = 1;
              if (seq + 2 >= 4 && !peer_ready) {           // images of the peer's channel half
                if (!mbar_wait_cluster(&ms.bar_vready, 0)) ms.abort_flag = 1;
                asm volatile("fence.proxy.async;" ::: "memory");
                peer_ready = true;
              }
              load_image(seq + 2);
            }
          }
        }
        __syncwarp();
      } else {
        // rounds 0-3: the own channel half.  z, z_lo (= the khat region) and the ring are still operands: one 32 KB staging buffer
        // in the X region, reused round after round (per column block: its leader waits for its copy to have been read).  Warp 5
        // (quadrant 1 of block 1) also covers quadrant 1 of block 0, whose warp is issuing.
        if (!warp_wait(&ms.bar_nu[0], 0, 0)) ms.abort_flag = 1;
        tc_fence_after_sync();
        EMR_STAMP(9);                                     // nu GEMM of the own channel half done
        float* ns = reinterpret_cast<float*>(smem);
#pragma unroll 1
        for (int r4 = 0; r4 < 4; ++r4) {
          const int dcol = rank * 256 + r4 * 64;
          // (r4 = 0: the staged M-step partial of this iteration was read from the same 8 KB of the block)
          if (q == 0 && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          bar_sync(9 + cb, 128);
          if (warp == 5) bar_sync(9, 128);
          drain_block(dcol, cb, ns);
          if (warp == 5) drain_block(dcol, 0, ns);
          fence_proxy_async_smem();
          bar_sync(5 + cb, 128);
          if (warp == 5) bar_sync(5, 128);
          drain_issue(dcol, ns);
        }
      }
      __syncthreads();                                    // every MMA is issued; the issuing warp is back
      // rounds 4-7: the peer's channel half, all 16 warps, a buffer per round (everything below the Misc block is dead by now)
      if (!warp_wait(&ms.bar_nu[1], 0, 0)) ms.abort_flag = 1;
      tc_fence_after_sync();
      EMR_STAMP(10);                                      // nu GEMM done
#pragma unroll 1
      for (int r4 = 0; r4 < 4; ++r4) {
        const int dcol = (rank ^ 1) * 256 + r4 * 64;
        float* ns = reinterpret_cast<float*>(smem + 32768u * (1 + r4));
        drain_block(dcol, cb, ns);
        fence_proxy_async_smem();
        bar_sync(5 + cb, 128);
        drain_issue(dcol, ns);
      }
      if (tid == 0) {
        if (!wait_counter_fast(counter, 4u * (unsigned)p.T)) ms.abort_flag = 1;   // the kappa all-reduce completed long ago
        EMR_STAMP(7);
      }
      __syncthreads();
      if (!ms.abort_flag) finalize(acc, true);
      EMR_STAMP(8);
      unsigned* counter_nu = cnt_nu;
      // prior rows of this CTA's nu slice: loaded now, under the wait for the other tiles' drains
      {
        const float4* pri4 = reinterpret_cast<const float4*>(p.nu_prior + (size_t)gs * kCv * L);
        const int l4n = (L < kL ? L : kL) / 4;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int k = d0 * l4n + tid + j * kThreads;
          nu_pre[j] = (k < d1 * l4n) ? __ldg(pri4 + (k / l4n) * (L / 4) + k % l4n) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      if (q == 0 && lane == 0) {
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // full completion (not just .read) before the arrival
        atomicAdd(counter_nu, 1u);
      }
      __syncwarp();
      if (tid == 0) {
        EMR_STAMP(11);                                  // nu drained
        if (!wait_counter_fast(counter_nu, 8u * (unsigned)p.T)) ms.abort_flag = 1;
        EMR_STAMP(12);                                  // every CTA of the unit has drained
      }
      tc_fence_before_sync();
      __syncthreads();
    }
    if (ms.abort_flag) {
      if (tid == 0) atomicExch(p.status, 1 + it);
      failed = true;
      break;
    }
  }

  if (!failed) {
    // ---- nu = (zita_ nu_ + sum / 2^14) / zita (reference :164-165) for side sd, this tile's slice of value channels --------
    const float4* acc4 = reinterpret_cast<const float4*>(p.acc_nu + (size_t)gs * kCv * kL);     // [d][128]
    const float4* pri4 = reinterpret_cast<const float4*>(p.nu_prior + (size_t)gs * kCv * L);    // [d][L]
    float4* out4 = reinterpret_cast<float4*>(p.nu + (size_t)gs * kCv * L);
    const int l4n = (L < kL ? L : kL) / 4;
    int jj = 0;
    for (int k = d0 * l4n + tid; k < d1 * l4n; k += kThreads, ++jj) {
      const int d = k / l4n, l4 = k % l4n, l = l4 * 4;
      const int i = d * (L / 4) + l4;
      const float4 a = __ldcg(acc4 + d * (kL / 4) + l4);
      const float4 pr = jj == 0 ? nu_pre[0] : jj == 1 ? nu_pre[1] : jj == 2 ? nu_pre[2] : __ldg(pri4 + i);
      float4 o;
      o.x = (ms.zp[l + 0] * pr.x + a.x * kInvZ) * ms.rz[l + 0];
      o.y = (ms.zp[l + 1] * pr.y + a.y * kInvZ) * ms.rz[l + 1];
      o.z = (ms.zp[l + 2] * pr.z + a.z * kInvZ) * ms.rz[l + 2];
      o.w = (ms.zp[l + 3] * pr.w + a.w * kInvZ) * ms.rz[l + 3];
      out4[i] = o;
      if (p.img_v != nullptr) {
        // the readout's nu operand of these bases: fp16 hi/lo, k-step kk = column / 16 of (u, channel half), [256 d][16 j] per step
        const int j = sd * p.img_Lt + p.img_bank * L + l;                 // column of the memory (side-major, then bank)
        const int ks2 = 2 * p.img_Lt / 16;
        uint8_t* dst = p.img_v + (((size_t)u * 2 + (d >> 8)) * ks2 + (j >> 4)) * 16384 + ((d & 255) % 8) * 16 + ((d & 255) / 8) * 128 +
                       ((j >> 3) & 1) * 4096 + (j & 7) * 2;
        __align__(8) __half hi[4];
        __align__(8) __half lo[4];
        split_half(o.x, hi[0], lo[0]); split_half(o.y, hi[1], lo[1]); split_half(o.z, hi[2], lo[2]); split_half(o.w, hi[3], lo[3]);
        *reinterpret_cast<uint2*>(dst) = *reinterpret_cast<uint2*>(hi);
        *reinterpret_cast<uint2*>(dst + 8192) = *reinterpret_cast<uint2*>(lo);
      }
    }
    EMR_STAMP(13);                   // nu slice written
  }
  tc_fence_before_sync();
  __syncthreads();
  if (tid == 0 && !failed) {
    // the last CTA of the unit to get here clears the unit's counters for the next launch (every other CTA has passed its waits)
    if (atomicAdd(cnt_dep, 1u) == 2u * (unsigned)p.T - 1u) {
      for (int i = 0; i < I * 2; ++i) p.counters[(size_t)u * I * 2 + i] = 0u;
      *cnt_zero = 0u;
      *cnt_nu = 0u;
      *cnt_dep = 0u;
    }
  }
  if (p.prof != nullptr && blockIdx.x == 0 && tid == 0) p.prof[0] = -(long long)n_stamp;
  if (warp == 0) tmem_dealloc(tmem, 512);
  cluster_arrive();                  // a CTA must not exit while its peer may still write into its shared memory
  cluster_wait();
  if (failed || ms.abort_flag) __trap();
}

// ------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------
template <bool VPM>
static int res_clusters_resident() {
  static PerDevice cache;                                // occupancy and the shared-memory attribute are per device
  const int dev_id = current_device();
  std::lock_guard<std::mutex> lock(cache.mu);
  int& n = cache.value[dev_id];
  if (n < 0) {
    cudaFuncSetAttribute(em_res_kernel<VPM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)emr::kSmemBytes);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2, 1, 1);
    cfg.blockDim = dim3(emr::kThreads, 1, 1);
    cfg.dynamicSmemBytes = emr::kSmemBytes;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&clusters, em_res_kernel<VPM>, &cfg) != cudaSuccess || clusters <= 0) {
      cudaGetLastError();
      int dev = 0, sms = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      clusters = sms / 2 - 4;                            // conservative guess
    }
    n = clusters;
  }
  return n;
}

constexpr int kBarWords = 4096;
static size_t bar_words_needed(const SwemDims& d) { return (size_t)d.B * d.N * (d.n_iters * 2 + 3) + 1; }

// SwemEmArgs.image_workspace: can the kernel write the readout's operand images itself?  (one block of rows per side: Lt <= 256)
bool fused_em_res_emits_images(const SwemEmArgs& a) {
  return a.image_workspace != nullptr && fused_em_res_covers(a.dims, a.v_pixel_major != 0) && a.image_n_banks >= 1 &&
         a.image_n_banks * a.dims.L <= 256 && a.image_bank >= 0 && a.image_bank < a.image_n_banks;
}

// Shapes of the V-resident kernel: Ck = 64, L <= 128, and a unit's tile pairs co-resident (single-launch form)
bool fused_em_res_covers(const SwemDims& d, bool v_pixel_major) {
  if (d.Ck != emr::kCk || (d.L != 64 && d.L != 128) || d.Cv != emr::kCv || d.n_iters < 1 || d.n_iters > 16) return false;
  if (bar_words_needed(d) > (size_t)kBarWords) return false;
  const char* off = getenv("SWEM_EM_RES");
  if (off != nullptr && off[0] == '0') return false;     // A/B switch: SWEM_EM_RES=0 runs em_pair_kernel on these shapes too
  const char* force_w = getenv("SWEM_EM_WINDOWED");
  if (force_w != nullptr && force_w[0] == '1') return false;
  const int T = (d.HW + emr::kTP - 1) / emr::kTP;
  const int resident = v_pixel_major ? res_clusters_resident<true>() : res_clusters_resident<false>();
  return T >= 1 && T <= resident;
}

// ---- library-owned arrival counters ---------------------------------------------------------------------------------------
// One zero-initialised buffer per device, cut into kBarRanges ranges of kBarWords counters; every EM call takes the next range
// (calls in flight on different streams never share one) and its kernel leaves the words it used zero again, so there is no
// memset in front of the kernel and nothing for the caller's workspace to preserve between calls.
constexpr int kBarRanges = 16;
static unsigned* bar_range(unsigned** base_out = nullptr) {
  static std::mutex mu;
  static unsigned* base[kMaxDevices] = {};
  static unsigned ticket[kMaxDevices] = {};
  const int dev = current_device();
  std::lock_guard<std::mutex> lock(mu);
  if (base[dev] == nullptr) {
    unsigned* ptr = nullptr;
    if (cudaMalloc(&ptr, (size_t)kBarRanges * kBarWords * sizeof(unsigned)) != cudaSuccess ||
        cudaMemset(ptr, 0, (size_t)kBarRanges * kBarWords * sizeof(unsigned)) != cudaSuccess) {
      cudaGetLastError();
      return nullptr;                                    // (e.g. first call ever made under stream capture)
    }
    base[dev] = ptr;
  }
  if (base_out != nullptr) *base_out = base[dev];
  return base[dev] + (size_t)(ticket[dev]++ % kBarRanges) * kBarWords;
}
template <bool VPM>
static int fused_em_res_forward_t(const SwemEmArgs& a, cudaStream_t st) {
  const SwemDims& d = a.dims;
  const int U = d.B * d.N;
  const int T = (d.HW + emr::kTP - 1) / emr::kTP;
  Arena ws(a.workspace);
  float* acc_k = ws.take<float>((size_t)U * d.n_iters * 2 * (emr::kCk + 1) * emr::kL);
  float* acc_nu = ws.take<float>((size_t)U * 2 * emr::kCv * emr::kL);
  uint8_t* vblob = ws.take<uint8_t>((size_t)U * T * emr::kImages * emr::kImgBytes);
  unsigned* counters = bar_range();
  if (counters == nullptr) {
    set_error("fused EM: cannot allocate the arrival counters (the first swem_em_forward of a device must not run under stream capture)");
    return SWEM_ERR_CUDA;
  }

  EmResParams p{};
  p.x = a.x; p.v = a.v; p.masks = a.masks;
  p.kappa_prior = a.kappa_prior; p.nu_prior = a.nu_prior; p.zita_prior = a.zita_prior;
  p.kappa = a.kappa; p.nu = a.nu; p.zita = a.zita; p.z_last = a.z_last;
  p.acc_k = acc_k; p.acc_nu = acc_nu; p.counters = counters; p.vblob = vblob;
  if (fused_em_res_emits_images(a)) {
    SwemDims rd = d;
    rd.n_banks = a.image_n_banks;
    const ReadoutImages im = readout_image_layout(a.image_workspace, rd);
    p.img_k = im.kblob; p.img_v = im.vblob; p.img_bank = a.image_bank; p.img_Lt = a.image_n_banks * d.L;
  }
  p.status = reinterpret_cast<int*>(counters + kBarWords - 1);
  p.N = d.N; p.HW = d.HW; p.T = T; p.n_iters = d.n_iters; p.L = d.L; p.U = U;
  p.c1s = kLog2e / (d.tau * emr::kKScale);
  p.prof = get_profile_buffer();
  {
    const char* dbg = getenv("SWEM_EM_DBG");
    p.dbg = dbg ? atoi(dbg) : 0;
  }
  // all CTAs of a launch spin on each other, so every launch must be co-resident (1 CTA per SM); units that do not fit are
  // spread evenly over the fewest launches
  const int upl_max = res_clusters_resident<VPM>() / T;
  const int n_launch = (U + upl_max - 1) / upl_max;
  const int upl = (U + n_launch - 1) / n_launch;
  for (int u0 = 0; u0 < U; u0 += upl) {
    const int nu = (U - u0 < upl) ? (U - u0) : upl;
    p.u0 = u0;
    em_res_kernel<VPM><<<nu * T * 2, emr::kThreads, emr::kSmemBytes, st>>>(p);
    SWEM_LAUNCH_CHECK();
  }
  return SWEM_OK;
}

int fused_em_res_forward(const SwemEmArgs& a, cudaStream_t st) {
  return a.v_pixel_major ? fused_em_res_forward_t<true>(a, st) : fused_em_res_forward_t<false>(a, st);
}

}  // namespace swem
