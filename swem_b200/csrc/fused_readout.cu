// Fused readout (attention from query pixels to the memory bases) for sm_100a: tcgen05 + TMEM + TMA.
//
// Covers Ck = 64 / 128, Cv = 512, Lt = banks x L columns per side in {64, 128, 256, 512, 1024}, any HW / B*N.  Lt = 512 /
// 1024 (L = 256 / 512 with both banks) run as 2- / 4-CTA clusters: each CTA takes 256 columns of either side, the cluster
// exchanges the row max / row sum and the un-normalised output chunks over distributed shared memory.  Reference semantics:
// methods/SWEM/modules.py:278-293 (matching), :232-276 (get_affinity), :198-208 (perm_inv_feat).
//
// Two launches + the shared top-l kernel:
//   1. readout_prep_kernel: memory banks -> tensor-core operand blobs in global memory:
//        khat = l2norm(kappa)*256 as fp16 hi/lo K-major B operands per (unit, side)   (:283)
//        nu as fp16 hi/lo K-major B operands per (unit, value-channel half, 16-column k-step)
//      (blobs are byte-for-byte the shared-memory images, so the main kernel stages them with
//       plain cp.async.bulk copies)
//   2. readout_fused_kernel: one CTA per (unit, 128-pixel tile, value-channel half):
//        scores a[p, j] = q_p . khat_j         24 x tcgen05.mma M128 N256 K16 (hi/lo split, 3 terms)
//        t = a / (||q_p|| + eps) (:282), joint max over both sides (:248-249), E = exp((t - max)/tau)
//        E -> fp16 written back INTO TMEM over the scores (packed, 2 per column) = A operand of
//        mem_out[p, d] = sum_j E[p, j] nu[d, j]  64..128 x tcgen05.mma (A from TMEM, B = nu hi/lo via an
//        8-stage TMA ring), then divided by the row sum of the rounded E (:265, :272-273)
//        The un-normalised E (fp32) also goes to a scratch buffer for
//   3. perm_inv_kernel (generic.cu): sorted top-l running sums -> S channels (:198-208)
// mem_out lands in channels [mem_channel, +Cv) and S in [s_channel, +2*topl) of the caller's
// concat buffer (:291), so no torch.cat of those pieces is needed.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "fused_common.cuh"
#include "tc05.cuh"

namespace swem {

using namespace tc05;

namespace ro {
constexpr int kTP = 128;
constexpr int kCv = 512;
constexpr int kDH = 256;                 // value channels per CTA
constexpr float kKScale = 256.f;         // khat staged as khat*256 (lo half stays normal fp16)
constexpr float kEScale = 1024.f;        // E (<= 1) staged as E*2^10
constexpr int kStages = 8;
constexpr uint32_t kStageBytes = 16384;  // one k-step of nu: [256 d][16 j] fp16 hi (8 KB) + lo (8 KB)

// shared memory map, key channels CK = 64 or 128
//   QH/QL : [c CK][p 128] chunks (MN-major A: SBO 128, LBO 2048), CK/8 * 2048 bytes each
//   KB    : khat blobs (hi + lo planes, banks * CK/64 * 32 KB per side).  CK = 64: both sides resident; CK = 128: one side
//           at a time (the second side is loaded once the MMAs of the first have retired); later the nu ring (8 x 16 KB)
template <int CK>
struct Lay {
  static constexpr uint32_t kOffQH = 0;
  static constexpr uint32_t kOffQL = kOffQH + (CK / 8) * 2048;
  static constexpr uint32_t kOffKB = kOffQL + (CK / 8) * 2048;
  static constexpr uint32_t kKBSide = 65536 * (CK / 64);       // room for 2 banks
  static constexpr int kSidesResident = (CK == 64) ? 2 : 1;
  static constexpr uint32_t kOffRing = kOffKB;                 // 8 x 16 KB, aliases the khat blobs once the scores are done
  static constexpr uint32_t kOffMisc = kOffKB + 131072;
};
struct Misc {
  float inv_nq[kTP];
  float ex_max[2][kTP];
  float ex_sum[2][kTP];
  float peer_max[3][kTP];   // (column-split clusters) row max / row sum of the peer CTAs' columns, written by the peers
  float peer_sum[3][kTP];
  uint64_t bar_k[2];
  uint64_t bar_mma;
  uint64_t bar_full[kStages];
  uint64_t bar_empty[kStages];
  uint32_t tmem_base;
  int abort_flag;
};
template <int CK>
constexpr uint32_t smem_bytes() { return Lay<CK>::kOffMisc + sizeof(Misc) + 128; }
static_assert(smem_bytes<64>() <= 227 * 1024 && smem_bytes<128>() <= 227 * 1024, "shared memory budget");
}  // namespace ro

struct ReadoutFusedParams {
  const float* qk;        // [B][64][HW]
  const uint8_t* kblob;   // [U][2 sides][column blocks][hi | lo planes of min(Lt, 256) rows]
  const uint8_t* vblob;   // [U][2 halves][KS2 k-steps][16 KB]
  float* out;             // [U][out_channels][HW]
  float* escratch;        // [U][HW][2*Lt] fp32 un-normalised E (for the top-l feature)
  long long* prof;        // optional phase stamps of CTA 0 (slots 128..), see swem_set_profile_buffer
  int N, HW, T, n_banks, out_channels, mem_channel, pixel_major;
  int Lt_total;           // columns per side in memory (= columns per side of one CTA x column blocks)
  float c1s;              // log2(e) / (tau * kKScale)
};

#define RO_STAMP()                                                                                   \
  do {                                                                                               \
    if (p.prof != nullptr && blockIdx.x == 0 && tid == 0 && n_stamp < 100) p.prof[129 + n_stamp++] = global_ns(); \
  } while (0)

// ---- prep: banks -> operand blobs ------------------------------------------------------------------
// khat blob of (u, s, column block): K-major rows jl = (bank*L + l) % R, R = min(Lt, 256) rows per block,
// byte = (jl%8)*16 + (jl/8)*128 + (c/8)*LBO + (c%8)*2, LBO = R*16; hi plane then lo plane.
template <int kCk>
__device__ __forceinline__ void prep_kappa_rows(const float* __restrict__ k0, const float* __restrict__ k1, int U, int n_banks, int bank0,
                                                int bank1, int kL, uint8_t* __restrict__ kblob, int i) {   // i <-> (u, s, bank - bank0, l)
  using namespace ro;
  const int nbk = bank1 - bank0;
  if (i >= U * 2 * nbk * kL) return;
  const int l = i % kL, bank = bank0 + (i / kL) % nbk, s = (i / (kL * nbk)) % 2, u = i / (kL * nbk * 2);
  const float* kp = (bank ? k1 : k0) + (((size_t)u * 2 + s) * kCk) * kL + l;
  float v[kCk];
  // squared norm as four quarter sums, (q0 + q1) + (q2 + q3): the order em_res_kernel uses when it emits these images itself
  // (fused_em_res.cu, "khat operand of these bases") -- emitted and converted images are bit-identical
  float qs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int c = 0; c < kCk; ++c) {
    v[c] = __ldg(kp + (size_t)c * kL);
    qs[c / (kCk / 4)] = fmaf(v[c], v[c], qs[c / (kCk / 4)]);
  }
  const float ss = (qs[0] + qs[1]) + (qs[2] + qs[3]);
  const float sc = kKScale / (sqrtf(ss) + kEpsNorm);
  const int Lt = n_banks * kL, R = Lt < 256 ? Lt : 256, nblk = Lt / R;
  const uint32_t plane = R * kCk * 2;                          // bytes of one (hi or lo) plane of a block
  const uint32_t lbo = R * 16;
  const int jt = bank * kL + l;
  uint8_t* base = kblob + (((size_t)u * 2 + s) * nblk + jt / R) * 2 * plane;
  const int j = jt % R;
#pragma unroll
  for (int g = 0; g < kCk / 8; ++g) {
    __align__(16) __half hi[8];
    __align__(16) __half lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) split_half(v[g * 8 + e] * sc, hi[e], lo[e]);
    const uint32_t off = (j % 8) * 16 + (j / 8) * 128 + g * lbo;
    *reinterpret_cast<uint4*>(base + off) = *reinterpret_cast<uint4*>(hi);
    *reinterpret_cast<uint4*>(base + plane + off) = *reinterpret_cast<uint4*>(lo);
  }
}

// nu blob of (u, half h, k-step kk): rows d (256), 16 columns j = 16*kk..; byte = (d%8)*16 + (d/8)*128 +
// (jj/8)*4096 + (jj%8)*2; hi plane (8 KB) then lo plane.  Column order j = s*Lt + bank*128 + l (:272, :295-306).
__device__ __forceinline__ void prep_nu_groups(const float* __restrict__ n0, const float* __restrict__ n1, int U, int n_banks, int bank0,
                                               int bank1, int kL, uint8_t* __restrict__ vblob, long long i) {   // i <-> (u, s, bank - bank0, d, l-group of 8)
  using namespace ro;
  const int nbk = bank1 - bank0;
  const long long total = (long long)U * 2 * nbk * kCv * (kL / 8);
  if (i >= total) return;
  const int lg = (int)(i % (kL / 8));
  const int d = (int)((i / (kL / 8)) % kCv);
  const int bank = bank0 + (int)((i / ((kL / 8) * kCv)) % nbk);
  const int s = (int)((i / ((long long)(kL / 8) * kCv * nbk)) % 2);
  const int u = (int)(i / ((long long)(kL / 8) * kCv * nbk * 2));
  const float4* src = reinterpret_cast<const float4*>((bank ? n1 : n0) + (((size_t)u * 2 + s) * kCv + d) * kL + lg * 8);
  const float4 a = __ldg(src), b = __ldg(src + 1);
  const float vals[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  __align__(16) __half hi[8];
  __align__(16) __half lo[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) split_half(vals[e], hi[e], lo[e]);
  const int Lt = n_banks * kL;
  const int j = s * Lt + bank * kL + lg * 8;
  const int kk = j / 16, jg = (j / 8) % 2;
  const int h = d / kDH, dl = d % kDH;
  const int ks2 = 2 * Lt / 16;
  uint8_t* base = vblob + (((size_t)u * 2 + h) * ks2 + kk) * kStageBytes;
  const uint32_t off = (dl % 8) * 16 + (dl / 8) * 128 + jg * 4096;
  *reinterpret_cast<uint4*>(base + off) = *reinterpret_cast<uint4*>(hi);
  *reinterpret_cast<uint4*>(base + 8192 + off) = *reinterpret_cast<uint4*>(lo);
}

// One launch for both conversions, banks [bank0, bank1) only (SwemReadArgs.bank_images_valid: a bank whose images are still
// in the workspace is skipped): the first `kappa_blocks` blocks convert khat rows, the rest nu column groups.
template <int kCk>
__global__ void __launch_bounds__(256) readout_prep_kernel(const float* __restrict__ k0, const float* __restrict__ k1,
                                                           const float* __restrict__ n0, const float* __restrict__ n1, int U, int n_banks,
                                                           int bank0, int bank1, int kL, int kappa_blocks, uint8_t* __restrict__ kblob,
                                                           uint8_t* __restrict__ vblob) {
  if ((int)blockIdx.x < kappa_blocks)
    prep_kappa_rows<kCk>(k0, k1, U, n_banks, bank0, bank1, kL, kblob, blockIdx.x * 256 + threadIdx.x);
  else
    prep_nu_groups(n0, n1, U, n_banks, bank0, bank1, kL, vblob, (long long)(blockIdx.x - kappa_blocks) * 256 + threadIdx.x);
}

// ---- main kernel --------------------------------------------------------------------------------------
// LT = columns per side handled by one CTA (64, 128 or 256), CK = key channels (64 or 128): they fix every loop count, so
// the MMA issue loops unroll.  NS = column blocks = CTAs per cluster (1; 2 / 4 when banks x L = 512 / 1024, launched with a
// cluster attribute): CTA `rank` takes columns [rank * LT, +LT) of either side and finalises value channels
// [rank * 256 / NS, +256 / NS) of its half.
template <int LT, int CK, int NS>
__global__ void __launch_bounds__(256, 1) readout_fused_kernel(const ReadoutFusedParams p) {
  using namespace ro;
  using LY = Lay<CK>;
  constexpr int kCk = CK;
  constexpr uint32_t kOffQH = LY::kOffQH, kOffQL = LY::kOffQL, kOffKB = LY::kOffKB, kKBSide = LY::kKBSide, kOffRing = LY::kOffRing,
                     kOffMisc = LY::kOffMisc;
  extern __shared__ __align__(1024) uint8_t smem[];
  Misc& ms = *reinterpret_cast<Misc*>(smem + kOffMisc);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rank = (NS == 1) ? 0 : (int)cluster_ctarank();
  const int bid = blockIdx.x / NS;
  const int h = bid & 1;
  const int tile = (bid >> 1) % p.T;
  const int u = (bid >> 1) / p.T;
  constexpr int kRing = kStages / NS;              // NS > 1: the rest of the 128 KB ring region receives the peers' partials (64 / 96 KB)
  const int b = u / p.N;
  const int p0 = tile * kTP;
  const int HW = p.HW;
  constexpr int Lt = LT;                  // columns per side = banks x bases per bank
  constexpr int ks_side = Lt / 16;        // PV k-steps per side
  constexpr int ks2 = 2 * ks_side;
  const uint32_t sbase = smem_u32(smem);
  constexpr uint32_t kplane = Lt * CK * 2;   // bytes of one khat plane (hi or lo) of one side: [Lt rows][CK] fp16
  int n_stamp = 0;
  RO_STAMP();

  if (warp == 0) tmem_alloc(&ms.tmem_base, 512);
  if (tid == 0) {
    mbar_init(&ms.bar_k[0], 1);
    mbar_init(&ms.bar_k[1], 1);
    mbar_init(&ms.bar_mma, 1);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&ms.bar_full[i], 1);
      mbar_init(&ms.bar_empty[i], 1);
    }
    ms.abort_flag = 0;
    fence_mbar_init();
    // khat blobs (hi + lo planes are contiguous): one bulk copy per resident side
    for (int s = 0; s < LY::kSidesResident; ++s) {
      mbar_expect_tx(&ms.bar_k[s], 2 * kplane);
      bulk_g2s(smem + kOffKB + s * kKBSide, p.kblob + (((size_t)u * 2 + s) * NS + rank) * 2 * kplane, 2 * kplane, &ms.bar_k[s]);
    }
  }
  if constexpr (NS > 1) cluster_arrive();   // the peer must be running before anything is stored into its shared memory
  // query tile: norms (thread <-> pixel) and fp16 hi/lo MN-major A operand
  if (tid < kTP) {
    const int px = p0 + tid;
    float ss = 0.f;
    if (px < HW) {
      const float* qp = p.qk + (size_t)b * kCk * HW + px;
#pragma unroll 8
      for (int c = 0; c < kCk; ++c) {
        const float t = __ldg(qp + (size_t)c * HW);
        ss = fmaf(t, t, ss);
      }
    }
    ms.inv_nq[tid] = 1.f / (sqrtf(ss) + kEpsNorm);
  }
#pragma unroll
  for (int cc = 0; cc < kCk / 64; ++cc) {
    const int c = cc * 64 + (tid >> 2);
    const float* qrow = p.qk + ((size_t)b * kCk + c) * HW;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int pg = (tid & 3) * 4 + j;
      __align__(16) __half hi[8];
      __align__(16) __half lo[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int px = p0 + pg * 8 + e;
        split_half(px < HW ? __ldg(qrow + px) : 0.f, hi[e], lo[e]);
      }
      const uint32_t off = (c % 8) * 16 + (c / 8) * 2048 + pg * 128;
      *reinterpret_cast<uint4*>(smem + kOffQH + off) = *reinterpret_cast<uint4*>(hi);
      *reinterpret_cast<uint4*>(smem + kOffQL + off) = *reinterpret_cast<uint4*>(lo);
    }
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = ms.tmem_base;
  bool ok = true;
  RO_STAMP();   // setup (query tile staged)

  // ---- scores: side s -> TMEM columns [256 s, 256 s + Lt) ---------------------------------------------
  // (single-thread sections: warp 0 / lane 0 + __syncwarp, so the rest of warp 0 parks at the warp
  //  barrier instead of spinning on an mbarrier in a divergent branch and starving lane 0)
  if (warp == 0) {
    if (lane == 0) {
    const uint32_t idesc = make_idesc(128, Lt, kFmtF16, kFmtF16, kMajorMN, kMajorK);
    constexpr uint32_t lbo_k = Lt * 16;
    for (int s = 0; s < 2; ++s) {
      const int slot = (LY::kSidesResident == 2) ? s : 0;
      if (LY::kSidesResident == 1 && s == 1) {
        // one side resident: wait until the MMAs that read side 0 have retired, then reuse its buffer (bar_k[1] tracks both)
        mma_commit(&ms.bar_k[1]);
        ok = mbar_wait(&ms.bar_k[1], 0) && ok;
        mbar_expect_tx(&ms.bar_k[1], 2 * kplane);
        bulk_g2s(smem + kOffKB, p.kblob + (((size_t)u * 2 + 1) * NS + rank) * 2 * kplane, 2 * kplane, &ms.bar_k[1]);
        ok = mbar_wait(&ms.bar_k[1], 1) && ok;
      } else {
        ok = mbar_wait(&ms.bar_k[s], 0) && ok;
      }
      const uint32_t kb = sbase + kOffKB + slot * kKBSide;
#pragma unroll
      for (int term = 0; term < 3; ++term) {
        const uint32_t qa = sbase + (term == 2 ? kOffQL : kOffQH);
        const uint32_t kbt = kb + (term == 1 ? kplane : 0);
#pragma unroll
        for (int kk = 0; kk < kCk / 16; ++kk) {
          const uint64_t ad = make_sdesc(qa + kk * 2 * 2048, /*lbo*/ 2048, /*sbo*/ 128);
          const uint64_t bd = make_sdesc(kbt + kk * 2 * lbo_k, /*lbo*/ lbo_k, /*sbo*/ 128);
          mma_f16_ss(tmem + s * 256, ad, bd, idesc, (term | kk) ? 1u : 0u);
        }
      }
    }
    mma_commit(&ms.bar_mma);
    }
    __syncwarp();
  }
  SWEM_CTA_WAIT(&ms.bar_mma, 0, ms.abort_flag);
  tc_fence_after_sync();
  RO_STAMP();   // scores done

  // the khat blobs are dead: start streaming nu k-steps into the ring (aliases them)
  // (k-steps of the blob follow the memory's column order j = s * Lt_total + ...; this CTA's k-step kk is number
  //  src_step(kk) of it)
  const uint8_t* vsrc = p.vblob + ((size_t)u * 2 + h) * (ks2 * NS) * kStageBytes;
  auto src_step = [&](int kk) { return (kk / ks_side) * (ks_side * NS) + rank * ks_side + kk % ks_side; };
  if (tid == 0) {
    fence_proxy_async_smem();
#pragma unroll
    for (int kk = 0; kk < kRing && kk < ks2; ++kk) {
      mbar_expect_tx(&ms.bar_full[kk], kStageBytes);
      bulk_g2s(smem + kOffRing + kk * kStageBytes, vsrc + (size_t)src_step(kk) * kStageBytes, kStageBytes, &ms.bar_full[kk]);
    }
  }

  // ---- softmax epilogue: thread <-> (pixel px, side sd) ------------------------------------------------
  const int px = (warp & 3) * 32 + lane, sd = warp >> 2;
  const uint32_t lane_base = (warp & 3) * 32;
  float inv_total;
  {
    constexpr int nchunk = Lt / 32;
    float mx = -3.0e38f;
    for (int q = 0; q < nchunk; ++q) {
      uint32_t r[32];
      tmem_ld32(tmem_addr(tmem, lane_base, sd * 256 + q * 32), r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(r[j]));
    }
    ms.ex_max[sd][px] = mx;
    __syncthreads();
    RO_STAMP(); // max pass
    float gm = fmaxf(ms.ex_max[0][px], ms.ex_max[1][px]);           // inv_nq > 0: max of a*inv = inv * max a
    if constexpr (NS > 1) {     // joint max over the column blocks: one float per pixel through the peer's shared memory
      cluster_wait();           // (start-up barrier: the peer is resident)
      for (int k = sd; k < NS - 1; k += 2)   // peer rank + 1 + k receives into its slot k (the two threads of a pixel share the peers)
        st_cluster_f32(map_to_peer(smem_u32(&ms.peer_max[k][px]), (uint32_t)((rank + 1 + k) % NS)), gm);
      cluster_arrive();
      cluster_wait();
#pragma unroll
      for (int k = 0; k < NS - 1; ++k) gm = fmaxf(gm, ms.peer_max[k][px]);
    }
    const float cw = ms.inv_nq[px] * p.c1s;
    float sum = 0.f;
    for (int q = 0; q < nchunk; ++q) {
      uint32_t r[32];
      tmem_ld32(tmem_addr(tmem, lane_base, sd * 256 + q * 32), r);
      tmem_ld_wait();
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float e0 = fast_exp2((__uint_as_float(r[2 * j]) - gm) * cw);
        const float e1 = fast_exp2((__uint_as_float(r[2 * j + 1]) - gm) * cw);
        const __half2 hh = __floats2half2_rn(e0 * kEScale, e1 * kEScale);
        const float2 back = __half22float2(hh);
        sum += back.x + back.y;                                      // row sum of the ROUNDED operand
        pk[j] = *reinterpret_cast<const uint32_t*>(&hh);
        r[2 * j] = __float_as_uint(e0);
        r[2 * j + 1] = __float_as_uint(e1);
      }
      if (h == 0) {
        // E chunk [32 px of this warp][32 cols] -> scratch rows.  Transposed through shared memory (the dead
        // query tile; XOR-swizzled float4 slots) so that one store instruction covers 4 rows x 128 B.
        float4* tbuf = reinterpret_cast<float4*>(smem + kOffQH) + warp * 256;      // [32 rows][8 float4]
#pragma unroll
        for (int j = 0; j < 8; ++j)
          tbuf[lane * 8 + (j ^ (lane & 7))] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                                          __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
        __syncwarp();
        const int c4 = lane & 7;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = i * 4 + (lane >> 3);
          const int pxr = p0 + (warp & 3) * 32 + row;
          const float4 val = tbuf[row * 8 + (c4 ^ (row & 7))];
          if (pxr < HW)
            *reinterpret_cast<float4*>(p.escratch + ((size_t)u * HW + pxr) * (2 * p.Lt_total) + sd * p.Lt_total + rank * Lt + q * 32 + c4 * 4) = val;
        }
        __syncwarp();
      }
      // packed E overwrites columns this thread has already consumed: [256 sd + 16 q, +16)
      tmem_st16(tmem_addr(tmem, lane_base, sd * 256 + q * 16), pk);
    }
    tmem_st_wait();
    ms.ex_sum[sd][px] = sum;
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    inv_total = ms.ex_sum[0][px] + ms.ex_sum[1][px];               // (NS = 1: inverted right here)
    if constexpr (NS > 1) {
      for (int k = sd; k < NS - 1; k += 2)
        st_cluster_f32(map_to_peer(smem_u32(&ms.peer_sum[k][px]), (uint32_t)((rank + 1 + k) % NS)), inv_total);
    } else {
      inv_total = 1.f / inv_total;
    }
  }
  RO_STAMP();   // exp pass + E packed

  // ---- mem_out = E nu^T: A from TMEM (packed E), B from the ring; accumulators at columns 128.. and 384.. ----
  if (warp == 0) {
    if (lane == 0) {
    const uint32_t idesc = make_idesc(128, 128, kFmtF16, kFmtF16, kMajorK, kMajorK);
    // One thread issues everything, so this loop is latency-bound on its own instruction stream: keep it
    // lean (descriptors advanced by constant adds, every index a compile-time constant after unrolling).
    constexpr int kLag = (NS == 1) ? 3 : (NS == 2) ? 2 : 1;   // refill the stage consumed kLag steps ago: its MMAs have retired, no issue stall
    const uint64_t bdesc0 = make_sdesc(sbase + kOffRing, /*lbo*/ 4096, /*sbo*/ 128);
#pragma unroll
    for (int kk = 0; kk < ks2; ++kk) {
      const int st = kk % kRing;
      ok = mbar_wait(&ms.bar_full[st], (kk / kRing) & 1) && ok;
      tc_fence_after_sync();
      const uint32_t a_tmem = tmem + (kk / ks_side) * 256 + (kk % ks_side) * 8;
#pragma unroll
      for (int term = 0; term < 2; ++term)
#pragma unroll
        for (int nh = 0; nh < 2; ++nh)
          mma_f16_ts(tmem + 128 + nh * 256, a_tmem, bdesc0 + ((st * kStageBytes + term * 8192 + nh * 2048) >> 4), idesc,
                     (kk | term) ? 1u : 0u);
      mma_commit(&ms.bar_empty[st]);
      if (kk >= kLag && kk - kLag + kRing < ks2) {
        const int prev = kk - kLag, nxt = prev + kRing;
        const int ps = prev % kRing;
        ok = mbar_wait(&ms.bar_empty[ps], (prev / kRing) & 1) && ok;
        mbar_expect_tx(&ms.bar_full[ps], kStageBytes);
        bulk_g2s(smem + kOffRing + ps * kStageBytes, vsrc + (size_t)src_step(nxt) * kStageBytes, kStageBytes, &ms.bar_full[ps]);
      }
    }
    mma_commit(&ms.bar_mma);
    }
    __syncwarp();
  }
  RO_STAMP();   // PV issued
  SWEM_CTA_WAIT(&ms.bar_mma, 1, ms.abort_flag);
  tc_fence_after_sync();
  RO_STAMP();   // PV done

  if constexpr (NS > 1) {
    // ---- column-split cluster: out = sum_r O_r / sum_r rowsum_r.  CTA `rank` finalises kChunk = 256 / NS value channels of
    // this half, [rank * kChunk, +kChunk); the other chunks of its partial go to their owners' receive buffers
    // [sender slot][ch][px 128] (the upper part of the ring region: dead since the owner's score MMAs, which it waited for
    // before the max exchange).
    constexpr int kChunk = kDH / NS;
    const int nh = warp >> 2;
    const bool in_range = p0 + px < HW;
    float* rbuf = reinterpret_cast<float*>(smem + kOffRing + kRing * kStageBytes);
    const uint32_t rbuf_u32 = smem_u32(rbuf);
#pragma unroll
    for (int cc = 0; cc < NS / 2; ++cc) {
      const int c = nh * (NS / 2) + cc;                      // chunk held by this thread in TMEM columns 128 + nh * 256 + cc * kChunk
      if (c != rank) {
        const int slot = ((rank - c + NS) % NS) - 1;
        const uint32_t dst = map_to_peer(rbuf_u32, (uint32_t)c) + (uint32_t)((slot * kChunk * kTP + px) * 4);
#pragma unroll
        for (int q = 0; q < kChunk / 32; ++q) {
          uint32_t r[32];
          tmem_ld32(tmem_addr(tmem, lane_base, 128 + nh * 256 + cc * kChunk + q * 32), r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) st_cluster_f32(dst + (q * 32 + j) * (kTP * 4), __uint_as_float(r[j]));
        }
      }
    }
    cluster_arrive();
    cluster_wait();
    // all 8 warps: pixel px, half of the chunk's channels (warps 0-3 the first half, warps 4-7 the rest)
    float total = ms.ex_sum[0][px] + ms.ex_sum[1][px];
#pragma unroll
    for (int k = 0; k < NS - 1; ++k) total += ms.peer_sum[k][px];
    const float scale = 1.f / total;
    const int cbase = nh * (kChunk / 2);                     // first channel (inside the chunk) of this thread
    const uint32_t tcol = 128 + ((rank * kChunk) / 128) * 256 + (rank * kChunk) % 128 + cbase;
#pragma unroll
    for (int q = 0; q < kChunk / 64; ++q) {
      uint32_t r[32];
      tmem_ld32(tmem_addr(tmem, lane_base, tcol + q * 32), r);
      tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float acc = __uint_as_float(r[j]);
#pragma unroll
        for (int k = 0; k < NS - 1; ++k) acc += rbuf[(k * kChunk + cbase + q * 32 + j) * kTP + px];
        v[j] = acc * scale;
      }
      if (in_range) {
        const int ch0 = p.mem_channel + h * kDH + rank * kChunk + cbase + q * 32;
        if (p.pixel_major) {
          float4* o4 = reinterpret_cast<float4*>(p.out + ((size_t)u * HW + p0 + px) * p.out_channels + ch0);
#pragma unroll
          for (int j = 0; j < 8; ++j) o4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        } else {
          float* obase = p.out + ((size_t)u * p.out_channels + ch0) * HW + p0 + px;
#pragma unroll
          for (int j = 0; j < 32; ++j) obase[(size_t)j * HW] = v[j];
        }
      }
    }
  } else {
    // ---- normalise and store: warps 0-3 -> channels [0,128) of this half, warps 4-7 -> [128,256) ---------------
    const int nh = warp >> 2;
    const float scale = inv_total;             // the 2^10 of E cancels against the row sum of the same operand
    const bool in_range = p0 + px < HW;
    if (p.pixel_major) {
      // [U][HW][out_channels]: a thread owns 128 consecutive channels of its pixel -> 16-byte stores
      float4* obase = reinterpret_cast<float4*>(p.out + ((size_t)u * HW + p0 + px) * p.out_channels + p.mem_channel + h * kDH + nh * 128);
      for (int q = 0; q < 4; ++q) {
        uint32_t r[32];
        tmem_ld32(tmem_addr(tmem, lane_base, 128 + nh * 256 + q * 32), r);
        tmem_ld_wait();
        if (in_range) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            obase[q * 8 + j] = make_float4(__uint_as_float(r[4 * j]) * scale, __uint_as_float(r[4 * j + 1]) * scale,
                                           __uint_as_float(r[4 * j + 2]) * scale, __uint_as_float(r[4 * j + 3]) * scale);
        }
      }
    } else {
      float* obase = p.out + ((size_t)u * p.out_channels + p.mem_channel + h * kDH + nh * 128) * HW + p0 + px;
      for (int q = 0; q < 4; ++q) {
        uint32_t r[32];
        tmem_ld32(tmem_addr(tmem, lane_base, 128 + nh * 256 + q * 32), r);
        tmem_ld_wait();
        if (in_range) {
#pragma unroll
          for (int j = 0; j < 32; ++j) obase[(size_t)(q * 32 + j) * HW] = __uint_as_float(r[j]) * scale;
        }
      }
    }
  }
  tc_fence_before_sync();
  const int bad = __syncthreads_or((!ok) || ms.abort_flag);
  RO_STAMP();   // stored
  if (p.prof != nullptr && blockIdx.x == 0 && tid == 0) p.prof[128] = n_stamp;
  if (warp == 0) tmem_dealloc(tmem, 512);
  if (bad) __trap();
}

// ------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------
bool fused_readout_supported(const SwemDims& d) {
  const int Lt = d.L * d.n_banks;
  return (d.Ck == 64 || d.Ck == 128) && (d.L == 64 || d.L == 128 || d.L == 256 || d.L == 512) &&
         (Lt == 64 || Lt == 128 || Lt == 256 || Lt == 512 || Lt == 1024) &&
         d.Cv == ro::kCv && (d.n_banks == 1 || d.n_banks == 2) && d.topl >= 1 && d.topl <= 64 && d.HW >= 1;
}

size_t fused_readout_workspace(const SwemDims& d) {
  const size_t U = (size_t)d.B * d.N;
  const size_t Lt = (size_t)d.L * d.n_banks;
  size_t bytes = 0;
  bytes += align_up(U * 2 * 2 * Lt * d.Ck * 2, 256);                     // khat blobs
  bytes += align_up(U * 2 * (2 * Lt / 16) * ro::kStageBytes, 256);       // nu blobs
  bytes += align_up(U * d.HW * 2 * Lt * 4, 256);                         // E scratch
  return bytes + 256;
}

// Where the operand images of a memory of d.n_banks banks live in a readout workspace (one source of truth for the readout and for
// the EM kernels that emit the images of the bases they produce)
ReadoutImages readout_image_layout(void* workspace, const SwemDims& d) {
  const size_t U = (size_t)d.B * d.N, Lt = (size_t)d.L * d.n_banks;
  Arena ws(workspace);
  ReadoutImages im;
  im.kblob = ws.take<uint8_t>(U * 2 * 2 * Lt * d.Ck * 2);
  im.vblob = ws.take<uint8_t>(U * 2 * (2 * Lt / 16) * ro::kStageBytes);
  im.end_offset = ws.off;
  return im;
}

// convert banks [bank0, bank1) of a memory of d.n_banks banks (kappa / nu: one pointer per bank, only those of the range are read)
int launch_bank_images(const SwemDims& d, const float* const kappa[2], const float* const nu[2], int bank0, int bank1,
                       const ReadoutImages& im, cudaStream_t st) {
  if (bank0 >= bank1) return SWEM_OK;
  const int U = d.B * d.N, nbk = bank1 - bank0;
  const int kappa_blocks = (U * 2 * nbk * d.L + 255) / 256;
  const long long m = (long long)U * 2 * nbk * ro::kCv * (d.L / 8);
  const unsigned grid = (unsigned)kappa_blocks + (unsigned)((m + 255) / 256);
  if (d.Ck == 64)
    readout_prep_kernel<64><<<grid, 256, 0, st>>>(kappa[0], kappa[1], nu[0], nu[1], U, d.n_banks, bank0, bank1, d.L, kappa_blocks, im.kblob, im.vblob);
  else
    readout_prep_kernel<128><<<grid, 256, 0, st>>>(kappa[0], kappa[1], nu[0], nu[1], U, d.n_banks, bank0, bank1, d.L, kappa_blocks, im.kblob, im.vblob);
  SWEM_LAUNCH_CHECK();
  return SWEM_OK;
}

int fused_readout_forward(const SwemReadArgs& a, cudaStream_t st) {
  const SwemDims& d = a.dims;
  const int U = d.B * d.N, nb = d.n_banks, Lt = d.L * nb;
  const int T = (d.HW + ro::kTP - 1) / ro::kTP;
  const ReadoutImages im = readout_image_layout(a.workspace, d);
  uint8_t* kblob = im.kblob;
  uint8_t* vblob = im.vblob;
  Arena ws(a.workspace);
  ws.off = im.end_offset;
  float* escr = ws.take<float>((size_t)U * d.HW * 2 * Lt);

  {
    // operand images: banks whose images the caller vouches for (bank_images_valid: the reference's fixed 'first' bank from the
    // second readout of a sequence on, a bank whose images the EM kernel emitted) are skipped; with two banks the banks to
    // convert always form one range
    int bank0 = 0, bank1 = nb;
    while (bank0 < bank1 && ((a.bank_images_valid >> bank0) & 1)) ++bank0;
    while (bank1 > bank0 && ((a.bank_images_valid >> (bank1 - 1)) & 1)) --bank1;
    const float* const kk[2] = {a.kappa[0], a.kappa[1]};
    const float* const nn[2] = {a.nu[0], a.nu[1]};
    if (int rc = launch_bank_images(d, kk, nn, bank0, bank1, im, st)) return rc;
  }
  if (fused_readout_topl_covers(d)) return fused_readout_topl_launch(a, kblob, vblob, st);   // one kernel: scores, softmax, PV and top-l
#define SWEM_RO_ATTR(LT_, CK_, NS_) \
  SWEM_CUDA(cudaFuncSetAttribute(readout_fused_kernel<LT_, CK_, NS_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ro::smem_bytes<CK_>()))
  static PerDevice once;                                 // function attributes are per device
  const int dev_id = current_device();
  {
  std::lock_guard<std::mutex> once_lock(once.mu);
  if (!once.done[dev_id]) {
    SWEM_RO_ATTR(64, 64, 1); SWEM_RO_ATTR(128, 64, 1); SWEM_RO_ATTR(256, 64, 1); SWEM_RO_ATTR(256, 64, 2); SWEM_RO_ATTR(256, 64, 4);
    SWEM_RO_ATTR(64, 128, 1); SWEM_RO_ATTR(128, 128, 1); SWEM_RO_ATTR(256, 128, 1); SWEM_RO_ATTR(256, 128, 2); SWEM_RO_ATTR(256, 128, 4);
    once.done[dev_id] = true;
  }
  }
#undef SWEM_RO_ATTR
  ReadoutFusedParams p{};
  p.qk = a.qk; p.kblob = kblob; p.vblob = vblob; p.out = a.out; p.escratch = escr;
  p.N = d.N; p.HW = d.HW; p.T = T; p.n_banks = nb; p.out_channels = a.out_channels; p.mem_channel = a.mem_channel;
  p.pixel_major = a.out_pixel_major;
  p.Lt_total = Lt;
  p.c1s = kLog2e / (d.tau * ro::kKScale);
  p.prof = get_profile_buffer();
#define SWEM_RO_LAUNCH(LT_, CK_) readout_fused_kernel<LT_, CK_, 1><<<U * T * 2, 256, ro::smem_bytes<CK_>(), st>>>(p)
  if (Lt >= 512) {
    // 2- / 4-CTA clusters (column blocks of 256 per side); CTAs of a cluster only wait for each other, so any number of
    // clusters may be queued
    const unsigned ns = (unsigned)(Lt / 256);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(U * T * 2) * ns, 1, 1);
    cfg.blockDim = dim3(256, 1, 1);
    cfg.dynamicSmemBytes = d.Ck == 64 ? ro::smem_bytes<64>() : ro::smem_bytes<128>();
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = ns;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (d.Ck == 64 && ns == 2) SWEM_CUDA(cudaLaunchKernelEx(&cfg, readout_fused_kernel<256, 64, 2>, p));
    else if (d.Ck == 64) SWEM_CUDA(cudaLaunchKernelEx(&cfg, readout_fused_kernel<256, 64, 4>, p));
    else if (ns == 2) SWEM_CUDA(cudaLaunchKernelEx(&cfg, readout_fused_kernel<256, 128, 2>, p));
    else SWEM_CUDA(cudaLaunchKernelEx(&cfg, readout_fused_kernel<256, 128, 4>, p));
  } else if (d.Ck == 64) {
    if (Lt == 64) SWEM_RO_LAUNCH(64, 64);
    else if (Lt == 128) SWEM_RO_LAUNCH(128, 64);
    else SWEM_RO_LAUNCH(256, 64);
  } else {
    if (Lt == 64) SWEM_RO_LAUNCH(64, 128);
    else if (Lt == 128) SWEM_RO_LAUNCH(128, 128);
    else SWEM_RO_LAUNCH(256, 128);
  }
#undef SWEM_RO_LAUNCH
  SWEM_LAUNCH_CHECK();
  return launch_perm_inv(escr, U, d.HW, Lt, d.topl, a.out, a.out_channels, a.s_channel, a.out_pixel_major, st);
}

}  // namespace swem
