// Fused readout with the sorted top-l feature computed in the same kernel (sm_100a: tcgen05 + TMEM + bulk-async copies), for
// Ck = 64 and Lt = banks x L <= 256 columns per side -- the BASELINE shape family; every other covered shape runs
// readout_fused_kernel + perm_inv_kernel (fused_readout.cu).  Reference semantics: methods/SWEM/modules.py:278-293 (matching),
// :232-276 (get_affinity), :198-208 (perm_inv_feat).
//
// One CTA of 16 warps per (unit, 128-pixel tile, value-channel half h):
//   scores a[p, j] = q_p . khat_j               hi/lo split, 3 products, both sides -> TMEM columns [256 s, 256 s + Lt)
//   softmax epilogue, 4 threads per pixel       (side, column half): joint max over both sides (:248-249), E = exp((t - max) / tau)
//                                               packed to fp16 over the consumed scores = A operand of the PV product; the
//                                               UN-ROUNDED fp32 E of the CTA's 64 "own" pixels (half h of the tile) goes to a
//                                               shared-memory table [64 px][2 Lt + 1]
//   mem_out[p, d] = sum_j E[p, j] nu[d, j]      one thread issues (A from TMEM, B = nu hi/lo through a 5-stage bulk-copy ring) ...
//   S (sorted top-l running sums, :198-208)     ... while the other 15 warps sort: one warp per own pixel, bitonic network on
//                                               packed (value, column) words in registers, exact fp32 running sums over rank
//   normalise + store mem_out, S                straight into the caller's concat buffer (:291), either layout
// Against readout_fused_kernel + perm_inv_kernel this removes the 16.6 MB fp32 E scratch (written, then re-read by a second
// kernel), one launch, and hides the sort under the PV product (profiles/r2_readout_phases.txt).
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "fused_common.cuh"
#include "tc05.cuh"
#include "topl.cuh"

namespace swem {

using namespace tc05;

namespace rot {
constexpr int kTP = 128;
constexpr int kCk = 64;
constexpr int kCv = 512;
constexpr int kDH = 256;                 // value channels per CTA
constexpr int kThreads = 512;
constexpr float kKScale = 256.f;
constexpr float kEScale = 1024.f;
constexpr int kRing = 5;
constexpr int kLag = 1;                  // refill the stage consumed kLag steps ago (1: four stages in flight; the MMAs of the current step are queued behind those waited for)
constexpr uint32_t kStageBytes = 16384;  // one k-step of nu: [256 d][16 j] fp16 hi (8 KB) + lo (8 KB)
constexpr int kOwn = 64;                 // pixels whose top-l feature this CTA computes

// shared memory map (bytes)
//   KB   [0, 128 KB)          khat blobs, side s at s * 64 KB (hi + lo planes); dead after the scores
//   ring [0, 80 KB)           5 x 16 KB nu k-steps (aliases KB)
//   E    [80 KB, 80 KB + 64 (2 Lt + 1) 4)   fp32 exp-affinities of the own pixels (aliases KB / Q: written after the scores)
//   QH / QL [128 KB, 160 KB)  query tile, MN-major A (SBO 128, LBO 2048); dead after the scores
constexpr uint32_t kOffKB = 0;
constexpr uint32_t kKBSide = 65536;
constexpr uint32_t kOffRing = 0;
constexpr uint32_t kOffE = kRing * kStageBytes;
constexpr uint32_t kOffQH = 131072;
constexpr uint32_t kOffQL = kOffQH + 16384;
constexpr uint32_t kEBytesMax = kOwn * (2 * 256 + 17) * 4;
constexpr uint32_t kOffMisc = ((kOffE + kEBytesMax + 127) / 128) * 128;
static_assert(kOffQL + 16384 <= kOffMisc, "query tile inside the aliased region");
constexpr int kTileStride = 68;          // floats per pixel row of a store-phase transpose tile (64 channels + 4: 16-byte aligned, conflict-free)
static_assert(16 * 32 * kTileStride * 4 <= kOffE + kEBytesMax, "transpose tiles of the store phase fit in the dead ring + E table");
struct Misc {
  float inv_nq[kTP];
  float ex_max[4][kTP];
  float ex_sum[4][kTP];
  __align__(16) uint32_t top[16][2][64];      // per warp: the top-l packed words of either side
  uint64_t bar_k[2];
  uint64_t bar_mma;
  uint64_t bar_full[kRing];
  uint64_t bar_empty[kRing];
  uint32_t tmem_base;
  int abort_flag;
};
constexpr uint32_t kSmemBytes = kOffMisc + sizeof(Misc) + 128;
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
}  // namespace rot

struct ReadoutToplParams {
  const float* qk;        // [B][64][HW]
  const uint8_t* kblob;   // [U][2 sides][hi | lo planes of Lt rows]
  const uint8_t* vblob;   // [U][2 halves][2 Lt / 16 k-steps][16 KB]
  float* out;             // [U][out_channels][HW] or [U][HW][out_channels]
  long long* prof;
  int N, HW, T, out_channels, mem_channel, s_channel, pixel_major, topl;
  int dbg;                // measurement switch (SWEM_RO_DBG): 1 = skip the top-l sort (S is garbage) -- times the PV product alone
  float c1s;              // log2(e) / (tau * kKScale)
};

#define ROT_STAMP()                                                                                          \
  do {                                                                                                       \
    if (p.prof != nullptr && blockIdx.x == 0 && tid == 0 && n_stamp < 100) p.prof[129 + n_stamp++] = global_ns(); \
  } while (0)

template <int LT>
__global__ void __block_size__((32, 16, 1)) __maxnreg__(128) readout_topl_kernel(const ReadoutToplParams p) {   // (fixed block shape: threadIdx.y is warp-uniform for ptxas)
  using namespace rot;
  constexpr int Lt = LT;
  constexpr int ks_side = Lt / 16, ks2 = 2 * ks_side;
  constexpr int nchunk = Lt / 32, nstep = nchunk / 2;      // 32-column chunks per side; a thread takes every other one
  constexpr int R = Lt / 16;                               // sorted words per lane (a half-warp sorts one side)
  constexpr int IDXB = (Lt == 64 ? 6 : Lt == 128 ? 7 : 8);
  constexpr int kESide = Lt + 16;                          // side 1 of a pixel sits 16 banks away from side 0
  constexpr int kERow = 2 * Lt + 17;                       // floats per own pixel in the E table (odd: conflict-free over pixels)
  constexpr uint32_t kplane = Lt * kCk * 2;                // bytes of one khat plane (hi or lo) of one side
  extern __shared__ __align__(1024) uint8_t smem[];
  Misc& ms = *reinterpret_cast<Misc*>(smem + kOffMisc);
  float* const etab = reinterpret_cast<float*>(smem + kOffE);
  const int warp = threadIdx.y, lane = threadIdx.x, tid = warp * 32 + lane;
  const int h = blockIdx.x & 1;
  const int tile = (blockIdx.x >> 1) % p.T;
  const int u = (blockIdx.x >> 1) / p.T;
  const int b = u / p.N;
  const int p0 = tile * kTP;
  const int HW = p.HW;
  const uint32_t sbase = smem_u32(smem);
  // thread roles: TMEM lane quadrant q, (side, column half) block cb
  const int q = warp & 3, cb = warp >> 2, sd = cb >> 1, ch = cb & 1;
  const int px = q * 32 + lane;
  const uint32_t lane_base = q * 32;
  const bool mma_thread = (tid == 32);
  int n_stamp = 0;
  ROT_STAMP();

  if (warp == 0) tmem_alloc(&ms.tmem_base, 512);
  if (tid == 0) {
    mbar_init(&ms.bar_k[0], 1);
    mbar_init(&ms.bar_k[1], 1);
    mbar_init(&ms.bar_mma, 1);
    for (int i = 0; i < kRing; ++i) {
      mbar_init(&ms.bar_full[i], 1);
      mbar_init(&ms.bar_empty[i], 1);
    }
    ms.abort_flag = 0;
    fence_mbar_init();
    for (int s = 0; s < 2; ++s) {       // khat blobs (hi + lo planes are contiguous): one bulk copy per side
      mbar_expect_tx(&ms.bar_k[s], 2 * kplane);
      bulk_g2s(smem + kOffKB + s * kKBSide, p.kblob + ((size_t)u * 2 + s) * 2 * kplane, 2 * kplane, &ms.bar_k[s]);
    }
  }
  // query tile: norms (4 threads per pixel, 16 channels each) and fp16 hi/lo MN-major A operand
  {
    const int pq = tid & 127, cq = tid >> 7, pp = p0 + pq;
    float ss = 0.f;
    if (pp < HW) {
      const float* qp = p.qk + ((size_t)b * kCk + cq * 16) * HW + pp;
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const float t = __ldg(qp + (size_t)c * HW);
        ss = fmaf(t, t, ss);
      }
    }
    ms.ex_sum[cq][pq] = ss;
  }
  {
    const int c = tid >> 3;
    const float* qrow = p.qk + ((size_t)b * kCk + c) * HW;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int pg = (tid & 7) * 2 + j;
      const int pp0 = p0 + pg * 8;
      float t[8];
      if (((HW & 3) == 0) && pp0 + 7 < HW) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(qrow + pp0));
        const float4 bq = __ldg(reinterpret_cast<const float4*>(qrow + pp0) + 1);
        t[0] = a.x; t[1] = a.y; t[2] = a.z; t[3] = a.w; t[4] = bq.x; t[5] = bq.y; t[6] = bq.z; t[7] = bq.w;
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) t[e] = (pp0 + e < HW) ? __ldg(qrow + pp0 + e) : 0.f;
      }
      __align__(16) __half hi[8];
      __align__(16) __half lo[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) split_half(t[e], hi[e], lo[e]);
      const uint32_t off = (c % 8) * 16 + (c / 8) * 2048 + pg * 128;
      *reinterpret_cast<uint4*>(smem + kOffQH + off) = *reinterpret_cast<uint4*>(hi);
      *reinterpret_cast<uint4*>(smem + kOffQL + off) = *reinterpret_cast<uint4*>(lo);
    }
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (tid < kTP) ms.inv_nq[tid] = 1.f / (sqrtf(ms.ex_sum[0][tid] + ms.ex_sum[1][tid] + ms.ex_sum[2][tid] + ms.ex_sum[3][tid]) + kEpsNorm);
  const uint32_t tmem = ms.tmem_base;
  bool ok = true;
  ROT_STAMP();   // setup (query tile staged)

  // ---- scores: side s -> TMEM columns [256 s, 256 s + Lt) ---------------------------------------------------------------
  if (warp == 1) {                      // (converged warp, elected lane issues: see the PV product below)
    const uint32_t idesc = make_idesc(128, Lt, kFmtF16, kFmtF16, kMajorMN, kMajorK);
    constexpr uint32_t lbo_k = Lt * 16;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      ok = mbar_wait(&ms.bar_k[s], 0) && ok;
      const uint32_t kb = sbase + kOffKB + s * kKBSide;
      if (elect_one()) {
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
      __syncwarp();
    }
    if (elect_one()) mma_commit(&ms.bar_mma);
    __syncwarp();
  }
  if (tid == 0 && !mbar_wait(&ms.bar_mma, 0)) ms.abort_flag = 1;
  __syncthreads();
  tc_fence_after_sync();
  ROT_STAMP();   // scores done

  // the khat blobs are dead: start streaming nu k-steps into the ring (aliases them)
  const uint8_t* vsrc = p.vblob + ((size_t)u * 2 + h) * ks2 * kStageBytes;
  if (mma_thread) {
    fence_proxy_async_smem();
#pragma unroll
    for (int kk = 0; kk < kRing && kk < ks2; ++kk) {
      mbar_expect_tx(&ms.bar_full[kk], kStageBytes);
      bulk_g2s(smem + kOffRing + kk * kStageBytes, vsrc + (size_t)kk * kStageBytes, kStageBytes, &ms.bar_full[kk]);
    }
  }

  // ---- softmax epilogue: thread <-> (pixel px, side sd, chunks ch, ch + 2, ... of 32 columns) ------------------------------
  float inv_total;
  {
    float mx = -3.0e38f;
#pragma unroll
    for (int k = 0; k < nstep; ++k) {
      uint32_t r[32];
      tmem_ld32(tmem_addr(tmem, lane_base, sd * 256 + (2 * k + ch) * 32), r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(r[j]));
    }
    ms.ex_max[cb][px] = mx;
    bar_sync(1 + q, 128);                               // the 4 warps of this lane quadrant
    const float gm = fmaxf(fmaxf(ms.ex_max[0][px], ms.ex_max[1][px]), fmaxf(ms.ex_max[2][px], ms.ex_max[3][px]));   // inv_nq > 0: max of a * inv = inv * max a
    ROT_STAMP(); // max pass
    const float cw = ms.inv_nq[px] * p.c1s;
    const float bw = -gm * cw;
    const float fix = fast_exp2(fmaf(-gm, cw, -bw));    // exact residual of the rounded offset (see fused_em_res.cu)
    const bool own = (q >> 1) == h;                     // this pixel's top-l feature is computed by this CTA
    float* erow = etab + ((q & 1) * 32 + lane) * kERow + sd * kESide;
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < nstep; ++k) {
      const int c = 2 * k + ch;
      uint32_t r[32];
      tmem_ld32(tmem_addr(tmem, lane_base, sd * 256 + c * 32), r);
      tmem_ld_wait();
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float e0 = fast_exp2(fmaf(__uint_as_float(r[2 * j]), cw, bw)) * fix;
        const float e1 = fast_exp2(fmaf(__uint_as_float(r[2 * j + 1]), cw, bw)) * fix;
        const __half2 hh = __floats2half2_rn(e0 * kEScale, e1 * kEScale);
        const float2 back = __half22float2(hh);
        sum += back.x + back.y;                         // row sum of the ROUNDED operand
        pk[j] = *reinterpret_cast<const uint32_t*>(&hh);
        r[2 * j] = __float_as_uint(e0);
        r[2 * j + 1] = __float_as_uint(e1);
      }
      if (own) {
#pragma unroll
        for (int j = 0; j < 32; ++j) erow[c * 32 + j] = __uint_as_float(r[j]);
      }
      // packed chunk c lands on columns [16 c, 16 c + 16) of the side, i.e. inside score chunk c / 2 -- which one of the two
      // threads of this (pixel, side) read in step <= k: both have finished the reads of this step at the barrier
      bar_sync(5 + q * 2 + sd, 64);
      tmem_st16(tmem_addr(tmem, lane_base, sd * 256 + c * 16), pk);
    }
    tmem_st_wait();
    ms.ex_sum[cb][px] = sum;
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    inv_total = 1.f / ((ms.ex_sum[0][px] + ms.ex_sum[1][px]) + (ms.ex_sum[2][px] + ms.ex_sum[3][px]));
  }
  ROT_STAMP();   // exp pass + E packed

  if (warp == 1) {
    // ---- mem_out = E nu^T: A from TMEM (packed E), B from the ring; accumulators at columns 128.. and 384.. -----------------
    // The whole warp runs the loop converged and one ELECTED lane issues: addresses and descriptors stay in uniform registers.
    // (Issued from inside an `if (lane == 0)` branch every UTCHMMA was wrapped in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop, ~100
    //  instructions per k-step of a single thread that shares its scheduler with three sorting warps: the product ran at the
    //  pace of that thread, 820 cycles per k-step against 520 for its four MMAs; profiles/r2_readout_phases.txt.)
    const uint32_t idesc = make_idesc(128, 128, kFmtF16, kFmtF16, kMajorK, kMajorK);
    const uint64_t bdesc0 = make_sdesc(sbase + kOffRing, /*lbo*/ 4096, /*sbo*/ 128);
#pragma unroll 1
    for (int kk = 0; kk < ks2; ++kk) {
      const int st = kk % kRing;
      ok = mbar_wait(&ms.bar_full[st], (kk / kRing) & 1) && ok;
      tc_fence_after_sync();
      const uint32_t a_tmem = tmem + (kk / ks_side) * 256 + (kk % ks_side) * 8;
      const uint64_t bd = bdesc0 + ((uint32_t)(st * kStageBytes) >> 4);
      if (elect_one()) {
        mma_f16_ts(tmem + 128, a_tmem, bd, idesc, kk ? 1u : 0u);                            // E nu_hi, channels 0..127
        mma_f16_ts(tmem + 384, a_tmem, bd + (2048 >> 4), idesc, kk ? 1u : 0u);              //          channels 128..255
        mma_f16_ts(tmem + 128, a_tmem, bd + (8192 >> 4), idesc, 1u);                        // E nu_lo
        mma_f16_ts(tmem + 384, a_tmem, bd + ((8192 + 2048) >> 4), idesc, 1u);
        mma_commit(&ms.bar_empty[st]);
      }
      __syncwarp();
      if (kk >= kLag && kk - kLag + kRing < ks2) {
        const int prev = kk - kLag, nxt = prev + kRing;
        const int ps = prev % kRing;
        ok = mbar_wait(&ms.bar_empty[ps], (prev / kRing) & 1) && ok;
        if (elect_one()) {
          mbar_expect_tx(&ms.bar_full[ps], kStageBytes);
          bulk_g2s(smem + kOffRing + ps * kStageBytes, vsrc + (size_t)nxt * kStageBytes, kStageBytes, &ms.bar_full[ps]);
        }
        __syncwarp();
      }
    }
    if (elect_one()) mma_commit(&ms.bar_mma);
    if (!ok) ms.abort_flag = 1;
  } else {
    // ---- S: sorted top-l running sums of the own pixels (reference :198-208), one warp per pixel, under the PV product --------
    // Static split by scheduler (warp & 3), 16 pixels each: the scheduler that also hosts the issuing warp has three sorting
    // warps (6 + 5 + 5 pixels) instead of four (4 each) -- the phase is bound by the min / max issue rate per scheduler.  (Pixel
    // counts derived from threadIdx.y only: ptxas must see the loop as warp-uniform, or every shuffle below becomes a
    // WARPSYNC.COLLECTIVE call.)
    const int topl = p.topl;
    uint32_t* mytop = &ms.top[warp][0][0];
    const int sch = warp & 3, wj = warp >> 2;
    const int px_first = sch * 16 + (sch == 1 ? (wj == 1 ? 0 : wj == 2 ? 6 : 11) : wj * 4);
    const int px_count = sch == 1 ? (wj == 1 ? 6 : 5) : 4;
#pragma unroll 1
    for (int pi = 0; pi < ((p.dbg & 1) ? 0 : px_count); ++pi) {
      const int pl = px_first + pi;
      const int pp = p0 + h * kOwn + pl;
      if (pp >= HW) break;
      const float* row = etab + pl * kERow;
      float* srow = etab + pl * kERow;
      const int side = lane >> 4, l16 = lane & 15;        // lanes 0-15 sort side 0, lanes 16-31 side 1
      float a[R];                                     // packed words, compared as floats (topl.cuh)
#pragma unroll
      for (int k = 0; k < R; ++k) {
        const int i = l16 + 16 * k;                     // (any assignment of columns to lanes does: the word carries the column)
        a[k] = __uint_as_float(((__float_as_uint(row[side * kESide + i]) >> (IDXB - 1)) << IDXB) | (uint32_t)i);
      }
      if constexpr (Lt == 256) {                        // selection network: the order of the other 192 words is never established
        float t4[4];
        top64_of_256_half(a, l16, t4);
        *reinterpret_cast<uint4*>(mytop + side * 64 + top64_rank_base(l16)) =
            make_uint4(__float_as_uint(t4[0]), __float_as_uint(t4[1]), __float_as_uint(t4[2]), __float_as_uint(t4[3]));
      } else {
        sort_desc_half<R>(a, l16);
#pragma unroll
        for (int k = 0; k < R; ++k) {
          const int r = l16 * R + k;                    // rank r sits in lane16 r / R, register r % R
          if (r < 64) mytop[side * 64 + r] = __float_as_uint(a[k]);
        }
      }
      __syncwarp();
      // lane handles ranks lane and lane + 32: exact values, inclusive running sums over rank
      float c0[2], c1[2];
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int r = lane + 32 * hf;
        float x0 = 0.f, x1 = 0.f;
        if (r < topl) {
          x0 = row[mytop[r] & (Lt - 1)];
          x1 = row[kESide + (mytop[64 + r] & (Lt - 1))];
        }
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const float y0 = __shfl_up_sync(0xffffffffu, x0, o);
          const float y1 = __shfl_up_sync(0xffffffffu, x1, o);
          if (lane >= o) { x0 += y0; x1 += y1; }
        }
        c0[hf] = x0;
        c1[hf] = x1;
      }
      const float t0 = __shfl_sync(0xffffffffu, c0[0], 31), t1 = __shfl_sync(0xffffffffu, c1[0], 31);
      c0[1] += t0;
      c1[1] += t1;
      __syncwarp();                                     // every lane has read its exact values: the row is dead, S takes its place
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int r = lane + 32 * hf;
        if (r < topl) {
          const float f = c0[hf] / (c0[hf] + c1[hf]);
          srow[r] = f;
          srow[topl + r] = 1.f - f;
        }
      }
    }
  }
  ROT_STAMP();   // S written (this warp's pixels)
  if (tid == 0 && !mbar_wait(&ms.bar_mma, 1)) ms.abort_flag = 1;
  __syncthreads();
  tc_fence_after_sync();
  ROT_STAMP();   // PV done

  // ---- S of the own pixels: table rows -> the caller's buffer, coalesced in either layout ------------------------------------
  {
    const int topl2 = 2 * p.topl;
    const int npx = min(kOwn, HW - (p0 + h * kOwn));      // own pixels that exist
    if (p.pixel_major) {                                  // [U][HW][C]: a warp writes the 2 topl channels of one pixel
      for (int pl = warp; pl < npx; pl += 16) {
        float* o = p.out + ((size_t)u * HW + p0 + h * kOwn + pl) * p.out_channels + p.s_channel;
        for (int c = lane; c < topl2; c += 32) o[c] = etab[pl * kERow + c];
      }
    } else {                                              // [U][C][HW]: a warp writes 32 consecutive pixels of one channel
      for (int c = warp; c < topl2; c += 16) {
        float* o = p.out + ((size_t)u * p.out_channels + p.s_channel + c) * HW + p0 + h * kOwn;
        for (int pl = lane; pl < npx; pl += 32) o[pl] = etab[pl * kERow + c];
      }
    }
  }
  __syncthreads();                                        // (the store tiles below overlay the E table the loop above has just read)
  // ---- normalise and store mem_out: thread <-> (pixel px, 64 channels [64 cb, +64) of this half) ------------------------------
  {
    const float scale = inv_total;             // the 2^10 of E cancels against the row sum of the same operand
    const bool in_range = p0 + px < HW;
    const uint32_t tcol = 128 + (cb >> 1) * 256 + (cb & 1) * 64;
    const int ch0 = p.mem_channel + h * kDH + cb * 64;
    // (pixel-major) [32 pixels][64 channels + 4] transpose tile of this warp: the ring and the E table are both dead by now
    float* tbuf = reinterpret_cast<float*>(smem + kOffRing) + warp * (32 * kTileStride);
#pragma unroll
    for (int qq = 0; qq < 2; ++qq) {
      uint32_t r[32];
      tmem_ld32(tmem_addr(tmem, lane_base, tcol + qq * 32), r);
      tmem_ld_wait();
      if (p.pixel_major) {
        // [U][HW][C]: the thread's 32 channels of its pixel go into the tile as 8 x 16 bytes (row stride 272 bytes: conflict-free) ...
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4)
          *reinterpret_cast<float4*>(tbuf + lane * kTileStride + qq * 32 + j4 * 4) =
              make_float4(__uint_as_float(r[4 * j4]) * scale, __uint_as_float(r[4 * j4 + 1]) * scale, __uint_as_float(r[4 * j4 + 2]) * scale,
                          __uint_as_float(r[4 * j4 + 3]) * scale);
      } else if (in_range) {
        float* o = p.out + ((size_t)u * p.out_channels + ch0 + qq * 32) * HW + p0 + px;
#pragma unroll
        for (int j = 0; j < 32; ++j) o[(size_t)j * HW] = __uint_as_float(r[j]) * scale;
      }
    }
    if (p.pixel_major) {
      // ... and leave as 16-byte stores: half a warp writes the 256 contiguous bytes of one pixel, 16 instructions for the 32 pixels
      // (the first version wrote 4 bytes per lane: 64 store + 64 shared-load instructions per warp, store phase 4.4 us)
      __syncwarp();
      const int npx = min(32, HW - (p0 + q * 32));
      float* o = p.out + ((size_t)u * HW + p0 + q * 32) * p.out_channels + ch0 + (lane & 15) * 4;
#pragma unroll 4
      for (int it = 0; it < 16; ++it) {
        const int pl = it * 2 + (lane >> 4);
        const float4 v = *reinterpret_cast<const float4*>(tbuf + pl * kTileStride + (lane & 15) * 4);
        if (pl < npx) *reinterpret_cast<float4*>(o + (size_t)pl * p.out_channels) = v;
      }
    }
  }
  tc_fence_before_sync();
  const int bad = __syncthreads_or(ms.abort_flag);
  ROT_STAMP();   // stored
  if (p.prof != nullptr && blockIdx.x == 0 && tid == 0) p.prof[128] = n_stamp;
  if (warp == 0) tmem_dealloc(tmem, 512);
  if (bad) __trap();
}

// ------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------
bool fused_readout_topl_covers(const SwemDims& d) {
  const int Lt = d.L * d.n_banks;
  if (d.Ck != rot::kCk || d.Cv != rot::kCv || (Lt != 64 && Lt != 128 && Lt != 256) || d.topl < 1 || d.topl > 64 || d.topl > Lt) return false;
  const char* off = getenv("SWEM_RO_TOPL");
  return !(off != nullptr && off[0] == '0');             // A/B switch: SWEM_RO_TOPL=0 runs readout_fused_kernel + perm_inv_kernel
}

int fused_readout_topl_launch(const SwemReadArgs& a, const uint8_t* kblob, const uint8_t* vblob, cudaStream_t st) {
  const SwemDims& d = a.dims;
  const int U = d.B * d.N, Lt = d.L * d.n_banks;
  const int T = (d.HW + rot::kTP - 1) / rot::kTP;
  static PerDevice once;
  const int dev_id = current_device();
  {
    std::lock_guard<std::mutex> lock(once.mu);
    if (!once.done[dev_id]) {
      SWEM_CUDA(cudaFuncSetAttribute(readout_topl_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rot::kSmemBytes));
      SWEM_CUDA(cudaFuncSetAttribute(readout_topl_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rot::kSmemBytes));
      SWEM_CUDA(cudaFuncSetAttribute(readout_topl_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rot::kSmemBytes));
      once.done[dev_id] = true;
    }
  }
  ReadoutToplParams p{};
  p.qk = a.qk; p.kblob = kblob; p.vblob = vblob; p.out = a.out;
  p.N = d.N; p.HW = d.HW; p.T = T; p.out_channels = a.out_channels; p.mem_channel = a.mem_channel; p.s_channel = a.s_channel;
  p.pixel_major = a.out_pixel_major; p.topl = d.topl;
  p.c1s = kLog2e / (d.tau * rot::kKScale);
  p.prof = get_profile_buffer();
  {
    const char* dbg = getenv("SWEM_RO_DBG");
    p.dbg = dbg ? atoi(dbg) : 0;
  }
  if (Lt == 64) readout_topl_kernel<64><<<U * T * 2, dim3(32, 16, 1), rot::kSmemBytes, st>>>(p);
  else if (Lt == 128) readout_topl_kernel<128><<<U * T * 2, dim3(32, 16, 1), rot::kSmemBytes, st>>>(p);
  else readout_topl_kernel<256><<<U * T * 2, dim3(32, 16, 1), rot::kSmemBytes, st>>>(p);
  SWEM_LAUNCH_CHECK();
  return SWEM_OK;
}

}  // namespace swem
