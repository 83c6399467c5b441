// Fused sequential-weighted-EM kernel for sm_100a (tcgen05 + TMEM), one launch per memorize call.
//
// Covers the BASELINE shape family: Ck = 64, L = 128 bases per side, Cv = 512, any HW, any B*N.
// Reference semantics: methods/SWEM/modules.py:129-168 (swem), :112-120 (E), :122-127 (M),
// :93-110 (W), :164-165 (nu).  Arithmetic: operands fp16, with x and the unit bases split into
// hi + lo halves on the Ck contraction (3 MMAs: hi*hi + hi*lo + lo*hi ~ fp32-accurate logits),
// fp32 accumulation in TMEM, fp32 softmax / normalisation.
//
// Decomposition: one CTA per (unit u = (b,n), pixel tile of 128 px), both sides [bg | fg] = 256
// basis columns.  All CTAs of a launch are co-resident (grid <= #SMs, 1 CTA/SM); the M-step sum
// over pixel tiles is a bulk-async reduce-add of each CTA's partial into an L2-resident
// accumulator followed by a per-unit arrival counter; every CTA then reads the total back and
// re-derives the unit bases for its next E-step locally (no second exchange).
//
// Per EM iteration in a CTA (256 threads, thread <-> (pixel, side) in the epilogues and
// thread <-> basis row in the finalize):
//   1. a[p, sl] = x_p . khat_sl        12 x tcgen05.mma M128 N256 K16 (A = X^T MN-major, B = khat K-major)
//      The W-step logits l2norm(x).khat equal a / (||x_p|| + eps): the same accumulators serve both.
//   2. epilogue: W-step weights (iteration > 0) and per-side softmax * weight -> z (fp16) into
//      shared memory as the MN-major A operand of the M-step
//   3. [sum_p z x | sum_p z] = Z^T [X^T | 1]   32 x tcgen05.mma M128 N80/64 K16 per CTA
//   4. partial -> smem -> cp.reduce.async.bulk (add.f32) -> L2 accumulator; arrive; wait; bulk load total
//   5. kappa = (zita_ kappa_ + sum)/zita ; khat = l2norm(kappa) -> fp16 hi/lo K-major B operand
// After the last E-step: nu partial = Z^T V^T (two passes over Cv halves, V streamed through a
// 3-stage fp16 ring), reduce-added the same way, then each CTA normalises a slice of nu.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc05.cuh"
#include "fused_common.cuh"

namespace swem {

using namespace tc05;

namespace em {
constexpr int kTP = 128;    // pixels per CTA
constexpr int kCk = 64;
constexpr int kL = 128;     // bases per side
constexpr int kSL = 256;    // both sides
constexpr int kCv = 512;
constexpr int kAccRow = 73; // floats per row of the kappa accumulator blob: 64 kappa sums, 1 zita sum, pad (odd stride)
constexpr uint32_t kAccBytes = kSL * kAccRow * 4;  // 74752
constexpr float kKScale = 256.f;                   // khat is staged as khat*256 so its lo half stays a normal fp16
constexpr float kZScale = 16384.f;                 // z (<= 1) is staged as z*2^14: responsibilities down to ~4e-9 stay normal fp16

// ---- shared memory map (bytes) ---------------------------------------------------------------
// XH : [c 0..79][p] chunks  : byte = (c%8)*16 + (c/8)*2048 + (p/8)*128 + (p%8)*2   (rows 64 = ones, 65..79 = 0)
// XL : [c 0..63][p]
//      as E-step A (MN-major, r=p, k=c): SBO=128,  LBO=2048 ; as M-step B (K-major, r=c, k=p): SBO=2048, LBO=128
// KH/KL : [sl 0..255][c] K-major: byte = (sl%8)*16 + (sl/8)*128 + (c/8)*4096 + (c%8)*2  -> SBO=128, LBO=4096
// Z  : [sl 0..255][p] MN-major A: byte = (sl%8)*2 + (p%8)*16 + (sl/8)*2048 + (p/8)*128   -> SBO=2048, LBO=128
// P  : fp32 [256][73] staging of the M-step partial / total (aliases Z)
// VS : V stage [d 0..255][p 0..31] K-major, hi plane (16 KB) then lo plane: byte = (d%8)*16 + (d/8)*128 + (p/8)*4096 + (p%8)*2
//      -> SBO=128, LBO=4096.  Two stages of 32 KB: stage 0 at kOffVS, stage 1 in the X region (idle during the nu GEMMs)
// ZL : lo half of z, same layout as Z (aliases KH/KL: khat is dead between the logits GEMM and the finalize)
// NS : fp32 [2 sides][32 d][128 l] staging of nu partials (aliases XH/XL, dead after the last M-step GEMM)
constexpr uint32_t kOffXH = 0;
constexpr uint32_t kOffXL = kOffXH + 10 * 2048;
constexpr uint32_t kOffKH = kOffXL + 8 * 2048;
constexpr uint32_t kOffKL = kOffKH + 8 * 4096;
constexpr uint32_t kOffZ = kOffKL + 8 * 4096;
constexpr uint32_t kOffVS = kOffZ + kAccBytes;            // 74752 is a multiple of 128
constexpr uint32_t kVPlane = 4 * 4096;                    // 16 KB: one fp16 plane of a V stage
constexpr uint32_t kOffMisc = kOffVS + 3 * kVPlane;       // (48 KB reserved: stage 0 uses 32 KB, the drain staging 32 KB)
constexpr uint32_t kOffZL = kOffKH;                       // 64 KB
constexpr uint32_t kOffNS = kOffXH;                       // 32 KB of the 36 KB X region
struct Misc {
  float inv_nx[kTP];
  float mask[2][kTP];
  float ex_max[2][kTP];
  float ex_sum[2][kTP];
  float zita[kSL];
  uint64_t bar_mma;
  uint64_t bar_tma;
  uint64_t bar_stage[2];
  uint32_t tmem_base;
  int abort_flag;
};
constexpr uint32_t kSmemBytes = kOffMisc + sizeof(Misc) + 128;
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");

// TMEM columns
constexpr uint32_t kColE = 0;      // [128 px][256]   E / W logits
constexpr uint32_t kColM = 256;    // side s at 256 + 80*s : [128 sl][80]
constexpr uint32_t kColNu = 0;     // nu pass: side s at 256*s : [128 sl][256 d]
}  // namespace em

struct EmFusedParams {
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
  float* acc_k;        // [U][n_iters][256][73], zeroed before launch
  float* acc_nu;       // [U][2][512][128], zeroed before launch
  unsigned* counters;  // [U][n_iters + 1], zeroed before launch
  int* status;         // device error word (0 = ok)
  long long* prof;     // optional: phase time stamps (ns) of CTA 0, see swem_set_profile_buffer
  int N, HW, T, n_iters, u0;
  float c1s;           // log2(e) / (tau * kKScale): scales staged logits into exp2 arguments
};

// phase stamp: CTA 0 / thread 0 only, when a profile buffer is installed
#define EM_STAMP()                                                   \
  do {                                                               \
    if (p.prof != nullptr && blockIdx.x == 0 && tid == 0 && n_stamp < 250) p.prof[1 + n_stamp++] = global_ns(); \
  } while (0)

__global__ void __launch_bounds__(256, 1) em_fused_kernel(const EmFusedParams p) {
  using namespace em;
  extern __shared__ __align__(1024) uint8_t smem[];
  Misc& ms = *reinterpret_cast<Misc*>(smem + kOffMisc);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x % p.T;
  const int u = p.u0 + blockIdx.x / p.T;
  const int b = u / p.N;
  const int p0 = tile * kTP;
  const int HW = p.HW;
  const int I = p.n_iters;
  const uint32_t sbase = smem_u32(smem);
  int n_stamp = 0;
  EM_STAMP();

  // ---- one-time setup -------------------------------------------------------------------------
  if (warp == 0) tmem_alloc(&ms.tmem_base, 512);
  if (tid == 0) {
    mbar_init(&ms.bar_mma, 1);
    mbar_init(&ms.bar_tma, 1);
    for (int i = 0; i < 2; ++i) mbar_init(&ms.bar_stage[i], 1);
    ms.abort_flag = 0;
    fence_mbar_init();
  }
  // thread <-> basis row r = tid (side = r / 128, l = r % 128) in the finalize steps
  const int row_s = tid >> 7, row_l = tid & 127;
  const float zita_p = __ldg(p.zita_prior + ((size_t)u * 2 + row_s) * kL + row_l);
  const float* kprior = p.kappa_prior + (((size_t)u * 2 + row_s) * kCk) * kL + row_l;   // + c*kL
  // khat = l2norm(kappa) * 256 -> fp16 hi/lo rows of the K-major B operand (reference :115)
  auto stage_khat = [&](const float (&kap)[kCk]) {
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < kCk; ++c) ss = fmaf(kap[c], kap[c], ss);
    const float sc = kKScale / (sqrtf(ss) + kEpsNorm);
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      __align__(16) __half hi[8];
      __align__(16) __half lo[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) split_half(kap[g * 8 + e] * sc, hi[e], lo[e]);
      const uint32_t off = (tid % 8) * 16 + (tid / 8) * 128 + g * 4096;
      *reinterpret_cast<uint4*>(smem + kOffKH + off) = *reinterpret_cast<uint4*>(hi);
      *reinterpret_cast<uint4*>(smem + kOffKL + off) = *reinterpret_cast<uint4*>(lo);
    }
  };
  float kap0[kCk];                   // kappa^0 = prior: loads issued first, consumed after the X tile is staged
#pragma unroll
  for (int c = 0; c < kCk; ++c) kap0[c] = __ldg(kprior + (size_t)c * kL);
  // pixel norms + masks (thread <-> pixel)
  if (tid < kTP) {
    const int px = p0 + tid;
    float ss = 0.f;
    if (px < HW) {
      const float* xp = p.x + (size_t)b * kCk * HW + px;
#pragma unroll 8
      for (int c = 0; c < kCk; ++c) {
        const float t = __ldg(xp + (size_t)c * HW);
        ss = fmaf(t, t, ss);
      }
    }
    ms.inv_nx[tid] = 1.f / (sqrtf(ss) + kEpsNorm);
    ms.mask[0][tid] = px < HW ? __ldg(p.masks + ((size_t)u * 2 + 0) * HW + px) : 0.f;
    ms.mask[1][tid] = px < HW ? __ldg(p.masks + ((size_t)u * 2 + 1) * HW + px) : 0.f;
  }
  // X tile -> fp16 hi/lo chunks.  thread -> (channel c = tid/4 [+64 for the aug rows], 4 pixel-groups)
  {
    const int c = tid >> 2;                  // 0..63
    const float* xrow = p.x + ((size_t)b * kCk + c) * HW;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int pg = (tid & 3) * 4 + j;      // pixel group of 8
      __align__(16) __half hi[8];
      __align__(16) __half lo[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int px = p0 + pg * 8 + e;
        const float val = px < HW ? __ldg(xrow + px) : 0.f;
        split_half(val, hi[e], lo[e]);
      }
      const uint32_t off = (c % 8) * 16 + (c / 8) * 2048 + pg * 128;
      *reinterpret_cast<uint4*>(smem + kOffXH + off) = *reinterpret_cast<uint4*>(hi);
      *reinterpret_cast<uint4*>(smem + kOffXL + off) = *reinterpret_cast<uint4*>(lo);
    }
    // augmented rows 64..79 of XH: row 64 = 1 (column-sum of z -> zita), rows 65..79 = 0
    for (int i = tid; i < 16 * 16; i += 256) {
      const int r = 64 + (i >> 4), pg = i & 15;
      const __half one = __float2half_rn(r == 64 ? 1.f : 0.f);
      __align__(16) __half vals[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) vals[e] = one;
      *reinterpret_cast<uint4*>(smem + kOffXH + (r % 8) * 16 + (r / 8) * 2048 + pg * 128) = *reinterpret_cast<uint4*>(vals);
    }
  }
  stage_khat(kap0);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = ms.tmem_base;
  uint32_t ph_mma = 0, ph_tma = 0;   // mbarrier phase parities

  bool failed = false;               // block-uniform
  EM_STAMP();                        // setup done

  const uint32_t idesc_e = make_idesc(128, 256, kFmtF16, kFmtF16, kMajorMN, kMajorK);
  const uint32_t idesc_m80 = make_idesc(128, 80, kFmtF16, kFmtF16, kMajorMN, kMajorK);
  const uint32_t idesc_m64 = make_idesc(128, 64, kFmtF16, kFmtF16, kMajorMN, kMajorK);
  const uint32_t idesc_nu = make_idesc(128, 256, kFmtF16, kFmtF16, kMajorMN, kMajorK);

  for (int it = 0; it < I; ++it) {
    // KH / KL hold khat of the current kappa (staged before the loop / at the end of the previous iteration)
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();

    // ---- (1) logits GEMM -----------------------------------------------------------------------
    // (single-thread sections are written as warp 0 / lane 0 + __syncwarp so that the other lanes of
    //  warp 0 park at the warp barrier instead of spinning on an mbarrier in a divergent branch)
    if (warp == 0) {
      if (lane == 0) {
#pragma unroll
      for (int term = 0; term < 3; ++term) {
        const uint32_t xa = sbase + (term == 2 ? kOffXL : kOffXH);
        const uint32_t kb = sbase + (term == 1 ? kOffKL : kOffKH);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint64_t ad = make_sdesc(xa + kk * 2 * 2048, /*lbo*/ 2048, /*sbo*/ 128);
          const uint64_t bd = make_sdesc(kb + kk * 2 * 4096, /*lbo*/ 4096, /*sbo*/ 128);
          mma_f16_ss(tmem + kColE, ad, bd, idesc_e, (term | kk) ? 1u : 0u);
        }
      }
      mma_commit(&ms.bar_mma);
      }
      __syncwarp();
    }
    SWEM_CTA_WAIT(&ms.bar_mma, ph_mma, ms.abort_flag);
    ph_mma ^= 1;
    tc_fence_after_sync();
    EM_STAMP();                      // logits GEMM done

    // ---- (2) epilogue: thread <-> (pixel px, side sd) ----------------------------------------------
    {
      const int px = (warp & 3) * 32 + lane, sd = warp >> 2;
      float a[kL];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint32_t r[32];
        tmem_ld32(tmem_addr(tmem, (warp & 3) * 32, kColE + sd * kL + q * 32), r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) a[q * 32 + j] = __uint_as_float(r[j]);
      }
      float mx = a[0];
#pragma unroll
      for (int i = 1; i < kL; ++i) mx = fmaxf(mx, a[i]);
      const bool do_w = it > 0;
      if (do_w) ms.ex_max[sd][px] = mx;
      __syncthreads();
      if (do_w) {
        // W-step (reference :93-110): t = a * inv_nx ; joint max over both sides
        const float gm = fmaxf(ms.ex_max[0][px], ms.ex_max[1][px]);
        const float cw = ms.inv_nx[px] * p.c1s;
        float e = 0.f;
#pragma unroll
        for (int i = 0; i < kL; ++i) e += fast_exp2((a[i] - gm) * cw);
        ms.ex_sum[sd][px] = e;
      }
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < kL; ++i) {
        a[i] = fast_exp2((a[i] - mx) * p.c1s);
        sum += a[i];
      }
      __syncthreads();
      float w = ms.mask[sd][px];
      if (do_w) {
        const float e0 = ms.ex_sum[0][px], e1 = ms.ex_sum[1][px];
        w *= 1.f - (sd ? e1 : e0) / (e0 + e1);
      }
      const float scale = w / sum;
      const float zs = scale * kZScale;
      // z -> fp16 hi + lo, MN-major A operands: 16-byte chunk = 8 consecutive bases of this pixel
#pragma unroll
      for (int g = 0; g < kL / 8; ++g) {
        __align__(16) __half hi[8];
        __align__(16) __half lo[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) split_half(a[g * 8 + e] * zs, hi[e], lo[e]);
        const uint32_t off = (px % 8) * 16 + (px / 8) * 128 + (sd * 16 + g) * 2048;
        *reinterpret_cast<uint4*>(smem + kOffZ + off) = *reinterpret_cast<uint4*>(hi);
        *reinterpret_cast<uint4*>(smem + kOffZL + off) = *reinterpret_cast<uint4*>(lo);
      }
      if (p.z_last != nullptr && it == I - 1 && p0 + px < HW) {
        float4* dst = reinterpret_cast<float4*>(p.z_last + (((size_t)u * 2 + sd) * HW + p0 + px) * kL);
#pragma unroll
        for (int g = 0; g < kL / 4; ++g)
          dst[g] = make_float4(a[g * 4] * scale, a[g * 4 + 1] * scale, a[g * 4 + 2] * scale, a[g * 4 + 3] * scale);
      }
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();

    EM_STAMP();                      // epilogue done
    // ---- (3) M-step GEMM: [kappa sums | zita sum] per side ------------------------------------------
    if (warp == 0) {
      if (lane == 0) {
#pragma unroll
      for (int sd = 0; sd < 2; ++sd) {
        const uint32_t za = sbase + kOffZ + sd * 16 * 2048;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint64_t ad = make_sdesc(za + kk * 2 * 128, /*lbo*/ 128, /*sbo*/ 2048);
          const uint64_t al = make_sdesc(za + (kOffZL - kOffZ) + kk * 2 * 128, /*lbo*/ 128, /*sbo*/ 2048);
          const uint64_t bh = make_sdesc(sbase + kOffXH + kk * 2 * 128, /*lbo*/ 128, /*sbo*/ 2048);
          const uint64_t bl = make_sdesc(sbase + kOffXL + kk * 2 * 128, /*lbo*/ 128, /*sbo*/ 2048);
          mma_f16_ss(tmem + kColM + sd * 80, ad, bh, idesc_m80, kk ? 1u : 0u);   // z_hi x_hi (+ zita column)
          mma_f16_ss(tmem + kColM + sd * 80, ad, bl, idesc_m64, 1u);             // z_hi x_lo
          mma_f16_ss(tmem + kColM + sd * 80, al, bh, idesc_m80, 1u);             // z_lo x_hi
        }
      }
      mma_commit(&ms.bar_mma);
      }
      __syncwarp();
    }
    SWEM_CTA_WAIT(&ms.bar_mma, ph_mma, ms.abort_flag);
    ph_mma ^= 1;
    tc_fence_after_sync();

    EM_STAMP();                      // M GEMM done
    // partial of this tile, row r = tid: 64 kappa sums + zita sum (kept in registers for now)
    float part[kCk + 1];
    {
      uint32_t r[32];
      const uint32_t base = tmem_addr(tmem, (warp & 3) * 32, kColM + row_s * 80);
      tmem_ld32(base, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) part[j] = __uint_as_float(r[j]);
      tmem_ld32(base + 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) part[32 + j] = __uint_as_float(r[j]);
      uint32_t r16[16];
      tmem_ld16(base + 64, r16);
      tmem_ld_wait();
      part[64] = __uint_as_float(r16[0]);
    }
    tc_fence_before_sync();

    const bool last = (it == I - 1);
    if (last) {
      // ---- nu partial = Z^T V^T, two passes over value-channel halves (Z must still be intact) --------
      __syncthreads();
      tc_fence_after_sync();
      // V chunk (seq = half*4 + ch): [256 d][32 px] fp32.  A warp reads 4 rows x 128 B per instruction
      // (lane -> row lane/8, pixels 4*(lane%8)..+3); 8 instructions cover its 32 rows.  Loads of chunk
      // seq+1 are issued into registers before chunk seq is converted and staged (software prefetch).
      const int px4 = (lane & 7) * 4;
      const bool vec_ok = (HW & 3) == 0;
      auto load_chunk = [&](int seq, float4 (&buf)[8]) {
        const int half = seq >> 2, ch = seq & 3;
        const int pxg = p0 + ch * 32 + px4;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int dl = warp * 32 + j * 4 + (lane >> 3);
          const float* src = p.v + ((size_t)u * kCv + half * 256 + dl) * HW + pxg;
          float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
          if (vec_ok && pxg + 3 < HW) {
            f = __ldg(reinterpret_cast<const float4*>(src));
          } else {
            if (pxg < HW) f.x = __ldg(src);
            if (pxg + 1 < HW) f.y = __ldg(src + 1);
            if (pxg + 2 < HW) f.z = __ldg(src + 2);
            if (pxg + 3 < HW) f.w = __ldg(src + 3);
          }
          buf[j] = f;
        }
      };
      auto stage_ptr = [&](int st) -> uint8_t* { return smem + (st ? kOffXH : kOffVS); };
      auto store_chunk = [&](int st, const float4 (&buf)[8]) {
        uint8_t* stage = stage_ptr(st);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int dl = warp * 32 + j * 4 + (lane >> 3);
          __half h0, h1, h2, h3, l0, l1, l2, l3;
          split_half(buf[j].x, h0, l0);
          split_half(buf[j].y, h1, l1);
          split_half(buf[j].z, h2, l2);
          split_half(buf[j].w, h3, l3);
          const __half2 ha = __halves2half2(h0, h1), hb = __halves2half2(h2, h3);
          const __half2 la = __halves2half2(l0, l1), lb = __halves2half2(l2, l3);
          uint2 ph, pl;
          ph.x = *reinterpret_cast<const uint32_t*>(&ha);
          ph.y = *reinterpret_cast<const uint32_t*>(&hb);
          pl.x = *reinterpret_cast<const uint32_t*>(&la);
          pl.y = *reinterpret_cast<const uint32_t*>(&lb);
          const uint32_t off = (dl % 8) * 16 + (dl / 8) * 128 + (px4 / 8) * 4096 + (px4 % 8) * 2;
          *reinterpret_cast<uint2*>(stage + off) = ph;
          *reinterpret_cast<uint2*>(stage + kVPlane + off) = pl;
        }
      };
      float4 vbuf[2][8];
      load_chunk(0, vbuf[0]);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const int seq = half * 4 + ch;
          const int st = seq & 1;
          if (seq + 1 < 8) load_chunk(seq + 1, vbuf[(seq + 1) & 1]);
          if (seq >= 2) SWEM_CTA_WAIT(&ms.bar_stage[st], ((seq / 2) - 1) & 1, ms.abort_flag);   // stage reuse: its MMAs retired
          store_chunk(st, vbuf[seq & 1]);
          fence_proxy_async_smem();
          tc_fence_before_sync();
          __syncthreads();
          tc_fence_after_sync();
          if (warp == 0) {
            if (lane == 0) {
            const uint32_t vb = sbase + (st ? kOffXH : kOffVS);
#pragma unroll
            for (int sd = 0; sd < 2; ++sd) {
              const uint32_t za = sbase + kOffZ + sd * 16 * 2048;
#pragma unroll
              for (int kk = 0; kk < 2; ++kk) {
                const uint64_t ad = make_sdesc(za + (ch * 2 + kk) * 2 * 128, /*lbo*/ 128, /*sbo*/ 2048);
                const uint64_t al = make_sdesc(za + (kOffZL - kOffZ) + (ch * 2 + kk) * 2 * 128, /*lbo*/ 128, /*sbo*/ 2048);
                const uint64_t bh = make_sdesc(vb + kk * 2 * 4096, /*lbo*/ 4096, /*sbo*/ 128);
                const uint64_t bl = make_sdesc(vb + kVPlane + kk * 2 * 4096, /*lbo*/ 4096, /*sbo*/ 128);
                mma_f16_ss(tmem + kColNu + sd * 256, ad, bh, idesc_nu, (ch | kk) ? 1u : 0u);   // z_hi v_hi
                mma_f16_ss(tmem + kColNu + sd * 256, ad, bl, idesc_nu, 1u);                     // z_hi v_lo
                mma_f16_ss(tmem + kColNu + sd * 256, al, bh, idesc_nu, 1u);                     // z_lo v_hi
              }
            }
            mma_commit(&ms.bar_stage[st]);
            if (ch == 3) mma_commit(&ms.bar_mma);
            }
            __syncwarp();
          }
        }
        SWEM_CTA_WAIT(&ms.bar_mma, ph_mma, ms.abort_flag);
        ph_mma ^= 1;
        tc_fence_after_sync();
        EM_STAMP();                  // nu pass GEMMs done
        // drain: TMEM [side][128 l][256 d] -> smem [side][32 d][128 l] -> bulk reduce-add into acc_nu.  Two staging
        // buffers (the dead X region and the first V stage, both idle now): round q only waits for the reads of q-2.
        {
          const int sd = warp >> 2, l = (warp & 3) * 32 + lane;
#pragma unroll 1
          for (int q = 0; q < 8; ++q) {
            float* ns = reinterpret_cast<float*>(smem + ((q & 1) ? kOffVS : kOffNS));
            if (q >= 2) {
              if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
              __syncthreads();
            }
            {
              uint32_t r[32];
              tmem_ld32(tmem_addr(tmem, (warp & 3) * 32, kColNu + sd * 256 + q * 32), r);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) ns[(sd * 32 + j) * 128 + l] = __uint_as_float(r[j]);
            }
            fence_proxy_async_smem();
            __syncthreads();
            if (tid == 0) {
#pragma unroll
              for (int s2 = 0; s2 < 2; ++s2) {
                float* dst = p.acc_nu + (((size_t)u * 2 + s2) * kCv + half * 256 + q * 32) * kL;
                asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst),
                             "r"(smem_u32(ns + s2 * 32 * 128)), "r"(32 * 128 * 4)
                             : "memory");
              }
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
          }
          if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // staging may be reused now
        }
        tc_fence_before_sync();
        __syncthreads();
        tc_fence_after_sync();
        EM_STAMP();                  // nu pass drained
      }
      // drain the stage barriers' outstanding phases is unnecessary: the kernel ends after this phase
    }

    // ---- (4) cross-tile reduction of the M-step partial -----------------------------------------------
    __syncthreads();                       // Z is dead now (all MMAs reading it have completed): P may alias it
    {
      float* P = reinterpret_cast<float*>(smem + kOffZ);
#pragma unroll
      for (int c = 0; c <= kCk; ++c) P[tid * kAccRow + c] = part[c];
#pragma unroll
      for (int c = kCk + 1; c < kAccRow; ++c) P[tid * kAccRow + c] = 0.f;
    }
    fence_proxy_async_smem();
    __syncthreads();
    float* acc = p.acc_k + ((size_t)(u * I + it)) * (kSL * kAccRow);
    unsigned* counter = p.counters + (size_t)u * (I + 1) + it;
    // prior row of this thread's basis: issue the 64 loads now so they fly during the cross-tile wait
    float kpr[kCk];
#pragma unroll
    for (int c = 0; c < kCk; ++c) kpr[c] = __ldg(kprior + (size_t)c * kL);
    if (warp == 0) {
      if (lane == 0) {
      asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(acc),
                   "r"(sbase + kOffZ), "r"(kAccBytes)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      EM_STAMP();                    // partial reduce-added
      __threadfence();
      atomicAdd(counter, 1u);
      const bool arrived = wait_counter(counter, (unsigned)p.T);
      EM_STAMP();                    // all tiles arrived
      if (!arrived) ms.abort_flag = 1;
      __threadfence();
      asm volatile("fence.proxy.async;" ::: "memory");
      mbar_expect_tx(&ms.bar_tma, kAccBytes);
      bulk_g2s(smem + kOffZ, acc, kAccBytes, &ms.bar_tma);
      }
      __syncwarp();
    }
    SWEM_CTA_WAIT(&ms.bar_tma, ph_tma, ms.abort_flag);
    ph_tma ^= 1;
    if (ms.abort_flag) {
      if (tid == 0) atomicExch(p.status, 1 + it);
      failed = true;
      break;
    }
    EM_STAMP();                      // total loaded
    // ---- (5) finalize row r = tid from the prior (reference :125-126) ------------------------------------
    {
      const float* P = reinterpret_cast<const float*>(smem + kOffZ) + tid * kAccRow;
      constexpr float kInvZ = 1.f / kZScale;
      const float zita_cur = zita_p + P[kCk] * kInvZ;
      const float rz = 1.f / zita_cur;
      float kap[kCk];
#pragma unroll
      for (int c = 0; c < kCk; ++c) kap[c] = (zita_p * kpr[c] + P[c] * kInvZ) * rz;
      if (last) {
        ms.zita[tid] = zita_cur;
        if (tile == 0) {
          p.zita[((size_t)u * 2 + row_s) * kL + row_l] = zita_cur;
          float* kout = p.kappa + (((size_t)u * 2 + row_s) * kCk) * kL + row_l;
#pragma unroll
          for (int c = 0; c < kCk; ++c) kout[(size_t)c * kL] = kap[c];
        }
      } else {
        stage_khat(kap);
      }
    }
    __syncthreads();
    EM_STAMP();                      // finalize done
  }

  // nu = (zita_ nu_ + acc_nu / 2^14) / zita (reference :164-165) is applied by nu_finalize_kernel, launched right
  // after this kernel: stream order replaces a third cross-CTA wait here.
  tc_fence_before_sync();
  __syncthreads();
  EM_STAMP();                        // done
  if (p.prof != nullptr && blockIdx.x == 0 && tid == 0) p.prof[0] = n_stamp;
  if (warp == 0) tmem_dealloc(tmem, 512);
  if (failed || ms.abort_flag) __trap();   // surface a protocol time-out as a CUDA error, never as silent garbage
}

// nu[g][d][l] = (zita_prior[g][l] * nu_prior[g][d][l] + acc_nu[g][d][l] / 2^14) / zita[g][l],  g = (b, n, s)
__global__ void nu_finalize_kernel(const float* __restrict__ acc_nu, const float* __restrict__ nu_prior,
                                   const float* __restrict__ zita_prior, const float* __restrict__ zita,
                                   float* __restrict__ nu, int G) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;          // float4 index
  if (i >= G * em::kCv * em::kL / 4) return;
  const int l4 = i % (em::kL / 4);
  const int g = i / (em::kCv * em::kL / 4);
  const float4 a = __ldcg(reinterpret_cast<const float4*>(acc_nu) + i);
  const float4 pr = __ldg(reinterpret_cast<const float4*>(nu_prior) + i);
  const float4 zp = __ldg(reinterpret_cast<const float4*>(zita_prior) + g * (em::kL / 4) + l4);
  const float4 z = __ldg(reinterpret_cast<const float4*>(zita) + g * (em::kL / 4) + l4);
  constexpr float k = 1.f / em::kZScale;
  float4 o;
  o.x = (zp.x * pr.x + a.x * k) / z.x;
  o.y = (zp.y * pr.y + a.y * k) / z.y;
  o.z = (zp.z * pr.z + a.z * k) / z.z;
  o.w = (zp.w * pr.w + a.w * k) / z.w;
  reinterpret_cast<float4*>(nu)[i] = o;
}

// ------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------
static long long* g_prof = nullptr;
void set_profile_buffer(void* dev) { g_prof = static_cast<long long*>(dev); }
long long* get_profile_buffer() { return g_prof; }

static int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

// SWEM_EM_KERNEL=v1 keeps the first-generation kernel of this file (one CTA per tile, both sides); the default is the
// pair kernel of fused_em2.cu.
static bool use_v1() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SWEM_EM_KERNEL");
    v = (e != nullptr && e[0] == 'v' && e[1] == '1') ? 1 : 0;
  }
  return v == 1;
}

int launch_nu_finalize(const float* acc_nu, const float* nu_prior, const float* zita_prior, const float* zita, float* nu,
                       int G, cudaStream_t st) {
  const int n4 = G * em::kCv * em::kL / 4;
  nu_finalize_kernel<<<(n4 + 255) / 256, 256, 0, st>>>(acc_nu, nu_prior, zita_prior, zita, nu, G);
  SWEM_LAUNCH_CHECK();
  return SWEM_OK;
}

bool fused_em_supported(const SwemDims& d) {
  if (!use_v1()) return fused_em2_supported(d);
  if (d.Ck != em::kCk || d.L != em::kL || d.Cv != em::kCv || d.n_iters < 1 || d.n_iters > 16) return false;
  const int T = (d.HW + em::kTP - 1) / em::kTP;
  return T >= 1 && T <= 128;
}

size_t fused_em_workspace(const SwemDims& d) {
  if (!use_v1()) return fused_em2_workspace(d);
  const size_t U = (size_t)d.B * d.N;
  size_t bytes = 0;
  bytes += align_up(U * d.n_iters * em::kAccBytes, 256);
  bytes += align_up(U * 2 * em::kCv * em::kL * 4, 256);
  bytes += align_up(U * (d.n_iters + 1) * 4 + 4, 256);
  return bytes + 256;
}

int fused_em_forward(const SwemEmArgs& a, cudaStream_t st) {
  if (!use_v1()) return fused_em2_forward(a, st);
  const SwemDims& d = a.dims;
  const int U = d.B * d.N;
  const int T = (d.HW + em::kTP - 1) / em::kTP;
  Arena ws(a.workspace);
  float* acc_k = ws.take<float>((size_t)U * d.n_iters * em::kSL * em::kAccRow);
  float* acc_nu = ws.take<float>((size_t)U * 2 * em::kCv * em::kL);
  unsigned* counters = ws.take<unsigned>((size_t)U * (d.n_iters + 1) + 1);
  int* status = reinterpret_cast<int*>(counters + (size_t)U * (d.n_iters + 1));
  SWEM_CUDA(cudaMemsetAsync(a.workspace, 0, ws.off, st));
  count_launch();

  static bool attr_set = false;
  if (!attr_set) {
    SWEM_CUDA(cudaFuncSetAttribute(em_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)em::kSmemBytes));
    attr_set = true;
  }
  EmFusedParams p{};
  p.x = a.x; p.v = a.v; p.masks = a.masks;
  p.kappa_prior = a.kappa_prior; p.nu_prior = a.nu_prior; p.zita_prior = a.zita_prior;
  p.kappa = a.kappa; p.nu = a.nu; p.zita = a.zita; p.z_last = a.z_last;
  p.acc_k = acc_k; p.acc_nu = acc_nu; p.counters = counters; p.status = status;
  p.N = d.N; p.HW = d.HW; p.T = T; p.n_iters = d.n_iters;
  p.c1s = kLog2e / (d.tau * em::kKScale);
  p.prof = g_prof;
  // all CTAs of a launch spin on each other: keep every launch co-resident (<= 1 CTA per SM)
  const int units_per_launch = sm_count() / T > 0 ? sm_count() / T : 1;
  for (int u0 = 0; u0 < U; u0 += units_per_launch) {
    const int nu = (U - u0 < units_per_launch) ? (U - u0) : units_per_launch;
    p.u0 = u0;
    em_fused_kernel<<<nu * T, 256, em::kSmemBytes, st>>>(p);
    SWEM_LAUNCH_CHECK();
  }
  return launch_nu_finalize(acc_nu, a.nu_prior, a.zita_prior, a.zita, a.nu, U * 2, st);
}

}  // namespace swem
