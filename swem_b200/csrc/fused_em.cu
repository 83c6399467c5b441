// Fused sequential-weighted-EM kernel for sm_100a (tcgen05 + TMEM + bulk-async copies), one launch per memorize call:
// one CTA per (unit u = (b,n), pixel tile of 128 px, side s), the two sides of a tile forming a 2-CTA cluster.
//
// Covers Ck = 64 (BASELINE) or 128 (the reference's CLI default), L = 64 / 128 bases per side (64: half of the rows /
// columns are zero padding) and L = 256 / 512 (a side's bases spread over LB = L / 128 CTAs: clusters of 2 * LB CTAs
// that exchange their softmax statistics, see the epilogue), Cv = 512, any B*N, any HW whose clusters are co-resident
// (128 x 64 / ~34 / ~16 pixels for L <= 128 / 256 / 512 on a B200).  Reference semantics: methods/SWEM/modules.py:129-168 (swem), :112-120 (E), :122-127 (M), :93-110 (W),
// :164-165 (nu).  Arithmetic: operands fp16 with x, the unit bases, the responsibilities and v split into hi + lo
// halves (3 MMAs per product: hi*hi + hi*lo + lo*hi ~ fp32-accurate), fp32 accumulation in TMEM, fp32 softmax /
// normalisation.  The W-step logits l2norm(x).khat equal the E-step logits x.khat / (||x_p|| + eps) -- same khat --
// so one GEMM per iteration serves both steps.
//
// Decomposition (a 5-object 480p frame fills 130 of the 148 SMs; the first-generation kernel, one CTA per tile with
// both sides, filled 65 and took 101 us against 70 us -- profiles/r1_phases_em_v1.txt vs r1_phases_em_pair.txt):
//
//   * every per-CTA phase handles one side only: 128 logits columns, 128 basis rows, half the exps, half
//     of the M-step / nu partials that have to be reduce-added through L2;
//   * the only coupling between the sides -- the W-step's share of exp-affinity per side (:101-108) -- is
//     one (max, sum) pair per pixel exchanged through distributed shared memory and a cluster barrier;
//   * nu = Z^T V^T needs a single pass (its [128 bases][512 channels] accumulator is exactly the 512 TMEM
//     columns); during set-up each CTA converts half of its tile's V to fp16 hi/lo operand images in global
//     memory (L2-resident; one chunk in set-up, the others by warps 4-7 in the shadow of the cross-tile
//     reductions), and after the last E-step both CTAs stream all of them back with plain bulk-async copies
//     (3-stage ring: one thread issues the MMAs, another refills stages as they retire) -- no register staging
//     on the critical path, and the conversion is shared by the pair;
//   * nu is normalised in this kernel: the last M-step barrier also covers the nu reduce-adds, after it every
//     CTA finalises a slice of value channels (no separate kernel, no third cross-tile wait).
//
// Cross-tile sums go through L2-resident accumulators with a per-(unit, side, iteration) arrival counter (bounded
// spin, co-resident grid): M-step partials as fp32 reductions straight from registers (red.global.add), the 256 KB
// nu partial of a CTA as bulk reduce-adds from shared memory.
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc05.cuh"
#include "fused_common.cuh"

namespace swem {

using namespace tc05;

namespace em {
constexpr int kTP = 128;    // pixels per CTA
constexpr int kL = 128;     // bases per side = rows owned by one CTA
constexpr int kCv = 512;
constexpr float kKScale = 256.f;
constexpr float kZScale = 16384.f;
constexpr uint32_t kStageBytes = 32768;            // one V operand image: [256 d][32 px] fp16, hi plane then lo plane
constexpr uint32_t kVPlane = 16384;
constexpr int kChunks = 8;                         // images per tile: 2 channel halves x 4 pixel quarters
constexpr int kStages = 3;                         // bulk-copy ring, one image per stage

// ---- shared memory map (bytes), key channels CK = 64 (BASELINE) or 128 (the reference's CLI default) ------------
// XH : [c 0..CK+15][p] : byte = (c%8)*16 + (c/8)*2048 + (p/8)*128 + (p%8)*2      (row CK = ones, CK+1.. = 0)
// XL : [c 0..CK-1][p]   E-step A (MN-major, r=p, k=c): SBO=128, LBO=2048; M-step B (K-major, r=c, k=p): SBO=2048, LBO=128
// KH/KL : [l 0..127][c] K-major: byte = (l%8)*16 + (l/8)*128 + (c/8)*2048 + (c%8)*2   -> SBO=128, LBO=2048
// Z  : [l 0..127][p] MN-major A: byte = (l%8)*2 + (p%8)*16 + (l/8)*2048 + (p/8)*128   -> SBO=2048, LBO=128
// ZL : lo half of z (aliases KH/KL: khat is dead between the logits GEMM and the finalize)
// VS : ring of V operand images, each [d 0..255][p 0..31] K-major: byte = (d%8)*16 + (d/8)*128 + (p/8)*4096 + (p%8)*2
//      -> SBO=128, LBO=4096; hi plane then lo plane.  CK = 64: three stages of their own.  CK = 128 (X and khat twice as
//      large): one stage of its own + two in the X region, which is dead once the last M-step GEMM has completed.
//      The nu drain staging (2 x 32 KB fp32) uses the same two/three buffers.
template <int CK>
struct Lay {
  static constexpr int kG = CK / 8;                               // 16-byte channel groups
  static constexpr uint32_t kOffXH = 0;
  static constexpr uint32_t kOffXL = kOffXH + (kG + 2) * 2048;
  static constexpr uint32_t kOffKH = kOffXL + kG * 2048;
  static constexpr uint32_t kOffKL = kOffKH + kG * 2048;
  static constexpr uint32_t kOffZ = kOffKL + kG * 2048;
  static constexpr uint32_t kOffZL = kOffKH;
  static constexpr uint32_t kOffVS = kOffZ + 16 * 2048;
  static constexpr int kOwnStages = (CK == 64) ? 3 : 1;
  static constexpr uint32_t kOffMisc = kOffVS + kOwnStages * kStageBytes;
  static constexpr uint32_t kAccBytes = (CK + 1) * kL * 4;        // M-step accumulator of one (unit, iteration, side): [CK sums + zita sum][128 l]
  static_assert(2 * kG * 2048 >= 16 * 2048, "ZL must fit over KH/KL");
  static_assert(CK == 64 || (2 * kG + 2) * 2048 >= 2 * kStageBytes, "two ring stages must fit in the X region");
  __host__ __device__ static constexpr uint32_t stage_off(int st) {
    return (st < kOwnStages) ? kOffVS + st * kStageBytes : kOffXH + (st - kOwnStages) * kStageBytes;
  }
};
template <int LB>
struct Misc {
  float inv_nx[kTP];
  float mask[kTP];
  float hmax[2][kTP];
  float hsum[2][kTP];
  float hew[2][kTP];
  float2 mbox[LB == 1 ? 2 : 1][LB == 1 ? 2 : 1][LB == 1 ? kTP : 1];   // LB = 1: [iteration parity][side][pixel] = (side max of the logits, side sum of W-step exps)
  float mstat[LB > 1 ? 2 : 1][3][LB > 1 ? 2 * LB : 1][LB > 1 ? kTP : 1];   // LB > 1: [parity][max | E-step sum | W-step sum][cluster rank][pixel]
  uint64_t bar_mma;
  uint64_t bar_full[kStages];
  uint64_t bar_empty[kStages];
  uint32_t tmem_base;
  int abort_flag;
};
template <int CK, int LB>
constexpr uint32_t smem_bytes() { return Lay<CK>::kOffMisc + sizeof(Misc<LB>) + 128; }
static_assert(smem_bytes<64, 1>() <= 227 * 1024 && smem_bytes<128, 1>() <= 227 * 1024 && smem_bytes<64, 2>() <= 227 * 1024 &&
              smem_bytes<128, 2>() <= 227 * 1024 && smem_bytes<64, 4>() <= 227 * 1024 && smem_bytes<128, 4>() <= 227 * 1024,
              "shared memory budget");

// TMEM columns
constexpr uint32_t kColE = 0;      // [128 px][128]     E / W logits of this side
constexpr uint32_t kColM = 128;    // [128 l][CK + 16]  M-step sums
constexpr uint32_t kColNu = 0;     // [128 l][512 d]    nu sums (after the last M-step partial has been read out)
}  // namespace em

struct EmPairParams {
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
  uint8_t* vblob;        // [U][T][8][32 KB] scratch: operand images of V (written in set-up, read after the last E-step)
  float* acc_k;          // [U][n_iters][2][LB][CK + 1][128], zeroed before launch   (LB = basis blocks of 128 per side)
  float* acc_nu;         // [U][2][LB][512][128], zeroed before launch
  unsigned* counters;    // [U][n_iters][2][LB], zeroed before launch
  int* status;
  long long* prof;
  int N, HW, T, n_iters, u0;
  int windowed;          // 0: the whole EM in one co-resident launch (cross-tile waits on the arrival counters);
  int it_begin;          // 1: one launch per iteration -- this launch finalises iteration it_begin - 1 from its completed
                         //    accumulators, then runs iteration it_begin up to the reduce-adds (it_begin = n_iters: outputs only)
  int L;                 // bases per side in global memory (64 or 128); a CTA always works on 128 rows, the rest are zero
  float c1s;             // log2(e) / (tau * kKScale)
};

#define EM_STAMP()                                                  \
  do {                                                               \
    if (p.prof != nullptr && blockIdx.x == 0 && tid == 0 && n_stamp < 120) p.prof[1 + n_stamp++] = global_ns(); \
  } while (0)

// ------------------------------------------------------------------------------------------------------
// V -> fp16 hi/lo operand image of ONE chunk ([256 d][32 px] fp32 in, 32 KB out; pixels past HW are zero), done by
// NW warps (`w` = warp index inside the group).  All loads of a thread are in flight before the first conversion.
// ------------------------------------------------------------------------------------------------------
template <int NW>
__device__ __forceinline__ void convert_v_chunk(const float* __restrict__ vsrc /* [256][HW] rows of this channel half */,
                                                uint8_t* __restrict__ image, int px_base, int HW, int w, int lane) {
  using namespace em;
  constexpr int J = 256 / (NW * 8);                     // rows of 8 channels per warp
  const int g = lane & 3;                               // group of 8 pixels
  const int px0 = px_base + g * 8;
  const bool vec = ((HW & 3) == 0) && (px0 + 7 < HW);
  float f[J][8];
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int d = w * (J * 8) + j * 8 + (lane >> 2);    // 0..255
    const float* src = vsrc + (size_t)d * HW + px0;
    if (vec) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(src));
      const float4 b = __ldg(reinterpret_cast<const float4*>(src) + 1);
      f[j][0] = a.x; f[j][1] = a.y; f[j][2] = a.z; f[j][3] = a.w; f[j][4] = b.x; f[j][5] = b.y; f[j][6] = b.z; f[j][7] = b.w;
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) f[j][e] = (px0 + e < HW) ? __ldg(src + e) : 0.f;
    }
  }
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int d = w * (J * 8) + j * 8 + (lane >> 2);
    __align__(16) __half hi[8];
    __align__(16) __half lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) split_half(f[j][e], hi[e], lo[e]);
    const uint32_t off = (d % 8) * 16 + (d / 8) * 128 + g * 4096;
    *reinterpret_cast<uint4*>(image + off) = *reinterpret_cast<uint4*>(hi);
    *reinterpret_cast<uint4*>(image + kVPlane + off) = *reinterpret_cast<uint4*>(lo);
  }
}

// Same image from a pixel-major (NHWC) value tensor: `vsrc` = [HW][512] rows of the unit, channels [dbase, dbase + 256).
// A warp step takes 8 pixels x 128 channels: 8 float4 loads per lane (512 contiguous bytes per pixel and warp), the lane
// then owns 4 channels x 8 pixels = four 16-byte operand units per plane, contiguous across the warp.
template <int NW>
__device__ __forceinline__ void convert_v_chunk_nhwc(const float* __restrict__ vsrc, int dbase, uint8_t* __restrict__ image,
                                                     int px_base, int HW, int w, int lane) {
  using namespace em;
#pragma unroll
  for (int step = w; step < 8; step += NW) {
    const int g = step >> 1, dh = step & 1;               // pixel group of 8, half of the image's 256 channels
    const int d = dh * 128 + lane * 4;
    float4 f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int px = px_base + g * 8 + e;
      f[e] = px < HW ? __ldg(reinterpret_cast<const float4*>(vsrc + (size_t)px * kCv + dbase + d)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    uint8_t* dst = image + d * 16 + g * 4096;             // (d % 8) * 16 + (d / 8) * 128 = d * 16
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      __align__(16) __half hi[8];
      __align__(16) __half lo[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) split_half(k == 0 ? f[e].x : k == 1 ? f[e].y : k == 2 ? f[e].z : f[e].w, hi[e], lo[e]);
      *reinterpret_cast<uint4*>(dst + k * 16) = *reinterpret_cast<uint4*>(hi);
      *reinterpret_cast<uint4*>(dst + kVPlane + k * 16) = *reinterpret_cast<uint4*>(lo);
    }
  }
}

// ------------------------------------------------------------------------------------------------------
// CK key channels; LB basis blocks of 128 per side (L = 64 / 128: 1, L = 256: 2, L = 512: 4) -> cluster of 2 * LB CTAs;
// VPM: value features pixel-major (NHWC)
template <int CK, int LB, bool VPM>
__global__ void __cluster_dims__(2 * LB, 1, 1) __launch_bounds__(256, 1) em_pair_kernel(const EmPairParams p) {
  using namespace em;
  using LY = Lay<CK>;
  constexpr int kCk = CK;
  constexpr int CS = 2 * LB;
  constexpr uint32_t kOffXH = LY::kOffXH, kOffXL = LY::kOffXL, kOffKH = LY::kOffKH, kOffKL = LY::kOffKL, kOffZ = LY::kOffZ,
                     kOffZL = LY::kOffZL, kOffMisc = LY::kOffMisc;
  extern __shared__ __align__(1024) uint8_t smem[];
  Misc<LB>& ms = *reinterpret_cast<Misc<LB>*>(smem + kOffMisc);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rank = (int)cluster_ctarank();
  const int sd = rank & 1;                              // side handled by this CTA (0 = background, 1 = foreground)
  const int lb = rank >> 1;                             // block of 128 bases of that side
  const int pair = blockIdx.x / CS;
  const int tile = pair % p.T;
  const int u = p.u0 + pair / p.T;
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
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&ms.bar_full[i], 1);
      mbar_init(&ms.bar_empty[i], 1);
    }
    ms.abort_flag = 0;
    fence_mbar_init();
  }
  // V operand images: this CTA converts channel half `sd` of the tile = 4 chunks (the peer converts the other half); both
  // read all 8 images back through the bulk-copy ring after the last E-step.  Only as many chunks as cannot be hidden
  // later are converted here by the whole CTA; the rest is done one per iteration by warps 4-7 while the row threads
  // (warps 0-3) run the cross-tile reduction and the finalize -- the conversion is HBM-bound and would otherwise be
  // 6 us of exposed set-up.  The W-step cluster barrier of the last iteration publishes all of them to the pair.
  uint8_t* const vimg = p.vblob + ((size_t)u * p.T + tile) * kChunks * kStageBytes;
  constexpr int kMine = kChunks / CS;                   // images converted by this CTA: rank * kMine + j, image c = (half c/4, quarter c%4)
  auto convert_mine = [&](int j, bool whole_cta) {
    const int c = rank * kMine + j;
    if constexpr (VPM) {
      const float* src = p.v + (size_t)u * HW * kCv;
      if (whole_cta) convert_v_chunk_nhwc<8>(src, (c >> 2) * 256, vimg + (size_t)c * kStageBytes, p0 + (c & 3) * 32, HW, warp, lane);
      else convert_v_chunk_nhwc<4>(src, (c >> 2) * 256, vimg + (size_t)c * kStageBytes, p0 + (c & 3) * 32, HW, warp - 4, lane);
    } else {
      const float* src = p.v + ((size_t)u * kCv + (c >> 2) * 256) * HW;
      if (whole_cta) convert_v_chunk<8>(src, vimg + (size_t)c * kStageBytes, p0 + (c & 3) * 32, HW, warp, lane);
      else convert_v_chunk<4>(src, vimg + (size_t)c * kStageBytes, p0 + (c & 3) * 32, HW, warp - 4, lane);
    }
  };
  int chunks_done = (I - 1 >= kMine - 1) ? 1 : kMine - (I - 1);   // what cannot be hidden behind iterations 0 .. I-2 is done here
  const bool windowed = p.windowed != 0;
  const bool fin_only = windowed && p.it_begin >= I;              // windowed: outputs-only launch
  if (windowed) chunks_done = (p.it_begin == I - 1) ? kMine : 0;  // the launch of the last iteration converts its images
  for (int j = 0; j < chunks_done; ++j) convert_mine(j, true);
  __threadfence();
  asm volatile("fence.proxy.async;" ::: "memory");     // generic-proxy global stores -> visible to the bulk-copy (async proxy) reads
  // rows: thread tid < 128 <-> basis l = tid of side sd (finalize steps)
  const bool row_thread = tid < kL;
  const int L = p.L;                                    // L = 64: rows / columns 64..127 of every operand are zero padding
  const int lrow = lb * kL + tid;                       // this row thread's basis inside the side
  const bool valid_row = lrow < L;
  const int gs = u * 2 + sd;                            // (b, n, s) index
  const int gsl = gs * LB + lb;                         // (b, n, s, basis block): index of this CTA's accumulators
  const float zita_p = valid_row ? __ldg(p.zita_prior + (size_t)gs * L + lrow) : 0.f;
  const float* kprior = p.kappa_prior + ((size_t)gs * kCk) * L + (valid_row ? lrow : 0);   // + c*L
  auto stage_khat = [&](const float (&kap)[kCk]) {      // khat = l2norm(kappa) * 256 -> fp16 hi/lo K-major rows (reference :115)
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < kCk; ++c) ss = fmaf(kap[c], kap[c], ss);
    const float sc = kKScale / (sqrtf(ss) + kEpsNorm);
#pragma unroll
    for (int g = 0; g < kCk / 8; ++g) {
      __align__(16) __half hi[8];
      __align__(16) __half lo[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) split_half(kap[g * 8 + e] * sc, hi[e], lo[e]);
      const uint32_t off = (tid % 8) * 16 + (tid / 8) * 128 + g * 2048;
      *reinterpret_cast<uint4*>(smem + kOffKH + off) = *reinterpret_cast<uint4*>(hi);
      *reinterpret_cast<uint4*>(smem + kOffKL + off) = *reinterpret_cast<uint4*>(lo);
    }
  };
  float kap0[kCk];
  if (row_thread) {
#pragma unroll
    for (int c = 0; c < kCk; ++c) kap0[c] = valid_row ? __ldg(kprior + (size_t)c * L) : 0.f;
    if (windowed && p.it_begin > 0) {
      // kappa of iteration it_begin - 1 from the prior and the totals the previous launch left in acc_k (reference :125-126)
      constexpr float kInvZ = 1.f / kZScale;
      const float* accp = p.acc_k + ((size_t)((u * I + p.it_begin - 1) * 2 + sd) * LB + lb) * ((kCk + 1) * kL);
      const float zita_cur = zita_p + __ldcg(accp + kCk * kL + tid) * kInvZ;
      const float rz = valid_row ? 1.f / zita_cur : 0.f;
#pragma unroll
      for (int c = 0; c < kCk; ++c) kap0[c] = (zita_p * kap0[c] + __ldcg(accp + c * kL + tid) * kInvZ) * rz;
      if (fin_only) {
        ms.hsum[0][tid] = rz;         // for the nu slice
        ms.hsum[1][tid] = zita_p;
        if (tile == 0 && valid_row) {
          p.zita[(size_t)gs * L + lrow] = zita_cur;
          float* kout = p.kappa + ((size_t)gs * kCk) * L + lrow;
#pragma unroll
          for (int c = 0; c < kCk; ++c) kout[(size_t)c * L] = kap0[c];
        }
      }
    }
  }
  // pixel norms + this side's mask (threads 128..255 <-> pixel, so they overlap with the prior loads of the row threads)
  if (!row_thread && !fin_only) {
    const int q = tid - kL, px = p0 + q;
    float ss = 0.f;
    if (px < HW) {
      const float* xp = p.x + (size_t)b * kCk * HW + px;
#pragma unroll 8
      for (int c = 0; c < kCk; ++c) {
        const float t = __ldg(xp + (size_t)c * HW);
        ss = fmaf(t, t, ss);
      }
    }
    ms.inv_nx[q] = 1.f / (sqrtf(ss) + kEpsNorm);
    ms.mask[q] = px < HW ? __ldg(p.masks + (size_t)gs * HW + px) : 0.f;
  }
  // X tile -> fp16 hi/lo chunks.  thread -> (channel c = tid/4 (+64), 4 pixel groups of 8)
#pragma unroll
  for (int cc = 0; cc < (fin_only ? 0 : kCk / 64); ++cc) {
    const int c = cc * 64 + (tid >> 2);
    const float* xrow = p.x + ((size_t)b * kCk + c) * HW;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int pg = (tid & 3) * 4 + j;
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
  }
  auto stage_aug_rows = [&]() {       // augmented rows CK..CK+15 of XH: row CK = 1 (-> zita), rest 0
    for (int i = tid; i < 16 * 16; i += 256) {
      const int r = kCk + (i >> 4), pg = i & 15;
      const __half one = __float2half_rn(r == kCk ? 1.f : 0.f);
      __align__(16) __half vals[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) vals[e] = one;
      *reinterpret_cast<uint4*>(smem + kOffXH + (r % 8) * 16 + (r / 8) * 2048 + pg * 128) = *reinterpret_cast<uint4*>(vals);
    }
  };
  stage_aug_rows();
  if (row_thread) stage_khat(kap0);
  tc_fence_before_sync();
  cluster_arrive();                   // both CTAs of the pair are running before any remote shared-memory store,
  cluster_wait();                     // and both halves of the tile's V images are written
  tc_fence_after_sync();
  auto load_image = [&](int k) {       // ring stage k % kStages <- image k of the tile
    const int st = k % kStages;
    mbar_expect_tx(&ms.bar_full[st], kStageBytes);
    bulk_g2s(smem + LY::stage_off(st), vimg + (size_t)k * kStageBytes, kStageBytes, &ms.bar_full[st]);
  };
  // the images that have a stage of their own start flying early (they are consumed after the last E-step); stages that
  // live in the X region (CK = 128) are filled by the consumer once the last M-step GEMM has released it
  auto prefetch_images = [&]() {
    asm volatile("fence.proxy.async;" ::: "memory");
    for (int k = 0; k < LY::kOwnStages; ++k) load_image(k);
  };
  if (LB == 1 && I == 1 && tid == 0 && !fin_only) prefetch_images();   // (LB > 1: every iteration has a cluster barrier, see the epilogue)
  const uint32_t tmem = ms.tmem_base;
  uint32_t ph_mma = 0;
  bool failed = false;
  EM_STAMP();                        // setup done

  const uint32_t idesc_e = make_idesc(128, 128, kFmtF16, kFmtF16, kMajorMN, kMajorK);
  const uint32_t idesc_mhi = make_idesc(128, kCk + 16, kFmtF16, kFmtF16, kMajorMN, kMajorK);
  const uint32_t idesc_mlo = make_idesc(128, kCk, kFmtF16, kFmtF16, kMajorMN, kMajorK);
  const uint32_t idesc_nu = make_idesc(128, 256, kFmtF16, kFmtF16, kMajorMN, kMajorK);
  const uint32_t peer_mbox = map_to_peer(smem_u32(&ms.mbox[0][0][0]), (uint32_t)(rank ^ 1));   // (LB = 1 only)

  auto nu_slice = [&]() {
    // ---- nu = (zita_ nu_ + sum / 2^14) / zita (reference :164-165) for this tile's slice of value channels: the counter
    // wait above ordered every tile's nu reduce-adds (completed before its arrival) before these loads.
    const int dper = (kCv + p.T - 1) / p.T;
    const int d0 = tile * dper, d1 = min(kCv, d0 + dper);
    constexpr float kInvZ = 1.f / kZScale;
    const float4* acc4 = reinterpret_cast<const float4*>(p.acc_nu + (size_t)gsl * kCv * kL);    // [d][128] of this basis block
    const float4* pri4 = reinterpret_cast<const float4*>(p.nu_prior + (size_t)gs * kCv * L);    // [d][L]
    float4* out4 = reinterpret_cast<float4*>(p.nu + (size_t)gs * kCv * L);
    const int l4n = (L < kL ? L : kL) / 4;              // float4 columns of this block that exist
    for (int k = d0 * l4n + tid; k < d1 * l4n; k += 256) {
      const int d = k / l4n, l4 = k % l4n, l = l4 * 4;
      const int i = d * (L / 4) + lb * (kL / 4) + l4;   // position in the [d][L] tensors
      const float4 a = __ldcg(acc4 + d * (kL / 4) + l4);
      const float4 pr = __ldg(pri4 + i);
      float4 o;
      o.x = (ms.hsum[1][l + 0] * pr.x + a.x * kInvZ) * ms.hsum[0][l + 0];
      o.y = (ms.hsum[1][l + 1] * pr.y + a.y * kInvZ) * ms.hsum[0][l + 1];
      o.z = (ms.hsum[1][l + 2] * pr.z + a.z * kInvZ) * ms.hsum[0][l + 2];
      o.w = (ms.hsum[1][l + 3] * pr.w + a.w * kInvZ) * ms.hsum[0][l + 3];
      out4[i] = o;
    }
    EM_STAMP();                    // nu slice written
  };
  const int it_first = windowed ? p.it_begin : 0, it_stop = windowed ? (fin_only ? I : p.it_begin + 1) : I;
  for (int it = it_first; it < it_stop; ++it) {
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();

    // ---- (1) logits of this side: a[p, l] = x_p . khat_l ---------------------------------------------
    if (warp == 0) {
      if (lane == 0) {
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
      __syncwarp();
    }
    SWEM_CTA_WAIT(&ms.bar_mma, ph_mma, ms.abort_flag);
    ph_mma ^= 1;
    tc_fence_after_sync();
    EM_STAMP();                      // logits GEMM done

    // ---- (2) epilogue: thread <-> (pixel px, half hb of this side's bases) ------------------------------
    {
      const int px = (warp & 3) * 32 + lane, hb = warp >> 2;
      const bool do_w = it > 0;
      float a[64];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        uint32_t r[32];
        tmem_ld32(tmem_addr(tmem, (warp & 3) * 32, kColE + hb * 64 + q * 32), r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) a[q * 32 + j] = __uint_as_float(r[j]);
      }
      const bool active = lb * kL + hb * 64 < L;        // L = 64: the second half of the columns is padding
      if (!active) {
#pragma unroll
        for (int i = 0; i < 64; ++i) a[i] = -3.0e38f;   // exp2((a - max) * c) underflows to exactly 0 everywhere below
      }
      float mx = a[0];
#pragma unroll
      for (int i = 1; i < 64; ++i) mx = fmaxf(mx, a[i]);
      ms.hmax[hb][px] = mx;
      __syncthreads();
      mx = fmaxf(ms.hmax[0][px], ms.hmax[1][px]);       // max over this side's 128 bases
      // W-step (reference :93-110) works on t = a * inv_nx with the max over BOTH sides; each side sums its exps
      // against its own max and the pair rescales after the exchange: exp(t - M) = exp(t - m_s) * exp(m_s - M).
      const float cw = ms.inv_nx[px] * p.c1s;
      const int par = it & 1;
      float sum = 0.f, w = ms.mask[px];
      if constexpr (LB == 1) {
        if (do_w) {                                     // W-step sums first: the exchange flies while the E-step exps run
          float e = 0.f;
#pragma unroll
          for (int i = 0; i < 64; ++i) e += fast_exp2((a[i] - mx) * cw);
          ms.hew[hb][px] = e;
          __syncthreads();
          if (hb == 0) {
            const float es = ms.hew[0][px] + ms.hew[1][px];
            ms.mbox[par][sd][px] = make_float2(mx, es);
            st_cluster_f2(peer_mbox + (uint32_t)(((par * 2 + sd) * kTP + px) * sizeof(float2)), mx, es);
          }
          cluster_arrive();
        }
#pragma unroll
        for (int i = 0; i < 64; ++i) {
          a[i] = fast_exp2((a[i] - mx) * p.c1s);
          sum += a[i];
        }
        ms.hsum[hb][px] = sum;
        __syncthreads();
        sum = ms.hsum[0][px] + ms.hsum[1][px];
        if (do_w) {
          cluster_wait();
          if (it == I - 1 && tid == 0) prefetch_images();  // every chunk of the pair was converted before this barrier
          const float2 m0 = ms.mbox[par][0][px], m1 = ms.mbox[par][1][px];
          const float gm = fmaxf(m0.x, m1.x);
          const float e0 = m0.y * fast_exp2((m0.x - gm) * cw), e1 = m1.y * fast_exp2((m1.x - gm) * cw);
          w *= 1.f - (sd ? e1 : e0) / (e0 + e1);
        }
      } else {
        // L = 128 * LB: the bases of a side are spread over LB CTAs.  Every CTA works against the max m of ITS 128
        // columns, publishes (m, E-step sum, W-step sum) to the whole cluster, and rescales after one barrier:
        //   softmax denominator of the side = sum_c s_c 2^((m_c - M_side) c1),  z *= 2^((m_own - M_side) c1)
        float e = 0.f;
        if (do_w) {
#pragma unroll
          for (int i = 0; i < 64; ++i) e += fast_exp2((a[i] - mx) * cw);
        }
#pragma unroll
        for (int i = 0; i < 64; ++i) {
          a[i] = fast_exp2((a[i] - mx) * p.c1s);
          sum += a[i];
        }
        ms.hsum[hb][px] = sum;
        ms.hew[hb][px] = e;
        __syncthreads();
        {
          const float sc = ms.hsum[0][px] + ms.hsum[1][px], ec = ms.hew[0][px] + ms.hew[1][px];
          const uint32_t slot = smem_u32(&ms.mstat[par][0][rank][px]);
          constexpr uint32_t kPlane = CS * kTP * sizeof(float);
          for (int rr = hb; rr < CS; rr += 2) {           // the two threads of a pixel share the CS destinations
            const uint32_t dst = map_to_peer(slot, (uint32_t)rr);
            st_cluster_f32(dst, mx);
            st_cluster_f32(dst + kPlane, sc);
            st_cluster_f32(dst + 2 * kPlane, ec);
          }
        }
        cluster_arrive();
        cluster_wait();
        if (it == I - 1 && tid == 0) prefetch_images();    // every chunk of the cluster was converted before this barrier
        float Ms = -3.0e38f, gm = -3.0e38f;
#pragma unroll
        for (int rr = 0; rr < CS; ++rr) {
          const float m = ms.mstat[par][0][rr][px];
          gm = fmaxf(gm, m);
          if ((rr & 1) == sd) Ms = fmaxf(Ms, m);
        }
        float S = 0.f, e0 = 0.f, e1 = 0.f;
#pragma unroll
        for (int rr = 0; rr < CS; ++rr) {
          const float tm = ms.mstat[par][0][rr][px];
          if ((rr & 1) == sd) S += ms.mstat[par][1][rr][px] * fast_exp2((tm - Ms) * p.c1s);
          const float ew = ms.mstat[par][2][rr][px] * fast_exp2((tm - gm) * cw);
          if (rr & 1) e1 += ew; else e0 += ew;
        }
        if (do_w) w *= 1.f - (sd ? e1 : e0) / (e0 + e1);
        sum = S;
        w *= fast_exp2((mx - Ms) * p.c1s);                 // this CTA's exps were taken against its own max
      }
      const float scale = w / sum;
      const float zs = scale * kZScale;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        __align__(16) __half hi[8];
        __align__(16) __half lo[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) split_half(a[g * 8 + k] * zs, hi[k], lo[k]);
        const uint32_t off = (px % 8) * 16 + (px / 8) * 128 + (hb * 8 + g) * 2048;
        *reinterpret_cast<uint4*>(smem + kOffZ + off) = *reinterpret_cast<uint4*>(hi);
        *reinterpret_cast<uint4*>(smem + kOffZL + off) = *reinterpret_cast<uint4*>(lo);
      }
      if (p.z_last != nullptr && it == I - 1 && p0 + px < HW && active) {
        float4* dst = reinterpret_cast<float4*>(p.z_last + ((size_t)gs * HW + p0 + px) * L + lb * kL + hb * 64);
#pragma unroll
        for (int g = 0; g < 16; ++g)
          dst[g] = make_float4(a[g * 4] * scale, a[g * 4 + 1] * scale, a[g * 4 + 2] * scale, a[g * 4 + 3] * scale);
      }
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    EM_STAMP();                      // epilogue done

    // ---- (3) M-step GEMM: [sum_p z x | sum_p z] for this side's 128 bases --------------------------------
    if (warp == 0) {
      if (lane == 0) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint64_t ad = make_sdesc(sbase + kOffZ + kk * 2 * 128, /*lbo*/ 128, /*sbo*/ 2048);
          const uint64_t al = make_sdesc(sbase + kOffZL + kk * 2 * 128, /*lbo*/ 128, /*sbo*/ 2048);
          const uint64_t bh = make_sdesc(sbase + kOffXH + kk * 2 * 128, /*lbo*/ 128, /*sbo*/ 2048);
          const uint64_t bl = make_sdesc(sbase + kOffXL + kk * 2 * 128, /*lbo*/ 128, /*sbo*/ 2048);
          mma_f16_ss(tmem + kColM, ad, bh, idesc_mhi, kk ? 1u : 0u);   // z_hi x_hi (+ zita column)
          mma_f16_ss(tmem + kColM, ad, bl, idesc_mlo, 1u);             // z_hi x_lo
          mma_f16_ss(tmem + kColM, al, bh, idesc_mhi, 1u);             // z_lo x_hi
        }
        mma_commit(&ms.bar_mma);
      }
      __syncwarp();
    }
    SWEM_CTA_WAIT(&ms.bar_mma, ph_mma, ms.abort_flag);
    ph_mma ^= 1;
    tc_fence_after_sync();
    EM_STAMP();                      // M GEMM done

    // partial of this tile for row l = tid -> fp32 reductions straight from TMEM into the L2-resident accumulator
    // [c][l] (coalesced over l; fire-and-forget, the fence comes later so that they fly during what follows)
    float* acc = p.acc_k + ((size_t)((u * I + it) * 2 + sd) * LB + lb) * ((kCk + 1) * kL);
    if (row_thread) {
      const uint32_t base = tmem_addr(tmem, (warp & 3) * 32, kColM);
#pragma unroll
      for (int q = 0; q < kCk / 32; ++q) {
        uint32_t r[32];
        tmem_ld32(base + q * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) atomicAdd(acc + (q * 32 + j) * kL + tid, __uint_as_float(r[j]));
      }
      uint32_t r16[16];
      tmem_ld16(base + kCk, r16);
      tmem_ld_wait();
      atomicAdd(acc + kCk * kL + tid, __uint_as_float(r16[0]));
    }
    tc_fence_before_sync();

    const bool last = (it == I - 1);
    if (last) {
      // ---- nu partial = Z^T V^T for this side: one pass, V images streamed through the ring ---------------
      __syncthreads();
      tc_fence_after_sync();
      if (warp == 0) {
        if (lane == 0) {
          for (int k = LY::kOwnStages; k < kStages; ++k) load_image(k);   // (CK = 128) stages in the X region, free now
#pragma unroll 1
          for (int seq = 0; seq < kChunks; ++seq) {
            const int st = seq % kStages;
            if (!mbar_wait(&ms.bar_full[st], (seq / kStages) & 1)) ms.abort_flag = 1;
            tc_fence_after_sync();
            const int h = seq >> 2, q = seq & 3;
            const uint32_t vb = sbase + LY::stage_off(st);
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
              const uint64_t ad = make_sdesc(sbase + kOffZ + (q * 2 + kk) * 2 * 128, /*lbo*/ 128, /*sbo*/ 2048);
              const uint64_t al = make_sdesc(sbase + kOffZL + (q * 2 + kk) * 2 * 128, /*lbo*/ 128, /*sbo*/ 2048);
              const uint64_t bh = make_sdesc(vb + kk * 2 * 4096, /*lbo*/ 4096, /*sbo*/ 128);
              const uint64_t bl = make_sdesc(vb + kVPlane + kk * 2 * 4096, /*lbo*/ 4096, /*sbo*/ 128);
              mma_f16_ss(tmem + kColNu + h * 256, ad, bh, idesc_nu, (q | kk) ? 1u : 0u);   // z_hi v_hi
              mma_f16_ss(tmem + kColNu + h * 256, ad, bl, idesc_nu, 1u);                    // z_hi v_lo
              mma_f16_ss(tmem + kColNu + h * 256, al, bh, idesc_nu, 1u);                    // z_lo v_hi
            }
            mma_commit(&ms.bar_empty[st]);              // -> the producer thread (warp 1) refills this stage
          }
          mma_commit(&ms.bar_mma);
        }
        __syncwarp();
      } else if (warp == 1) {
        if (lane == 0) {              // producer: refill a stage as soon as the MMAs that read it have retired
#pragma unroll 1
          for (int k = kStages; k < kChunks; ++k) {
            const int st = k % kStages;
            if (!mbar_wait(&ms.bar_empty[st], ((k - kStages) / kStages) & 1)) ms.abort_flag = 1;
            load_image(k);
          }
        }
        __syncwarp();
      }
      SWEM_CTA_WAIT(&ms.bar_mma, ph_mma, ms.abort_flag);
      ph_mma ^= 1;
      tc_fence_after_sync();
      EM_STAMP();                    // nu GEMMs done
      // drain: TMEM [128 l][512 d] -> smem [64 d][128 l] fp32 -> bulk reduce-add into acc_nu, 8 rounds, two staging
      // buffers in the (now idle) V ring: round q only waits for the reads of round q-2.
      {
        const int l = (warp & 3) * 32 + lane, cg = warp >> 2;
#pragma unroll 1
        for (int q = 0; q < 8; ++q) {
          float* ns = reinterpret_cast<float*>(smem + LY::stage_off(kStages - 1 - (q & 1)));
          if (q >= 2) {
            if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncthreads();
          }
          {
            uint32_t r[32];
            tmem_ld32(tmem_addr(tmem, (warp & 3) * 32, kColNu + q * 64 + cg * 32), r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) ns[(cg * 32 + j) * 128 + l] = __uint_as_float(r[j]);
          }
          fence_proxy_async_smem();
          __syncthreads();
          if (tid == 0) {
            float* dst = p.acc_nu + ((size_t)gsl * kCv + q * 64) * kL;
            asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst),
                         "r"(smem_u32(ns)), "r"(64 * 128 * 4)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
        // full completion (not just .read): the arrival on this iteration's counter below also publishes the nu sums
        if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      }
      tc_fence_before_sync();
      __syncthreads();
      tc_fence_after_sync();
      EM_STAMP();                    // nu drained
    }

    // ---- (4) cross-tile reduction of the M-step partial: fence the reductions issued above, arrive; after the last tile
    // arrived every CTA reads the total back the same way it was accumulated.
    unsigned* counter = p.counters + (((size_t)u * I + it) * 2 + sd) * LB + lb;
    if (windowed) {
      // the kernel boundary is the cross-tile barrier: the next launch finalises this iteration from acc_k / acc_nu
    } else if (row_thread) {          // warps 0-3; they synchronise among themselves on named barrier 1
      __threadfence();                // (the prior-row loads come after it: a fence waits for every earlier access)
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (tid == 0) {
        EM_STAMP();                   // partial reduce-added
        atomicAdd(counter, 1u);
      }
      float kap[kCk];                 // prior row first: the loads fly during the cross-tile wait (CK = 64; 128 would not fit)
      if constexpr (kCk == 64) {
#pragma unroll
        for (int c = 0; c < kCk; ++c) kap[c] = valid_row ? __ldg(kprior + (size_t)c * L) : 0.f;
      }
      if (tid == 0) {
        const bool arrived = wait_counter(counter, (unsigned)p.T);
        EM_STAMP();                   // all tiles arrived
        if (!arrived) ms.abort_flag = 1;
        __threadfence();
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      // ---- (5) finalize row l = tid from the prior (reference :125-126) -------------------------------------
      if (!ms.abort_flag) {
        constexpr float kInvZ = 1.f / kZScale;
        const float zita_cur = zita_p + __ldcg(acc + kCk * kL + tid) * kInvZ;
        const float rz = valid_row ? 1.f / zita_cur : 0.f;
#pragma unroll
        for (int c = 0; c < kCk; ++c) {
          const float prior = (kCk == 64) ? kap[c] : (valid_row ? __ldg(kprior + (size_t)c * L) : 0.f);
          kap[c] = (zita_p * prior + __ldcg(acc + c * kL + tid) * kInvZ) * rz;
        }
        if (last) {
          ms.hsum[0][tid] = rz;       // (dead E-step scratch) 1 / zita and the prior zita of row l, for the nu slice below
          ms.hsum[1][tid] = zita_p;
          if (tile == 0 && valid_row) {
            p.zita[(size_t)gs * L + lrow] = zita_cur;
            float* kout = p.kappa + ((size_t)gs * kCk) * L + lrow;
#pragma unroll
            for (int c = 0; c < kCk; ++c) kout[(size_t)c * L] = kap[c];
          }
        } else {
          stage_khat(kap);
        }
      }
    } else if (chunks_done < kMine && !last) {   // (never in windowed mode: chunks_done is 0 or kMine there, see set-up)
      // warps 4-7: one more V chunk, hidden behind the row threads' reduction, cross-tile wait and finalize
      convert_mine(chunks_done, false);
      __threadfence();
      asm volatile("fence.proxy.async;" ::: "memory");
    }
    if (chunks_done < kMine && !last) ++chunks_done;
    __syncthreads();
    if (ms.abort_flag) {
      if (tid == 0) atomicExch(p.status, 1 + it);
      failed = true;
      break;
    }
    EM_STAMP();                      // finalize done
    if (last && !windowed) nu_slice();
  }

  if (fin_only) nu_slice();           // windowed mode, last launch: the totals of every tile are complete
  tc_fence_before_sync();
  __syncthreads();
  EM_STAMP();
  if (p.prof != nullptr && blockIdx.x == 0 && tid == 0) p.prof[0] = n_stamp;
  if (warp == 0) tmem_dealloc(tmem, 512);
  if (failed || ms.abort_flag) __trap();
}

// ------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------
static long long* g_prof = nullptr;
void set_profile_buffer(void* dev) { g_prof = static_cast<long long*>(dev); }
long long* get_profile_buffer() { return g_prof; }

template <int CK, int LB, bool VPM>
static int max_clusters_resident() {
  static PerDevice cache;                                // occupancy and the shared-memory attribute are per device
  const int dev_id = current_device();
  std::lock_guard<std::mutex> lock(cache.mu);
  int& n = cache.value[dev_id];
  if (n < 0) {
    cudaFuncSetAttribute(em_pair_kernel<CK, LB, VPM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)em::smem_bytes<CK, LB>());
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * LB, 1, 1);
    cfg.blockDim = dim3(256, 1, 1);
    cfg.dynamicSmemBytes = em::smem_bytes<CK, LB>();
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2 * LB;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&clusters, em_pair_kernel<CK, LB, VPM>, &cfg) != cudaSuccess || clusters <= 0) {
      cudaGetLastError();
      int dev = 0, sms = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      clusters = sms / (2 * LB) - 4;                   // conservative guess
    }
    n = clusters;
  }
  return n;
}

bool fused_em_supported(const SwemDims& d) {
  if ((d.Ck != 64 && d.Ck != 128) || (d.L != 64 && d.L != 128 && d.L != 256 && d.L != 512) || d.Cv != em::kCv || d.n_iters < 1 ||
      d.n_iters > 16)
    return false;
  const int T = (d.HW + em::kTP - 1) / em::kTP;
  // Up to the co-residency limit of a unit's clusters (74 pairs / ~34 quads / ~16 octets on a B200, asked from the
  // occupancy API at launch) the whole EM is one launch; beyond it the windowed form (one launch per iteration) runs.
  return T >= 1 && T <= 4096;
}

size_t fused_em_workspace(const SwemDims& d) {
  const size_t U = (size_t)d.B * d.N;
  const size_t T = (d.HW + em::kTP - 1) / em::kTP;
  size_t bytes = 0;
  const size_t LB = d.L > 128 ? d.L / 128 : 1;
  bytes += align_up(U * d.n_iters * 2 * LB * (size_t)(d.Ck + 1) * em::kL * 4, 256);
  bytes += align_up(U * 2 * LB * em::kCv * em::kL * 4, 256);
  bytes += align_up(U * d.n_iters * 2 * LB * 4 + 4, 256);
  bytes += align_up(U * T * em::kChunks * em::kStageBytes, 256);
  return bytes + 256;
}

template <int CK, int LB, bool VPM>
static int fused_em_forward_t(const SwemEmArgs& a, cudaStream_t st) {
  const SwemDims& d = a.dims;
  const int U = d.B * d.N;
  const int T = (d.HW + em::kTP - 1) / em::kTP;
  Arena ws(a.workspace);
  float* acc_k = ws.take<float>((size_t)U * d.n_iters * 2 * LB * (CK + 1) * em::kL);
  float* acc_nu = ws.take<float>((size_t)U * 2 * LB * em::kCv * em::kL);
  unsigned* counters = ws.take<unsigned>((size_t)U * d.n_iters * 2 * LB + 1);
  int* status = reinterpret_cast<int*>(counters + (size_t)U * d.n_iters * 2 * LB);
  SWEM_CUDA(cudaMemsetAsync(a.workspace, 0, ws.off, st));
  count_launch();
  uint8_t* vblob = ws.take<uint8_t>((size_t)U * T * em::kChunks * em::kStageBytes);

  (void)max_clusters_resident<CK, LB, VPM>();            // sets the shared-memory attribute on this device (once per device)
  EmPairParams p{};
  p.x = a.x; p.v = a.v; p.masks = a.masks;
  p.kappa_prior = a.kappa_prior; p.nu_prior = a.nu_prior; p.zita_prior = a.zita_prior;
  p.kappa = a.kappa; p.nu = a.nu; p.zita = a.zita; p.z_last = a.z_last;
  p.vblob = vblob;
  p.acc_k = acc_k; p.acc_nu = acc_nu; p.counters = counters; p.status = status;
  p.N = d.N; p.HW = d.HW; p.T = T; p.n_iters = d.n_iters; p.L = d.L;
  p.c1s = kLog2e / (d.tau * em::kKScale);
  p.prof = get_profile_buffer();
  // Windowed form (one launch per EM iteration + one for the outputs; the kernel boundary is the cross-tile barrier):
  // taken when the clusters of a unit cannot all be co-resident (large HW x L), or forced by SWEM_EM_WINDOWED=1 (tests).
  const char* force_w = getenv("SWEM_EM_WINDOWED");
  if (max_clusters_resident<CK, LB, VPM>() < T || (force_w != nullptr && force_w[0] == '1')) {
    p.windowed = 1;
    p.u0 = 0;
    for (int it = 0; it <= d.n_iters; ++it) {
      p.it_begin = it;
      em_pair_kernel<CK, LB, VPM><<<U * T * 2 * LB, 256, em::smem_bytes<CK, LB>(), st>>>(p);
      SWEM_LAUNCH_CHECK();
    }
    return SWEM_OK;
  }
  // Single-launch form: all CTAs of a launch spin on each other, so every launch must be co-resident (1 CTA per SM);
  // units that do not fit are spread evenly over the fewest launches
  const int upl_max = max_clusters_resident<CK, LB, VPM>() / T;
  const int n_launch = (U + upl_max - 1) / upl_max;
  const int upl = (U + n_launch - 1) / n_launch;
  for (int u0 = 0; u0 < U; u0 += upl) {
    const int nu = (U - u0 < upl) ? (U - u0) : upl;
    p.u0 = u0;
    em_pair_kernel<CK, LB, VPM><<<nu * T * 2 * LB, 256, em::smem_bytes<CK, LB>(), st>>>(p);
    SWEM_LAUNCH_CHECK();
  }
  return SWEM_OK;
}

template <bool VPM>
static int fused_em_forward_v(const SwemEmArgs& a, cudaStream_t st) {
  if (a.dims.L == 512) return a.dims.Ck == 128 ? fused_em_forward_t<128, 4, VPM>(a, st) : fused_em_forward_t<64, 4, VPM>(a, st);
  if (a.dims.L == 256) return a.dims.Ck == 128 ? fused_em_forward_t<128, 2, VPM>(a, st) : fused_em_forward_t<64, 2, VPM>(a, st);
  return a.dims.Ck == 128 ? fused_em_forward_t<128, 1, VPM>(a, st) : fused_em_forward_t<64, 1, VPM>(a, st);
}

int fused_em_forward(const SwemEmArgs& a, cudaStream_t st) {
  if (fused_em_res_covers(a.dims, a.v_pixel_major != 0)) return fused_em_res_forward(a, st);
  return a.v_pixel_major ? fused_em_forward_v<true>(a, st) : fused_em_forward_v<false>(a, st);
}

}  // namespace swem
