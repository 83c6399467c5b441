// Sorted top-l helper shared by perm_inv_kernel (generic.cu) and the fused readout (fused_readout_topl.cu).
#pragma once

#include <stdint.h>

namespace swem {

// Bitonic sort, descending, of two independent sequences of 32 * NPL packed words held NPL per lane (rank r ends in lane r / NPL,
// register r % NPL).  word = (fp32 bits of a non-negative value, top 32 - B bits) << B | column index: one unsigned min / max
// moves key and index together.
template <int NPL>
__device__ __forceinline__ void bitonic_desc2(uint32_t (&a)[NPL], uint32_t (&b)[NPL], int lane) {
  constexpr int N = 32 * NPL;
#pragma unroll
  for (int sz = 2; sz <= N; sz <<= 1) {
#pragma unroll
    for (int d = sz >> 1; d > 0; d >>= 1) {
      if (d < NPL) {            // partner lives in this lane
#pragma unroll
        for (int k = 0; k < NPL; ++k) {
          if ((k & d) == 0) {
            const bool desc = (sz < NPL) ? ((k & sz) == 0) : (((lane * NPL) & sz) == 0);
            const uint32_t alo = min(a[k], a[k | d]), ahi = max(a[k], a[k | d]);
            a[k] = desc ? ahi : alo;
            a[k | d] = desc ? alo : ahi;
            const uint32_t blo = min(b[k], b[k | d]), bhi = max(b[k], b[k | d]);
            b[k] = desc ? bhi : blo;
            b[k | d] = desc ? blo : bhi;
          }
        }
      } else {                  // partner lives in lane ^ (d / NPL), same register
        const int ld = d / NPL;
        const bool desc = (((lane * NPL) & sz) == 0);
        const bool take_max = (desc == ((lane & ld) == 0));
#pragma unroll
        for (int k = 0; k < NPL; ++k) {
          const uint32_t ya = __shfl_xor_sync(0xffffffffu, a[k], ld);
          const uint32_t yb = __shfl_xor_sync(0xffffffffu, b[k], ld);
          a[k] = take_max ? max(a[k], ya) : min(a[k], ya);
          b[k] = take_max ? max(b[k], yb) : min(b[k], yb);
        }
      }
    }
  }
}


// ---- the same sort as a direction-free network: every merge stage starts with a MIRRORED exchange (position i with position
// i ^ (sz - 1)) and continues with half-cleaners (i with i ^ d); in all of them the lower position keeps the larger word.  No
// per-block direction flags: an in-register compare-exchange is one max + one min, an exchange between lanes one shuffle + one
// predicated max / min pair.  ~1300 instructions for 2 x 256 words against ~3100 of bitonic_desc2 (SASS counts,
// profiles/r2_readout_phases.txt).  Position of a word = lane * NPL + register.
__device__ __forceinline__ void keep_max_or_min(uint32_t& a, uint32_t y, uint32_t take_max) {
  asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p max.u32 %0, %0, %1;\n\t@!p min.u32 %0, %0, %1;\n\t}" : "+r"(a) : "r"(y), "r"(take_max));
}
template <int NPL>
__device__ __forceinline__ void sort_desc2(uint32_t (&a)[NPL], uint32_t (&b)[NPL], int lane) {
  constexpr int N = 32 * NPL;
#pragma unroll
  for (int sz = 2; sz <= N; sz <<= 1) {
    // ---- mirrored first step: i <-> i ^ (sz - 1)
    if (sz <= NPL) {
#pragma unroll
      for (int k = 0; k < NPL; ++k) {
        const int k2 = k ^ (sz - 1);
        if (k < k2) {
          const uint32_t ahi = max(a[k], a[k2]), alo = min(a[k], a[k2]);
          a[k] = ahi; a[k2] = alo;
          const uint32_t bhi = max(b[k], b[k2]), blo = min(b[k], b[k2]);
          b[k] = bhi; b[k2] = blo;
        }
      }
    } else {
      const int lmask = sz / NPL - 1;                    // partner lane; its register NPL - 1 - k pairs with register k
      const uint32_t take_max = ((lane & (sz / NPL / 2)) == 0) ? 1u : 0u;
      {
        uint32_t y[NPL];
#pragma unroll
        for (int k = 0; k < NPL; ++k) y[k] = __shfl_xor_sync(0xffffffffu, a[NPL - 1 - k], lmask);
#pragma unroll
        for (int k = 0; k < NPL; ++k) keep_max_or_min(a[k], y[k], take_max);
#pragma unroll
        for (int k = 0; k < NPL; ++k) y[k] = __shfl_xor_sync(0xffffffffu, b[NPL - 1 - k], lmask);
#pragma unroll
        for (int k = 0; k < NPL; ++k) keep_max_or_min(b[k], y[k], take_max);
      }
    }
    // ---- half-cleaners: i <-> i ^ d
#pragma unroll
    for (int d = sz >> 2; d > 0; d >>= 1) {
      if (d < NPL) {
#pragma unroll
        for (int k = 0; k < NPL; ++k) {
          if ((k & d) == 0) {
            const uint32_t ahi = max(a[k], a[k | d]), alo = min(a[k], a[k | d]);
            a[k] = ahi; a[k | d] = alo;
            const uint32_t bhi = max(b[k], b[k | d]), blo = min(b[k], b[k | d]);
            b[k] = bhi; b[k | d] = blo;
          }
        }
      } else {
        const int ld = d / NPL;
        const uint32_t take_max = ((lane & ld) == 0) ? 1u : 0u;
#pragma unroll
        for (int k = 0; k < NPL; ++k) {
          const uint32_t ya = __shfl_xor_sync(0xffffffffu, a[k], ld);
          const uint32_t yb = __shfl_xor_sync(0xffffffffu, b[k], ld);
          keep_max_or_min(a[k], ya, take_max);
          keep_max_or_min(b[k], yb, take_max);
        }
      }
    }
  }
}

// One sequence of 16 * R words held R per lane by the 16 lanes of a half-warp (lane16 = lane & 15; position = lane16 * R +
// register): the two sides of a pixel sort side by side in the two halves of a warp.  Against two sequences over the full warp
// (sort_desc2) the exchanges between lanes drop from 15 steps x 16 words to 10 x 16, the rest stays in registers.
template <int R>
__device__ __forceinline__ void sort_desc_half(uint32_t (&a)[R], int lane16) {
  constexpr int N = 16 * R;
#pragma unroll
  for (int sz = 2; sz <= N; sz <<= 1) {
    if (sz <= R) {
#pragma unroll
      for (int k = 0; k < R; ++k) {
        const int k2 = k ^ (sz - 1);
        if (k < k2) {
          const uint32_t hi = max(a[k], a[k2]), lo = min(a[k], a[k2]);
          a[k] = hi; a[k2] = lo;
        }
      }
    } else {
      const int lmask = sz / R - 1;
      const bool take_max = (lane16 & (sz / R / 2)) == 0;
      uint32_t y[R];
#pragma unroll
      for (int k = 0; k < R; ++k) y[k] = __shfl_xor_sync(0xffffffffu, a[R - 1 - k], lmask);
#pragma unroll
      for (int k = 0; k < R; ++k) {
        const uint32_t hi = max(a[k], y[k]), lo = min(a[k], y[k]);
        a[k] = take_max ? hi : lo;
      }
    }
#pragma unroll
    for (int d = sz >> 2; d > 0; d >>= 1) {
      if (d < R) {
#pragma unroll
        for (int k = 0; k < R; ++k) {
          if ((k & d) == 0) {
            const uint32_t hi = max(a[k], a[k | d]), lo = min(a[k], a[k | d]);
            a[k] = hi; a[k | d] = lo;
          }
        }
      } else {
        const int ld = d / R;
        const bool take_max = (lane16 & ld) == 0;
#pragma unroll
        for (int k = 0; k < R; ++k) {
          const uint32_t y = __shfl_xor_sync(0xffffffffu, a[k], ld);
          const uint32_t hi = max(a[k], y), lo = min(a[k], y);
          a[k] = take_max ? hi : lo;
        }
      }
    }
  }
}

// The same network on the words REINTERPRETED AS FLOATS: a word whose upper bits come from a non-negative finite float orders
// the same way as an unsigned integer and as a float (exponent < 255: never a NaN; denormal words are compared exactly, fminf /
// fmaxf do not flush).  FMNMX issues at twice the rate of VIMNMX.U32 on sm_100 (measured: the sort phase of the fused readout
// is bound by the min / max pipe, profiles/r2_readout_phases.txt).
template <int R>
__device__ __forceinline__ void sort_desc_half(float (&a)[R], int lane16) {
  constexpr int N = 16 * R;
#pragma unroll
  for (int sz = 2; sz <= N; sz <<= 1) {
    if (sz <= R) {
#pragma unroll
      for (int k = 0; k < R; ++k) {
        const int k2 = k ^ (sz - 1);
        if (k < k2) {
          const float hi = fmaxf(a[k], a[k2]), lo = fminf(a[k], a[k2]);
          a[k] = hi; a[k2] = lo;
        }
      }
    } else {
      const int lmask = sz / R - 1;
      const bool take_max = (lane16 & (sz / R / 2)) == 0;
      float y[R];
#pragma unroll
      for (int k = 0; k < R; ++k) y[k] = __shfl_xor_sync(0xffffffffu, a[R - 1 - k], lmask);
#pragma unroll
      for (int k = 0; k < R; ++k) {
        const float hi = fmaxf(a[k], y[k]), lo = fminf(a[k], y[k]);
        a[k] = take_max ? hi : lo;
      }
    }
#pragma unroll
    for (int d = sz >> 2; d > 0; d >>= 1) {
      if (d < R) {
#pragma unroll
        for (int k = 0; k < R; ++k) {
          if ((k & d) == 0) {
            const float hi = fmaxf(a[k], a[k | d]), lo = fminf(a[k], a[k | d]);
            a[k] = hi; a[k | d] = lo;
          }
        }
      } else {
        const int ld = d / R;
        const bool take_max = (lane16 & ld) == 0;
#pragma unroll
        for (int k = 0; k < R; ++k) {
          const float y = __shfl_xor_sync(0xffffffffu, a[k], ld);
          const float hi = fmaxf(a[k], y), lo = fminf(a[k], y);
          a[k] = take_max ? hi : lo;
        }
      }
    }
  }
}

// ---- sorted top-64 of 256 words, 16 per lane over the 16 lanes of a half-warp: selection instead of a full sort --------------
// The reference's feature needs the 64 largest affinities of a side in order (modules.py:200-204), not the order of the other
// 192.  (1) every lane sorts its 16 words with the 60-comparator / 10-layer network for 16 inputs (20 fewer compare-exchanges
// than the bitonic network), (2) lanes merge by mirrored bitonic merges into four sorted runs of 64 (4 lanes each), (3) two
// prune-merges: of two sorted runs only max(x_i, y_63-i) -- a bitonic sequence holding the 64 largest of the 128 -- is kept,
// and the survivors are spread over ALL lanes again (half as many registers per lane) before they are sorted, so that no
// instruction is issued for discarded words.  The sort phase of the fused readout is bound by the min / max issue rate (2
// cycles per warp instruction and scheduler): 480 min / max / select + 100 shuffles per pixel against 736 + 160 for the full
// sort (sort_desc_half<16>); emulated lane by lane in tools/top64_emul.py.
// Result: rank r (0 = largest) of the top 64 sits in register r % 4 of the lane for which top64_rank_base(lane16) = r - r % 4.
#define SWEM_CE(i, j)                                        \
  {                                                          \
    const float hi_ = fmaxf(a[i], a[j]), lo_ = fminf(a[i], a[j]); \
    a[i] = hi_;                                              \
    a[j] = lo_;                                              \
  }
__device__ __forceinline__ void sort16_desc(float (&a)[16]) {
  SWEM_CE(0, 13) SWEM_CE(1, 12) SWEM_CE(2, 15) SWEM_CE(3, 14) SWEM_CE(4, 8) SWEM_CE(5, 6) SWEM_CE(7, 11) SWEM_CE(9, 10)
  SWEM_CE(0, 5) SWEM_CE(1, 7) SWEM_CE(2, 9) SWEM_CE(3, 4) SWEM_CE(6, 13) SWEM_CE(8, 14) SWEM_CE(10, 15) SWEM_CE(11, 12)
  SWEM_CE(0, 1) SWEM_CE(2, 3) SWEM_CE(4, 5) SWEM_CE(6, 8) SWEM_CE(7, 9) SWEM_CE(10, 11) SWEM_CE(12, 13) SWEM_CE(14, 15)
  SWEM_CE(0, 2) SWEM_CE(1, 3) SWEM_CE(4, 10) SWEM_CE(5, 11) SWEM_CE(6, 7) SWEM_CE(8, 9) SWEM_CE(12, 14) SWEM_CE(13, 15)
  SWEM_CE(1, 2) SWEM_CE(3, 12) SWEM_CE(4, 6) SWEM_CE(5, 7) SWEM_CE(8, 10) SWEM_CE(9, 11) SWEM_CE(13, 14)
  SWEM_CE(1, 4) SWEM_CE(2, 6) SWEM_CE(5, 8) SWEM_CE(7, 10) SWEM_CE(9, 13) SWEM_CE(11, 14)
  SWEM_CE(2, 4) SWEM_CE(3, 6) SWEM_CE(9, 12) SWEM_CE(11, 13)
  SWEM_CE(3, 5) SWEM_CE(6, 8) SWEM_CE(7, 9) SWEM_CE(10, 12)
  SWEM_CE(3, 4) SWEM_CE(5, 6) SWEM_CE(7, 8) SWEM_CE(9, 10) SWEM_CE(11, 12)
  SWEM_CE(6, 7) SWEM_CE(8, 9)
}
// exchange with lane ^ lmask, register by register: the lane keeps the larger (take_max) or the smaller word
template <int R>
__device__ __forceinline__ void xlane_keep(float (&a)[R], int lmask, bool take_max) {
  float y[R];
#pragma unroll
  for (int k = 0; k < R; ++k) y[k] = __shfl_xor_sync(0xffffffffu, a[k], lmask);
#pragma unroll
  for (int k = 0; k < R; ++k) {
    const float hi = fmaxf(a[k], y[k]), lo = fminf(a[k], y[k]);
    a[k] = take_max ? hi : lo;
  }
}
// half-cleaners over register distances d = R / 2 ... 1 (the lower register keeps the larger word)
template <int R>
__device__ __forceinline__ void inreg_clean(float (&a)[R]) {
#pragma unroll
  for (int d = R / 2; d > 0; d >>= 1) {
#pragma unroll
    for (int k = 0; k < R; ++k) {
      if ((k & d) == 0) {
        const float hi = fmaxf(a[k], a[k | d]), lo = fminf(a[k], a[k | d]);
        a[k] = hi;
        a[k | d] = lo;
      }
    }
  }
}
__device__ __forceinline__ int top64_rank_base(int lane16) {
  const int hi2 = (lane16 >> 3) & 1;
  const int lb = hi2 ? (lane16 ^ 12) : lane16;
  const int hi = (lb >> 2) & 1;
  const int b45 = (lb & 3) ^ (hi ? 3 : 0);
  return b45 * 16 + hi * 8 + hi2 * 4;
}
__device__ __forceinline__ void top64_of_256_half(float (&a)[16], int lane16, float (&out)[4]) {
  // (1) + (2): four sorted runs of 64, run r in lanes 4 r .. 4 r + 3, position (lane16 & 3) * 16 + register
  sort16_desc(a);
#pragma unroll
  for (int sz = 32; sz <= 64; sz <<= 1) {
    const int lmask = sz / 16 - 1;
    const bool take_max = (lane16 & (sz / 32)) == 0;
    float y[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) y[k] = __shfl_xor_sync(0xffffffffu, a[15 - k], lmask);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float hi = fmaxf(a[k], y[k]), lo = fminf(a[k], y[k]);
      a[k] = take_max ? hi : lo;
    }
    if (sz == 64) xlane_keep<16>(a, 1, (lane16 & 1) == 0);
    inreg_clean<16>(a);
  }
  // (3a) runs (0, 1) and (2, 3): survivor i = max(x_i, y_63-i) of a pair lives in BOTH lanes l and l ^ 7 (registers k and 15 - k);
  // lanes 0-3 of a group of 8 keep i % 16 < 8, lanes 4-7 take i % 16 >= 8 of their partner: 8 registers per lane.
  // Position of (lane, register m): [(lane & 3) ^ (hi1 ? 3 : 0)] * 16 + hi1 * 8 + m
  const bool hi1 = (lane16 & 4) != 0;
  float b[8];
  {
    float t[8], c[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) t[m] = __shfl_xor_sync(0xffffffffu, a[8 + m], 7);
#pragma unroll
    for (int x = 0; x < 8; ++x) c[x] = fmaxf(a[x], t[7 - x]);
#pragma unroll
    for (int m = 0; m < 8; ++m) b[m] = hi1 ? c[7 - m] : c[m];
  }
  {
    const int b45 = (lane16 & 3) ^ (hi1 ? 3 : 0);
    xlane_keep<8>(b, 2, (b45 & 2) == 0);
    xlane_keep<8>(b, 1, (b45 & 1) == 0);
    xlane_keep<8>(b, 7, !hi1);
    inreg_clean<8>(b);
  }
  // (3b) the two sorted runs of 64 (lanes 0-7, lanes 8-15; partner lane ^ 12, registers m and 7 - m): 4 registers per lane.
  // Position: top64_rank_base(lane16) + register
  const bool hi2 = (lane16 & 8) != 0;
  {
    float t[4], c[4];
#pragma unroll
    for (int x = 0; x < 4; ++x) t[x] = __shfl_xor_sync(0xffffffffu, b[4 + x], 12);
#pragma unroll
    for (int x = 0; x < 4; ++x) c[x] = fmaxf(b[x], t[3 - x]);
#pragma unroll
    for (int n = 0; n < 4; ++n) out[n] = hi2 ? c[3 - n] : c[n];
  }
  {
    const int lb = hi2 ? (lane16 ^ 12) : lane16;
    const bool hi = (lb & 4) != 0;
    const int b45 = (lb & 3) ^ (hi ? 3 : 0);
    xlane_keep<4>(out, 2, (b45 & 2) == 0);
    xlane_keep<4>(out, 1, (b45 & 1) == 0);
    xlane_keep<4>(out, 7, !hi);
    xlane_keep<4>(out, 12, !hi2);
    inreg_clean<4>(out);
  }
}
#undef SWEM_CE

}  // namespace swem
