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

}  // namespace swem
