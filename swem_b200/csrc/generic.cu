// Generic-shape kernel family of libswem_b200.so: any Ck / Cv / L / HW, fp32 throughout.
//
// This family covers every shape the reference accepts (the fused tcgen05 family in fused_*.cu
// covers the BASELINE shapes and is what SWEM_PATH_AUTO picks for them).  It is organised as a
// strided batched tile GEMM plus row kernels, with the reference's operation order:
//
//   memorize (reference modules.py:129-168), per EM iteration i:
//     khat   = kappa / (||kappa||_c + eps)                      kappa_unit_kernel      (:115)
//     a      = x^T khat                                         sgemm                  (:116)
//     w      = masks * (1 - side share of exp((a/||x|| - max)/tau))   em_assign_kernel (:93-110, i>0)
//     z      = softmax_l((a - max_l a)/tau) * w                 em_assign_kernel       (:117-119)
//     zita   = zita_ + sum_p z                                  colsum + zita kernels  (:125)
//     kappa  = (zita_ kappa_ + x z) / zita                      sgemm + bases_finalize (:126)
//   nu = (zita_ nu_ + v z_last) / zita                          sgemm + bases_finalize (:164-165)
//
// Note the W-step logits l2norm(x)^T khat (:98-100) equal a / (||x||+eps) with the SAME khat the
// next E-step uses, so one GEMM per iteration serves both steps.
//
//   readout (modules.py:232-293): scores = q^T khat / (||q||+eps) -> P = softmax over (side, j)
//   -> mem_out = V P (sgemm, accumulated over banks and sides) ; S from sorted top-l of P.
#include <float.h>

#include "common.cuh"
#include "topl.cuh"

namespace swem {

// ------------------------------------------------------------------------------------------
// strided, 3-level batched SGEMM: C[z][m][n] (+)= sum_k A[z][m][k] * B[z][k][n]
// ------------------------------------------------------------------------------------------
struct GemmShape {
  int M, N, K;
  long long sAm, sAk, sBk, sBn, sCm, sCn;   // element strides
  int n0, n1, n2;                           // batch grid; blockIdx.z -> (i0, i1, i2)
  long long bA[3], bB[3], bC[3];            // batch strides (0 = broadcast)
  int accumulate;                           // 1: C += ...
};

template <int BM, int BN, int BK, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
sgemm_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C, GemmShape g) {
  constexpr int NT = (BM / TM) * (BN / TN);
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int z = blockIdx.z;
  const int i2 = z % g.n2, i1 = (z / g.n2) % g.n1, i0 = z / (g.n2 * g.n1);
  A += i0 * g.bA[0] + i1 * g.bA[1] + i2 * g.bA[2];
  B += i0 * g.bB[0] + i1 * g.bB[1] + i2 * g.bB[2];
  C += i0 * g.bC[0] + i1 * g.bC[1] + i2 * g.bC[2];
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  const bool a_mfast = (g.sAm == 1), b_nfast = (g.sBn == 1);
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < g.K; k0 += BK) {
    for (int e = tid; e < BM * BK; e += NT) {
      int m, k;
      if (a_mfast) { m = e % BM; k = e / BM; } else { k = e % BK; m = e / BK; }
      float val = 0.f;
      if (m0 + m < g.M && k0 + k < g.K) val = A[(long long)(m0 + m) * g.sAm + (long long)(k0 + k) * g.sAk];
      As[k][m] = val;
    }
    for (int e = tid; e < BN * BK; e += NT) {
      int n, k;
      if (b_nfast) { n = e % BN; k = e / BN; } else { k = e % BK; n = e / BK; }
      float val = 0.f;
      if (n0 + n < g.N && k0 + k < g.K) val = B[(long long)(k0 + k) * g.sBk + (long long)(n0 + n) * g.sBn];
      Bs[k][n] = val;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[k][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[k][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * TM + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n >= g.N) continue;
      float* p = C + (long long)m * g.sCm + (long long)n * g.sCn;
      *p = g.accumulate ? (*p + acc[i][j]) : acc[i][j];
    }
  }
}

static int launch_gemm(const float* A, const float* B, float* C, const GemmShape& g, cudaStream_t st) {
  const int nb = g.n0 * g.n1 * g.n2;
  const long long big_tiles = (long long)((g.M + 63) / 64) * ((g.N + 63) / 64) * nb;
  if (big_tiles >= 120) {
    dim3 grid((g.N + 63) / 64, (g.M + 63) / 64, nb);
    sgemm_kernel<64, 64, 16, 4, 4><<<grid, 256, 0, st>>>(A, B, C, g);
  } else {
    dim3 grid((g.N + 31) / 32, (g.M + 31) / 32, nb);
    sgemm_kernel<32, 32, 16, 2, 2><<<grid, 256, 0, st>>>(A, B, C, g);
  }
  SWEM_LAUNCH_CHECK();
  return SWEM_OK;
}

// ------------------------------------------------------------------------------------------
// row / column kernels
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// inv[b][p] = 1 / (||x[b,:,p]||_2 + eps)
__global__ void pixel_inv_norm_kernel(const float* __restrict__ x, float* __restrict__ inv, int B, int C, int HW) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * HW) return;
  const int b = i / HW, p = i % HW;
  const float* xp = x + (long long)b * C * HW + p;
  float ss = 0.f;
  for (int c = 0; c < C; ++c) { const float t = xp[(long long)c * HW]; ss = fmaf(t, t, ss); }
  inv[i] = 1.f / (sqrtf(ss) + kEpsNorm);
}

// khat[g][c][l] = kappa[g][c][l] / (||kappa[g,:,l]||_2 + eps),  g over (b,n,s)
__global__ void kappa_unit_kernel(const float* __restrict__ kappa, float* __restrict__ khat, int G, int C, int L) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= G * L) return;
  const int g = i / L, l = i % L;
  const float* kp = kappa + (long long)g * C * L + l;
  float ss = 0.f;
  for (int c = 0; c < C; ++c) { const float t = kp[(long long)c * L]; ss = fmaf(t, t, ss); }
  const float nrm = sqrtf(ss) + kEpsNorm;
  float* op = khat + (long long)g * C * L + l;
  for (int c = 0; c < C; ++c) op[(long long)c * L] = kp[(long long)c * L] / nrm;
}

// One warp per (unit u, pixel p).  a: [U][2][HW][L] logits x^T khat, overwritten with z.
__global__ void em_assign_kernel(float* __restrict__ a, const float* __restrict__ inv_nx,
                                 const float* __restrict__ masks, int U, int N, int HW, int L,
                                 float inv_tau, int do_w) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= U * HW) return;
  const int u = warp / HW, p = warp % HW, b = u / N;
  float* row0 = a + ((long long)(u * 2 + 0) * HW + p) * L;
  float* row1 = a + ((long long)(u * 2 + 1) * HW + p) * L;
  float w0 = masks[(long long)(u * 2 + 0) * HW + p];
  float w1 = masks[(long long)(u * 2 + 1) * HW + p];
  if (do_w) {
    // W-step (reference :93-110) on t = a / (||x_p|| + eps)
    const float inv = inv_nx[b * HW + p];
    float mx = -FLT_MAX;
    for (int l = lane; l < L; l += 32) mx = fmaxf(mx, fmaxf(row0[l], row1[l]) * inv);
    // (a*inv) is monotone in a for inv > 0, but take the max of the products like the reference
    mx = warp_max(mx);
    float e0 = 0.f, e1 = 0.f;
    for (int l = lane; l < L; l += 32) {
      e0 += expf((row0[l] * inv - mx) * inv_tau);
      e1 += expf((row1[l] * inv - mx) * inv_tau);
    }
    e0 = warp_sum(e0);
    e1 = warp_sum(e1);
    const float tot = e0 + e1;
    w0 *= (1.f - e0 / tot);
    w1 *= (1.f - e1 / tot);
  }
  // E-step (reference :112-120), each side separately
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    float* row = s ? row1 : row0;
    const float w = s ? w1 : w0;
    float mx = -FLT_MAX;
    for (int l = lane; l < L; l += 32) mx = fmaxf(mx, row[l]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int l = lane; l < L; l += 32) sum += expf((row[l] - mx) * inv_tau);
    sum = warp_sum(sum);
    for (int l = lane; l < L; l += 32) row[l] = expf((row[l] - mx) * inv_tau) / sum * w;
  }
}

// part[g][chunk][l] = sum over the chunk's pixels of z[g][p][l]
__global__ void colsum_partial_kernel(const float* __restrict__ z, float* __restrict__ part, int HW, int L,
                                      int chunk_px, int n_chunks) {
  const int g = blockIdx.y, ch = blockIdx.x;
  const int p0 = ch * chunk_px, p1 = min(HW, p0 + chunk_px);
  for (int l = threadIdx.x; l < L; l += blockDim.x) {
    const float* zp = z + ((long long)g * HW + p0) * L + l;
    float s = 0.f;
    for (int p = p0; p < p1; ++p, zp += L) s += *zp;
    part[((long long)g * n_chunks + ch) * L + l] = s;
  }
}

// zita[g][l] = zita_prior[g][l] + sum_chunks part[g][chunk][l]   (fixed order -> deterministic)
__global__ void zita_kernel(const float* __restrict__ zita_prior, const float* __restrict__ part,
                            float* __restrict__ zita, int G, int L, int n_chunks) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= G * L) return;
  const int g = i / L, l = i % L;
  float s = 0.f;
  for (int c = 0; c < n_chunks; ++c) s += part[((long long)g * n_chunks + c) * L + l];
  zita[i] = zita_prior[i] + s;
}

// out[g][r][l] = (zita_prior[g][l] * prior[g][r][l] + acc[g][r][l]) / zita[g][l]; acc may alias out
__global__ void bases_finalize_kernel(const float* __restrict__ prior, const float* acc,
                                      const float* __restrict__ zita_prior, const float* __restrict__ zita,
                                      float* out, int G, int R, int L) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)G * R * L) return;
  const int l = (int)(i % L);
  const int g = (int)(i / ((long long)R * L));
  out[i] = (zita_prior[g * L + l] * prior[i] + acc[i]) / zita[g * L + l];
}

// One warp per (u, p): scores row [2Lt] -> P = softmax over both sides of (score/||q||)/tau, in place.
__global__ void readout_rows_kernel(float* __restrict__ sc, const float* __restrict__ inv_nq, int U, int N,
                                    int HW, int W2, float inv_tau, float* __restrict__ zsum = nullptr) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= U * HW) return;
  const int u = warp / HW, p = warp % HW, b = u / N;
  float* row = sc + ((long long)u * HW + p) * W2;
  const float inv = inv_nq[b * HW + p];
  float mx = -FLT_MAX;
  for (int j = lane; j < W2; j += 32) mx = fmaxf(mx, row[j] * inv);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < W2; j += 32) sum += expf((row[j] * inv - mx) * inv_tau);
  sum = warp_sum(sum);
  for (int j = lane; j < W2; j += 32) row[j] = expf((row[j] * inv - mx) * inv_tau) / sum;
  if (zsum != nullptr && lane == 0) zsum[(long long)u * HW + p] = sum;       // row sum of the exp-affinities (max = 1)
}

// ---- kernelised memory (reference gen_kernels, modules.py:210-230; inference only, off by default) -----------------------------
// One warp per (u, column j of the 2 Lt bases): the K pixels with the largest affinity a[p] = score[p][j] / ||q_p|| (:213; ties: the
// lower pixel index).  Every lane keeps the K best of its pixels p = lane, lane + 32, ... in a sorted register list, then K rounds of
// a warp arg-max pop the winners.
constexpr int kMkmMax = 16;
__global__ void __launch_bounds__(256) mkm_topk_kernel(const float* __restrict__ sc, const float* __restrict__ inv_nq, int U, int N, int HW,
                                                       int W2, int K, int* __restrict__ centers) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= U * W2) return;
  const int u = warp / W2, j = warp % W2, b = u / N;
  float val[kMkmMax];
  int idx[kMkmMax];
#pragma unroll
  for (int k = 0; k < kMkmMax; ++k) { val[k] = -FLT_MAX; idx[k] = 0x7fffffff; }
  for (int p = lane; p < HW; p += 32) {
    float v = sc[((long long)u * HW + p) * W2 + j] * inv_nq[b * HW + p];
    int pi = p;
#pragma unroll
    for (int k = 0; k < kMkmMax; ++k) {              // insertion into the descending list (pixels arrive in ascending order: strict >)
      if (k < K && v > val[k]) {
        const float tv = val[k]; const int ti = idx[k];
        val[k] = v; idx[k] = pi;
        v = tv; pi = ti;
      }
    }
  }
  for (int r = 0; r < K; ++r) {
    const float head = val[0];
    const int hidx = idx[0];
    float best = head;
    int bidx = hidx;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
      if (ov > best || (ov == best && oi < bidx)) { best = ov; bidx = oi; }
    }
    if (hidx == bidx) {                              // this lane held the winner: pop it
#pragma unroll
      for (int k = 0; k + 1 < kMkmMax; ++k) { val[k] = val[k + 1]; idx[k] = idx[k + 1]; }
      val[kMkmMax - 1] = -FLT_MAX; idx[kMkmMax - 1] = 0x7fffffff;
    }
    if (lane == 0) centers[((long long)u * W2 + j) * kMkmMax + r] = bidx;
  }
}

// One warp per (u, p): P[j] <- P[j] G[j] / (sum_j P[j] G[j] + 1e-8 / Z), G[j] = exp(-min_k d^2(p, center_jk) / (2 sigma^2 tau)), Z the row
// sum of the exp-affinities -- i.e. E G / (sum E G + 1e-8) of the reference (:254-256) written on the normalised P = E / Z.
// The same kernel applies the memory dropout of the training branch (:258-263): centers = NULL, colmask [U][Lt] of 0 / 1 (one mask
// for both sides), eps = 1e-6; `out` may be P itself (in place) or a second buffer (the backward keeps the plain P); `srow`
// (optional) receives the denominator sum_j P[j] w[j] + eps / Z of every row.
__global__ void __launch_bounds__(256) mkm_apply_kernel(const float* __restrict__ P, const float* __restrict__ zsum, const int* __restrict__ centers,
                                                        const float* __restrict__ colmask, int U, int HW, int W2, int K, int width,
                                                        float inv_2s2tau, float eps, float* __restrict__ out, float* __restrict__ srow) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= U * HW) return;
  const int u = warp / HW, p = warp % HW, Lt = W2 / 2;
  const float px = (float)(p % width), py = (float)(p / width);
  const float* row = P + ((long long)u * HW + p) * W2;
  float* orow = out + ((long long)u * HW + p) * W2;
  float sum = 0.f;
  for (int j = lane; j < W2; j += 32) {
    float w = row[j];
    if (centers != nullptr) {
      const int* c = centers + ((long long)u * W2 + j) * kMkmMax;
      float d2 = FLT_MAX;
      for (int k = 0; k < K; ++k) {
        const int q = c[k];
        const float dx = px - (float)(q % width), dy = py - (float)(q / width);
        d2 = fminf(d2, dx * dx + dy * dy);
      }
      w *= expf(-d2 * inv_2s2tau);
    }
    if (colmask != nullptr) w *= colmask[(long long)u * Lt + j % Lt];
    orow[j] = w;
    sum += w;
  }
  sum = warp_sum(sum);
  const float den = sum + eps / zsum[(long long)u * HW + p];
  const float inv = 1.f / den;
  for (int j = lane; j < W2; j += 32) orow[j] *= inv;
  if (srow != nullptr && lane == 0) srow[(long long)u * HW + p] = den;
}

// ------------------------------------------------------------------------------------------
// permutation-invariant feature (reference :198-208).  One warp per (u, p), both sides at once.
// The Lt values of a side are sorted in registers by a bitonic network on packed words
//   word = (fp32 bits of E, top 32-B bits) << B | column index      (E >= 0: bit order = value order)
// so one unsigned min/max moves key and index together; the top-l columns are then re-read in
// exact fp32 for the running sums.  Keys keep >= 14 mantissa bits, so two values can only swap
// ranks when they agree to ~1e-4 relative, which perturbs a running sum by less than that times
// the smaller of the two.  f is scale invariant, so any positive multiple of exp-affinity works.
// ------------------------------------------------------------------------------------------
template <int NPL>
__global__ void __launch_bounds__(256) perm_inv_kernel(const float* __restrict__ P, int U, int HW, int Lt,
                                                       int topl, float* __restrict__ out, int out_channels,
                                                       int s_channel, int pixel_major) {
  constexpr int N = 32 * NPL;
  constexpr int IDXB = (NPL == 1 ? 5 : NPL == 2 ? 6 : NPL == 4 ? 7 : NPL == 8 ? 8 : NPL == 16 ? 9 : 10);
  constexpr int PXB = 8;                                        // pixels per block: one per warp (many small blocks balance the SMs)
  extern __shared__ float tile[];                               // [2*topl][PXB+1] then uint32 top[8 warps][2][64]
  uint32_t* top = reinterpret_cast<uint32_t*>(tile + 2 * topl * (PXB + 1));
  const int u = blockIdx.y;
  const int p_base = blockIdx.x * PXB;
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t* mytop = top + wid * 128;
  for (int q = wid; q < PXB; q += 8) {
    const int p = p_base + q;
    if (p >= HW) break;
    const float* row = P + ((long long)u * HW + p) * (2 * Lt);
    uint32_t a[NPL], b[NPL];
#pragma unroll
    for (int k = 0; k < NPL; ++k) {
      const int i = lane * NPL + k;
      const uint32_t va = i < Lt ? __float_as_uint(row[i]) : 0u;
      const uint32_t vb = i < Lt ? __float_as_uint(row[Lt + i]) : 0u;
      a[k] = ((va >> (IDXB - 1)) << IDXB) | (uint32_t)i;
      b[k] = ((vb >> (IDXB - 1)) << IDXB) | (uint32_t)i;
    }
    bitonic_desc2<NPL>(a, b, lane);
    // rank r sits in lane r / NPL, register r % NPL; publish the top-l words
#pragma unroll
    for (int k = 0; k < NPL; ++k) {
      const int r = lane * NPL + k;
      if (r < 64) {
        mytop[r] = a[k];
        mytop[64 + r] = b[k];
      }
    }
    __syncwarp();
    // lane handles ranks lane and lane + 32: exact values, inclusive running sums over rank
    float c0[2], c1[2];
#pragma unroll
    for (int hlf = 0; hlf < 2; ++hlf) {
      const int r = lane + 32 * hlf;
      float x0 = 0.f, x1 = 0.f;
      if (r < topl) {
        const int i0 = mytop[r] & (N - 1), i1 = mytop[64 + r] & (N - 1);
        x0 = i0 < Lt ? row[i0] : 0.f;          // padding words can only surface if real values are exact zeros
        x1 = i1 < Lt ? row[Lt + i1] : 0.f;
      }
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float y0 = __shfl_up_sync(0xffffffffu, x0, o);
        const float y1 = __shfl_up_sync(0xffffffffu, x1, o);
        if (lane >= o) { x0 += y0; x1 += y1; }
      }
      c0[hlf] = x0;
      c1[hlf] = x1;
    }
    const float t0 = __shfl_sync(0xffffffffu, c0[0], 31), t1 = __shfl_sync(0xffffffffu, c1[0], 31);
    c0[1] += t0;
    c1[1] += t1;
#pragma unroll
    for (int hlf = 0; hlf < 2; ++hlf) {
      const int r = lane + 32 * hlf;
      if (r < topl) {
        const float f = c0[hlf] / (c0[hlf] + c1[hlf]);
        tile[r * (PXB + 1) + q] = f;
        tile[(topl + r) * (PXB + 1) + q] = 1.f - f;
      }
    }
    __syncwarp();
  }
  __syncthreads();
  const int npx = min(PXB, HW - p_base);
  if (pixel_major) {                                            // [U][HW][out_channels]: channels of a pixel are contiguous
    for (int e = threadIdx.x; e < 2 * topl * PXB; e += blockDim.x) {
      const int q = e / (2 * topl), ch = e % (2 * topl);
      if (q < npx) out[((long long)u * HW + p_base + q) * out_channels + s_channel + ch] = tile[ch * (PXB + 1) + q];
    }
  } else {
    for (int e = threadIdx.x; e < 2 * topl * PXB; e += blockDim.x) {
      const int ch = e / PXB, q = e % PXB;
      if (q < npx) out[((long long)u * out_channels + s_channel + ch) * HW + p_base + q] = tile[ch * (PXB + 1) + q];
    }
  }
}

int launch_perm_inv(const float* P, int U, int HW, int Lt, int topl, float* out, int out_channels,
                    int s_channel, int pixel_major, cudaStream_t st) {
  if (topl > 64) {
    set_error("perm_inv: topl=%d > 64 unsupported", topl);
    return SWEM_ERR_UNSUPPORTED;
  }
  dim3 grid((HW + 7) / 8, U);
  const size_t smem = (size_t)2 * topl * 9 * sizeof(float) + 8 * 128 * sizeof(uint32_t);
  const int npl = (Lt + 31) / 32;
#define SWEM_PI(NPL_) perm_inv_kernel<NPL_><<<grid, 256, smem, st>>>(P, U, HW, Lt, topl, out, out_channels, s_channel, pixel_major)
  if (npl <= 1) SWEM_PI(1);
  else if (npl <= 2) SWEM_PI(2);
  else if (npl <= 4) SWEM_PI(4);
  else if (npl <= 8) SWEM_PI(8);
  else if (npl <= 16) SWEM_PI(16);
  else if (npl <= 32) SWEM_PI(32);
  else {
    set_error("perm_inv: Lt=%d > 1024 unsupported", Lt);
    return SWEM_ERR_UNSUPPORTED;
  }
#undef SWEM_PI
  SWEM_LAUNCH_CHECK();
  return SWEM_OK;
}

// ------------------------------------------------------------------------------------------
// memorize
// ------------------------------------------------------------------------------------------
static constexpr int kChunkPx = 128;

size_t generic_em_workspace(const SwemDims& d) {
  const size_t U = (size_t)d.B * d.N, G = U * 2;
  const size_t n_chunks = (d.HW + kChunkPx - 1) / kChunkPx;
  size_t bytes = 0;
  bytes += align_up(G * d.HW * d.L * 4, 256);         // a / z
  bytes += align_up(G * d.Ck * d.L * 4, 256);         // khat
  bytes += align_up(G * d.Ck * d.L * 4, 256);         // kappa (iterate)
  bytes += align_up((size_t)d.B * d.HW * 4, 256);     // inv ||x||
  bytes += align_up(G * n_chunks * d.L * 4, 256);     // column-sum partials
  return bytes + 256;
}

int generic_em_forward(const SwemEmArgs& a, cudaStream_t st) {
  const SwemDims& d = a.dims;
  const int U = d.B * d.N, G = U * 2;
  const int n_chunks = (d.HW + kChunkPx - 1) / kChunkPx;
  Arena ws(a.workspace);
  float* z = a.z_last ? a.z_last : ws.take<float>((size_t)G * d.HW * d.L);
  if (a.z_last) (void)ws.take<float>((size_t)G * d.HW * d.L);
  float* khat = ws.take<float>((size_t)G * d.Ck * d.L);
  float* kcur = ws.take<float>((size_t)G * d.Ck * d.L);
  float* inv_nx = ws.take<float>((size_t)d.B * d.HW);
  float* part = ws.take<float>((size_t)G * n_chunks * d.L);
  const float inv_tau = 1.f / d.tau;

  pixel_inv_norm_kernel<<<(d.B * d.HW + 255) / 256, 256, 0, st>>>(a.x, inv_nx, d.B, d.Ck, d.HW);
  SWEM_LAUNCH_CHECK();

  for (int it = 0; it < d.n_iters; ++it) {
    const float* ksrc = (it == 0) ? a.kappa_prior : kcur;
    kappa_unit_kernel<<<(G * d.L + 127) / 128, 128, 0, st>>>(ksrc, khat, G, d.Ck, d.L);
    SWEM_LAUNCH_CHECK();
    {  // a[b,n,s][p][l] = sum_c x[b][c][p] khat[b,n,s][c][l]
      GemmShape g{};
      g.M = d.HW; g.N = d.L; g.K = d.Ck;
      g.sAm = 1; g.sAk = d.HW; g.sBk = d.L; g.sBn = 1; g.sCm = d.L; g.sCn = 1;
      g.n0 = d.B; g.n1 = d.N; g.n2 = 2;
      g.bA[0] = (long long)d.Ck * d.HW; g.bA[1] = 0; g.bA[2] = 0;
      g.bB[0] = (long long)d.N * 2 * d.Ck * d.L; g.bB[1] = 2LL * d.Ck * d.L; g.bB[2] = (long long)d.Ck * d.L;
      g.bC[0] = (long long)d.N * 2 * d.HW * d.L; g.bC[1] = 2LL * d.HW * d.L; g.bC[2] = (long long)d.HW * d.L;
      if (int rc = launch_gemm(a.x, khat, z, g, st)) return rc;
    }
    {
      const long long threads = (long long)U * d.HW * 32;
      em_assign_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(z, inv_nx, a.masks, U, d.N, d.HW, d.L,
                                                                          inv_tau, it > 0);
      SWEM_LAUNCH_CHECK();
    }
    colsum_partial_kernel<<<dim3(n_chunks, G), min(d.L, 256), 0, st>>>(z, part, d.HW, d.L, kChunkPx, n_chunks);
    SWEM_LAUNCH_CHECK();
    zita_kernel<<<(G * d.L + 127) / 128, 128, 0, st>>>(a.zita_prior, part, a.zita, G, d.L, n_chunks);
    SWEM_LAUNCH_CHECK();
    float* kdst = (it == d.n_iters - 1) ? a.kappa : kcur;
    {  // acc[b,n,s][c][l] = sum_p x[b][c][p] z[b,n,s][p][l]
      GemmShape g{};
      g.M = d.Ck; g.N = d.L; g.K = d.HW;
      g.sAm = d.HW; g.sAk = 1; g.sBk = d.L; g.sBn = 1; g.sCm = d.L; g.sCn = 1;
      g.n0 = d.B; g.n1 = d.N; g.n2 = 2;
      g.bA[0] = (long long)d.Ck * d.HW; g.bA[1] = 0; g.bA[2] = 0;
      g.bB[0] = (long long)d.N * 2 * d.HW * d.L; g.bB[1] = 2LL * d.HW * d.L; g.bB[2] = (long long)d.HW * d.L;
      g.bC[0] = (long long)d.N * 2 * d.Ck * d.L; g.bC[1] = 2LL * d.Ck * d.L; g.bC[2] = (long long)d.Ck * d.L;
      if (int rc = launch_gemm(a.x, z, kdst, g, st)) return rc;
    }
    {
      const long long n = (long long)G * d.Ck * d.L;
      bases_finalize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a.kappa_prior, kdst, a.zita_prior, a.zita,
                                                                         kdst, G, d.Ck, d.L);
      SWEM_LAUNCH_CHECK();
    }
  }
  {  // nu acc[b,n,s][dch][l] = sum_p v[b,n][dch][p] z[b,n,s][p][l]
    GemmShape g{};
    g.M = d.Cv; g.N = d.L; g.K = d.HW;
    g.sAm = d.HW; g.sAk = 1; g.sBk = d.L; g.sBn = 1; g.sCm = d.L; g.sCn = 1;
    g.n0 = d.B; g.n1 = d.N; g.n2 = 2;
    g.bA[0] = (long long)d.N * d.Cv * d.HW; g.bA[1] = (long long)d.Cv * d.HW; g.bA[2] = 0;
    g.bB[0] = (long long)d.N * 2 * d.HW * d.L; g.bB[1] = 2LL * d.HW * d.L; g.bB[2] = (long long)d.HW * d.L;
    g.bC[0] = (long long)d.N * 2 * d.Cv * d.L; g.bC[1] = 2LL * d.Cv * d.L; g.bC[2] = (long long)d.Cv * d.L;
    if (int rc = launch_gemm(a.v, z, a.nu, g, st)) return rc;
    const long long n = (long long)G * d.Cv * d.L;
    bases_finalize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a.nu_prior, a.nu, a.zita_prior, a.zita, a.nu,
                                                                       G, d.Cv, d.L);
    SWEM_LAUNCH_CHECK();
  }
  return SWEM_OK;
}

// ------------------------------------------------------------------------------------------
// memorize, backward (reference: autograd through modules.py:164-165, the only differentiable line of swem)
// ------------------------------------------------------------------------------------------
// w[g][d][l] = grad_nu[g][d][l] / zita[g][l] ;  grad_nu_prior[g][d][l] = w * zita_prior[g][l]
__global__ void nu_grad_scale_kernel(const float* __restrict__ gnu, const float* __restrict__ zita_prior,
                                     const float* __restrict__ zita, float* __restrict__ w, float* __restrict__ gprior,
                                     int G, int R, int L) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)G * R * L) return;
  const int l = (int)(i % L);
  const int g = (int)(i / ((long long)R * L));
  const float t = gnu[i] / zita[g * L + l];
  w[i] = t;
  if (gprior != nullptr) gprior[i] = t * zita_prior[g * L + l];
}

size_t generic_em_backward_workspace(const SwemDims& d) {
  return align_up((size_t)d.B * d.N * 2 * d.Cv * d.L * 4, 256) + 256;
}

int generic_em_backward(const SwemEmBwdArgs& a, cudaStream_t st) {
  const SwemDims& d = a.dims;
  const int G = d.B * d.N * 2;
  Arena ws(a.workspace);
  float* w = ws.take<float>((size_t)G * d.Cv * d.L);
  {
    const long long n = (long long)G * d.Cv * d.L;
    nu_grad_scale_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a.grad_nu, a.zita_prior, a.zita, w, a.grad_nu_prior,
                                                                      G, d.Cv, d.L);
    SWEM_LAUNCH_CHECK();
  }
  if (a.grad_v == nullptr) return SWEM_OK;
  for (int s = 0; s < 2; ++s) {   // grad_v[b,n][dch][p] (+)= sum_l w[b,n,s][dch][l] z[b,n,s][p][l]
    GemmShape g{};
    g.M = d.Cv; g.N = d.HW; g.K = d.L;
    g.sAm = d.L; g.sAk = 1; g.sBk = 1; g.sBn = d.L; g.sCm = d.HW; g.sCn = 1;
    g.n0 = d.B; g.n1 = d.N; g.n2 = 1;
    g.bA[0] = (long long)d.N * 2 * d.Cv * d.L; g.bA[1] = 2LL * d.Cv * d.L; g.bA[2] = 0;
    g.bB[0] = (long long)d.N * 2 * d.HW * d.L; g.bB[1] = 2LL * d.HW * d.L; g.bB[2] = 0;
    g.bC[0] = (long long)d.N * d.Cv * d.HW; g.bC[1] = (long long)d.Cv * d.HW; g.bC[2] = 0;
    g.accumulate = s;
    if (int rc = launch_gemm(w + (size_t)s * d.Cv * d.L, a.z_last + (size_t)s * d.HW * d.L, a.grad_v, g, st)) return rc;
  }
  return SWEM_OK;
}

// ------------------------------------------------------------------------------------------
// readout
// ------------------------------------------------------------------------------------------
size_t generic_readout_workspace(const SwemDims& d) {
  const size_t U = (size_t)d.B * d.N, G = U * 2;
  const size_t Lt = (size_t)d.L * d.n_banks;
  size_t bytes = 0;
  bytes += align_up(U * d.HW * 2 * Lt * 4, 256);              // scores / P
  bytes += align_up(G * d.Ck * d.L * 4, 256) * d.n_banks;     // khat per bank
  bytes += align_up((size_t)d.B * d.HW * 4, 256);             // inv ||q||
  bytes += align_up(U * 2 * Lt * kMkmMax * 4, 256);           // kernelised memory: best-matching pixels per basis
  bytes += align_up(U * d.HW * 4, 256);                       //                    row sums of the exp-affinities
  return bytes + 256;
}

int generic_readout_forward(const SwemReadArgs& a, cudaStream_t st) {
  const SwemDims& d = a.dims;
  const int U = d.B * d.N, G = U * 2;
  const int Lt = d.L * d.n_banks, W2 = 2 * Lt;
  Arena ws(a.workspace);
  float* P = ws.take<float>((size_t)U * d.HW * W2);
  float* khat[2] = {nullptr, nullptr};
  for (int k = 0; k < d.n_banks; ++k) khat[k] = ws.take<float>((size_t)G * d.Ck * d.L);
  float* inv_nq = ws.take<float>((size_t)d.B * d.HW);
  int* centers = ws.take<int>((size_t)U * W2 * kMkmMax);
  float* zsum = ws.take<float>((size_t)U * d.HW);
  const bool mkm = a.mkm_kernels > 0, drop = a.drop_mask != nullptr, weighted = mkm || drop;

  pixel_inv_norm_kernel<<<(d.B * d.HW + 255) / 256, 256, 0, st>>>(a.qk, inv_nq, d.B, d.Ck, d.HW);
  SWEM_LAUNCH_CHECK();
  for (int k = 0; k < d.n_banks; ++k) {
    kappa_unit_kernel<<<(G * d.L + 127) / 128, 128, 0, st>>>(a.kappa[k], khat[k], G, d.Ck, d.L);
    SWEM_LAUNCH_CHECK();
    // P[u][p][s*Lt + k*L + l] = sum_c q[b][c][p] khat_k[u,s][c][l]
    GemmShape g{};
    g.M = d.HW; g.N = d.L; g.K = d.Ck;
    g.sAm = 1; g.sAk = d.HW; g.sBk = d.L; g.sBn = 1; g.sCm = W2; g.sCn = 1;
    g.n0 = d.B; g.n1 = d.N; g.n2 = 2;
    g.bA[0] = (long long)d.Ck * d.HW; g.bA[1] = 0; g.bA[2] = 0;
    g.bB[0] = (long long)d.N * 2 * d.Ck * d.L; g.bB[1] = 2LL * d.Ck * d.L; g.bB[2] = (long long)d.Ck * d.L;
    g.bC[0] = (long long)d.N * d.HW * W2; g.bC[1] = (long long)d.HW * W2; g.bC[2] = Lt;
    if (int rc = launch_gemm(a.qk, khat[k], P + (size_t)k * d.L, g, st)) return rc;
  }
  if (mkm) {                                                  // best-matching pixels of every basis, from the raw affinities
    const long long threads = (long long)U * W2 * 32;
    mkm_topk_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(P, inv_nq, U, d.N, d.HW, W2, a.mkm_kernels, centers);
    SWEM_LAUNCH_CHECK();
  }
  {
    const long long threads = (long long)U * d.HW * 32;
    readout_rows_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(P, inv_nq, U, d.N, d.HW, W2, 1.f / d.tau, weighted ? zsum : nullptr);
    SWEM_LAUNCH_CHECK();
  }
  if (weighted) {                                             // S from the plain affinities (:269), then the weights on the attention
    if (int rc = launch_perm_inv(P, U, d.HW, Lt, d.topl, a.out, a.out_channels, a.s_channel, a.out_pixel_major, st)) return rc;
    const long long threads = (long long)U * d.HW * 32;
    mkm_apply_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(
        P, zsum, mkm ? centers : nullptr, a.drop_mask, U, d.HW, W2, a.mkm_kernels, mkm ? a.mkm_width : 1,
        mkm ? 1.f / (2.f * a.mkm_sigma * a.mkm_sigma * d.tau) : 0.f, mkm ? 1e-8f : 1e-6f, P, nullptr);
    SWEM_LAUNCH_CHECK();
  }
  // mem_out[u][dch][p] = sum_{s,k,l} nu_k[u,s][dch][l] P[u][p][s*Lt + k*L + l]
  bool first = true;
  for (int s = 0; s < 2; ++s)
    for (int k = 0; k < d.n_banks; ++k) {
      GemmShape g{};
      g.M = d.Cv; g.N = d.HW; g.K = d.L;
      g.sAm = d.L; g.sAk = 1; g.sBk = 1; g.sBn = W2;
      g.sCm = a.out_pixel_major ? 1 : d.HW; g.sCn = a.out_pixel_major ? a.out_channels : 1;
      g.n0 = d.B; g.n1 = d.N; g.n2 = 1;
      g.bA[0] = (long long)d.N * 2 * d.Cv * d.L; g.bA[1] = 2LL * d.Cv * d.L; g.bA[2] = 0;
      g.bB[0] = (long long)d.N * d.HW * W2; g.bB[1] = (long long)d.HW * W2; g.bB[2] = 0;
      g.bC[0] = (long long)d.N * a.out_channels * d.HW; g.bC[1] = (long long)a.out_channels * d.HW; g.bC[2] = 0;
      g.accumulate = first ? 0 : 1;
      first = false;
      if (int rc = launch_gemm(a.nu[k] + (size_t)s * d.Cv * d.L, P + (size_t)s * Lt + (size_t)k * d.L,
                               a.out + (a.out_pixel_major ? (size_t)a.mem_channel : (size_t)a.mem_channel * d.HW), g, st))
        return rc;
    }
  if (weighted) return SWEM_OK;
  return launch_perm_inv(P, U, d.HW, Lt, d.topl, a.out, a.out_channels, a.s_channel, a.out_pixel_major, st);
}


// ------------------------------------------------------------------------------------------
// readout, backward (training).  Reference: autograd through modules.py:232-293 -- gradient to the raw query key
// (through the l2norm :282, the affinity, exp, both the attention P and the sorted-prefix feature S) and to the memory
// values nu of every bank; the memory keys are constants.  With a = khat^T qhat, P = softmax over (side, j) of a / tau:
//   gP      = nu^T g_mem  (+ the gradient of S, below)                  batched GEMM + perm_inv_backward_kernel
//   g_nu    = g_mem P^T                                                  batched GEMM
//   g_a     = P (gP - sum_j P gP) / tau                                  softmax_backward_rows_kernel
//   g_qhat  = sum over objects, sides, banks of khat g_a                 batched GEMMs, accumulated
//   g_q     = g_qhat / d - q (q . g_qhat) / (n d^2),  n = ||q||, d = n + eps      l2norm_backward_kernel
// S depends on the exp-affinities only through ratios, so it is the same function of P; its gradient is taken with
// respect to P and joins gP before the softmax backward (the max-subtraction cancels analytically).
// ------------------------------------------------------------------------------------------
// One warp per (u, p): redo the descending sort of both sides (same packed-key network as the forward), then
//   f_r = c0_r / (c0_r + c1_r),  g_f[r] = gS[r] - gS[topl + r]   (S = [f, 1 - f])
//   g_c0[r] = g_f c1 / (c0 + c1)^2,  g_c1[r] = -g_f c0 / (c0 + c1)^2
//   g_e(side, rank r) = sum_{r' >= r} g_c(side, r')   (the running sum at r' contains every rank <= r')
// and add it to gP at the column that holds rank r.
template <int NPL>
__global__ void __launch_bounds__(256) perm_inv_backward_kernel(const float* __restrict__ P, int U, int HW, int Lt, int topl,
                                                                const float* __restrict__ gout, int out_channels, int s_channel,
                                                                float* __restrict__ gP) {
  constexpr int N = 32 * NPL;
  constexpr int IDXB = (NPL == 1 ? 5 : NPL == 2 ? 6 : NPL == 4 ? 7 : NPL == 8 ? 8 : NPL == 16 ? 9 : 10);
  __shared__ uint32_t top[8][128];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long w = (long long)blockIdx.x * 8 + wid;
  if (w >= (long long)U * HW) return;
  const int u = (int)(w / HW), p = (int)(w % HW);
  const float* row = P + ((long long)u * HW + p) * (2 * Lt);
  uint32_t a[NPL], b[NPL];
#pragma unroll
  for (int k = 0; k < NPL; ++k) {
    const int i = lane * NPL + k;
    const uint32_t va = i < Lt ? __float_as_uint(row[i]) : 0u;
    const uint32_t vb = i < Lt ? __float_as_uint(row[Lt + i]) : 0u;
    a[k] = ((va >> (IDXB - 1)) << IDXB) | (uint32_t)i;
    b[k] = ((vb >> (IDXB - 1)) << IDXB) | (uint32_t)i;
  }
  bitonic_desc2<NPL>(a, b, lane);
  uint32_t* mytop = top[wid];
#pragma unroll
  for (int k = 0; k < NPL; ++k) {
    const int r = lane * NPL + k;
    if (r < 64) {
      mytop[r] = a[k];
      mytop[64 + r] = b[k];
    }
  }
  __syncwarp();
  // lane handles ranks lane and lane + 32 (as the forward): inclusive running sums c0, c1
  float c0[2], c1[2];
  int i0[2], i1[2];
#pragma unroll
  for (int hlf = 0; hlf < 2; ++hlf) {
    const int r = lane + 32 * hlf;
    float x0 = 0.f, x1 = 0.f;
    i0[hlf] = i1[hlf] = -1;
    if (r < topl) {
      const int j0 = mytop[r] & (N - 1), j1 = mytop[64 + r] & (N - 1);
      if (j0 < Lt) { x0 = row[j0]; i0[hlf] = j0; }
      if (j1 < Lt) { x1 = row[Lt + j1]; i1[hlf] = j1; }
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float y0 = __shfl_up_sync(0xffffffffu, x0, o);
      const float y1 = __shfl_up_sync(0xffffffffu, x1, o);
      if (lane >= o) { x0 += y0; x1 += y1; }
    }
    c0[hlf] = x0;
    c1[hlf] = x1;
  }
  const float t0 = __shfl_sync(0xffffffffu, c0[0], 31), t1 = __shfl_sync(0xffffffffu, c1[0], 31);
  c0[1] += t0;
  c1[1] += t1;
  // gradients of the running sums, then suffix sums over rank (ranks 32..63 first, their total joins ranks 0..31)
  float g0[2], g1[2];
#pragma unroll
  for (int hlf = 0; hlf < 2; ++hlf) {
    const int r = lane + 32 * hlf;
    g0[hlf] = g1[hlf] = 0.f;
    if (r < topl) {
      const float* gs = gout + ((long long)u * out_channels + s_channel) * HW + p;
      const float gf = gs[(long long)r * HW] - gs[(long long)(topl + r) * HW];
      const float den = c0[hlf] + c1[hlf];
      const float inv2 = 1.f / (den * den);
      g0[hlf] = gf * c1[hlf] * inv2;
      g1[hlf] = -gf * c0[hlf] * inv2;
    }
  }
#pragma unroll
  for (int hlf = 1; hlf >= 0; --hlf) {
    float x0 = g0[hlf], x1 = g1[hlf];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float y0 = __shfl_down_sync(0xffffffffu, x0, o);
      const float y1 = __shfl_down_sync(0xffffffffu, x1, o);
      if (lane + o < 32) { x0 += y0; x1 += y1; }
    }
    g0[hlf] = x0;
    g1[hlf] = x1;
  }
  const float h0 = __shfl_sync(0xffffffffu, g0[1], 0), h1 = __shfl_sync(0xffffffffu, g1[1], 0);
  g0[0] += h0;
  g1[0] += h1;
  float* grow = gP + ((long long)u * HW + p) * (2 * Lt);
#pragma unroll
  for (int hlf = 0; hlf < 2; ++hlf) {
    if (i0[hlf] >= 0) grow[i0[hlf]] += g0[hlf];            // distinct columns per lane: no race
    if (i1[hlf] >= 0) grow[Lt + i1[hlf]] += g1[hlf];
  }
}

// One warp per (u, p): gP row [2Lt] -> g_a = P (gP - sum P gP) / tau, in place.
__global__ void softmax_backward_rows_kernel(const float* __restrict__ P, float* __restrict__ gP, int U, int HW, int W2,
                                             float inv_tau) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= U * HW) return;
  const float* prow = P + (long long)warp * W2;
  float* grow = gP + (long long)warp * W2;
  float dot = 0.f;
  for (int j = lane; j < W2; j += 32) dot = fmaf(prow[j], grow[j], dot);
  dot = warp_sum(dot);
  for (int j = lane; j < W2; j += 32) grow[j] = prow[j] * (grow[j] - dot) * inv_tau;
}

// Memory dropout (forward: Q = E m / (sum E m + eps), E = exp((a - max a) / tau)): one warp per (u, p),
//   g_a[k] = Q[k] (gQ[k] - sum_j Q[j] gQ[j]) / tau  -  [k = argmax a] (eps / D) (sum_j Q[j] gQ[j]) / tau,   D = sum E m + eps = Z srow
// (the second term is what the reference's autograd sends through its max subtraction, which no longer cancels once eps is in
// the denominator), in place on gQ.
__global__ void drop_softmax_backward_rows_kernel(const float* __restrict__ Q, const float* __restrict__ P, float* __restrict__ gQ,
                                                  const float* __restrict__ zsum, const float* __restrict__ srow, int U, int HW, int W2,
                                                  float inv_tau, float eps) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= U * HW) return;
  const float* qrow = Q + (long long)warp * W2;
  const float* prow = P + (long long)warp * W2;
  float* grow = gQ + (long long)warp * W2;
  float dot = 0.f, best = -1.f;
  int bj = 0x7fffffff;
  for (int j = lane; j < W2; j += 32) {
    dot = fmaf(qrow[j], grow[j], dot);
    if (prow[j] > best) { best = prow[j]; bj = j; }
  }
  dot = warp_sum(dot);
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
    if (ov > best || (ov == best && oj < bj)) { best = ov; bj = oj; }
  }
  const float corr = eps / (zsum[warp] * srow[warp]) * dot;
  for (int j = lane; j < W2; j += 32) grow[j] = (qrow[j] * (grow[j] - dot) - (j == bj ? corr : 0.f)) * inv_tau;
}
__global__ void add_inplace_kernel(float* __restrict__ dst, const float* __restrict__ src, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] += src[i];
}

// g_q[b][c][p] = g_qhat / d - q (q . g_qhat) / (n d^2)   (l2norm of modules.py:7-9 along the channel axis)
__global__ void l2norm_backward_kernel(const float* __restrict__ q, const float* __restrict__ gqh, float* __restrict__ gq,
                                       int B, int C, int HW) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * HW) return;
  const int b = i / HW, p = i % HW;
  const float* qp = q + (long long)b * C * HW + p;
  const float* gp = gqh + (long long)b * C * HW + p;
  float ss = 0.f, dot = 0.f;
  for (int c = 0; c < C; ++c) {
    const float t = qp[(long long)c * HW];
    ss = fmaf(t, t, ss);
    dot = fmaf(t, gp[(long long)c * HW], dot);
  }
  const float n = sqrtf(ss), d = n + kEpsNorm;
  const float k = n > 0.f ? dot / (n * d * d) : 0.f;
  float* op = gq + (long long)b * C * HW + p;
  for (int c = 0; c < C; ++c) op[(long long)c * HW] = gp[(long long)c * HW] / d - qp[(long long)c * HW] * k;
}

size_t generic_readout_backward_workspace(const SwemDims& d) {
  const size_t U = (size_t)d.B * d.N, G = U * 2;
  const size_t Lt = (size_t)d.L * d.n_banks;
  size_t bytes = 0;
  bytes += 2 * align_up(U * d.HW * 2 * Lt * 4, 256);          // P, gP
  bytes += align_up(G * d.Ck * d.L * 4, 256) * d.n_banks;     // khat per bank
  bytes += align_up((size_t)d.B * d.HW * 4, 256);             // inv ||q||
  bytes += align_up((size_t)d.B * d.Ck * d.HW * 4, 256);      // g_qhat
  bytes += align_up(U * d.HW * 2 * Lt * 4, 256);              // memory dropout: the dropped attention Q / gradient of S
  bytes += 2 * align_up(U * d.HW * 4, 256);                   //                 row sums Z, denominators
  return bytes + 256;
}

int generic_readout_backward(const SwemReadBwdArgs& a, cudaStream_t st) {
  const SwemDims& d = a.dims;
  const int U = d.B * d.N, G = U * 2;
  const int Lt = d.L * d.n_banks, W2 = 2 * Lt;
  Arena ws(a.workspace);
  float* P = ws.take<float>((size_t)U * d.HW * W2);
  float* gP = ws.take<float>((size_t)U * d.HW * W2);
  float* khat[2] = {nullptr, nullptr};
  for (int k = 0; k < d.n_banks; ++k) khat[k] = ws.take<float>((size_t)G * d.Ck * d.L);
  float* inv_nq = ws.take<float>((size_t)d.B * d.HW);
  float* gqh = ws.take<float>((size_t)d.B * d.Ck * d.HW);
  float* Qb = ws.take<float>((size_t)U * d.HW * W2);
  float* zsum = ws.take<float>((size_t)U * d.HW);
  float* srow = ws.take<float>((size_t)U * d.HW);
  const bool drop = a.drop_mask != nullptr;
  const float* g_mem = a.grad_out + (size_t)a.mem_channel * d.HW;
  const long long so = (long long)a.out_channels * d.HW;       // per-unit stride of grad_out

  // ---- recompute P (as generic_readout_forward) ----
  pixel_inv_norm_kernel<<<(d.B * d.HW + 255) / 256, 256, 0, st>>>(a.qk, inv_nq, d.B, d.Ck, d.HW);
  SWEM_LAUNCH_CHECK();
  for (int k = 0; k < d.n_banks; ++k) {
    kappa_unit_kernel<<<(G * d.L + 127) / 128, 128, 0, st>>>(a.kappa[k], khat[k], G, d.Ck, d.L);
    SWEM_LAUNCH_CHECK();
    GemmShape g{};
    g.M = d.HW; g.N = d.L; g.K = d.Ck;
    g.sAm = 1; g.sAk = d.HW; g.sBk = d.L; g.sBn = 1; g.sCm = W2; g.sCn = 1;
    g.n0 = d.B; g.n1 = d.N; g.n2 = 2;
    g.bA[0] = (long long)d.Ck * d.HW; g.bA[1] = 0; g.bA[2] = 0;
    g.bB[0] = (long long)d.N * 2 * d.Ck * d.L; g.bB[1] = 2LL * d.Ck * d.L; g.bB[2] = (long long)d.Ck * d.L;
    g.bC[0] = (long long)d.N * d.HW * W2; g.bC[1] = (long long)d.HW * W2; g.bC[2] = Lt;
    if (int rc = launch_gemm(a.qk, khat[k], P + (size_t)k * d.L, g, st)) return rc;
  }
  {
    const long long threads = (long long)U * d.HW * 32;
    readout_rows_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(P, inv_nq, U, d.N, d.HW, W2, 1.f / d.tau, drop ? zsum : nullptr);
    SWEM_LAUNCH_CHECK();
    if (drop) {                                                // the attention the forward used: Q = P m / (sum P m + 1e-6 / Z)
      mkm_apply_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(P, zsum, nullptr, a.drop_mask, U, d.HW, W2, 0, 1, 0.f, 1e-6f, Qb, srow);
      SWEM_LAUNCH_CHECK();
    }
  }
  const float* Pm = drop ? Qb : P;                             // what multiplied nu in the forward
  // ---- gP = nu^T g_mem ;  g_nu = g_mem P^T ----
  for (int s = 0; s < 2; ++s)
    for (int k = 0; k < d.n_banks; ++k) {
      const float* nu_sk = a.nu[k] + (size_t)s * d.Cv * d.L;
      float* col = gP + (size_t)s * Lt + (size_t)k * d.L;
      {  // gP[u][p][col + l] = sum_dch g_mem[u][dch][p] nu_k[u,s][dch][l]
        GemmShape g{};
        g.M = d.HW; g.N = d.L; g.K = d.Cv;
        g.sAm = 1; g.sAk = d.HW; g.sBk = d.L; g.sBn = 1; g.sCm = W2; g.sCn = 1;
        g.n0 = d.B; g.n1 = d.N; g.n2 = 1;
        g.bA[0] = (long long)d.N * so; g.bA[1] = so; g.bA[2] = 0;
        g.bB[0] = (long long)d.N * 2 * d.Cv * d.L; g.bB[1] = 2LL * d.Cv * d.L; g.bB[2] = 0;
        g.bC[0] = (long long)d.N * d.HW * W2; g.bC[1] = (long long)d.HW * W2; g.bC[2] = 0;
        if (int rc = launch_gemm(g_mem, nu_sk, col, g, st)) return rc;
      }
      if (a.grad_nu[k] != nullptr) {  // g_nu_k[u,s][dch][l] = sum_p g_mem[u][dch][p] P[u][p][col + l]
        GemmShape g{};
        g.M = d.Cv; g.N = d.L; g.K = d.HW;
        g.sAm = d.HW; g.sAk = 1; g.sBk = W2; g.sBn = 1; g.sCm = d.L; g.sCn = 1;
        g.n0 = d.B; g.n1 = d.N; g.n2 = 1;
        g.bA[0] = (long long)d.N * so; g.bA[1] = so; g.bA[2] = 0;
        g.bB[0] = (long long)d.N * d.HW * W2; g.bB[1] = (long long)d.HW * W2; g.bB[2] = 0;
        g.bC[0] = (long long)d.N * 2 * d.Cv * d.L; g.bC[1] = 2LL * d.Cv * d.L; g.bC[2] = 0;
        if (int rc = launch_gemm(g_mem, Pm + (size_t)s * Lt + (size_t)k * d.L, a.grad_nu[k] + (size_t)s * d.Cv * d.L, g, st)) return rc;
      }
    }
  if (a.grad_qk == nullptr) return SWEM_OK;
  // ---- + gradient of S, softmax backward ----
  {
    const long long warps = (long long)U * d.HW;
    const int npl = (Lt + 31) / 32;
    const unsigned grid = (unsigned)((warps + 7) / 8);
    float* gS = gP;                                            // where the gradient of S (with respect to the plain P) is added
    if (drop) {
      // the mem_out path goes through Q: finish it first (in place on gP), then the S path on its own buffer (Q is no longer needed)
      drop_softmax_backward_rows_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(Qb, P, gP, zsum, srow, U, d.HW, W2, 1.f / d.tau, 1e-6f);
      SWEM_LAUNCH_CHECK();
      SWEM_CUDA(cudaMemsetAsync(Qb, 0, (size_t)U * d.HW * W2 * sizeof(float), st));
      gS = Qb;
    }
#define SWEM_PIB(NPL_) perm_inv_backward_kernel<NPL_><<<grid, 256, 0, st>>>(P, U, d.HW, Lt, d.topl, a.grad_out, a.out_channels, a.s_channel, gS)
    if (npl <= 1) SWEM_PIB(1);
    else if (npl <= 2) SWEM_PIB(2);
    else if (npl <= 4) SWEM_PIB(4);
    else if (npl <= 8) SWEM_PIB(8);
    else if (npl <= 16) SWEM_PIB(16);
    else SWEM_PIB(32);
#undef SWEM_PIB
    SWEM_LAUNCH_CHECK();
    softmax_backward_rows_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(P, gS, U, d.HW, W2, 1.f / d.tau);
    SWEM_LAUNCH_CHECK();
    if (drop) {
      const long long n = (long long)U * d.HW * W2;
      add_inplace_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(gP, gS, n);
      SWEM_LAUNCH_CHECK();
    }
  }
  // ---- g_qhat[b][c][p] = sum_{n,s,k,l} khat_k[b,n,s][c][l] g_a[b,n][p][s*Lt + k*L + l] (objects accumulate serially) ----
  bool first = true;
  for (int n = 0; n < d.N; ++n)
    for (int s = 0; s < 2; ++s)
      for (int k = 0; k < d.n_banks; ++k) {
        GemmShape g{};
        g.M = d.Ck; g.N = d.HW; g.K = d.L;
        g.sAm = d.L; g.sAk = 1; g.sBk = 1; g.sBn = W2; g.sCm = d.HW; g.sCn = 1;
        g.n0 = d.B; g.n1 = 1; g.n2 = 1;
        g.bA[0] = (long long)d.N * 2 * d.Ck * d.L; g.bB[0] = (long long)d.N * d.HW * W2; g.bC[0] = (long long)d.Ck * d.HW;
        g.accumulate = first ? 0 : 1;
        first = false;
        if (int rc = launch_gemm(khat[k] + ((size_t)n * 2 + s) * d.Ck * d.L,
                                 gP + (size_t)n * d.HW * W2 + (size_t)s * Lt + (size_t)k * d.L, gqh, g, st))
          return rc;
      }
  l2norm_backward_kernel<<<(d.B * d.HW + 255) / 256, 256, 0, st>>>(a.qk, gqh, a.grad_qk, d.B, d.Ck, d.HW);
  SWEM_LAUNCH_CHECK();
  return SWEM_OK;
}

}  // namespace swem
