// Element-wise glue of the decoder around the memory path, channels-last (NHWC) fp32 -- SURVEY section 8f rank 3.
//
// The reference runs `UpsampleBlock.forward` (methods/basic_modules/networks.py:192-196) and the tail of
// `ResBlock.forward` (:25-32) as separate passes over (objects x 256 channels x 1/4-resolution) tensors: conv bias
// add, bilinear up-sampling, skip add, ReLU, residual add -- each a full read + write of up to 133 MB at 480p with 5
// objects.  These two kernels do the same arithmetic in one pass each:
//
//   upsample_add:  x = skip[b] + bilinear(lo_a [+ lo_b]) + bias      (also writes relu(x), the next conv's input)
//   bias_add_act:  out = act(a [+ b] [+ c shared by the n objects] + bias)
//   glu_gate:      the GLU of the fusion layer on [layer_f | layer_a] pre-activations, object-independent part added in
//   maxpool3x3s2:  the 3x3 / stride-2 pooling after both ResNet stems (ATen's NHWC pooling kernel runs at ~1 TB/s)
//
// `skip` is shared by the n objects of a batch element (the reference recomputes skip_conv per object), the biases of the
// convolutions that produced lo_a / lo_b / skip are passed here as one per-channel vector (bilinear interpolation
// reproduces constants), so those convolutions run without a bias pass.  Index arithmetic of the interpolation follows
// ATen's upsample_bilinear2d with align_corners = false.  HBM-bound: one thread per (pixel, 4 channels), float4 accesses.
#include <cuda_bf16.h>

#include "common.cuh"

namespace swem {

__device__ __forceinline__ float4 f4add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4fma(float s, float4 a, float4 acc) {
  return make_float4(fmaf(s, a.x, acc.x), fmaf(s, a.y, acc.y), fmaf(s, a.z, acc.z), fmaf(s, a.w, acc.w));
}
__device__ __forceinline__ float4 f4relu(float4 a) {
  return make_float4(fmaxf(a.x, 0.f), fmaxf(a.y, 0.f), fmaxf(a.z, 0.f), fmaxf(a.w, 0.f));
}

// grid (ceil(W * C4 / 256), H, BN): the row and the image come from the block index, so a thread's index arithmetic is
// one 32-bit division (the flat 64-bit index of the first version cost four 64-bit div/mods per float4: ALU-bound at 2.5 TB/s)
__global__ void __launch_bounds__(256) upsample_add_kernel(const float4* __restrict__ lo_a, const float4* __restrict__ lo_b,
                                                           const float4* __restrict__ bias, const float4* __restrict__ skip,
                                                           int BN, int n, int h, int w, int H, int W, int C4,
                                                           float4* __restrict__ x, float4* __restrict__ xr) {
  const int t = blockIdx.x * 256 + threadIdx.x;
  if (t >= W * C4) return;
  const int X = t / C4, c4 = t - X * C4;
  const int Y = blockIdx.y, bn = blockIdx.z;
  const float sh = (float)h / (float)H, sw = (float)w / (float)W;
  float sy = sh * (Y + 0.5f) - 0.5f, sx = sw * (X + 0.5f) - 0.5f;
  sy = sy < 0.f ? 0.f : sy;
  sx = sx < 0.f ? 0.f : sx;
  const int y0 = min((int)sy, h - 1), x0 = min((int)sx, w - 1);
  const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
  const float ly = sy - y0, lx = sx - x0, my = 1.f - ly, mx = 1.f - lx;
  const float4* pa = lo_a + (long long)bn * h * w * C4 + c4;
  const int o00 = (y0 * w + x0) * C4, o01 = (y0 * w + x1) * C4, o10 = (y1 * w + x0) * C4, o11 = (y1 * w + x1) * C4;
  float4 v00 = __ldg(pa + o00), v01 = __ldg(pa + o01), v10 = __ldg(pa + o10), v11 = __ldg(pa + o11);
  if (lo_b != nullptr) {
    const float4* pb = lo_b + (long long)bn * h * w * C4 + c4;
    v00 = f4add(v00, __ldg(pb + o00));
    v01 = f4add(v01, __ldg(pb + o01));
    v10 = f4add(v10, __ldg(pb + o10));
    v11 = f4add(v11, __ldg(pb + o11));
  }
  // my * (mx * v00 + lx * v01) + ly * (mx * v10 + lx * v11), in ATen's order
  float4 top = make_float4(mx * v00.x, mx * v00.y, mx * v00.z, mx * v00.w);
  top = f4fma(lx, v01, top);
  float4 bot = make_float4(mx * v10.x, mx * v10.y, mx * v10.z, mx * v10.w);
  bot = f4fma(lx, v11, bot);
  float4 out = make_float4(my * top.x, my * top.y, my * top.z, my * top.w);
  out = f4fma(ly, bot, out);
  const int b = bn / n;
  out = f4add(out, __ldg(skip + ((long long)b * H + Y) * W * C4 + t));
  if (bias != nullptr) out = f4add(out, __ldg(bias + c4));
  const long long idx = ((long long)bn * H + Y) * W * C4 + t;
  x[idx] = out;
  if (xr != nullptr) xr[idx] = f4relu(out);
}

// Exact 2x up-sampling (H = 2h, W = 2w -- both decoder stages): output rows {2i+1, 2i+2} x columns {2j+1, 2j+2} read the
// same four low-resolution pixels (i, i+1) x (j, j+1) with weights 3/4 and 1/4, so a thread computes that 2 x 2 block from
// 4 (+4) loads instead of 16 (+16): grid (ceil((w+1) * C4 / 256), h + 1, BN), block row i = blockIdx.y - 1.  Same
// operation order as the general kernel (the weights 0.25 / 0.75 are what its index arithmetic produces), so the results
// are identical.
__global__ void __launch_bounds__(256) upsample2x_add_kernel(const float4* __restrict__ lo_a, const float4* __restrict__ lo_b,
                                                             const float4* __restrict__ bias, const float4* __restrict__ skip,
                                                             int n, int h, int w, int C4, float4* __restrict__ x,
                                                             float4* __restrict__ xr) {
  const int t = blockIdx.x * 256 + threadIdx.x;
  if (t >= (w + 1) * C4) return;
  const int jb = t / C4, c4 = t - jb * C4;
  const int i = (int)blockIdx.y - 1, j = jb - 1, bn = blockIdx.z;
  const int H = 2 * h, W = 2 * w;
  const int ya = max(i, 0), yb = min(i + 1, h - 1), xa = max(j, 0), xb = min(j + 1, w - 1);
  const long long lbase = (long long)bn * h * w * C4 + c4;
  const int oaa = (ya * w + xa) * C4, oab = (ya * w + xb) * C4, oba = (yb * w + xa) * C4, obb = (yb * w + xb) * C4;
  float4 vaa = __ldg(lo_a + lbase + oaa), vab = __ldg(lo_a + lbase + oab), vba = __ldg(lo_a + lbase + oba), vbb = __ldg(lo_a + lbase + obb);
  if (lo_b != nullptr) {
    vaa = f4add(vaa, __ldg(lo_b + lbase + oaa));
    vab = f4add(vab, __ldg(lo_b + lbase + oab));
    vba = f4add(vba, __ldg(lo_b + lbase + oba));
    vbb = f4add(vbb, __ldg(lo_b + lbase + obb));
  }
  const float4 bs = bias != nullptr ? __ldg(bias + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
  const int b = bn / n;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int Y = 2 * i + 1 + r;
    if (Y < 0 || Y >= H) continue;
    // general kernel at this row: sy = max(i + 0.25 + 0.5 r, 0) -> weight 0 at Y = 0 (clamped); at Y = H - 1 both source
    // rows are row h - 1 (ya = yb) with weights 3/4 + 1/4, as there
    const float ly = (Y == 0) ? 0.f : (r ? 0.75f : 0.25f), my = 1.f - ly;
    const float4 r0a = vaa, r0b = vab;                   // source row y0 (left, right)
    const float4 r1a = vba, r1b = vbb;                   // source row y1
#pragma unroll
    for (int cidx = 0; cidx < 2; ++cidx) {
      const int X = 2 * j + 1 + cidx;
      if (X < 0 || X >= W) continue;
      const float lx = (X == 0) ? 0.f : (cidx ? 0.75f : 0.25f), mx = 1.f - lx;
      const float4 v00 = r0a, v01 = r0b, v10 = r1a, v11 = r1b;
      float4 top = make_float4(mx * v00.x, mx * v00.y, mx * v00.z, mx * v00.w);
      top = f4fma(lx, v01, top);
      float4 bot = make_float4(mx * v10.x, mx * v10.y, mx * v10.z, mx * v10.w);
      bot = f4fma(lx, v11, bot);
      float4 out = make_float4(my * top.x, my * top.y, my * top.z, my * top.w);
      out = f4fma(ly, bot, out);
      out = f4add(out, __ldg(skip + (((long long)b * H + Y) * W + X) * C4 + c4));
      if (bias != nullptr) out = f4add(out, bs);
      const long long idx = (((long long)bn * H + Y) * W + X) * C4 + c4;
      x[idx] = out;
      if (xr != nullptr) xr[idx] = f4relu(out);
    }
  }
}

// x [pixels][C4] -> hi [pixels][C4] and hl [pixels][2 C4] = [hi | lo]; hi = x rounded to TF32 (sign-magnitude bit pattern +
// half ulp, low 13 mantissa bits cleared), lo = x - hi (exact in fp32)
__device__ __forceinline__ float tf32_hi(float v) { return __int_as_float((__float_as_int(v) + 0x1000) & ~0x1fff); }
__global__ void __launch_bounds__(256) tf32_split_kernel(const float4* __restrict__ x, long long total4, int C4,
                                                         float4* __restrict__ hi_out, float4* __restrict__ hl_out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total4) return;
  const long long px = idx / C4;
  const int c4 = (int)(idx - px * C4);
  const float4 v = __ldg(x + idx);
  const float4 hi = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
  hi_out[idx] = hi;
  float4* o = hl_out + px * 2 * C4 + c4;
  o[0] = hi;
  o[C4] = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
}

__global__ void __launch_bounds__(256) tf32_split_bf16_kernel(const float4* __restrict__ x, long long total4, int C4,
                                                              float4* __restrict__ hi_out, uint2* __restrict__ xl_out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total4) return;
  const long long px = idx / C4;
  const int c4 = (int)(idx - px * C4);
  const float4 v = __ldg(x + idx);
  const float4 hi = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
  hi_out[idx] = hi;
  auto pack = [](float a, float b, float c, float d) {
    const __nv_bfloat162 p0 = __floats2bfloat162_rn(a, b), p1 = __floats2bfloat162_rn(c, d);
    return make_uint2(*reinterpret_cast<const uint32_t*>(&p0), *reinterpret_cast<const uint32_t*>(&p1));
  };
  uint2* o = xl_out + px * 2 * C4 + c4;
  o[0] = pack(v.x, v.y, v.z, v.w);
  o[C4] = pack(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
}

// out = float(x_bf16) (+ add) (+ add2): the bf16 cross-term convolution's output widened for the fused epilogue of the TF32 main-term
// convolution (which takes an fp32 addend) with the residual of the block folded in, or accumulated in place onto the main-term
// convolution's output (convolutions without a ReLU have no fused cuDNN epilogue) -- one vectorised pass (8 elements per thread)
// instead of ATen's converting copy followed by one or two in-place adds.
__global__ void __launch_bounds__(256) bf16_widen_add_kernel(const uint4* __restrict__ x, const float4* add, const float4* __restrict__ add2,
                                                             long long total8, float4* out) {   // (`out` may be `add`: in place)
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total8) return;
  const uint4 v = __ldg(x + idx);
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
  float f[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) {                      // bf16 -> fp32: the 16 bits are the upper half of the float
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
  float4 a = make_float4(f[0], f[1], f[2], f[3]), b = make_float4(f[4], f[5], f[6], f[7]);
  if (add != nullptr) {
    a = f4add(a, add[2 * idx]);
    b = f4add(b, add[2 * idx + 1]);
  }
  if (add2 != nullptr) {
    a = f4add(a, __ldg(add2 + 2 * idx));
    b = f4add(b, __ldg(add2 + 2 * idx + 1));
  }
  out[2 * idx] = a;
  out[2 * idx + 1] = b;
}

__global__ void __launch_bounds__(256) bias_add_act_kernel(const float4* __restrict__ a, const float4* __restrict__ b,
                                                           const float4* __restrict__ c, const float4* __restrict__ bias,
                                                           long long total4, long long per_image4, int n_share, int C4, int relu,
                                                           float4* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total4) return;
  float4 v = __ldg(a + idx);
  if (b != nullptr) v = f4add(v, __ldg(b + idx));
  if (c != nullptr) v = f4add(v, __ldg(c + (idx / (per_image4 * n_share)) * per_image4 + idx % per_image4));   // shared by n_share images
  if (bias != nullptr) v = f4add(v, __ldg(bias + (int)(idx % C4)));
  out[idx] = relu ? f4relu(v) : v;
}

// GLU gate of the fusion layer (modules.py:13-26): y [BN, pixels, 2C] = [layer_f | layer_a] pre-activations of the
// per-object conv, `shared` [BN / n, pixels, 2C] the object-independent part (+ bias [2C]):
//   out[bn, p, c] = (y_f + s_f + b_f) * sigmoid(y_a + s_a + b_a)
__global__ void __launch_bounds__(256) glu_gate_kernel(const float4* __restrict__ y, const float4* __restrict__ shared,
                                                       const float4* __restrict__ bias, long long total4, long long pixels, int n,
                                                       int C4, float4* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // over [BN][pixels][C4]
  if (idx >= total4) return;
  const int c4 = (int)(idx % C4);
  const long long px = idx / C4;                                               // bn * pixels + p
  const long long spx = (px / (pixels * n)) * pixels + px % pixels;
  float4 f = __ldg(y + px * 2 * C4 + c4), g = __ldg(y + px * 2 * C4 + C4 + c4);
  if (shared != nullptr) {
    f = f4add(f, __ldg(shared + spx * 2 * C4 + c4));
    g = f4add(g, __ldg(shared + spx * 2 * C4 + C4 + c4));
  }
  if (bias != nullptr) {
    f = f4add(f, __ldg(bias + c4));
    g = f4add(g, __ldg(bias + C4 + c4));
  }
  out[idx] = make_float4(f.x / (1.f + expf(-g.x)), f.y / (1.f + expf(-g.y)), f.z / (1.f + expf(-g.z)), f.w / (1.f + expf(-g.w)));
}

// 3x3 / stride 2 / padding 1 max pooling, NHWC (the stem of both ResNet trunks: torchvision resnet.py `maxpool`,
// used by methods/basic_modules/networks.py:150 and mod_resnet.py); out-of-range taps are skipped (= -inf padding).
__global__ void __launch_bounds__(256) maxpool3x3s2_kernel(const float4* __restrict__ in, int N, int H, int W, int C4, int Ho, int Wo,
                                                           float4* __restrict__ out) {
  // grid (ceil(Wo * C4 / 256), Ho, N): row and image from the block index, one 32-bit division per thread
  const int t = blockIdx.x * 256 + threadIdx.x;
  if (t >= Wo * C4) return;
  const int px = t / C4, c4 = t - px * C4;
  const int py = blockIdx.y, n = blockIdx.z;
  const long long idx = ((long long)n * Ho + py) * Wo * C4 + t;
  float4 m = make_float4(-3.4e38f, -3.4e38f, -3.4e38f, -3.4e38f);
#pragma unroll
  for (int ky = -1; ky <= 1; ++ky) {
    const int y = 2 * py + ky;
    if (y < 0 || y >= H) continue;
#pragma unroll
    for (int kx = -1; kx <= 1; ++kx) {
      const int x = 2 * px + kx;
      if (x < 0 || x >= W) continue;
      const float4 v = __ldg(in + (((long long)n * H + y) * W + x) * C4 + c4);
      m = make_float4(fmaxf(m.x, v.x), fmaxf(m.y, v.y), fmaxf(m.z, v.z), fmaxf(m.w, v.w));
    }
  }
  out[idx] = m;
}

// Tail of the decoder in one pass: the last ResBlock's residual add (networks.py:25-32), the ReLU and the 3x3 `pred`
// conv to ONE logit plane (networks.py:205-213):  out[y, x] = bp + sum_{dy,dx,c} wp[dy][dx][c] relu(a + b + bias)[y+dy-1, x+dx-1, c]
// (zero padding).  A 256 -> 1 conv is pure bandwidth (cuDNN: 106 us for 133 MB at 480p / 5 objects) on top of the 133 MB
// write + read of its input; here a and b are read once (+ halo) and nothing but the logit plane is written.
// CTA = 16 x 32 output pixels.  Stage 1 streams the 18 x 34 halo tile once: 8 lanes share a pixel (lane slot s takes the
// float4 channel groups s, s + 8, ...), four pixels per lane group in flight; every value is used straight from registers
// for its nine tap products (weights: 9 x C floats in shared memory, one broadcast LDS.128 per 16 FMAs), the lane group
// reduces the 4 x 9 partial dot products by shuffles and leaves them in shared memory [halo pixel][9].  Stage 2: each
// output pixel adds its nine neighbours' partials.  HBM-bound (the first version kept activations in shared memory
// and re-read them per tap: 143 us, LDS-bound).
constexpr int kTpTH = 16, kTpTW = 32, kTpHW = kTpTW + 2, kTpHalo = (kTpTH + 2) * kTpHW;
__global__ void __launch_bounds__(256, 2) resblock_tail_pred_kernel(const float4* __restrict__ a, const float4* __restrict__ b,
                                                                    const float4* __restrict__ bias, const float4* __restrict__ wp,
                                                                    float bp, int H, int W, int C4, float* __restrict__ out) {
  extern __shared__ float4 tp_smem[];
  float4* wsm = tp_smem;                                         // [9][C4] tap weights
  float* part = reinterpret_cast<float*>(tp_smem + 9 * C4);      // [kTpHalo][9] partial dot products
  const int tid = threadIdx.x, slot = tid & 7, grp = tid >> 3;   // 32 lane groups of 8
  const int x0 = blockIdx.x * kTpTW, y0 = blockIdx.y * kTpTH, bn = blockIdx.z;
  const long long img = (long long)bn * H * W;
  for (int i = tid; i < 9 * C4; i += 256) wsm[i] = __ldg(wp + i);
  __syncthreads();
  constexpr int P = 4;
  for (int base = 0; base < kTpHalo; base += 32 * P) {
    float acc[P][9];
    long long g[P];
    bool ok[P];
#pragma unroll
    for (int i = 0; i < P; ++i) {
      const int px = base + i * 32 + grp;
      const int y = y0 - 1 + px / kTpHW, x = x0 - 1 + px % kTpHW;
      ok[i] = px < kTpHalo && y >= 0 && y < H && x >= 0 && x < W;
      g[i] = (img + (long long)y * W + x) * C4 + slot;
#pragma unroll
      for (int t = 0; t < 9; ++t) acc[i][t] = 0.f;
    }
#pragma unroll 2
    for (int c4 = slot; c4 < C4; c4 += 8) {
      float4 v[P];
      const float4 bs = __ldg(bias + c4);
#pragma unroll
      for (int i = 0; i < P; ++i) {
        v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok[i]) v[i] = f4relu(f4add(f4add(__ldg(a + g[i] + (c4 - slot)), __ldg(b + g[i] + (c4 - slot))), bs));
      }
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const float4 w = wsm[t * C4 + c4];
#pragma unroll
        for (int i = 0; i < P; ++i)
          acc[i][t] = fmaf(w.x, v[i].x, fmaf(w.y, v[i].y, fmaf(w.z, v[i].z, fmaf(w.w, v[i].w, acc[i][t]))));
      }
    }
#pragma unroll
    for (int i = 0; i < P; ++i) {
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        float r = acc[i][t];
        r += __shfl_xor_sync(0xffffffffu, r, 1);
        r += __shfl_xor_sync(0xffffffffu, r, 2);
        r += __shfl_xor_sync(0xffffffffu, r, 4);
        acc[i][t] = r;
      }
      const int px = base + i * 32 + grp;
      if (px < kTpHalo) {
#pragma unroll
        for (int t = 0; t < 9; ++t)
          if ((t & 7) == slot) part[px * 9 + t] = acc[i][t];
      }
    }
  }
  __syncthreads();
  const int tx = tid & 31, ty = tid >> 5;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int oy = ty + 8 * half;
    float r = bp;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) r += part[((oy + dy) * kTpHW + tx + dx) * 9 + dy * 3 + dx];
    if (x0 + tx < W && y0 + oy < H) out[img + (long long)(y0 + oy) * W + x0 + tx] = r;
  }
}

// ------------------------------------------------------------------------------------------------------
// CBAM of the value encoder's fuser (methods/basic_modules/attentions.py:22-85, used at networks.py:35-52):
//   channel gate  g_c = sigmoid(mlp(avg_HW x) + mlp(max_HW x)),  mlp = Linear(C, C/16) -> ReLU -> Linear(C/16, C)
//   spatial gate  g_p = sigmoid(conv7x7([max_c (x g_c), mean_c (x g_c)]))
//   out = x + x g_c g_p            (the fuser adds the attention output to its input, networks.py:49)
// The reference spends ~14 launches on [objects, 512, H/16, W/16]; here: channel statistics, the tiny MLP, the
// per-pixel channel pooling and the final gate are one pass each (the 2 -> 1 channel 7x7 conv stays cuDNN).  NHWC.
// ------------------------------------------------------------------------------------------------------
// grid (C/32, images), 256 threads = 32 pixel lanes x 8 float4: mean / max over the pixels of 32 channels of one image
__global__ void __launch_bounds__(256) cbam_channel_stats_kernel(const float4* __restrict__ x, int pixels, int C4,
                                                                 float* __restrict__ mean, float* __restrict__ mx) {
  __shared__ float4 ssum[32][8], smax[32][8];
  const int c4 = blockIdx.x * 8 + (threadIdx.x & 7), pl = threadIdx.x >> 3, img = blockIdx.y;
  const float4* xp = x + (long long)img * pixels * C4 + c4;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), m = make_float4(-3.4e38f, -3.4e38f, -3.4e38f, -3.4e38f);
  for (int p = pl; p < pixels; p += 32) {
    const float4 v = __ldg(xp + (long long)p * C4);
    s = f4add(s, v);
    m = make_float4(fmaxf(m.x, v.x), fmaxf(m.y, v.y), fmaxf(m.z, v.z), fmaxf(m.w, v.w));
  }
  ssum[pl][threadIdx.x & 7] = s;
  smax[pl][threadIdx.x & 7] = m;
  __syncthreads();
  if (threadIdx.x < 8) {
    for (int i = 1; i < 32; ++i) {
      const float4 a = ssum[i][threadIdx.x], b = smax[i][threadIdx.x];
      s = f4add(s, a);
      m = make_float4(fmaxf(m.x, b.x), fmaxf(m.y, b.y), fmaxf(m.z, b.z), fmaxf(m.w, b.w));
    }
    const float inv = 1.f / (float)pixels;
    const int c = (blockIdx.x * 8 + threadIdx.x) * 4;
    float* mo = mean + (long long)img * C4 * 4 + c;
    float* xo = mx + (long long)img * C4 * 4 + c;
    mo[0] = s.x * inv; mo[1] = s.y * inv; mo[2] = s.z * inv; mo[3] = s.w * inv;
    xo[0] = m.x; xo[1] = m.y; xo[2] = m.z; xo[3] = m.w;
  }
}

// one block per image: gate[c] = sigmoid(W2 (relu(W1 avg + b1) + relu(W1 max + b1)) + 2 b2);  W1 [R][C], W2 [C][R], R <= 64
__global__ void __launch_bounds__(256) cbam_channel_mlp_kernel(const float* __restrict__ mean, const float* __restrict__ mx,
                                                               const float* __restrict__ w1, const float* __restrict__ b1,
                                                               const float* __restrict__ w2, const float* __restrict__ b2, int C, int R,
                                                               float* __restrict__ gate) {
  __shared__ float hid[64];
  const int img = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* a = mean + (long long)img * C;
  const float* m = mx + (long long)img * C;
  for (int j = warp; j < R; j += 8) {                    // one warp per hidden unit
    float sa = 0.f, sm = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float w = __ldg(w1 + (long long)j * C + c);
      sa = fmaf(w, a[c], sa);
      sm = fmaf(w, m[c], sm);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sa += __shfl_xor_sync(0xffffffffu, sa, o);
      sm += __shfl_xor_sync(0xffffffffu, sm, o);
    }
    if (lane == 0) hid[j] = fmaxf(sa + b1[j], 0.f) + fmaxf(sm + b1[j], 0.f);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    float t = 2.f * b2[c];
    for (int j = 0; j < R; ++j) t = fmaf(__ldg(w2 + (long long)c * R + j), hid[j], t);
    gate[(long long)img * C + c] = 1.f / (1.f + expf(-t));
  }
}

// one warp per pixel: pooled[img][0][p] = max_c (x g_c), pooled[img][1][p] = mean_c (x g_c)   (NCHW, the 7x7 conv's input)
__global__ void __launch_bounds__(256) cbam_spatial_stats_kernel(const float4* __restrict__ x, const float4* __restrict__ gate,
                                                                 long long total_px, int pixels, int C4, float* __restrict__ pooled) {
  const long long wp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wp >= total_px) return;
  const int img = (int)(wp / pixels), p = (int)(wp % pixels);
  const float4* xp = x + wp * C4;
  const float4* gp = gate + (long long)img * C4;
  float s = 0.f, m = -3.4e38f;
  for (int c4 = lane; c4 < C4; c4 += 32) {
    const float4 v = __ldg(xp + c4), g = __ldg(gp + c4);
    const float a = v.x * g.x, b = v.y * g.y, c = v.z * g.z, d = v.w * g.w;
    s += (a + b) + (c + d);
    m = fmaxf(fmaxf(m, fmaxf(a, b)), fmaxf(c, d));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  }
  if (lane == 0) {
    pooled[((long long)img * 2 + 0) * pixels + p] = m;
    pooled[((long long)img * 2 + 1) * pixels + p] = s / (float)(C4 * 4);
  }
}

// out = x (1 + g_c sigmoid(sp_p))
__global__ void __launch_bounds__(256) cbam_apply_kernel(const float4* __restrict__ x, const float4* __restrict__ gate,
                                                         const float* __restrict__ sp, long long total4, int pixels, int C4,
                                                         float4* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total4) return;
  const int c4 = (int)(idx % C4);
  const long long px = idx / C4;
  const int img = (int)(px / pixels);
  const float gs = 1.f / (1.f + expf(-__ldg(sp + px)));
  const float4 v = __ldg(x + idx), g = __ldg(gate + (long long)img * C4 + c4);
  out[idx] = make_float4(v.x * fmaf(g.x, gs, 1.f), v.y * fmaf(g.y, gs, 1.f), v.z * fmaf(g.z, gs, 1.f), v.w * fmaf(g.w, gs, 1.f));
}

// Input of a ResNet stem (7x7 / stride 2 / padding 3 conv) in "space-to-depth" form, NHWC with a zero border:
//   out[n, Y, X, 4*ci + 2*p + q] = plane_ci[2*(Y - 2) + p, 2*(X - 2) + q]        (0 outside the image / for pad channels)
// over planes ci = 0..2: (frame - mean) / std  (networks.py:77, :115, :154), ci = 3: the object's mask, ci = 4: the mask
// of the other objects 1 - mask - background (swem.py:55-56; networks.py:117).  On this tensor the stem is a 4x4 / stride-1
// conv without padding over 4*planes (padded to Cpad) channels -- a tensor-core implicit GEMM with K = 16*Cpad instead of
// cuDNN's scalar-indexed path for 3..5 input channels (352 -> ~110 us for 5 objects at 480p).  out: [B*N, H/2+3, W/2+3, Cpad].
__global__ void __launch_bounds__(256) stem_input_kernel(const float* __restrict__ frame, const float* __restrict__ masks,
                                                         float3 mean, float3 istd, int B, int N, int planes, int H, int W,
                                                         int C4, float4* __restrict__ out) {
  const int Ho = H / 2 + 3, Wo = W / 2 + 3;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * N * Ho * Wo * C4;
  if (idx >= total) return;
  const int ci = (int)(idx % C4);                         // one plane per float4: elements (p, q) = (0,0) (0,1) (1,0) (1,1)
  const int X = (int)((idx / C4) % Wo) - 2;
  const int Y = (int)((idx / ((long long)C4 * Wo)) % Ho) - 2;
  const int bn = (int)(idx / ((long long)C4 * Wo * Ho));
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (ci < planes && Y >= 0 && Y < H / 2 && X >= 0 && X < W / 2) {
    const int b = bn / N, n = bn % N;
    const long long off = (long long)(2 * Y) * W + 2 * X;
    if (ci < 3) {
      const float* src = frame + ((long long)b * 3 + ci) * H * W + off;
      const float2 r0 = __ldg(reinterpret_cast<const float2*>(src)), r1 = __ldg(reinterpret_cast<const float2*>(src + W));
      const float m = ci == 0 ? mean.x : (ci == 1 ? mean.y : mean.z), s = ci == 0 ? istd.x : (ci == 1 ? istd.y : istd.z);
      v = make_float4((r0.x - m) / s, (r0.y - m) / s, (r1.x - m) / s, (r1.y - m) / s);
    } else {
      const float* mo = masks + ((long long)b * (N + 1) + n + 1) * H * W + off;
      const float2 r0 = __ldg(reinterpret_cast<const float2*>(mo)), r1 = __ldg(reinterpret_cast<const float2*>(mo + W));
      if (ci == 3) {
        v = make_float4(r0.x, r0.y, r1.x, r1.y);
      } else {
        const float* bg = masks + ((long long)b * (N + 1)) * H * W + off;
        const float2 g0 = __ldg(reinterpret_cast<const float2*>(bg)), g1 = __ldg(reinterpret_cast<const float2*>(bg + W));
        v = make_float4(1.f - r0.x - g0.x, 1.f - r0.y - g0.y, 1.f - r1.x - g1.x, 1.f - r1.y - g1.y);
      }
    }
  }
  out[idx] = v;
}

}  // namespace swem

using namespace swem;

extern "C" {

int swem_upsample_add(const float* lo_a, const float* lo_b, const float* bias, const float* skip, int32_t BN, int32_t n,
                      int32_t h, int32_t w, int32_t H, int32_t W, int32_t C, float* x, float* x_relu, void* stream) {
  reset_launch_count();
  SWEM_CHECK_ARG(lo_a && skip && x, "NULL pointer");
  SWEM_CHECK_ARG(BN > 0 && n > 0 && BN % n == 0 && h > 0 && w > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0,
                 "bad sizes BN=%d n=%d h=%d w=%d H=%d W=%d C=%d (C must be a multiple of 4)", BN, n, h, w, H, W, C);
  SWEM_CHECK_ARG(H <= 65535 && BN <= 65535 && (long long)h * w * (C / 4) < (1ll << 31) && (long long)W * (C / 4) < (1ll << 31),
                 "sizes beyond the kernel's index range (H=%d BN=%d)", H, BN);
  if (H == 2 * h && W == 2 * w && h >= 2 && w >= 2) {
    upsample2x_add_kernel<<<dim3((unsigned)(((w + 1) * (C / 4) + 255) / 256), (unsigned)(h + 1), (unsigned)BN), 256, 0,
                            static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4*>(lo_a), reinterpret_cast<const float4*>(lo_b), reinterpret_cast<const float4*>(bias),
        reinterpret_cast<const float4*>(skip), n, h, w, C / 4, reinterpret_cast<float4*>(x), reinterpret_cast<float4*>(x_relu));
    SWEM_LAUNCH_CHECK();
    return SWEM_OK;
  }
  upsample_add_kernel<<<dim3((unsigned)((W * (C / 4) + 255) / 256), (unsigned)H, (unsigned)BN), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(lo_a), reinterpret_cast<const float4*>(lo_b), reinterpret_cast<const float4*>(bias),
      reinterpret_cast<const float4*>(skip), BN, n, h, w, H, W, C / 4, reinterpret_cast<float4*>(x),
      reinterpret_cast<float4*>(x_relu));
  SWEM_LAUNCH_CHECK();
  return SWEM_OK;
}

int swem_bias_add_act(const float* a, const float* b, const float* c_shared, const float* bias, int32_t images, int32_t n_share,
                      int64_t pixels, int32_t C, int32_t relu, float* out, void* stream) {
  reset_launch_count();
  SWEM_CHECK_ARG(a && out, "NULL pointer");
  SWEM_CHECK_ARG(images > 0 && n_share > 0 && images % n_share == 0 && pixels > 0 && C > 0 && C % 4 == 0,
                 "bad sizes images=%d n_share=%d pixels=%lld C=%d (C must be a multiple of 4)", images, n_share, (long long)pixels, C);
  const long long per_image4 = (long long)pixels * (C / 4), total4 = per_image4 * images;
  bias_add_act_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(b), reinterpret_cast<const float4*>(c_shared),
      reinterpret_cast<const float4*>(bias), total4, per_image4, n_share, C / 4, relu, reinterpret_cast<float4*>(out));
  SWEM_LAUNCH_CHECK();
  return SWEM_OK;
}

int swem_glu_gate(const float* y, const float* shared, const float* bias, int32_t images, int32_t n_share, int64_t pixels, int32_t C,
                  float* out, void* stream) {
  reset_launch_count();
  SWEM_CHECK_ARG(y && out, "NULL pointer");
  SWEM_CHECK_ARG(images > 0 && n_share > 0 && images % n_share == 0 && pixels > 0 && C > 0 && C % 4 == 0,
                 "bad sizes images=%d n_share=%d pixels=%lld C=%d (C must be a multiple of 4)", images, n_share, (long long)pixels, C);
  const long long total4 = (long long)images * pixels * (C / 4);
  glu_gate_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(y), reinterpret_cast<const float4*>(shared), reinterpret_cast<const float4*>(bias), total4,
      pixels, n_share, C / 4, reinterpret_cast<float4*>(out));
  SWEM_LAUNCH_CHECK();
  return SWEM_OK;
}

int swem_resblock_tail_pred(const float* a, const float* b, const float* bias, const float* wp, float bp, int32_t BN, int32_t H,
                            int32_t W, int32_t C, float* out, void* stream) {
  reset_launch_count();
  SWEM_CHECK_ARG(a && b && bias && wp && out, "NULL pointer");
  SWEM_CHECK_ARG(BN > 0 && BN <= 65535 && H > 0 && W > 0 && C > 0 && C % 32 == 0, "bad sizes BN=%d H=%d W=%d C=%d (C must be a multiple of 32)",
                 BN, H, W, C);
  const size_t smem = (size_t)9 * (C / 4) * sizeof(float4) + (size_t)kTpHalo * 9 * sizeof(float);
  SWEM_CHECK_ARG(smem <= 96 * 1024, "C=%d too large for the weight tile", C);
  {
    static PerDevice once;                               // the attribute is per device: one set per GPU this process uses
    const int dev = current_device();
    std::lock_guard<std::mutex> lock(once.mu);
    if (!once.done[dev]) {
      SWEM_CUDA(cudaFuncSetAttribute(resblock_tail_pred_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
      once.done[dev] = true;
    }
  }
  dim3 grid((W + kTpTW - 1) / kTpTW, (H + kTpTH - 1) / kTpTH, BN);
  resblock_tail_pred_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(b), reinterpret_cast<const float4*>(bias),
      reinterpret_cast<const float4*>(wp), bp, H, W, C / 4, out);
  SWEM_LAUNCH_CHECK();
  return SWEM_OK;
}

int swem_cbam_channel_gate(const float* x, const float* w1, const float* b1, const float* w2, const float* b2, int32_t images,
                           int64_t pixels, int32_t C, int32_t R, float* stats, float* gate, void* stream) {
  reset_launch_count();
  SWEM_CHECK_ARG(x && w1 && b1 && w2 && b2 && stats && gate, "NULL pointer");
  SWEM_CHECK_ARG(images > 0 && images <= 65535 && pixels > 0 && C > 0 && C % 32 == 0 && R > 0 && R <= 64,
                 "bad sizes images=%d pixels=%lld C=%d R=%d (C %% 32 == 0, R <= 64)", images, (long long)pixels, C, R);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* mean = stats;
  float* mx = stats + (size_t)images * C;
  cbam_channel_stats_kernel<<<dim3(C / 32, images), 256, 0, st>>>(reinterpret_cast<const float4*>(x), (int)pixels, C / 4, mean, mx);
  SWEM_LAUNCH_CHECK();
  cbam_channel_mlp_kernel<<<images, 256, 0, st>>>(mean, mx, w1, b1, w2, b2, C, R, gate);
  SWEM_LAUNCH_CHECK();
  return SWEM_OK;
}

int swem_cbam_spatial_pool(const float* x, const float* gate, int32_t images, int64_t pixels, int32_t C, float* pooled, void* stream) {
  reset_launch_count();
  SWEM_CHECK_ARG(x && gate && pooled, "NULL pointer");
  SWEM_CHECK_ARG(images > 0 && pixels > 0 && C > 0 && C % 4 == 0, "bad sizes images=%d pixels=%lld C=%d", images, (long long)pixels, C);
  const long long total_px = (long long)images * pixels;
  cbam_spatial_stats_kernel<<<(unsigned)((total_px * 32 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(gate), total_px, (int)pixels, C / 4, pooled);
  SWEM_LAUNCH_CHECK();
  return SWEM_OK;
}

int swem_cbam_apply(const float* x, const float* gate, const float* spatial_logit, int32_t images, int64_t pixels, int32_t C,
                    float* out, void* stream) {
  reset_launch_count();
  SWEM_CHECK_ARG(x && gate && spatial_logit && out, "NULL pointer");
  SWEM_CHECK_ARG(images > 0 && pixels > 0 && C > 0 && C % 4 == 0, "bad sizes images=%d pixels=%lld C=%d", images, (long long)pixels, C);
  const long long total4 = (long long)images * pixels * (C / 4);
  cbam_apply_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(gate), spatial_logit, total4, (int)pixels, C / 4,
      reinterpret_cast<float4*>(out));
  SWEM_LAUNCH_CHECK();
  return SWEM_OK;
}

int swem_stem_input(const float* frame, const float* masks, const float* mean3, const float* std3, int32_t B, int32_t N,
                    int32_t planes, int32_t H, int32_t W, int32_t Cpad, float* out, void* stream) {
  reset_launch_count();
  SWEM_CHECK_ARG(frame && mean3 && std3 && out, "NULL pointer");
  SWEM_CHECK_ARG(B > 0 && N > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0, "bad sizes B=%d N=%d H=%d W=%d (H, W must be even)", B, N, H, W);
  SWEM_CHECK_ARG(planes >= 3 && planes <= 5 && (planes == 3 || masks != nullptr) && Cpad % 4 == 0 && Cpad >= 4 * planes,
                 "bad planes=%d / Cpad=%d", planes, Cpad);
  const long long total = (long long)B * N * (H / 2 + 3) * (W / 2 + 3) * (Cpad / 4);
  stem_input_kernel<<<(unsigned)((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      frame, masks, make_float3(mean3[0], mean3[1], mean3[2]), make_float3(std3[0], std3[1], std3[2]), B, N, planes, H, W, Cpad / 4,
      reinterpret_cast<float4*>(out));
  SWEM_LAUNCH_CHECK();
  return SWEM_OK;
}

int swem_maxpool3x3s2(const float* in, int32_t N, int32_t H, int32_t W, int32_t C, float* out, void* stream) {
  reset_launch_count();
  SWEM_CHECK_ARG(in && out, "NULL pointer");
  SWEM_CHECK_ARG(N > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0, "bad sizes N=%d H=%d W=%d C=%d (C must be a multiple of 4)", N, H, W, C);
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  SWEM_CHECK_ARG(Ho <= 65535 && N <= 65535 && (long long)Wo * (C / 4) < (1ll << 31), "sizes beyond the kernel's index range (H=%d N=%d)", H, N);
  maxpool3x3s2_kernel<<<dim3((unsigned)((Wo * (C / 4) + 255) / 256), (unsigned)Ho, (unsigned)N), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(in), N, H, W, C / 4, Ho, Wo, reinterpret_cast<float4*>(out));
  SWEM_LAUNCH_CHECK();
  return SWEM_OK;
}

int swem_tf32_split(const float* x, int64_t pixels, int32_t C, float* hi, float* hl, void* stream) {
  reset_launch_count();
  SWEM_CHECK_ARG(x && hi && hl, "NULL pointer");
  SWEM_CHECK_ARG(pixels > 0 && C > 0 && C % 4 == 0 && pixels * (C / 4) < (1ll << 40), "bad sizes pixels=%lld C=%d (C must be a multiple of 4)",
                 (long long)pixels, C);
  const long long total4 = pixels * (C / 4);
  tf32_split_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(x), total4, C / 4, reinterpret_cast<float4*>(hi), reinterpret_cast<float4*>(hl));
  SWEM_LAUNCH_CHECK();
  return SWEM_OK;
}

int swem_tf32_split_bf16(const float* x, int64_t pixels, int32_t C, float* hi, void* xl_bf16, void* stream) {
  reset_launch_count();
  SWEM_CHECK_ARG(x && hi && xl_bf16, "NULL pointer");
  SWEM_CHECK_ARG(pixels > 0 && C > 0 && C % 4 == 0 && pixels * (C / 4) < (1ll << 40), "bad sizes pixels=%lld C=%d (C must be a multiple of 4)",
                 (long long)pixels, C);
  const long long total4 = pixels * (C / 4);
  tf32_split_bf16_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(x), total4, C / 4, reinterpret_cast<float4*>(hi), reinterpret_cast<uint2*>(xl_bf16));
  SWEM_LAUNCH_CHECK();
  return SWEM_OK;
}

int swem_bf16_widen_add(const void* x_bf16, const float* add, const float* add2, int64_t n, float* out, void* stream) {
  reset_launch_count();
  SWEM_CHECK_ARG(x_bf16 && out, "NULL pointer");
  SWEM_CHECK_ARG(n > 0 && n % 8 == 0, "n=%lld must be a positive multiple of 8", (long long)n);
  SWEM_CHECK_ARG((reinterpret_cast<uintptr_t>(x_bf16) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(add) & 15) == 0 && (reinterpret_cast<uintptr_t>(add2) & 15) == 0,
                 "pointers must be 16-byte aligned");
  SWEM_CHECK_ARG(add2 == nullptr || add2 != out, "only `add` may alias `out`");
  const long long total8 = n / 8;
  bf16_widen_add_kernel<<<(unsigned)((total8 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint4*>(x_bf16), reinterpret_cast<const float4*>(add), reinterpret_cast<const float4*>(add2), total8,
      reinterpret_cast<float4*>(out));
  SWEM_LAUNCH_CHECK();
  return SWEM_OK;
}

}  // extern "C"
