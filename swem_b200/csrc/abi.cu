// extern "C" surface of libswem_b200.so: argument validation, family dispatch, mask prep.
#include <stdarg.h>
#include <string.h>

#include <nvtx3/nvToolsExt.h>

#include "common.cuh"

namespace swem {

static thread_local char g_err[512] = "";
static thread_local int g_launches = 0;
static thread_local long long g_total_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) {
  g_launches += n;
  g_total_launches += n;
}
void reset_launch_count() { g_launches = 0; }

static int check_dims(const SwemDims& d, bool em) {
  SWEM_CHECK_ARG(d.B > 0 && d.N > 0 && d.Ck > 0 && d.Cv > 0 && d.HW > 0 && d.L > 0,
                 "non-positive dimension (B=%d N=%d Ck=%d Cv=%d HW=%d L=%d)", d.B, d.N, d.Ck, d.Cv, d.HW, d.L);
  SWEM_CHECK_ARG(d.tau > 0.f, "tau must be > 0 (got %g)", (double)d.tau);   // reference modules.py:70
  if (em) {
    SWEM_CHECK_ARG(d.n_iters >= 1, "n_iters must be >= 1 (got %d)", d.n_iters);
  } else {
    SWEM_CHECK_ARG(d.n_banks == 1 || d.n_banks == 2, "n_banks must be 1 or 2 (got %d)", d.n_banks);
    SWEM_CHECK_ARG(d.topl >= 1 && d.topl <= d.L * d.n_banks, "topl=%d outside [1, Lt=%d]", d.topl, d.L * d.n_banks);
  }
  return SWEM_OK;
}

// ------------------------------------------------------------------------------------------
// mask prep (reference swem.py:80-84): nearest-resized hard mask x bilinear-resized soft mask.
// Index arithmetic follows ATen's upsample kernels (float scale = in/out; nearest: floor(dst*scale);
// bilinear, align_corners=false: src = scale*(dst+0.5)-0.5 clamped at 0).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int nearest_src(int dst, int in, int out, float scale) {
  if (in == out) return dst;
  if (out == 2 * in) return dst >> 1;
  return min((int)floorf(dst * scale), in - 1);
}

__global__ void em_masks_kernel(const long long* __restrict__ hard, int Hm, int Wm,
                                const float* __restrict__ soft, int Hs, int Ws, int B, int N, int H, int W,
                                float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * N * H * W) return;
  const int w = i % W, h = (i / W) % H, n = (i / (W * H)) % N, b = i / (W * H * N);
  const float sh_n = (float)Hm / (float)H, sw_n = (float)Wm / (float)W;
  const int hy = nearest_src(h, Hm, H, sh_n), hx = nearest_src(w, Wm, W, sw_n);
  const float hv = (float)hard[(((long long)b * (N + 1) + n + 1) * Hm + hy) * Wm + hx];

  const float sh = (float)Hs / (float)H, sw = (float)Ws / (float)W;
  float sy = sh * (h + 0.5f) - 0.5f, sx = sw * (w + 0.5f) - 0.5f;
  sy = sy < 0.f ? 0.f : sy;
  sx = sx < 0.f ? 0.f : sx;
  const int y0 = min((int)sy, Hs - 1), x0 = min((int)sx, Ws - 1);
  const int y1 = y0 + (y0 < Hs - 1 ? 1 : 0), x1 = x0 + (x0 < Ws - 1 ? 1 : 0);
  const float ly = fminf(fmaxf(sy - y0, 0.f), 1.f), lx = fminf(fmaxf(sx - x0, 0.f), 1.f);
  const float* sp = soft + ((long long)b * (N + 1) + n + 1) * Hs * Ws;
  const float sv = (1.f - ly) * ((1.f - lx) * sp[y0 * Ws + x0] + lx * sp[y0 * Ws + x1]) +
                   ly * ((1.f - lx) * sp[y1 * Ws + x0] + lx * sp[y1 * Ws + x1]);
  float* op = out + (((long long)b * N + n) * 2) * H * W + h * W + w;
  op[0] = (1.f - hv) * (1.f - sv);
  op[(long long)H * W] = hv * sv;
}

// ------------------------------------------------------------------------------------------
// decoder tail (reference networks.py:214-215 + swem.py:92-116): bilinear up-sampling of the N per-object
// logit planes to the output size, sigmoid, soft aggregation with the background product, clamp, logit,
// softmax over the N+1 classes -- one pass over the output instead of ~12 element-wise launches.
// ------------------------------------------------------------------------------------------
constexpr int kMaxTailObjs = 16;

__global__ void decode_tail_kernel(const float* __restrict__ lr, int B, int N, int Hl, int Wl, int H, int W,
                                   const float* __restrict__ valid, float* __restrict__ logits_out,
                                   float* __restrict__ prob_out, long long* __restrict__ pred_out,
                                   long long* __restrict__ hard_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * H * W) return;
  const int x = i % W, y = (i / W) % H, b = i / (W * H);
  const float sh = (float)Hl / (float)H, sw = (float)Wl / (float)W;
  float sy = sh * (y + 0.5f) - 0.5f, sx = sw * (x + 0.5f) - 0.5f;
  sy = sy < 0.f ? 0.f : sy;
  sx = sx < 0.f ? 0.f : sx;
  const int y0 = min((int)sy, Hl - 1), x0 = min((int)sx, Wl - 1);
  const int y1 = y0 + (y0 < Hl - 1 ? 1 : 0), x1 = x0 + (x0 < Wl - 1 ? 1 : 0);
  const float ly = sy - y0, lx = sx - x0, my = 1.f - ly, mx = 1.f - lx;
  float prob[kMaxTailObjs + 1];
  float bg = 1.f;
  for (int n = 0; n < N; ++n) {
    const float* pl = lr + (size_t)(b * N + n) * Hl * Wl;
    const float v = my * (mx * pl[y0 * Wl + x0] + lx * pl[y0 * Wl + x1]) + ly * (mx * pl[y1 * Wl + x0] + lx * pl[y1 * Wl + x1]);
    float pr = 1.f / (1.f + expf(-v));
    if (valid != nullptr) pr *= valid[b * (N + 1) + n + 1];
    prob[n + 1] = pr;
    bg *= 1.f - pr;
  }
  prob[0] = bg;
  float mxl = -3.0e38f;
  for (int k = 0; k <= N; ++k) {
    const float c = fminf(fmaxf(prob[k], 1e-7f), 1.f - 1e-7f);
    const float lg = logf(c / (1.f - c));
    prob[k] = lg;
    mxl = fmaxf(mxl, lg);
    logits_out[((size_t)(b * (N + 1) + k) * H + y) * W + x] = lg;
  }
  float sum = 0.f;
  for (int k = 0; k <= N; ++k) {
    prob[k] = expf(prob[k] - mxl);
    sum += prob[k];
  }
  int best = 0;
  float best_p = -1.f;
  for (int k = 0; k <= N; ++k) {
    const float pk = prob[k] / sum;
    prob_out[((size_t)(b * (N + 1) + k) * H + y) * W + x] = pk;
    if (pk > best_p) {                       // first maximum, like torch.argmax
      best_p = pk;
      best = k;
    }
  }
  // the evaluator's next two ops (swem_evaluator.py:83-87: argmax over the classes, its one-hot) on the values just written
  if (pred_out != nullptr) pred_out[((size_t)b * H + y) * W + x] = best;
  if (hard_out != nullptr)
    for (int k = 0; k <= N; ++k) hard_out[((size_t)(b * (N + 1) + k) * H + y) * W + x] = (k == best) ? 1 : 0;
}

}  // namespace swem

using namespace swem;

extern "C" {

// NVTX ranges around the hot-path entry points (tracing): visible in nsys / ncu --nvtx timelines, free otherwise
// (header-only NVTX v3: the calls are no-ops unless a tool has injected itself).
namespace {
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};
}  // namespace

int swem_abi_version(void) { return SWEM_B200_ABI_VERSION; }
const char* swem_last_error(void) { return g_err; }
int swem_last_launch_count(void) { return g_launches; }
long long swem_total_launch_count(void) { return g_total_launches; }

int swem_device_check(int device) {
  cudaDeviceProp prop;
  SWEM_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("device %d is sm_%d%d; libswem_b200 is built for sm_100a only", device, prop.major, prop.minor);
    return SWEM_ERR_DEVICE;
  }
  return SWEM_OK;
}

int swem_set_profile_buffer(void* dev, size_t bytes) {
  if (dev != nullptr && bytes < 256 * sizeof(long long)) {
    set_error("profile buffer needs >= 2048 bytes");
    return SWEM_ERR_INVALID_ARG;
  }
  set_profile_buffer(dev);
  return SWEM_OK;
}

int swem_em_fused_supported(const SwemDims* d) { return d && fused_em_supported(*d) ? 1 : 0; }
int swem_readout_fused_supported(const SwemDims* d) { return d && fused_readout_supported(*d) ? 1 : 0; }

// SWEM_PATH_AUTO is the product path: the tcgen05 family, or SWEM_ERR_UNSUPPORTED when it does not cover the shape (SURVEY
// section 8(b): unsupported shapes fail loudly, nothing is dispatched silently to a slower family).  The generic fp32 family
// runs only when SWEM_PATH_GENERIC is asked for by name (tests, debugging, training shapes outside the fused coverage).
static bool use_fused_em(const SwemDims&, int path) { return path != SWEM_PATH_GENERIC; }
static bool use_fused_readout(const SwemDims&, int path) { return path != SWEM_PATH_GENERIC; }

size_t swem_em_workspace_bytes(const SwemDims* d, int32_t path) {
  if (!d || check_dims(*d, true)) return 0;
  if (path != SWEM_PATH_GENERIC && !fused_em_supported(*d)) return 0;
  return use_fused_em(*d, path) ? fused_em_workspace(*d) : generic_em_workspace(*d);
}

int swem_em_forward(const SwemEmArgs* a, void* stream) {
  NvtxRange nvtx("swem_em_forward");
  reset_launch_count();
  SWEM_CHECK_ARG(a != nullptr, "args is NULL");
  if (int rc = check_dims(a->dims, true)) return rc;
  SWEM_CHECK_ARG(a->x && a->v && a->masks && a->kappa_prior && a->nu_prior && a->zita_prior,
                 "a required input pointer is NULL");
  SWEM_CHECK_ARG(a->kappa && a->nu && a->zita, "a required output pointer is NULL");
  SWEM_CHECK_ARG(a->path >= SWEM_PATH_AUTO && a->path <= SWEM_PATH_FUSED, "bad path %d", a->path);
  if (a->path != SWEM_PATH_GENERIC && !fused_em_supported(a->dims)) {
    set_error("the tcgen05 EM kernels do not cover Ck=%d Cv=%d L=%d HW=%d n_iters=%d (Ck in {64,128}, L in {64,128,256,512}, Cv=512); "
              "the generic fp32 family runs only on request (SWEM_PATH_GENERIC)", a->dims.Ck, a->dims.Cv, a->dims.L, a->dims.HW, a->dims.n_iters);
    return SWEM_ERR_UNSUPPORTED;
  }
  const size_t need = swem_em_workspace_bytes(&a->dims, a->path);
  if (!a->workspace || a->workspace_bytes < need || (reinterpret_cast<uintptr_t>(a->workspace) & 255)) {
    set_error("EM workspace: need %zu bytes 256-aligned, got %zu at %p", need, a->workspace_bytes, a->workspace);
    return SWEM_ERR_WORKSPACE;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (a->v_pixel_major && !use_fused_em(a->dims, a->path)) {
    set_error("v_pixel_major is implemented by the fused EM kernels only (Ck=%d Cv=%d L=%d path=%d)", a->dims.Ck, a->dims.Cv, a->dims.L, a->path);
    return SWEM_ERR_UNSUPPORTED;
  }
  if (a->image_workspace != nullptr) {
    SWEM_CHECK_ARG(a->image_n_banks >= 1 && a->image_n_banks <= 2 && a->image_bank >= 0 && a->image_bank < a->image_n_banks,
                   "image_bank=%d / image_n_banks=%d", a->image_bank, a->image_n_banks);
    SWEM_CHECK_ARG((reinterpret_cast<uintptr_t>(a->image_workspace) & 255) == 0, "image_workspace must be 256-byte aligned");
  }
  const int rc = use_fused_em(a->dims, a->path) ? fused_em_forward(*a, st) : generic_em_forward(*a, st);
  if (rc != SWEM_OK || a->image_workspace == nullptr) return rc;
  if (use_fused_em(a->dims, a->path) && fused_em_res_emits_images(*a)) return rc;          // written by the EM kernel itself
  // every other kernel: append the conversion launch for this bank (same stream, after the bases are complete)
  SwemDims rd = a->dims;
  rd.n_banks = a->image_n_banks;
  rd.topl = 1;                                                                                 // (irrelevant to the images)
  if (!fused_readout_supported(rd)) return rc;                                               // (no tcgen05 readout will read images of this shape)
  const ReadoutImages im = readout_image_layout(a->image_workspace, rd);
  const float* kk[2] = {nullptr, nullptr};
  const float* nn[2] = {nullptr, nullptr};
  kk[a->image_bank] = a->kappa;
  nn[a->image_bank] = a->nu;
  const int rc2 = launch_bank_images(rd, kk, nn, a->image_bank, a->image_bank + 1, im, st);
  return rc2;
}

size_t swem_em_backward_workspace_bytes(const SwemDims* d) {
  if (!d || d->B <= 0 || d->N <= 0 || d->Cv <= 0 || d->HW <= 0 || d->L <= 0) return 0;
  return generic_em_backward_workspace(*d);
}

int swem_em_backward(const SwemEmBwdArgs* a, void* stream) {
  NvtxRange nvtx("swem_em_backward");
  reset_launch_count();
  SWEM_CHECK_ARG(a != nullptr, "args is NULL");
  const SwemDims& d = a->dims;
  SWEM_CHECK_ARG(d.B > 0 && d.N > 0 && d.Cv > 0 && d.HW > 0 && d.L > 0,
                 "non-positive dimension (B=%d N=%d Cv=%d HW=%d L=%d)", d.B, d.N, d.Cv, d.HW, d.L);
  SWEM_CHECK_ARG(a->z_last && a->zita_prior && a->zita && a->grad_nu, "a required input pointer is NULL");
  SWEM_CHECK_ARG(a->grad_v || a->grad_nu_prior, "both output pointers are NULL");
  const size_t need = swem_em_backward_workspace_bytes(&d);
  if (!a->workspace || a->workspace_bytes < need || (reinterpret_cast<uintptr_t>(a->workspace) & 255)) {
    set_error("EM backward workspace: need %zu bytes 256-aligned, got %zu at %p", need, a->workspace_bytes, a->workspace);
    return SWEM_ERR_WORKSPACE;
  }
  return generic_em_backward(*a, static_cast<cudaStream_t>(stream));
}

size_t swem_readout_workspace_bytes(const SwemDims* d, int32_t path) {
  if (!d || check_dims(*d, false)) return 0;
  if (path != SWEM_PATH_GENERIC && !fused_readout_supported(*d)) return 0;
  return use_fused_readout(*d, path) ? fused_readout_workspace(*d) : generic_readout_workspace(*d);
}

int swem_readout_forward(const SwemReadArgs* a, void* stream) {
  NvtxRange nvtx("swem_readout_forward");
  reset_launch_count();
  SWEM_CHECK_ARG(a != nullptr, "args is NULL");
  if (int rc = check_dims(a->dims, false)) return rc;
  const SwemDims& d = a->dims;
  SWEM_CHECK_ARG(a->qk && a->out, "qk/out pointer is NULL");
  for (int k = 0; k < d.n_banks; ++k) SWEM_CHECK_ARG(a->kappa[k] && a->nu[k], "bank %d pointer is NULL", k);
  SWEM_CHECK_ARG(a->mem_channel >= 0 && a->mem_channel + d.Cv <= a->out_channels &&
                 a->s_channel >= 0 && a->s_channel + 2 * d.topl <= a->out_channels,
                 "channel placement outside out_channels=%d", a->out_channels);
  SWEM_CHECK_ARG(a->path >= SWEM_PATH_AUTO && a->path <= SWEM_PATH_FUSED, "bad path %d", a->path);
  SWEM_CHECK_ARG(a->bank_images_valid >= 0 && a->bank_images_valid < (1 << d.n_banks), "bank_images_valid=%d names a bank beyond n_banks=%d",
                 a->bank_images_valid, d.n_banks);
  SWEM_CHECK_ARG(a->out_pixel_major == 0 || (a->out_pixel_major == 1 && a->out_channels % 4 == 0 && a->mem_channel % 4 == 0),
                 "out_pixel_major=%d needs out_channels and mem_channel to be multiples of 4", a->out_pixel_major);
  SWEM_CHECK_ARG(a->out_pixel_major == 0 || (reinterpret_cast<uintptr_t>(a->out) & 15) == 0, "out_pixel_major: `out` must be 16-byte aligned");
  if (a->mkm_kernels != 0) {
    SWEM_CHECK_ARG(a->mkm_kernels > 0 && a->mkm_kernels <= 16 && a->mkm_kernels <= d.HW, "mkm_kernels=%d (1..16, <= HW)", a->mkm_kernels);
    SWEM_CHECK_ARG(a->mkm_sigma > 0.f && a->mkm_width > 0 && d.HW % a->mkm_width == 0, "mkm_sigma=%g mkm_width=%d (HW=%d)",
                   a->mkm_sigma, a->mkm_width, d.HW);
    if (a->path != SWEM_PATH_GENERIC) {
      set_error("the kernelised-memory readout (mkm_kernels=%d) is implemented by the generic family only: ask for SWEM_PATH_GENERIC", a->mkm_kernels);
      return SWEM_ERR_UNSUPPORTED;
    }
  }
  if (a->drop_mask != nullptr) {
    SWEM_CHECK_ARG(a->mkm_kernels == 0, "drop_mask and mkm_kernels exclude each other (training / inference branches of the reference)");
    if (a->path != SWEM_PATH_GENERIC) {
      set_error("memory dropout (drop_mask) is implemented by the generic family only: ask for SWEM_PATH_GENERIC");
      return SWEM_ERR_UNSUPPORTED;
    }
  }
  if (a->path != SWEM_PATH_GENERIC && !fused_readout_supported(d)) {
    set_error("the tcgen05 readout kernels do not cover Ck=%d Cv=%d L=%d banks=%d topl=%d (Ck in {64,128}, L in {64,128,256,512}, Cv=512, "
              "topl<=64); the generic fp32 family runs only on request (SWEM_PATH_GENERIC)", d.Ck, d.Cv, d.L, d.n_banks, d.topl);
    return SWEM_ERR_UNSUPPORTED;
  }
  const size_t need = swem_readout_workspace_bytes(&d, a->path);
  if (!a->workspace || a->workspace_bytes < need || (reinterpret_cast<uintptr_t>(a->workspace) & 255)) {
    set_error("readout workspace: need %zu bytes 256-aligned, got %zu at %p", need, a->workspace_bytes, a->workspace);
    return SWEM_ERR_WORKSPACE;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return use_fused_readout(d, a->path) ? fused_readout_forward(*a, st) : generic_readout_forward(*a, st);
}

size_t swem_readout_backward_workspace_bytes(const SwemDims* d) {
  if (!d || check_dims(*d, false)) return 0;
  return generic_readout_backward_workspace(*d);
}

int swem_readout_backward(const SwemReadBwdArgs* a, void* stream) {
  NvtxRange nvtx("swem_readout_backward");
  reset_launch_count();
  SWEM_CHECK_ARG(a != nullptr, "args is NULL");
  if (int rc = check_dims(a->dims, false)) return rc;
  const SwemDims& d = a->dims;
  SWEM_CHECK_ARG(a->qk && a->grad_out, "qk/grad_out pointer is NULL");
  for (int k = 0; k < d.n_banks; ++k) SWEM_CHECK_ARG(a->kappa[k] && a->nu[k], "bank %d pointer is NULL", k);
  SWEM_CHECK_ARG(a->mem_channel >= 0 && a->mem_channel + d.Cv <= a->out_channels &&
                 a->s_channel >= 0 && a->s_channel + 2 * d.topl <= a->out_channels,
                 "channel placement outside out_channels=%d", a->out_channels);
  if (d.topl > 64 || d.L * d.n_banks > 1024) {
    set_error("readout backward: topl=%d / Lt=%d unsupported", d.topl, d.L * d.n_banks);
    return SWEM_ERR_UNSUPPORTED;
  }
  const size_t need = swem_readout_backward_workspace_bytes(&d);
  if (!a->workspace || a->workspace_bytes < need || (reinterpret_cast<uintptr_t>(a->workspace) & 255)) {
    set_error("readout backward workspace: need %zu bytes 256-aligned, got %zu at %p", need, a->workspace_bytes, a->workspace);
    return SWEM_ERR_WORKSPACE;
  }
  return generic_readout_backward(*a, static_cast<cudaStream_t>(stream));
}

int swem_em_masks(const int64_t* hard, int32_t Hm, int32_t Wm, const float* soft, int32_t Hs, int32_t Ws,
                  int32_t B, int32_t N, int32_t H16, int32_t W16, float* out, void* stream) {
  reset_launch_count();
  SWEM_CHECK_ARG(hard && soft && out, "NULL pointer");
  SWEM_CHECK_ARG(Hm > 0 && Wm > 0 && Hs > 0 && Ws > 0 && B > 0 && N > 0 && H16 > 0 && W16 > 0, "non-positive size");
  const int n = B * N * H16 * W16;
  em_masks_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const long long*>(hard), Hm, Wm, soft, Hs, Ws, B, N, H16, W16, out);
  SWEM_LAUNCH_CHECK();
  return SWEM_OK;
}

int swem_decode_tail_masks(const float* logits_lr, int32_t B, int32_t N, int32_t Hl, int32_t Wl, int32_t H, int32_t W,
                           const float* valid_obj, float* logits_out, float* prob_out, int64_t* pred_out, int64_t* hard_out,
                           void* stream) {
  reset_launch_count();
  SWEM_CHECK_ARG(logits_lr && logits_out && prob_out, "NULL pointer");
  SWEM_CHECK_ARG(B > 0 && N > 0 && Hl > 0 && Wl > 0 && H > 0 && W > 0, "non-positive size");
  if (N > kMaxTailObjs) {
    set_error("decode tail supports at most %d objects (got %d)", kMaxTailObjs, N);
    return SWEM_ERR_UNSUPPORTED;
  }
  const int n = B * H * W;
  decode_tail_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      logits_lr, B, N, Hl, Wl, H, W, valid_obj, logits_out, prob_out, reinterpret_cast<long long*>(pred_out),
      reinterpret_cast<long long*>(hard_out));
  SWEM_LAUNCH_CHECK();
  return SWEM_OK;
}

int swem_decode_tail(const float* logits_lr, int32_t B, int32_t N, int32_t Hl, int32_t Wl, int32_t H, int32_t W,
                     const float* valid_obj, float* logits_out, float* prob_out, void* stream) {
  return swem_decode_tail_masks(logits_lr, B, N, Hl, Wl, H, W, valid_obj, logits_out, prob_out, nullptr, nullptr, stream);
}

}  // extern "C"
