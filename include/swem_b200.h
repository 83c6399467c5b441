/*
 * swem_b200.h -- C ABI of the B200-native SWEM memory hot path (libswem_b200.so).
 *
 * The reference (lmm077/SWEM) is pure Python/PyTorch and has no FFI; the boundary it offers for
 * this path is the Python class methods/SWEM/modules.py::SWEMCore.  Each entry point below
 * replaces the body of one of its methods and is what a ctypes/pybind stub inside that class
 * binds:
 *
 *   swem_em_forward       <- SWEMCore.swem            methods/SWEM/modules.py:129-168
 *                            (= swe_step :112-120, swm_step :122-127, sww_step :93-110, nu :164-165)
 *   swem_readout_forward  <- SWEMCore.get_affinity    methods/SWEM/modules.py:232-276
 *                            + perm_inv_feat :198-208 + the l2norms of matching :282-283
 *                            + the concat placement of matching :291
 *   swem_em_masks         <- mask prep of SWEM.memorize  methods/SWEM/swem.py:80-84
 *   swem_upsample_add     <- UpsampleBlock.forward      methods/basic_modules/networks.py:192-196
 *   swem_bias_add_act     <- residual tail of ResBlock.forward  networks.py:25-32
 *   swem_decode_tail      <- final up-sampling of Decoder.forward (methods/basic_modules/networks.py:214-215)
 *                            + SWEM.decode / aggregate      methods/SWEM/swem.py:92-116
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to contiguous fp32 (row-major, last index fastest) unless
 *     stated; the caller (PyTorch) owns all memory, the library never allocates device memory;
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous on it and never
 *     synchronise the device;
 *   - return value 0 = ok, non-zero = SwemStatus; swem_last_error() gives a thread-local message;
 *   - unsupported shapes fail loudly (SWEM_ERR_UNSUPPORTED); there is no CPU fallback.
 *
 * Index names: b batch, n object, s side (0 = background, 1 = foreground), c key channel (Ck),
 * d value channel (Cv), p pixel (HW = H/16 * W/16), l basis (L per side per bank).
 */
#ifndef SWEM_B200_H_
#define SWEM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SWEM_B200_ABI_VERSION 3

typedef enum SwemStatus {
  SWEM_OK = 0,
  SWEM_ERR_INVALID_ARG = 1,   /* null pointer, non-positive dim, tau <= 0 ...            */
  SWEM_ERR_UNSUPPORTED = 2,   /* shape outside what the kernels implement                */
  SWEM_ERR_WORKSPACE = 3,     /* workspace missing or smaller than swem_*_workspace_bytes */
  SWEM_ERR_CUDA = 4,          /* a CUDA runtime call or launch failed                    */
  SWEM_ERR_DEVICE = 5         /* not running on an sm_100 device                         */
} SwemStatus;

/* Which kernel family runs.  AUTO picks the fused tcgen05 kernels when the shape is one they
 * cover and the generic tiled kernels otherwise; the other two values force one family (the
 * tests use them to check both against the oracle). */
typedef enum SwemPath {
  SWEM_PATH_AUTO = 0,
  SWEM_PATH_GENERIC = 1,
  SWEM_PATH_FUSED = 2
} SwemPath;

typedef struct SwemDims {
  int32_t B;        /* batch                                   */
  int32_t N;        /* objects                                 */
  int32_t Ck;       /* key channels                            */
  int32_t Cv;       /* value channels                          */
  int32_t HW;       /* pixels at stride 16                     */
  int32_t L;        /* bases per side per bank (n_bases)       */
  int32_t n_iters;  /* EM iterations (memorize only)           */
  int32_t n_banks;  /* 1 or 2 banks read (readout only)        */
  int32_t topl;     /* ranks of the permutation-invariant feature (readout only) */
  float   tau;      /* softmax temperature                     */
} SwemDims;

/* ---- sequential weighted EM update: reference SWEMCore.swem, modules.py:129-168 ------------- */
typedef struct SwemEmArgs {
  SwemDims dims;
  const float* x;            /* [B, Ck, HW]        raw key features (NOT normalised, :116)       */
  const float* v;            /* [B, N, Cv, HW]     value features                                 */
  const float* masks;        /* [B, N, 2, HW]      [bg, fg] pixel weights                         */
  const float* kappa_prior;  /* [B, N, 2, Ck, L]   prior key bases (random_init rows for new objects, :136-146) */
  const float* nu_prior;     /* [B, N, 2, Cv, L]                                                  */
  const float* zita_prior;   /* [B, N, 2, L]                                                      */
  float* kappa;              /* out [B, N, 2, Ck, L]                                              */
  float* nu;                 /* out [B, N, 2, Cv, L]                                              */
  float* zita;               /* out [B, N, 2, L]                                                  */
  float* z_last;             /* optional out [B, N, 2, HW, L]: responsibilities of the last E-step (NULL = skip) */
  void*  workspace;          /* >= swem_em_workspace_bytes(&dims, path) bytes, 256-byte aligned  */
  size_t workspace_bytes;
  int32_t path;              /* SwemPath                                                          */
  int32_t v_pixel_major;     /* 1: `v` is [B, N, HW, Cv] (channels-last, as a cuDNN NHWC value encoder leaves it) -- fused
                                family only (the generic family returns SWEM_ERR_UNSUPPORTED); 0: [B, N, Cv, HW] */
  /* Optional (ABI v3): also leave the readout's tensor-core operand images of the bases this call produces in the workspace of
   * the swem_readout_forward calls that will read them (same B, N, Ck, Cv, L; a memory of image_n_banks banks, these bases being
   * bank image_bank), so that those calls can set bit image_bank of bank_images_valid and skip the conversion.  The EM kernel
   * of the BASELINE shapes writes them from its finalize / nu slice; for every other shape the call appends the conversion
   * launch itself.  image_workspace = NULL: nothing.                                                                      */
  void*  image_workspace;
  int32_t image_bank;
  int32_t image_n_banks;
} SwemEmArgs;

size_t swem_em_workspace_bytes(const SwemDims* dims, int32_t path);
int    swem_em_forward(const SwemEmArgs* args, void* stream);

/* ---- backward of the EM update (training; BASELINE configs[4]).  In the reference only the last line of swem,
 * nu = (zita_ * nu_ + v z) / zita (modules.py:164-165), is differentiable -- E / M / W steps run under
 * @torch.no_grad() (:93,:112,:122) -- so with z = z_last saved by swem_em_forward:
 *     grad_v[b,n,d,p]        = sum_{s,l} grad_nu[b,n,s,d,l] / zita[b,n,s,l] * z[b,n,s,p,l]
 *     grad_nu_prior[b,n,s,d,l] = grad_nu[b,n,s,d,l] * zita_prior[b,n,s,l] / zita[b,n,s,l]
 * Uses dims B, N, Cv, HW, L (Ck / n_iters / tau are ignored).                                         */
typedef struct SwemEmBwdArgs {
  SwemDims dims;
  const float* z_last;        /* [B, N, 2, HW, L]  saved by the forward                              */
  const float* zita_prior;    /* [B, N, 2, L]                                                        */
  const float* zita;          /* [B, N, 2, L]      output of the forward                             */
  const float* grad_nu;       /* [B, N, 2, Cv, L]  incoming gradient                                 */
  float* grad_v;              /* out [B, N, Cv, HW]           (NULL = skip)                          */
  float* grad_nu_prior;       /* out [B, N, 2, Cv, L]         (NULL = skip)                          */
  void*  workspace;           /* >= swem_em_backward_workspace_bytes(&dims)                          */
  size_t workspace_bytes;
} SwemEmBwdArgs;

size_t swem_em_backward_workspace_bytes(const SwemDims* dims);
int    swem_em_backward(const SwemEmBwdArgs* args, void* stream);

/* ---- readout: reference SWEMCore.matching -> get_affinity -> perm_inv_feat, modules.py:198-293 */
typedef struct SwemReadArgs {
  SwemDims dims;
  const float* qk;           /* [B, Ck, HW]  raw query key (the l2norm of :282 is fused)          */
  const float* kappa[2];     /* per bank [B, N, 2, Ck, L], order [first, update] (:295-306); raw, the l2norm of :283 is fused */
  const float* nu[2];        /* per bank [B, N, 2, Cv, L]                                         */
  float* out;                /* [B*N, out_channels, HW] caller-allocated concat buffer (:291)     */
  int32_t out_channels;      /* channel count of `out` (reference: 2*Cv + 2*topl)                 */
  int32_t mem_channel;       /* first channel of mem_out (Cv channels; reference: 0)              */
  int32_t s_channel;         /* first channel of S (2*topl channels; reference: 2*Cv)             */
  void*  workspace;          /* >= swem_readout_workspace_bytes(&dims, path)                      */
  size_t workspace_bytes;
  int32_t path;              /* SwemPath                                                          */
  int32_t out_pixel_major;   /* 0: out is [B*N, out_channels, HW] (reference, NCHW); 1: [B*N, HW, out_channels] (NHWC,
                                what a channels-last fusion conv consumes without a layout copy)          */
  int32_t bank_images_valid; /* bit k set: the tensor-core operand images of bank k (fp16 hi/lo of l2norm(kappa) and of nu) that an
                                earlier swem_readout_forward call built in THIS workspace, for the same dims and the same,
                                unmodified kappa[k] / nu[k], are still intact -- their conversion is skipped.  The reference
                                never changes its 'first' bank while a sequence runs (modules.py:44-60), so a caller that keeps a
                                workspace per sequence sets bit 0 from the second readout on.  0 = convert every bank (always
                                safe); ignored by the generic family.                                                  */
  /* Kernelised memory (the reference's `gen_kernels` branch of get_affinity, modules.py:210-230,252-256; off by default there and
   * here, inference only): mkm_kernels = n_kernel > 0 weights the attention of basis j at pixel p by
   * exp(-min_k d^2(p, p_jk) / (2 sigma^2 tau)), p_jk the n_kernel pixels that basis j matches best, and normalises with + 1e-8; S is
   * unchanged.  mkm_width = W of the H x W pixel grid (HW = H W).  SWEM_PATH_GENERIC only (anything else: SWEM_ERR_UNSUPPORTED).  */
  int32_t mkm_kernels;       /* 0 (off) .. 16                                                      */
  float   mkm_sigma;
  int32_t mkm_width;
  /* Memory dropout (the reference's training-only branch modules.py:258-263; p_drop is hard-wired to 0 there): drop_mask
   * [B, N, Lt] of 0 / 1 (Lt = n_banks L; one mask for both sides) multiplies the exp-affinities of the attention, normalised with
   * + 1e-6; S is unchanged.  NULL = off.  SWEM_PATH_GENERIC only; pass the same mask to swem_readout_backward.                    */
  const float* drop_mask;
} SwemReadArgs;

size_t swem_readout_workspace_bytes(const SwemDims* dims, int32_t path);
int    swem_readout_forward(const SwemReadArgs* args, void* stream);

/* ---- backward of the readout (training).  Reference: autograd through modules.py:232-293 -- gradient to the raw query
 * key (through l2norm :282, affinity, exp, both the attention P and the sorted-prefix feature S :198-208) and to the
 * memory values nu of every bank; the memory keys carry no gradient.  The forward is recomputed (nothing is saved).
 * grad_out has the layout of SwemReadArgs.out ([B*N, out_channels, HW]; only the mem_out and S channels are read).   */
typedef struct SwemReadBwdArgs {
  SwemDims dims;
  const float* qk;           /* [B, Ck, HW]  raw query key of the forward                                   */
  const float* kappa[2];     /* per bank [B, N, 2, Ck, L]                                                    */
  const float* nu[2];        /* per bank [B, N, 2, Cv, L]                                                    */
  const float* grad_out;     /* [B*N, out_channels, HW]                                                      */
  int32_t out_channels, mem_channel, s_channel;
  float* grad_qk;            /* out [B, Ck, HW]            (NULL = skip)                                     */
  float* grad_nu[2];         /* out per bank [B, N, 2, Cv, L] (NULL = skip)                                  */
  void*  workspace;          /* >= swem_readout_backward_workspace_bytes(&dims)                              */
  size_t workspace_bytes;
  const float* drop_mask;    /* the forward's SwemReadArgs.drop_mask (NULL = none)                           */
} SwemReadBwdArgs;

size_t swem_readout_backward_workspace_bytes(const SwemDims* dims);
int    swem_readout_backward(const SwemReadBwdArgs* args, void* stream);

/* ---- mask prep of SWEM.memorize, swem.py:80-84 ------------------------------------------------
 * hard: [B, N+1, Hm, Wm] int64 one-hot (channel 0 = background, skipped), nearest-resized;
 * soft: [B, N+1, Hs, Ws] fp32 probabilities, bilinear-resized (align_corners = false);
 * out : [B, N, 2, H16, W16] with out[:,:,0] = (1-hard)(1-soft), out[:,:,1] = hard*soft.        */
int swem_em_masks(const int64_t* hard, int32_t Hm, int32_t Wm,
                  const float* soft, int32_t Hs, int32_t Ws,
                  int32_t B, int32_t N, int32_t H16, int32_t W16,
                  float* out, void* stream);

/* ---- decoder tail: Decoder.forward's final F.interpolate (networks.py:214-215) + SWEM.decode / aggregate
 * (swem.py:92-116), one pass.  logits_lr: [B*N, Hl, Wl] output of the decoder's `pred` conv (1/4 resolution);
 * bilinear (align_corners = false) to (H, W), sigmoid, optional valid_obj [B, N+1] mask (:100-101),
 * background = prod(1 - p), clamp to [1e-7, 1-1e-7], logit, softmax over the N+1 classes.
 * logits_out, prob_out: [B, N+1, H, W].  N <= 16.                                                    */
int swem_decode_tail(const float* logits_lr, int32_t B, int32_t N, int32_t Hl, int32_t Wl, int32_t H, int32_t W,
                     const float* valid_obj, float* logits_out, float* prob_out, void* stream);
/* Same, plus the evaluator's next two ops on the probabilities just written (swem_evaluator.py:83-87):
 * pred_out [B, 1, H, W] int64 = argmax over the N+1 classes (first maximum, like torch.argmax), hard_out [B, N+1, H, W] int64 =
 * its one-hot -- what SWEM.memorize / swem_em_masks take as the hard masks.  Either may be NULL.                              */
int swem_decode_tail_masks(const float* logits_lr, int32_t B, int32_t N, int32_t Hl, int32_t Wl, int32_t H, int32_t W,
                           const float* valid_obj, float* logits_out, float* prob_out, int64_t* pred_out, int64_t* hard_out,
                           void* stream);

/* ---- decoder glue, channels-last (NHWC) fp32: UpsampleBlock.forward (networks.py:192-196) and the residual tail of
 * ResBlock.forward (:25-32) as single passes.
 *   swem_upsample_add : x[bn] = skip[bn / n] + bilinear(lo_a[bn] (+ lo_b[bn])) (+ bias[c]);  x_relu = relu(x) (optional)
 *                       lo_a, lo_b: [BN, h, w, C]; skip: [BN / n, H, W, C]; x, x_relu: [BN, H, W, C]; bilinear with
 *                       align_corners = false (ATen index arithmetic); `bias` carries the per-channel biases of the
 *                       convolutions that produced lo_a / lo_b / skip (interpolation reproduces constants).
 *   swem_bias_add_act : out = act(a (+ b) (+ c_shared) (+ bias[ch])); a, b, out: [images, pixels, C]; c_shared:
 *                       [images / n_share, pixels, C] (one row block per n_share consecutive images); act = relu or identity.
 *   swem_glu_gate     : FeatureFusionLayer.forward (modules.py:13-26) on stacked pre-activations y = [layer_f | layer_a]:
 *                       out[.., ch] = (y_f + s_f + b_f) * sigmoid(y_a + s_a + b_a); y: [images, pixels, 2C]; shared:
 *                       [images / n_share, pixels, 2C]; bias: [2C]; out: [images, pixels, C].
 * C must be a multiple of 4; optional pointers may be NULL.                                                       */
int swem_upsample_add(const float* lo_a, const float* lo_b, const float* bias, const float* skip, int32_t BN, int32_t n,
                      int32_t h, int32_t w, int32_t H, int32_t W, int32_t C, float* x, float* x_relu, void* stream);
int swem_bias_add_act(const float* a, const float* b, const float* c_shared, const float* bias, int32_t images,
                      int32_t n_share, int64_t pixels, int32_t C, int32_t relu, float* out, void* stream);
int swem_glu_gate(const float* y, const float* shared, const float* bias, int32_t images, int32_t n_share, int64_t pixels,
                  int32_t C, float* out, void* stream);
/* ---- GLU feature-fusion layer (FeatureFusionLayer, modules.py:13-26; input = cat[mem_out, qv, S], :291) as one implicit-GEMM
 * 3x3 / padding-1 convolution on the tensor cores with the gate in its epilogue (csrc/fusion_conv.cu):
 *   out[bn, p, c] = (conv_f(x)[c] + shared[bn / n_share, p, c] + bias[c]) * sigmoid(conv_a(x)[c] + shared[.., Cout + c] + bias[Cout + c])
 * x = feats [BN, H, W, Cin] channels-last fp32 (what swem_readout_forward writes with out_pixel_major = 1: Cin = Cv + 2 topl);
 * shared [BN / n_share, H, W, 2 Cout] (the convolution of the object-independent input channels, or NULL), bias [2 Cout] (or
 * NULL), out [BN, H, W, Cout].  fp32-accurate: operands split into fp16 hi + lo, three products, fp32 accumulation.
 *   swem_fusion_weight_bytes / swem_fusion_prepare_weights : once per model.  w = [2 Cout, Cin, 3, 3] (layer_f's weight stacked on
 *       layer_a's, torch layout, only the input channels of `feats`); scale = a power of two with max|w| * scale in [2^8, 2^14]
 *       (keeps the fp16 lo parts normal); the same scale is passed to swem_fusion_conv_glu.
 *   swem_fusion_workspace_bytes / swem_fusion_conv_glu : two launches (operand images of x, convolution); the workspace needs no
 *       initialisation and holds nothing between calls.
 * Cin must be a multiple of 32, Cout a multiple of 128; anything else is SWEM_ERR_INVALID_ARG (no fallback).                  */
size_t swem_fusion_weight_bytes(int32_t Cin, int32_t Cout);
int swem_fusion_prepare_weights(const float* w, int32_t Cin, int32_t Cout, float scale, void* wblob, void* stream);
size_t swem_fusion_workspace_bytes(int32_t BN, int32_t H, int32_t W, int32_t Cin);
int swem_fusion_conv_glu(const float* feats, const void* wblob, float scale, const float* shared, const float* bias, int32_t BN,
                         int32_t n_share, int32_t H, int32_t W, int32_t Cin, int32_t Cout, void* workspace, size_t workspace_bytes,
                         float* out, void* stream);
/* CBAM of the value encoder's fuser (attentions.py:22-85; networks.py:35-52), NHWC x [images, pixels, C]:
 *   swem_cbam_channel_gate : gate[img, c] = sigmoid(mlp(avg_p x) + mlp(max_p x)), mlp = Linear(C, R) -> ReLU -> Linear(R, C)
 *                            (w1 [R, C], b1 [R], w2 [C, R], b2 [C]); stats: scratch [2, images, C]
 *   swem_cbam_spatial_pool : pooled[img, 0, p] = max_c (x gate), pooled[img, 1, p] = mean_c (x gate)   (input of the 7x7 conv)
 *   swem_cbam_apply        : out = x * (1 + gate[img, c] * sigmoid(spatial_logit[img, p]))  = x + CBAM(x)                */
int swem_cbam_channel_gate(const float* x, const float* w1, const float* b1, const float* w2, const float* b2, int32_t images,
                           int64_t pixels, int32_t C, int32_t R, float* stats, float* gate, void* stream);
int swem_cbam_spatial_pool(const float* x, const float* gate, int32_t images, int64_t pixels, int32_t C, float* pooled,
                           void* stream);
int swem_cbam_apply(const float* x, const float* gate, const float* spatial_logit, int32_t images, int64_t pixels, int32_t C,
                    float* out, void* stream);
/* Tail of the decoder: out[bn, y, x] = bp + sum_{dy, dx, c} wp[dy][dx][c] * relu(a + b + bias[c])[bn, y+dy-1, x+dx-1, c]
 * -- the residual add of the last ResBlock (networks.py:25-32), the ReLU and the 3x3 `pred` conv to one logit plane
 * (networks.py:205-213) in one pass.  a, b: [BN, H, W, C] NHWC; wp: [3, 3, C]; out: [BN, H, W]; C % 32 == 0.          */
int swem_resblock_tail_pred(const float* a, const float* b, const float* bias, const float* wp, float bp, int32_t BN,
                            int32_t H, int32_t W, int32_t C, float* out, void* stream);
/* Input of a ResNet stem in space-to-depth form (see stages.cu): frame [B, 3, H, W]; masks [B, N+1, H, W] or NULL;
 * mean3 / std3: HOST pointers to the 3 normalisation constants (networks.py:72-73); planes = 3 (key encoder, N = 1),
 * 4 (+ object mask, single-object value encoder) or 5 (+ mask of the other objects, networks.py:115-117);
 * out [B*N, H/2+3, W/2+3, Cpad] NHWC, Cpad >= 4*planes, zero border and zero pad channels.                         */
int swem_stem_input(const float* frame, const float* masks, const float* mean3, const float* std3, int32_t B, int32_t N,
                    int32_t planes, int32_t H, int32_t W, int32_t Cpad, float* out, void* stream);
/* 3x3 / stride 2 / padding 1 max pooling of the ResNet stems (networks.py:150, mod_resnet.py), NHWC:
 * in [N, H, W, C] -> out [N, (H-1)/2+1, (W-1)/2+1, C].                                                           */
int swem_maxpool3x3s2(const float* in, int32_t N, int32_t H, int32_t W, int32_t C, float* out, void* stream);
/* Operand split for fp32-accurate convolutions on the TF32 tensor cores (FrameEngine(split_tf32=True)): per pixel of an NHWC
 * tensor x [pixels, C] writes hi [pixels, C] = x rounded to TF32 (10 mantissa bits, nearest) and hl [pixels, 2C] = [hi | lo],
 * lo = x - hi.  conv(x, w) = conv(hi, w_hi) + conv([hi | lo], [w_lo ; w_hi]) + O(2^-22): two cuDNN TF32 convolutions, the
 * small cross terms accumulated on their own (added to the main term in fp32 by the fused conv + add epilogue).  C % 4 == 0. */
int swem_tf32_split(const float* x, int64_t pixels, int32_t C, float* hi, float* hl, void* stream);
/* Same split with the cross-term operand in bf16: xl [pixels, 2C] bf16 = [bf16(x) | bf16(x - hi)].  The cross terms are
 * 2^-11 of the result, so 8 mantissa bits on their operands (and on their bf16 sum) still leave it at 2^-20, and their
 * convolution runs at twice the TF32 rate on half the bytes.                                                              */
int swem_tf32_split_bf16(const float* x, int64_t pixels, int32_t C, float* hi, void* xl_bf16, void* stream);
/* out[i] = float(x_bf16[i]) (+ add[i]) (+ add2[i]): widens the bf16 cross-term convolution's output (and folds the block's residual
 * in) for the fp32 addend of the main-term convolution's fused epilogue, or accumulates it onto the main-term output in place
 * (out == add); n a multiple of 8, flat (any memory format, equal strides), add / add2 may be NULL.                            */
int swem_bf16_widen_add(const void* x_bf16, const float* add, const float* add2, int64_t n, float* out, void* stream);

/* ---- misc ------------------------------------------------------------------------------------ */
int         swem_abi_version(void);          /* == SWEM_B200_ABI_VERSION                          */
const char* swem_last_error(void);           /* thread-local, never NULL                          */
int         swem_device_check(int device);   /* SWEM_OK iff `device` is compute capability 10.x   */
/* 1 if the fused tcgen05 kernels cover these dims (what SWEM_PATH_AUTO would pick), else 0.     */
int         swem_em_fused_supported(const SwemDims* dims);
int         swem_readout_fused_supported(const SwemDims* dims);
/* Diagnostics: install (dev != NULL) or remove (NULL) a device buffer of >= 256 int64.  While
 * installed, CTA 0 of the fused EM kernel writes buf[0] = number of phase stamps and buf[1..] =
 * %globaltimer nanoseconds at its phase boundaries; the fused readout kernel does the same at
 * buf[128] / buf[129..] (tools/profile_phases.py prints them).                                   */
int         swem_set_profile_buffer(void* dev, size_t bytes);
/* number of kernel launches the last call on this thread issued (for bench.py's gpu_launches)   */
int         swem_last_launch_count(void);
/* kernel / memset launches issued by every call on this thread since the library was loaded (monotonic)  */
long long   swem_total_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif  /* SWEM_B200_H_ */
