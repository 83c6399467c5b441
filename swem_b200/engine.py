"""``FrameEngine`` -- inference form of the torch stages that surround the SWEM memory.

The reference runs its encoders / fusion conv / decoder exactly as they are written for training
(``methods/SWEM/swem.py:45-116``, ``methods/basic_modules/networks.py``): BatchNorm as its own pass,
ReLU / residual adds as separate kernels, and -- because every object is treated as one more batch
sample -- the *object-independent* parts of several convolutions recomputed once per object.  After the
EM + readout path is fused these pieces are ~98 % of a frame.  This
engine evaluates the SAME function from the SAME ``SWEM`` module's parameters, reorganised for
inference (``eval()`` + ``no_grad`` only):

* BatchNorm folded into the preceding conv (running statistics, ``networks.py`` ResNet trunks);
* conv + bias + ReLU and conv + bias + residual + ReLU issued as single cuDNN fused ops;
* ``key_proj`` and ``key_comp`` (``swem.py:45-49``) share one conv over ``f16``;
* linearity splits for inputs that do not depend on the object:
    - value-encoder fuser ``block1`` (``networks.py:94-106``): ``conv(cat[x_obj, f16])`` =
      ``conv_a(x_obj) + conv_b(f16)``, the 1024-channel ``f16`` half computed once per frame; its
      ``conv1`` and ``downsample`` read the same input and are stacked into one conv;
    - GLU fusion (``modules.py:13-26``): ``layer_f`` / ``layer_a`` stacked, the ``qv`` third of the
      concat ``[mem_out | qv | S]`` (``modules.py:291``) convolved once per frame; the readout kernels
      write ``mem_out`` / ``S`` into a 640-channel buffer, the ``qv`` copy disappears;
    - decoder ``skip_conv`` of ``s8`` / ``s4`` (``networks.py:186-203``) computed once per frame instead
      of once per object, and the N-fold ``expand`` copies of ``s8`` / ``s4`` (``swem.py:93-94``) dropped;
* both ResNet stems (7x7 / stride 2, 3-5 input planes) as a 4x4 conv over a space-to-depth input that one kernel
  (``swem_stem_input``) builds from the frame and the masks -- normalisation, mask planes, the N-fold frame copy and
  the ``torch.cat`` disappear, and cuDNN runs a tensor-core implicit GEMM instead of its scalar path for tiny Cin;
  3x3 / stride-2 pooling by ``swem_maxpool3x3s2``;
* decoder glue as single passes (``swem_upsample_add`` / ``swem_bias_add_act`` of the C ABI, NHWC): bias add +
  bilinear up-sampling + skip add + ReLU, and bias + residual add + ReLU -- the reference spends five to six
  full passes over (objects x 256 x H/4 x W/4) tensors on them.

Only summation order changes (fp32 / TF32 rounding); tests compare the engine with the plain modules
and with the CPU reference port.  It is called like the model (``engine('encode_key', frame)`` ...), so the
sequence runners in ``evaluator.py`` take either.  There is no CPU fallback
of the memory kernels here: ``init`` / ``memorize`` / the readout go to ``SWEMCore``.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

import math
import os

import torch
import torch.nn.functional as F

from . import _lib
from .networks import _Basic, _Bottle

ConvP = Tuple[torch.Tensor, Optional[torch.Tensor], int, int]      # weight, bias, stride, padding


def _fold_bn(conv, bn) -> Tuple[torch.Tensor, torch.Tensor]:
    """conv -> BatchNorm(eval) as one affine conv: w' = w * g/sqrt(var+eps), b' = beta + (b - mean) * g/sqrt(var+eps)."""
    scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    w = conv.weight * scale.view(-1, 1, 1, 1)
    b0 = conv.bias if conv.bias is not None else torch.zeros_like(bn.running_mean)
    return w, bn.bias + (b0 - bn.running_mean) * scale


class FrameEngine:
    def __init__(self, model, channels_last: bool = True, fused_conv: bool = True, stage_kernels: bool = True,
                 split_tf32: bool = False, cross_bf16: bool = False, fusion_kernel: Optional[bool] = None):
        self.model = model
        # GLU fusion layer on libswem_b200's own tcgen05 implicit-GEMM kernel (swem_fusion_conv_glu: fp32-accurate, gate in the
        # epilogue) instead of cuDNN main + cross-term convolutions + swem_glu_gate.  Default: whenever fp32 accuracy is asked for
        # (split_tf32); a single cuDNN TF32 convolution is cheaper than three fp16 products when it is not.  SWEM_FUSION_KERNEL=0/1 overrides.
        env = os.environ.get('SWEM_FUSION_KERNEL')
        self.fusion_kernel = (split_tf32 if fusion_kernel is None else fusion_kernel) if env is None else env == '1'
        self.g_fused = None
        self.channels_last = channels_last
        self.fused_conv = fused_conv
        self.stage_kernels = stage_kernels and channels_last     # the decoder glue kernels of libswem_b200 are NHWC
        # fp32-accurate convolutions ON the tensor cores: every conv as three TF32 convs over hi / lo splits (see _conv_split)
        self.split_tf32 = split_tf32
        self.cross_bf16 = cross_bf16                              # (with split_tf32) the cross-term convolution in bf16
        self._wsplit_cache = {}
        self._split_memo = []                                     # the last two split activations: (tensor, version, hi, [hi | lo])
        self._built = False

    # ------------------------------------------------------------------------------------------
    # parameter preparation
    # ------------------------------------------------------------------------------------------
    def _w(self, w: torch.Tensor) -> torch.Tensor:
        w = w.detach().float()
        return w.contiguous(memory_format=torch.channels_last) if self.channels_last else w.contiguous()

    def _cp(self, w, b, stride=1, padding=1) -> ConvP:
        return (self._w(w), None if b is None else b.detach().float().contiguous(), stride, padding)

    def _folded(self, conv, bn) -> ConvP:
        w, b = _fold_bn(conv, bn)
        return self._cp(w, b, conv.stride[0], conv.padding[0])

    def _plain(self, conv) -> ConvP:
        return self._cp(conv.weight, conv.bias, conv.stride[0], conv.padding[0])

    def _s2d_stem(self, conv, bn, owner):
        """7x7 / stride-2 / padding-3 stem as a 4x4 / stride-1 conv over the space-to-depth input built by
        ``swem_stem_input``: W4[co, 4*ci + 2*p + q, a + 2, b + 2] = W7[co, ci, 2*a + p + 3, 2*b + q + 3], a, b in -2..1."""
        if conv.kernel_size != (7, 7) or conv.stride != (2, 2) or conv.padding != (3, 3):
            return None
        w7, bias = _fold_bn(conv, bn)
        co, planes = w7.shape[:2]
        if not 3 <= planes <= 5:
            return None
        cpad = (4 * planes + 7) // 8 * 8
        w4 = torch.zeros(co, cpad, 4, 4, dtype=torch.float32, device=w7.device)
        for ky in range(7):
            a, p_ = divmod(ky - 3, 2)
            for kx in range(7):
                b, q = divmod(kx - 3, 2)
                w4[:, p_ * 2 + q:4 * planes:4, a + 2, b + 2] = w7[:, :, ky, kx]
        return {'w': self._w(w4), 'b': bias.detach().float().contiguous(), 'planes': planes, 'cpad': cpad,
                'mean': (C.c_float * 3)(*owner.mean.flatten().tolist()), 'std': (C.c_float * 3)(*owner.std.flatten().tolist())}

    def _stem_s2d(self, st, frame, masks, n):
        """relu(bn(conv1(cat[normalised frame, mask planes]))) for n objects per frame -> (B*n, 64, H/2, W/2), or None when the
        space-to-depth path does not apply (CPU tensors, odd sizes): the caller then runs the plain stem."""
        if st is None or not self._glue_ok(frame, masks) or frame.shape[-2] % 2 or frame.shape[-1] % 2:
            return None
        bsz, _, h, w = frame.shape
        frame = frame.contiguous()
        masks = None if masks is None else masks.contiguous()
        x2 = torch.empty((bsz * n, st['cpad'], h // 2 + 3, w // 2 + 3), device=frame.device, dtype=torch.float32,
                         memory_format=torch.channels_last)
        with torch.cuda.device(frame.device):
            rc = _lib.load().swem_stem_input(frame.data_ptr(), None if masks is None else masks.data_ptr(), st['mean'], st['std'],
                                             bsz, n, st['planes'], h, w, st['cpad'], x2.data_ptr(),
                                             torch.cuda.current_stream(frame.device).cuda_stream)
        _lib.check(rc, 'swem_stem_input')
        return self._conv(x2, (st['w'], st['b'], 1, 0), relu=True)

    def _stage(self, stage) -> List[dict]:
        blocks = []
        for blk in stage:
            last_conv, last_bn = (blk.conv3, blk.bn3) if isinstance(blk, _Bottle) else (blk.conv2, blk.bn2)
            if not isinstance(blk, _Bottle):
                assert isinstance(blk, _Basic)
            w_last, b_last = _fold_bn(last_conv, last_bn)
            down = None
            if blk.downsample is not None:
                # relu(last(y) + b_last + down(x) + b_down): the shortcut's bias rides on the last conv's fused bias,
                # so the shortcut conv needs no bias pass of its own
                w_down, b_down = _fold_bn(blk.downsample[0], blk.downsample[1])
                b_last = b_last + b_down
                down = self._cp(w_down, None, blk.downsample[0].stride[0], blk.downsample[0].padding[0])
            d = {'kind': 'bottle' if isinstance(blk, _Bottle) else 'basic',
                 'c1': self._folded(blk.conv1, blk.bn1), 'down': down,
                 'last': self._cp(w_last, b_last, last_conv.stride[0], last_conv.padding[0])}
            if isinstance(blk, _Bottle):
                d['c2'] = self._folded(blk.conv2, blk.bn2)
            blocks.append(d)
        return blocks

    @torch.no_grad()
    def refresh(self) -> None:
        """(Re)derive the inference parameters from the model's current weights (call after loading a checkpoint)."""
        m = self.model
        if m.training:
            raise RuntimeError('FrameEngine is inference-only: call model.eval() first')
        self._wsplit_cache.clear()
        self._split_memo = []
        ke, ve, dec = m.key_encoder, m.value_encoder, m.decoder
        self.k_stem = self._folded(ke.conv1, ke.bn1)
        self.k_stem_s2d = self._s2d_stem(ke.conv1, ke.bn1, ke)
        self.k_stages = [self._stage(s) for s in (ke.res2, ke.layer2, ke.layer3)]
        kp, kc = m.key_proj.key_proj, m.key_comp
        self.keydim = kp.out_channels
        self.k_heads = self._cp(torch.cat([kp.weight, kc.weight], 0), torch.cat([kp.bias, kc.bias], 0), 1, 1)

        self.v_stem = self._folded(ve.conv1, ve.bn1)
        self.v_stem_s2d = self._s2d_stem(ve.conv1, ve.bn1, ve)
        self.v_stages = [self._stage(s) for s in (ve.layer1, ve.layer2, ve.layer3)]
        b1, b2 = ve.fuser.block1, ve.fuser.block2
        cx = ve.layer3[-1].conv2.out_channels                      # channels of the per-object half of cat[x, f16]
        self.f_has_down = b1.downsample is not None                 # reference R50: 1280 -> 512, yes; R18 trunk: 512 -> 512, no
        # block1's conv1 (and its downsample) over cat[x_obj, f16], split by linearity: per-object half without bias,
        # f16 half once per frame; the biases go to the glue kernel that adds the halves
        self.f_c1_obj = self._cp(b1.conv1.weight[:, :cx], None, 1, 1)
        self.f_c1_sh = self._cp(b1.conv1.weight[:, cx:], None, 1, 1)
        self.f_c1_bias = b1.conv1.bias.detach().float().contiguous()
        if self.f_has_down:
            self.f_dn_obj = self._cp(b1.downsample.weight[:, :cx], None, 1, 1)
            self.f_dn_sh = self._cp(b1.downsample.weight[:, cx:], None, 1, 1)
            self.f_b1_tail_bias = (b1.conv2.bias + b1.downsample.bias).detach().float().contiguous()
        else:
            self.f_b1_tail_bias = b1.conv2.bias.detach().float().contiguous()
        self.f_b1c2 = self._cp(b1.conv2.weight, None, 1, 1)
        self.f_b2c1 = self._plain(b2.conv1)
        self.f_b2c2 = self._cp(b2.conv2.weight, None, 1, 1)
        self.f_b2_tail_bias = b2.conv2.bias.detach().float().contiguous()

        core = m.swem_core
        fl = core.fusion_layer
        cv, tl = core.valdim, core.topl
        w = torch.cat([fl.layer_f.weight, fl.layer_a.weight], 0)   # [2*512, 2*Cv + 2*topl, 3, 3]
        self.g_obj = self._cp(torch.cat([w[:, :cv], w[:, 2 * cv:]], 1), None, 1, 1)     # [mem_out | S] channels
        self.g_shared = self._cp(w[:, cv:2 * cv], None, 1, 1)                            # qv channels, once per frame
        self.g_bias = torch.cat([fl.layer_f.bias, fl.layer_a.bias], 0).detach().float().contiguous()
        self.g_out = fl.layer_f.out_channels
        self.g_fused = None
        cin = cv + 2 * tl
        if (self.fusion_kernel and self.stage_kernels and w.is_cuda and cin % 32 == 0 and self.g_out % 128 == 0
                and _lib.load().swem_fusion_weight_bytes(cin, self.g_out)):
            # operand images of the per-object input channels' weights, once per refresh (swem_fusion_prepare_weights)
            lib = _lib.load()
            w_obj = torch.cat([w[:, :cv], w[:, 2 * cv:]], 1).detach().float().contiguous()       # [2 Cout, Cin, 3, 3], torch layout
            amax = float(w_obj.abs().max())
            scale = 2.0 ** (11 - math.ceil(math.log2(amax))) if amax > 0 else 1.0               # max |w| * scale in (2^10, 2^11]
            wblob = torch.empty(lib.swem_fusion_weight_bytes(cin, self.g_out), dtype=torch.uint8, device=w.device)
            with torch.cuda.device(w.device):
                rc = lib.swem_fusion_prepare_weights(w_obj.data_ptr(), cin, self.g_out, scale, wblob.data_ptr(),
                                                     torch.cuda.current_stream(w.device).cuda_stream)
            _lib.check(rc, 'swem_fusion_prepare_weights')
            self.g_fused = (wblob, scale, cin)

        if dec.compress.downsample is not None:
            raise RuntimeError('FrameEngine expects Decoder.compress to keep its width (reference: 512 -> 512)')
        self.d_c1, self.d_c2 = self._plain(dec.compress.conv1), self._plain(dec.compress.conv2)
        self.d_up = []
        for up in (dec.up_16_8, dec.up_8_4):
            rb = up.out_conv
            d = {'skip': self._plain(up.skip_conv), 'c1': self._plain(rb.conv1), 'c2': self._plain(rb.conv2),
                 'down': None if rb.downsample is None else self._plain(rb.downsample)}
            self.d_up.append(d)
        self.d_pred = self._plain(dec.pred)
        self.d_pred_w = dec.pred.weight.detach().float()[0].permute(1, 2, 0).contiguous()      # [3, 3, C] for the fused tail
        self.d_pred_b = float(dec.pred.bias.detach().float()[0]) if dec.pred.out_channels == 1 else None
        # per-channel constants handed to the fused glue kernels instead of separate bias passes: the biases of the convs
        # that feed each up-sampling (conv2 [+ downsample] of the previous ResBlock, skip_conv of this level), and of the last ResBlock
        def bias_of(p: ConvP):
            return p[1]
        prev = bias_of(self.d_c2)
        self.d_bias = []
        for d in self.d_up:
            self.d_bias.append((prev + bias_of(d['skip'])).contiguous())
            prev = bias_of(d['c2']) + (bias_of(d['down']) if d['down'] is not None else 0)
        self.d_bias.append(prev.contiguous())
        self._built = True

    def _ready(self):
        if not self._built:
            self.refresh()
        if torch.is_grad_enabled():
            raise RuntimeError('FrameEngine is inference-only: call it under torch.no_grad()')

    # ------------------------------------------------------------------------------------------
    # conv helpers
    # ------------------------------------------------------------------------------------------
    @staticmethod
    def _tf32_split(t: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """t = hi + lo with hi exactly representable in TF32 (10 mantissa bits, round to nearest) and lo the fp32 residual."""
        hi = ((t.view(torch.int32) + 0x1000) & -0x2000).view(torch.float32)
        return hi, t - hi

    def _split(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """x (N,C,H,W) -> hi (N,C,H,W), [hi | lo] (N,2C,H,W)  (``swem_tf32_split`` for NHWC tensors).  A block input feeds
        two convolutions (conv1 and the shortcut; the object-independent halves of the fuser): the last two splits are kept."""
        for ref, ver, hi, hl in self._split_memo:
            if ref is x and ver == x._version:
                return hi, hl
        hi, hl = self._split_uncached(x)
        self._split_memo = self._split_memo[-1:] + [(x, x._version, hi, hl)]
        return hi, hl

    def _split_uncached(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        n, c, h, w = x.shape
        if self._glue_ok(x) and c % 4 == 0 and x.is_contiguous(memory_format=torch.channels_last):
            hi = torch.empty_like(x)
            hl = torch.empty((n, 2 * c, h, w), device=x.device, dtype=torch.bfloat16 if self.cross_bf16 else torch.float32,
                             memory_format=torch.channels_last)
            fn = _lib.load().swem_tf32_split_bf16 if self.cross_bf16 else _lib.load().swem_tf32_split
            with torch.cuda.device(x.device):
                rc = fn(x.data_ptr(), n * h * w, c, hi.data_ptr(), hl.data_ptr(), torch.cuda.current_stream(x.device).cuda_stream)
            _lib.check(rc, 'swem_tf32_split')
            return hi, hl
        hi, lo = self._tf32_split(x)
        hl = torch.cat([x, lo], dim=1).bfloat16() if self.cross_bf16 else torch.cat([hi, lo], dim=1)
        return (self._cl(hi), self._cl(hl)) if self.channels_last else (hi, hl)

    def _conv_split(self, x, p: ConvP, relu: bool, add: Optional[torch.Tensor]):
        """conv(x, w) to fp32 accuracy on the tensor cores: x = xh + xl, w = wh + wl (hi parts exactly representable in
        TF32), conv(x, w) = conv(xh, wh) + [conv(xh, wl) + conv(xl, wh)] + O(2^-22) -- the hi x hi products are exact in the
        fp32 accumulator, the cross terms carry a relative rounding of 2^-11 on a term that is 2^-11 of the result.  Issued
        as TWO cuDNN TF32 convolutions: the cross terms as one conv over [xh | xl] against [wl ; wh], then the main term
        with the cross sum (and the residual, bias, ReLU) in its fused epilogue.  Cross terms accumulate on their own: fed
        through the same accumulator as the main term (one conv over [xh | xh | xl]) every small product is truncated at the
        big sum's ulp -- 3x the error, enough to cost 0.3 % of the masks (tools/split_conv_probe.py).  cuDNN's IEEE-fp32
        path has no tensor cores on sm_100 (100x slower than TF32 with its heuristic algorithms, 18x autotuned)."""
        w, b, s, pad = p
        key = w.data_ptr()
        if key not in self._wsplit_cache:
            wh, wl = self._tf32_split(w)
            zero = torch.zeros(w.shape[0], device=w.device, dtype=torch.float32)
            wc = self._w(torch.cat([wl, wh], dim=1))
            if self.cross_bf16:            # cross terms are 2^-11 of the result: 8 mantissa bits on their operands leave it at 2^-20
                wc = wc.bfloat16()
            self._wsplit_cache[key] = (w, self._w(wh), wc, zero)    # (keeps `w` alive: unique pointer)
        _, wh, wc, zero = self._wsplit_cache[key]
        hi, hl = self._split(x)
        old = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = True
        try:
            cross = F.conv2d(hl, wc, None, stride=s, padding=pad)
            fused_relu = relu and self.fused_conv and x.is_cuda
            glue = (self.cross_bf16 and self._glue_ok(add) and cross.is_cuda and cross.numel() % 8 == 0
                    and (add is None or (add.shape == cross.shape and add.stride() == cross.stride()))
                    and (cross.is_contiguous() or cross.is_contiguous(memory_format=torch.channels_last)))
            if glue:
                # swem_bf16_widen_add: one vectorised pass instead of a converting copy and one or two in-place adds
                lib, st = _lib.load(), torch.cuda.current_stream(cross.device).cuda_stream
                ptr = lambda t: None if t is None else t.data_ptr()
                if fused_relu:           # widened (+ residual) addend for the fused epilogue of the main-term convolution
                    wide = torch.empty_like(cross, dtype=torch.float32)
                    with torch.cuda.device(cross.device):
                        _lib.check(lib.swem_bf16_widen_add(cross.data_ptr(), ptr(add), None, cross.numel(), wide.data_ptr(), st), 'swem_bf16_widen_add')
                    return torch.cudnn_convolution_add_relu(hi, wh, wide, 1.0, zero if b is None else b, (s, s), (pad, pad), (1, 1), 1)
                y = F.conv2d(hi, wh, b, stride=s, padding=pad)
                if y.stride() == cross.stride():     # no fused epilogue without a ReLU: accumulate onto the main term in place
                    with torch.cuda.device(cross.device):
                        _lib.check(lib.swem_bf16_widen_add(cross.data_ptr(), y.data_ptr(), ptr(add), cross.numel(), y.data_ptr(), st), 'swem_bf16_widen_add')
                    return F.relu_(y) if relu else y
                y.add_(cross.float())
                if add is not None:
                    y.add_(add)
                return F.relu_(y) if relu else y
            if self.cross_bf16:
                cross = cross.float()
            if add is not None:
                cross.add_(add)
            if fused_relu:
                return torch.cudnn_convolution_add_relu(hi, wh, cross, 1.0, zero if b is None else b, (s, s), (pad, pad), (1, 1), 1)
            y = F.conv2d(hi, wh, b, stride=s, padding=pad).add_(cross)
            return F.relu_(y) if relu else y
        finally:
            torch.backends.cudnn.allow_tf32 = old

    def _conv(self, x, p: ConvP, relu: bool = False, add: Optional[torch.Tensor] = None):
        if self.split_tf32 and x.is_cuda:
            return self._conv_split(x, p, relu, add)
        return self._conv_plain(x, p, relu, add)

    def _conv_plain(self, x, p: ConvP, relu: bool = False, add: Optional[torch.Tensor] = None):
        w, b, s, pad = p
        if self.fused_conv and relu and x.is_cuda and b is not None:
            if add is None:
                return torch.cudnn_convolution_relu(x, w, b, (s, s), (pad, pad), (1, 1), 1)
            return torch.cudnn_convolution_add_relu(x, w, add, 1.0, b, (s, s), (pad, pad), (1, 1), 1)
        y = F.conv2d(x, w, b, stride=s, padding=pad)
        if add is not None:
            y.add_(add)
        return F.relu_(y) if relu else y

    def _run_stage(self, x, blocks):
        for d in blocks:
            y = self._conv(x, d['c1'], relu=True)
            if d['kind'] == 'bottle':
                y = self._conv(y, d['c2'], relu=True)
            skip = x if d['down'] is None else self._conv(x, d['down'])
            x = self._conv(y, d['last'], relu=True, add=skip)
        return x

    def _maxpool(self, x):
        """3x3 / stride-2 / padding-1 pooling after the stem (NHWC kernel of the library; ATen's runs at ~1 TB/s)."""
        if self._glue_ok(x) and x.shape[1] % 4 == 0:
            x = self._cl(x)
            n, c, h, w = x.shape
            out = torch.empty((n, c, (h - 1) // 2 + 1, (w - 1) // 2 + 1), device=x.device, dtype=torch.float32,
                              memory_format=torch.channels_last)
            with torch.cuda.device(x.device):
                rc = _lib.load().swem_maxpool3x3s2(x.data_ptr(), n, h, w, c, out.data_ptr(),
                                                   torch.cuda.current_stream(x.device).cuda_stream)
            _lib.check(rc, 'swem_maxpool3x3s2')
            return out
        return F.max_pool2d(x, 3, stride=2, padding=1)

    def _trunk(self, y, stages, taps=False):
        """y = the stem's output (conv1 + bn1 + relu) -> max-pool -> the three residual stages."""
        x = self._maxpool(y)
        feats = []
        for st in stages:
            x = self._run_stage(x, st)
            feats.append(x)
        return feats if taps else x

    # ------------------------------------------------------------------------------------------
    # the model's modes (swem.py:118-132)
    # ------------------------------------------------------------------------------------------
    def encode_key(self, frames):
        """frames (B,3,H,W) -> (qk16, qv16, f16, f8, f4), as SWEM.encode_key (swem.py:45-49)."""
        self._ready()
        ke = self.model.key_encoder
        y = self._stem_s2d(self.k_stem_s2d, frames, None, 1)
        if y is None:
            y = self._conv((frames - ke.mean) / ke.std, self.k_stem, relu=True)
        f4, f8, f16 = self._trunk(y, self.k_stages, taps=True)
        heads = self._conv(f16, self.k_heads)
        return heads[:, :self.keydim], heads[:, self.keydim:], f16, f8, f4

    def shared_parts(self, qv16, s16, s8, s4) -> dict:
        """The object-independent convolutions of a frame -- the f16 halves of the fuser's first block, the qv third of the
        GLU fusion conv and the decoder's skip convs: functions of the key encoder's outputs only.  ``match`` /
        ``encode_value`` / ``segment`` compute them on demand; a runner that encodes keys ahead of time (see
        ``PipelinedSequenceRunner``) computes them there too and passes them in as ``shared=``."""
        self._ready()
        parts = {'g': self._conv(qv16, self.g_shared), 'f_c1': self._conv(s16, self.f_c1_sh),
                 'skip': [self._conv(f, self._nobias(d['skip'])) for d, f in zip(self.d_up, (s8, s4))]}
        if self.f_has_down:
            parts['f_dn'] = self._conv(s16, self.f_dn_sh)
        return parts

    def encode_value(self, frame, masks, s16, shared=None):
        """frame (B,3,H,W), masks (B,N+1,H,W), s16 = f16 of the same frame -> (B,N,Cv,H16,W16) (swem.py:51-62)."""
        self._ready()
        m, ve = self.model, self.model.value_encoder
        n = masks.shape[1] - 1
        bsz = frame.shape[0]
        y = self._stem_s2d(self.v_stem_s2d, frame, masks.float(), n)
        if y is None:
            image = ((frame - ve.mean) / ve.std).unsqueeze(1).expand(-1, n, -1, -1, -1)
            planes = [image, masks[:, 1:].unsqueeze(2)]
            if not m.single_object:
                others = 1 - masks - masks[:, 0:1]
                planes.append(others[:, 1:].unsqueeze(2))
            x = torch.cat(planes, dim=2).flatten(end_dim=1)        # (B*N, 3 + extra, H, W)
            if self.channels_last:
                x = x.contiguous(memory_format=torch.channels_last)
            y = self._conv(x, self.v_stem, relu=True)
        x = self._trunk(y, self.v_stages)                          # (B*N, 256, H16, W16), post-ReLU
        # fuser.block1 on cat[x, f16]: both halves are post-ReLU, so block1's leading ReLU is the identity
        f_c1 = shared['f_c1'] if shared is not None else self._conv(s16, self.f_c1_sh)
        h1 = self._add_act(self._conv(x, self.f_c1_obj), None, f_c1, self.f_c1_bias, n, relu=True)
        r = self._conv(h1, self.f_b1c2)
        if self.f_has_down:
            f_dn = shared['f_dn'] if shared is not None else self._conv(s16, self.f_dn_sh)
            x = self._add_act(r, self._conv(x, self.f_dn_obj), f_dn, self.f_b1_tail_bias, n, relu=False)
        else:
            skip = torch.cat([x.view(bsz, n, *x.shape[1:]), s16.unsqueeze(1).expand(-1, n, -1, -1, -1)], 2).flatten(end_dim=1)
            x = self._add_act(r, skip, None, self.f_b1_tail_bias, n, relu=False)
        x = self._cbam_residual(x, ve.fuser.attention)
        r = self._conv(self._conv(F.relu(x), self.f_b2c1, relu=True), self.f_b2c2)
        x = self._add_act(r, x, None, self.f_b2_tail_bias, n, relu=False)
        return x.view(bsz, n, *x.shape[1:])

    def match(self, qk16, qv16, shared=None):
        """Readout + GLU fusion: (context (B*N, Cv, H, W), N), as SWEMCore.matching (modules.py:278-293)."""
        self._ready()
        core = self.model.swem_core
        n, cv = core._readout_objects()
        bsz, _, h, w = qk16.shape
        feats = torch.empty((bsz * n, cv + 2 * core.topl, h, w), device=qk16.device, dtype=torch.float32,
                            memory_format=torch.channels_last if self.channels_last else torch.contiguous_format)
        core.readout_into(qk16, feats, 0, cv)                      # [mem_out | S], written NHWC for the channels-last conv
        g = shared['g'] if shared is not None else self._conv(qv16, self.g_shared)   # (B, 1024, H, W): [layer_f | layer_a] of the qv third
        if (self.g_fused is not None and feats.is_cuda and feats.shape[1] == self.g_fused[2]
                and feats.is_contiguous(memory_format=torch.channels_last)):
            return self._fusion_conv_glu(feats, g, bsz * n, n, h, w), n
        return self._glu(self._conv(feats, self.g_obj), g, self.g_bias, n), n

    def _fusion_conv_glu(self, feats, g, bn, n, h, w):
        """FeatureFusionLayer on the per-object channels + shared term + biases + gate in two launches (swem_fusion_conv_glu)."""
        lib = _lib.load()
        wblob, scale, cin = self.g_fused
        g = self._cl(g)
        need = lib.swem_fusion_workspace_bytes(bn, h, w, cin)
        ws = self.model.swem_core._workspace.get(feats.device, need, 'fusion')   # (the core's scratch: pinned while a graph references it)
        out = torch.empty((bn, self.g_out, h, w), device=feats.device, dtype=torch.float32, memory_format=torch.channels_last)
        with torch.cuda.device(feats.device):
            rc = lib.swem_fusion_conv_glu(feats.data_ptr(), wblob.data_ptr(), scale, g.data_ptr(), self.g_bias.data_ptr(), bn, n, h, w,
                                          cin, self.g_out, ws.data_ptr(), ws.numel(), out.data_ptr(),
                                          torch.cuda.current_stream(feats.device).cuda_stream)
        _lib.check(rc, 'swem_fusion_conv_glu')
        return out

    @staticmethod
    def _nobias(p: ConvP) -> ConvP:
        return (p[0], None, p[2], p[3])

    def _glue_ok(self, *ts) -> bool:
        return self.stage_kernels and all(t is None or (t.is_cuda and t.dtype == torch.float32) for t in ts)

    def _cl(self, t):
        return t if t.is_contiguous(memory_format=torch.channels_last) else t.contiguous(memory_format=torch.channels_last)

    def _upsample_add(self, lo_a, lo_b, bias, skip, n):
        """x = skip[b] + bilinear(lo_a + lo_b) + bias  and relu(x)  (UpsampleBlock.forward, networks.py:192-196)."""
        if self._glue_ok(lo_a, lo_b, skip) and lo_a.shape[1] % 4 == 0:
            lo_a, skip = self._cl(lo_a), self._cl(skip)
            lo_b = None if lo_b is None else self._cl(lo_b)
            bn, c, h, w = lo_a.shape
            H, W = skip.shape[-2:]
            x = torch.empty((bn, c, H, W), device=lo_a.device, dtype=torch.float32, memory_format=torch.channels_last)
            xr = torch.empty_like(x)
            with torch.cuda.device(lo_a.device):
                rc = _lib.load().swem_upsample_add(lo_a.data_ptr(), None if lo_b is None else lo_b.data_ptr(), bias.data_ptr(),
                                                   skip.data_ptr(), bn, n, h, w, H, W, c, x.data_ptr(), xr.data_ptr(),
                                                   torch.cuda.current_stream(lo_a.device).cuda_stream)
            _lib.check(rc, 'swem_upsample_add')
            return x, xr
        lo = lo_a if lo_b is None else lo_a + lo_b
        up = F.interpolate(lo, size=skip.shape[-2:], mode='bilinear', align_corners=False)
        x = up.view(-1, n, *up.shape[1:]).add_(skip.unsqueeze(1)).flatten(end_dim=1).add_(bias.view(1, -1, 1, 1))
        return x, F.relu(x)

    def _add_act(self, a, b, shared, bias, n, relu):
        """act(a + b + shared[image // n] + bias): bias pass, object-independent half, residual add and ReLU in one kernel."""
        if self._glue_ok(a, b, shared) and a.shape[1] % 4 == 0:
            a = self._cl(a)
            b = None if b is None else self._cl(b)
            shared = None if shared is None else self._cl(shared)
            out = torch.empty_like(a)
            ptr = lambda t: None if t is None else t.data_ptr()
            with torch.cuda.device(a.device):
                rc = _lib.load().swem_bias_add_act(a.data_ptr(), ptr(b), ptr(shared), ptr(bias), a.shape[0], n if shared is not None else 1,
                                                   a.shape[2] * a.shape[3], a.shape[1], int(relu), out.data_ptr(),
                                                   torch.cuda.current_stream(a.device).cuda_stream)
            _lib.check(rc, 'swem_bias_add_act')
            return out
        out = a if b is None else a + b
        if shared is not None:
            out = (out.view(-1, n, *out.shape[1:]) + shared.unsqueeze(1)).flatten(end_dim=1)
        if bias is not None:
            out = out + bias.view(1, -1, 1, 1)
        return F.relu(out) if relu else out

    def _cbam_residual(self, x, cbam):
        """x + CBAM(x) (networks.py:49; attentions.py:22-85): channel statistics + MLP, channel pooling per pixel and the
        final gate as one kernel each; only the 2 -> 1 channel 7x7 conv of the spatial gate stays cuDNN."""
        mlp = cbam.ChannelGate.mlp
        lin1, lin2 = mlp[1], mlp[3]
        c, r = x.shape[1], lin1.out_features
        if not (self._glue_ok(x) and c % 32 == 0 and r <= 64 and x.shape[0] <= 65535):
            return x + cbam(x)
        x = self._cl(x)
        bn, _, h, w = x.shape
        dev = x.device
        stats = torch.empty(2, bn, c, device=dev, dtype=torch.float32)
        gate = torch.empty(bn, c, device=dev, dtype=torch.float32)
        pooled = torch.empty(bn, 2, h, w, device=dev, dtype=torch.float32)
        out = torch.empty_like(x)
        lib = _lib.load()
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(lib.swem_cbam_channel_gate(x.data_ptr(), lin1.weight.data_ptr(), lin1.bias.data_ptr(), lin2.weight.data_ptr(),
                                                  lin2.bias.data_ptr(), bn, h * w, c, r, stats.data_ptr(), gate.data_ptr(), stream),
                       'swem_cbam_channel_gate')
            _lib.check(lib.swem_cbam_spatial_pool(x.data_ptr(), gate.data_ptr(), bn, h * w, c, pooled.data_ptr(), stream),
                       'swem_cbam_spatial_pool')
            logit = cbam.SpatialGate.spatial(pooled).contiguous()                  # (bn, 1, h, w)
            _lib.check(lib.swem_cbam_apply(x.data_ptr(), gate.data_ptr(), logit.data_ptr(), bn, h * w, c, out.data_ptr(), stream),
                       'swem_cbam_apply')
        return out

    def _glu(self, y, shared, bias, n):
        """(y_f + s_f + b_f) * sigmoid(y_a + s_a + b_a) on stacked [layer_f | layer_a] pre-activations (modules.py:13-26)."""
        c = y.shape[1] // 2
        if self._glue_ok(y, shared) and c % 4 == 0:
            y, shared = self._cl(y), self._cl(shared)
            out = torch.empty((y.shape[0], c, y.shape[2], y.shape[3]), device=y.device, dtype=torch.float32,
                              memory_format=torch.channels_last)
            with torch.cuda.device(y.device):
                rc = _lib.load().swem_glu_gate(y.data_ptr(), shared.data_ptr(), bias.data_ptr(), y.shape[0], n, y.shape[2] * y.shape[3], c,
                                               out.data_ptr(), torch.cuda.current_stream(y.device).cuda_stream)
            _lib.check(rc, 'swem_glu_gate')
            return out
        t = (y.view(-1, n, *y.shape[1:]) + shared.unsqueeze(1)).flatten(end_dim=1) + bias.view(1, -1, 1, 1)
        return t[:, :c] * torch.sigmoid(t[:, c:])

    def decode(self, n, context, s8, s4, valid_obj, out_size, shared=None):
        """-> (logits, prob) (B, N+1, H, W), as SWEM.decode (swem.py:92-108).  Convolutions whose output only feeds an
        up-sampling or a residual add run without their bias pass; the constants go to the glue kernels (`d_bias`)."""
        self._ready()
        lo_a = self._conv(self._conv(F.relu(context), self.d_c1, relu=True), self._nobias(self.d_c2))
        lo_b = context
        for lvl, (d, skip_f, bias) in enumerate(zip(self.d_up, (s8, s4), self.d_bias)):
            skip = shared['skip'][lvl] if shared is not None else self._conv(skip_f, self._nobias(d['skip']))   # once per frame
            x, xr = self._upsample_add(lo_a, lo_b, bias, skip, n)
            lo_a = self._conv(self._conv(xr, d['c1'], relu=True), self._nobias(d['c2']))
            lo_b = x if d['down'] is None else self._conv(x, self._nobias(d['down']))
        lr = self._tail_pred(lo_a, lo_b, self.d_bias[-1], n)                                     # (B*n, 1, Hl, Wl)
        return self.model.decode_from_lowres(lr, n, valid_obj, out_size)

    def _tail_pred(self, a, b, bias, n):
        """pred(relu(a + b + bias)): residual tail of the last ResBlock + ReLU + the 3x3 conv to one logit plane."""
        if (self._glue_ok(a, b) and a.shape[1] % 32 == 0 and self.d_pred_b is not None and a.shape[0] <= 65535
                and self.d_pred[2] == 1 and self.d_pred[3] == 1):
            a, b = self._cl(a), self._cl(b)
            bn, c, h, w = a.shape
            out = torch.empty((bn, 1, h, w), device=a.device, dtype=torch.float32)
            with torch.cuda.device(a.device):
                rc = _lib.load().swem_resblock_tail_pred(a.data_ptr(), b.data_ptr(), bias.data_ptr(), self.d_pred_w.data_ptr(),
                                                         self.d_pred_b, bn, h, w, c, out.data_ptr(),
                                                         torch.cuda.current_stream(a.device).cuda_stream)
            _lib.check(rc, 'swem_resblock_tail_pred')
            return out
        return self._conv(self._add_act(a, b, None, bias, n, relu=True), self.d_pred)

    _MODES = {'encode_key': 'encode_key', 'encode_value': 'encode_value', 'match': 'match', 'segment': 'decode'}

    def __call__(self, mode, *args, **kwargs):
        if mode in self._MODES:
            return getattr(self, self._MODES[mode])(*args, **kwargs)
        return self.model(mode, *args, **kwargs)                   # 'init' / 'memorize': the memory itself

    @property
    def swem_core(self):
        return self.model.swem_core
