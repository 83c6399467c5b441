"""``SWEM`` -- same module tree, ``forward(mode, ...)`` dispatch and state-dict keys as the reference's
``methods/SWEM/swem.py::SWEM`` (:9-132), with ``swem_core`` being the CUDA-backed
:class:`swem_b200.core.SWEMCore`.  Encoders / decoder / fusion conv are plain torch modules.
"""
from __future__ import annotations


import torch
from torch import nn
from torch.nn import functional as F

from . import _lib
from .core import SWEMCore
from .networks import Decoder, KeyEncoder, KeyProjection, ValueEncoder, ValueEncoderSO


class SWEM(nn.Module):
    #: build the (B,N,2,H16,W16) EM masks with the library's kernel (True) or with torch ops (False)
    fused_mask_prep = True
    #: run up-sampling + sigmoid + aggregation + softmax of `decode` as one kernel (inference on CUDA only)
    fused_decode_tail = True

    def __init__(self, config_model):
        super().__init__()
        keydim, valdim = config_model.KEYDIM, config_model.VALDIM
        self.single_object = config_model.SINGLE_OBJ
        self.key_encoder = KeyEncoder(config_model.BACKBONE)
        f16_dim = self.key_encoder.num_features[0]
        self.value_encoder = ValueEncoderSO(f16_dim) if self.single_object else ValueEncoder(f16_dim)
        self.key_proj = KeyProjection(f16_dim, keydim=keydim)
        self.key_comp = nn.Conv2d(f16_dim, valdim, kernel_size=3, padding=1)
        self.swem_core = SWEMCore(n_bases=config_model.NUM_BASES, valdim=valdim,
                                  n_iters=config_model.NUM_EM_ITERS, tau=config_model.EM_TAU,
                                  topl=config_model.TOPL)
        self.decoder = Decoder([valdim, self.key_encoder.num_features[1], self.key_encoder.num_features[2]], 256)

    # torch-only stages ------------------------------------------------------------------------
    def encode_key(self, frames):
        s16, s8, s4 = self.key_encoder(frames)
        return self.key_proj(s16), self.key_comp(s16), s16, s8, s4

    def encode_value(self, frame, masks, s16):
        """frame (B,3,H,W), masks (B,N+1,H,W) -> (B,N,Cv,H16,W16); one encoder pass per object."""
        n = masks.shape[1] - 1
        others = 1 - masks - masks[:, 0:1]                       # everything that is neither bg nor this object
        fg = masks[:, 1:].flatten(end_dim=1).unsqueeze(1)
        ot = others[:, 1:].flatten(end_dim=1).unsqueeze(1)
        frame = frame.unsqueeze(1).expand(-1, n, -1, -1, -1).flatten(end_dim=1)
        s16 = s16.unsqueeze(1).expand(-1, n, -1, -1, -1).flatten(end_dim=1)
        mv16 = self.value_encoder(frame, s16, fg) if self.single_object else self.value_encoder(frame, s16, fg, ot)
        return mv16.view(-1, n, *mv16.shape[1:])

    # memory stages ----------------------------------------------------------------------------
    def _em_masks(self, masks_hard, masks_soft, h16, w16):
        if self.fused_mask_prep and masks_hard.is_cuda and masks_hard.dtype == torch.int64:
            hard = masks_hard.contiguous()
            soft = masks_soft.float().contiguous()
            b, n1 = hard.shape[:2]
            out = torch.empty(b, n1 - 1, 2, h16, w16, device=hard.device, dtype=torch.float32)
            lib = _lib.load()
            with torch.cuda.device(hard.device):
                rc = lib.swem_em_masks(hard.data_ptr(), hard.shape[2], hard.shape[3],
                                       soft.data_ptr(), soft.shape[2], soft.shape[3],
                                       b, n1 - 1, h16, w16, out.data_ptr(),
                                       torch.cuda.current_stream(hard.device).cuda_stream)
            _lib.check(rc, 'swem_em_masks')
            return out
        hard = F.interpolate(masks_hard[:, 1:].float(), size=(h16, w16), mode='nearest')
        soft = F.interpolate(masks_soft[:, 1:], size=(h16, w16), mode='bilinear')
        return torch.stack([(1 - hard) * (1 - soft), hard * soft], dim=2)

    def memorize(self, qk16, mv16, masks_hard, masks_soft):
        h16, w16 = qk16.shape[-2:]
        self.swem_core.memorize(qk16, mv16, self._em_masks(masks_hard, masks_soft, h16, w16))

    def init_mem(self, qk16, mv16, mask):
        self.swem_core.empty()
        return self.memorize(qk16, mv16, mask, mask.float())

    def match(self, qk16, qv16):
        return self.swem_core.matching(qk16, qv16)

    def decode(self, n, context, s8, s4, valid_obj, out_size):
        s8 = s8.unsqueeze(1).expand(-1, n, -1, -1, -1).flatten(end_dim=1)
        s4 = s4.unsqueeze(1).expand(-1, n, -1, -1, -1).flatten(end_dim=1)
        if self.fused_decode_tail and context.is_cuda and not torch.is_grad_enabled() and n <= 16:
            return self.decode_from_lowres(self.decoder.lowres_logits(context, s8, s4), n, valid_obj, out_size)
        preds = torch.sigmoid(self.decoder(context, s8, s4, out_size))
        return self._aggregate_objects(preds, n, valid_obj)

    def decode_from_lowres(self, lr, n, valid_obj, out_size):
        """(B*n, 1, Hl, Wl) decoder logits at 1/4 resolution -> (logits, prob) (B, n+1, H, W): the final
        up-sampling of ``Decoder.forward`` + sigmoid + aggregation + softmax (swem.py:92-116)."""
        if not (self.fused_decode_tail and lr.is_cuda and not torch.is_grad_enabled() and n <= 16):
            preds = torch.sigmoid(F.interpolate(lr, size=out_size, mode='bilinear', align_corners=False))
            return self._aggregate_objects(preds, n, valid_obj)
        lr = lr.float().contiguous()
        b = lr.shape[0] // n
        h, w = int(out_size[0]), int(out_size[1])
        logits = torch.empty(b, n + 1, h, w, device=lr.device, dtype=torch.float32)
        prob = torch.empty_like(logits)
        pred = torch.empty(b, 1, h, w, device=lr.device, dtype=torch.int64)
        hard = torch.empty(b, n + 1, h, w, device=lr.device, dtype=torch.int64)
        valid = None if valid_obj is None else valid_obj.float().contiguous()
        with torch.cuda.device(lr.device):
            rc = _lib.load().swem_decode_tail_masks(lr.data_ptr(), b, n, lr.shape[-2], lr.shape[-1], h, w,
                                                    None if valid is None else valid.data_ptr(),
                                                    logits.data_ptr(), prob.data_ptr(), pred.data_ptr(), hard.data_ptr(),
                                                    torch.cuda.current_stream(lr.device).cuda_stream)
        _lib.check(rc, 'swem_decode_tail_masks')
        # the evaluator's next two ops on `prob` (argmax over the classes + its one-hot, swem_evaluator.py:83-87) came out of the
        # same kernel; evaluator.hard_masks_from_scores picks them up from the tensor instead of launching three ATen kernels
        prob.swem_hard_masks = (pred, hard)
        return logits, prob

    def _aggregate_objects(self, preds, n, valid_obj):
        preds = preds.view(-1, n, *preds.shape[-2:])
        if valid_obj is not None:
            preds = preds * valid_obj[:, 1:].unsqueeze(2).unsqueeze(2)
        logits = self.aggregate(preds)
        return logits, F.softmax(logits, dim=1)

    @staticmethod
    def aggregate(prob):
        allp = torch.cat([torch.prod(1 - prob, dim=1, keepdim=True), prob], dim=1).clamp(1e-7, 1 - 1e-7)
        return torch.log(allp / (1 - allp))

    _MODES = {'encode_key': 'encode_key', 'encode_value': 'encode_value', 'init': 'init_mem',
              'memorize': 'memorize', 'match': 'match', 'segment': 'decode'}

    def forward(self, mode, *args, **kwargs):
        if mode not in self._MODES:
            raise NotImplementedError
        return getattr(self, self._MODES[mode])(*args, **kwargs)


def make_config(keydim=64, valdim=512, n_bases=128, n_iters=4, tau=0.05, topl=64, single_obj=False,
                backbone='resnet50'):
    """The subset of the reference's ``VOSConfig.MODEL`` (configs/config.py:51-62) that SWEM reads."""
    from types import SimpleNamespace
    return SimpleNamespace(KEYDIM=keydim, VALDIM=valdim, NUM_BASES=n_bases, NUM_EM_ITERS=n_iters, EM_TAU=tau,
                           TOPL=topl, SINGLE_OBJ=single_obj, BACKBONE=backbone)
