"""``SWEMCore`` -- the drop-in replacement of the reference's ``methods/SWEM/modules.py::SWEMCore``.

Same constructor, attributes and methods (``empty`` / ``memorize`` / ``matching`` / ``get_mem`` /
``swem`` / ``random_init``, ``memories['first'|'update'].bases`` dicts with keys
``kappa (B,N,2,Ck,L)``, ``nu (B,N,2,Cv,L)``, ``zita (B,N,2,1,L)``, ``fusion_layer.layer_f/layer_a``
state-dict keys), but the E / M / W steps, the nu update and the readout attention run in the
CUDA kernels of ``libswem_b200.so`` through its C ABI (``include/swem_b200.h``).

Training: ``swem_b200/autograd.py`` (nu is differentiable in v and the prior nu; the readout in qk and nu).

What stays torch, as in the reference: the RNG draw for new bases (``random_init`` uses the
global generator of the key's device, modules.py:170-178, so seeding behaves identically), the
bank bookkeeping (modules.py:29-60,183-193) and the GLU fusion conv (modules.py:13-26).

There is no CPU path: tensors must live on an sm_100 device and the library must be built.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Optional, Tuple

import torch
from torch import nn

from . import _lib
from .networks import FeatureFusionLayer

Bases = Dict[str, torch.Tensor]


class MemoryBank:
    """One bank of bases; 'fixed' only appends objects it has not seen, 'updated' is replaced."""

    def __init__(self, mode: str = 'updated'):
        assert mode in ('fixed', 'updated')
        self.mode = mode
        self.bases: Optional[Bases] = None
        self.n_objs = 0

    def initial_memory(self):
        self.bases, self.n_objs = None, 0

    def add_new(self, bases: Bases):
        if self.bases is None:
            self.bases, self.n_objs = bases, bases['kappa'].shape[1]
            return
        n = bases['kappa'].shape[1]
        if n > self.n_objs:
            self.bases = {k: torch.cat([self.bases[k], bases[k][:, self.n_objs:]], dim=1) for k in bases}
        self.n_objs = n

    def update(self, bases: Bases):
        if self.mode == 'fixed':
            self.add_new(bases)
        else:
            self.bases = bases


class _Workspace:
    """Scratch buffers handed to the kernels, owned by ONE ``SWEMCore`` and keyed by (purpose, device, stream): two cores,
    two streams or a forward and a backward never share (or regrow) each other's scratch.

    A buffer whose address has been baked into a CUDA graph must outlive that graph: while the owner is ``pinned`` (a
    runner is capturing / replaying, ``SWEMCore.static_banks``) a request that does not fit raises instead of
    reallocating, and a buffer that was ever handed out while pinned is retired (kept alive), never freed, when a later
    un-pinned call needs a bigger one."""

    def __init__(self):
        self._buf: Dict[tuple, torch.Tensor] = {}
        self._seen_pinned: Dict[tuple, bool] = {}
        self._retired = []
        self.pinned = False

    def get(self, device: torch.device, nbytes: int, purpose: str = 'fwd') -> torch.Tensor:
        key = (purpose, device, torch.cuda.current_stream(device).cuda_stream)
        buf = self._buf.get(key)
        if buf is None or buf.numel() < nbytes:
            if buf is not None and self.pinned:
                raise RuntimeError(f'swem_b200: the kernels need a {nbytes}-byte workspace but the {buf.numel()}-byte one is '
                                   'referenced by a captured CUDA graph (static_banks); re-capture for the larger shape')
            if buf is not None and self._seen_pinned.get(key):
                self._retired.append(buf)
            buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
            self._buf[key] = buf
            self._seen_pinned[key] = False
        if self.pinned:
            self._seen_pinned[key] = True
        return buf


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f'swem_b200: `{name}` is on {t.device}; the SWEM hot path runs on a B200 only '
                           '(no CPU fallback -- use oracle/ for CPU checking)')
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _wants_grad(*tensors) -> bool:
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


def _invoke(name, call):
    """Every C-ABI launch goes through here; profilers (bench.py) replace it to bracket the call with
    CUDA events on the launching stream."""
    return call()


class SWEMCore(nn.Module):
    # SWEM_PATH_* of the C ABI per entry point; tests flip these to exercise both kernel families
    em_path = _lib.PATH_AUTO
    readout_path = _lib.PATH_AUTO
    _static_banks = False

    @property
    def static_banks(self) -> bool:
        """True: the 'update' bank keeps its tensors and is overwritten in place, and the kernels' workspace may not be
        reallocated (needed when the frame step is replayed from a CUDA graph, where every address is baked in).
        False: bases are replaced wholesale like the reference."""
        return self._static_banks

    @static_banks.setter
    def static_banks(self, value: bool) -> None:
        self._static_banks = bool(value)
        self._workspace.pinned = bool(value)

    def __init__(self, n_bases=256, valdim=512, n_iters=4, tau=0.05, topl=64):
        super().__init__()
        assert tau > 0
        self.n_bases, self.n_iters, self.tau, self.valdim = n_bases, n_iters, tau, valdim
        self.memories = {'first': MemoryBank('fixed'), 'update': MemoryBank('updated')}
        self.p_drop = 0.0
        self.topl = int(min(n_bases, topl))
        self.fusion_layer = FeatureFusionLayer(valdim * 2 + self.topl * 2, valdim)
        self.launches = 0          # kernels launched by this object's last memorize/matching call
        self._workspace = _Workspace()
        self._image_key = None     # what the bank images in the readout workspace were built from (see _readout_launch)
        self._image1_src = None    # the 'update' bank tensors whose images the last memorize emitted (see memorize / _em_launch)
        self._emit_images = False
        self._emitted = None

    # -- bank state ------------------------------------------------------------------------
    def empty(self):
        for bank in self.memories.values():
            bank.initial_memory()
        self._image_key = None
        self._image1_src = None

    def get_mem(self) -> Tuple[torch.Tensor, torch.Tensor]:
        banks = [m.bases for m in self.memories.values() if m.bases is not None]
        return (torch.cat([b['kappa'] for b in banks], dim=-1), torch.cat([b['nu'] for b in banks], dim=-1))

    def random_init(self, size, norm_dim=-2, dtype=None, device=None):
        """Fresh bases for objects without a prior; draws from torch's global RNG like the reference."""
        B, N, _, _, L = size
        kappa = torch.zeros(size=size).type(dtype).to(device)
        kappa.normal_(0, math.sqrt(2.0 / size[-1]))
        kappa = kappa / (torch.linalg.norm(kappa, dim=norm_dim, keepdim=True) + 1e-6)
        nu = torch.zeros(B, N, 2, self.valdim, L).type(dtype).to(device)
        zita = torch.zeros(B, N, 2, 1, L).type(dtype).to(device) + 1e-6
        return kappa, nu, zita

    # -- memorize --------------------------------------------------------------------------
    def swem(self, x, v, masks, bases_: Optional[Bases] = None, return_z: bool = False) -> Bases:
        """x (B,Ck,H,W) raw key, v (B,N,Cv,H,W), masks (B,N,2,H,W) -> bases dict (one fused EM update).

        Gradients (training): nu is differentiable in ``v`` and in the prior ``nu`` (reference :164-165); kappa and
        zita are constants, like under the reference's ``@torch.no_grad()`` steps."""
        B, Ck, H, W = x.shape
        N = masks.shape[1]
        L = self.n_bases
        if bases_ is None:
            kappa_, nu_, zita_ = self.random_init((B, N, 2, Ck, L), dtype=x.type(), device=x.device)
        else:
            kappa_, nu_, zita_ = bases_['kappa'], bases_['nu'], bases_['zita']
        n_new = N - kappa_.shape[1]
        if n_new > 0:
            k2, n2, z2 = self.random_init((B, n_new, 2, Ck, L), dtype=x.type(), device=x.device)
            kappa_, nu_, zita_ = torch.cat([kappa_, k2], 1), torch.cat([nu_, n2], 1), torch.cat([zita_, z2], 1)

        x, masks = _f32c(x.detach(), 'x'), _f32c(masks.detach(), 'masks')
        kappa_, zita_ = _f32c(kappa_.detach(), 'kappa'), _f32c(zita_.detach(), 'zita')
        nu_ = _f32c(nu_, 'nu')
        if not (_wants_grad(v, nu_) or return_z) and self._takes_pixel_major(v, B, N, Ck, H * W):
            v = v.detach()                # channels-last values go to the kernel as they are (no layout copy)
        else:
            v = _f32c(v, 'v')
        if _wants_grad(v, nu_):
            from .autograd import EMFunction
            kappa, nu, zita = EMFunction.apply(self, x, v, masks, kappa_, nu_, zita_)
            return {'kappa': kappa, 'nu': nu, 'zita': zita}
        kappa, nu, zita, z = self._em_launch(x, v.detach(), masks, kappa_, nu_.detach(), zita_, return_z)
        bases = {'kappa': kappa, 'nu': nu, 'zita': zita}
        if return_z:
            bases['z'] = z
        return bases

    def _takes_pixel_major(self, v, B, N, Ck, HW) -> bool:
        """True when ``v`` (B,N,Cv,H,W) is dense in channels-last order per object -- what a cuDNN NHWC value encoder
        (``FrameEngine``) hands over -- and the fused EM kernels, which read that layout directly, will run."""
        if not (v.is_cuda and v.dtype == torch.float32 and v.dim() == 5 and self.em_path != _lib.PATH_GENERIC):
            return False
        _, _, Cv, H, W = v.shape
        if v.stride() != (N * HW * Cv, HW * Cv, 1, W * Cv, Cv) or v.data_ptr() % 16:
            return False
        dims = _lib.SwemDims(B, N, Ck, Cv, HW, self.n_bases, self.n_iters, 0, 0, self.tau)
        return bool(_lib.load().swem_em_fused_supported(C.byref(dims)))

    def _em_launch(self, x, v, masks, kappa_, nu_, zita_, return_z: bool):
        """One ``swem_em_forward`` call on fp32 CUDA tensors (contiguous; ``v`` possibly channels-last, see
        ``_takes_pixel_major``) -> (kappa, nu, zita, z_last | None)."""
        B, Ck, H, W = x.shape
        N, L, Cv = masks.shape[1], self.n_bases, v.shape[2]
        if v.shape[:2] != (B, N) or v.shape[-2:] != (H, W) or kappa_.shape != (B, N, 2, Ck, L) \
                or nu_.shape != (B, N, 2, Cv, L) or masks.shape != (B, N, 2, H, W):
            raise RuntimeError(f'swem: inconsistent shapes x{tuple(x.shape)} v{tuple(v.shape)} '
                               f'masks{tuple(masks.shape)} kappa{tuple(kappa_.shape)} nu{tuple(nu_.shape)}')
        dev = x.device
        kappa = torch.empty_like(kappa_)
        nu = torch.empty_like(nu_)
        zita = torch.empty_like(zita_)
        z_last = torch.empty(B, N, 2, H * W, L, device=dev, dtype=torch.float32) if return_z else None

        lib = _lib.load()
        dims = _lib.SwemDims(B, N, Ck, Cv, H * W, L, self.n_iters, 0, 0, self.tau)
        need = lib.swem_em_workspace_bytes(C.byref(dims), self.em_path)
        ws = self._workspace.get(dev, need, 'em')
        # SwemEmArgs.image_workspace: when this call produces the 'update' bank of a two-bank memory (memorize() says so), the
        # kernel also leaves the readout's operand images of the new bases in the readout workspace: the next readout converts
        # nothing (bank 0 cached, bank 1 emitted here)
        img_ws = None
        if self._emit_images and not return_z and self.readout_path != _lib.PATH_GENERIC:
            rdims = _lib.SwemDims(B, N, Ck, Cv, H * W, L, 0, 2, self.topl, self.tau)
            rneed = lib.swem_readout_workspace_bytes(C.byref(rdims), self.readout_path)
            if rneed:
                img_ws = self._workspace.get(dev, rneed, 'readout')
        self._emit_images = False
        args = _lib.SwemEmArgs(dims, x.data_ptr(), v.data_ptr(), masks.data_ptr(),
                               kappa_.data_ptr(), nu_.data_ptr(), zita_.data_ptr(),
                               kappa.data_ptr(), nu.data_ptr(), zita.data_ptr(),
                               z_last.data_ptr() if return_z else None,
                               ws.data_ptr(), ws.numel(), self.em_path, 0 if v.is_contiguous() else 1,
                               None if img_ws is None else img_ws.data_ptr(), 1, 2)
        self._emitted = None if img_ws is None else (img_ws.data_ptr(), img_ws.numel(), B, N, Ck, Cv, L)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            rc = _invoke('em', lambda: lib.swem_em_forward(C.byref(args), stream))
        _lib.check(rc, 'swem_em_forward')
        self.launches = lib.swem_last_launch_count()
        return kappa, nu, zita, z_last

    def memorize(self, qk, qv, masks):
        prior = self.memories['update'].bases
        if prior is None:
            prior = self.memories['first'].bases
        # from the second memorize on the result is the 'update' bank of a two-bank memory (modules.py:183-193): ask the EM call
        # for its readout operand images (see _em_launch)
        self._emit_images = self.memories['first'].bases is not None and not torch.is_grad_enabled()
        self._emitted = None
        bases = self.swem(qk, qv, masks, prior)
        self._emit_images = False
        if self.memories['first'].bases is None:
            self.memories['first'].update(bases)
        else:
            n_first = self.memories['first'].n_objs
            self.memories['first'].update(bases)
            if self.memories['first'].n_objs != n_first or bases['kappa'].shape[1] != n_first:
                self._image_key = None                      # new objects joined the 'first' bank: its images are stale
            upd = self.memories['update']
            if self.static_banks and upd.bases is not None and upd.bases['kappa'].shape == bases['kappa'].shape:
                for key in ('kappa', 'nu', 'zita'):
                    upd.bases[key].copy_(bases[key])
            else:
                upd.update(bases)
            # the tensors that now hold the bank whose images the kernel emitted (static banks: the copies, same values)
            self._image1_src = None
            if self._emitted is not None:
                k1, n1 = upd.bases['kappa'], upd.bases['nu']
                self._image1_src = (self._emitted, k1, n1, k1._version, n1._version)

    # -- readout ---------------------------------------------------------------------------
    def matching_features(self, qk, qv) -> Tuple[torch.Tensor, int]:
        """-> the concat buffer [mem_out | qv | S] (B*N, 2*Cv + 2*topl, H, W) and N."""
        N, Cv = self._readout_objects()
        banks = self._banks()
        B, _, H, W = qk.shape
        # memory dropout (modules.py:258-263; training only, p_drop is 0.0 in the reference): one 0 / 1 mask per (b, n, basis of the
        # concatenated memory), drawn like the reference does -- torch.rand on the CPU's global generator -- and handed to the kernels
        drop = None
        if self.training and self.p_drop > 0:
            Lt = sum(b['nu'].shape[-1] for b in banks)
            drop = (torch.rand(B, N, 1, Lt, 1) > self.p_drop).to(device=qk.device, dtype=torch.float32).reshape(B, N, Lt).contiguous()
        if _wants_grad(qk, qv, *[b['nu'] for b in banks]):       # training: kernels forward, autograd-visible concat
            from .autograd import ReadoutFunction
            qk = _f32c(qk, 'qk')
            ms = ReadoutFunction.apply(self, qk, len(banks), drop, *[_f32c(b['kappa'].detach(), 'kappa') for b in banks],
                                       *[_f32c(b['nu'], 'nu') for b in banks])
            qv = qv.unsqueeze(1).expand(-1, N, -1, -1, -1).flatten(end_dim=1)
            return torch.cat([ms[:, :Cv], qv, ms[:, Cv:]], dim=1), N
        qv = _f32c(qv, 'qv')
        chans = 2 * Cv + 2 * self.topl
        feats = torch.empty(B * N, chans, H, W, device=qk.device, dtype=torch.float32)
        feats.view(B, N, chans, H, W)[:, :, Cv:2 * Cv] = qv.unsqueeze(1)
        return self.readout_into(qk, feats, 0, 2 * Cv, drop_mask=drop), N

    def _banks(self):
        return [m.bases for m in self.memories.values() if m.bases is not None]

    def _readout_objects(self) -> Tuple[int, int]:
        banks = self._banks()
        if not banks:
            raise RuntimeError('matching() before any memorize(): memory is empty')
        return banks[0]['nu'].shape[1], banks[0]['nu'].shape[3]

    def readout_into(self, qk, feats, mem_channel: int, s_channel: int, drop_mask=None) -> torch.Tensor:
        """Readout kernels only (no autograd): write ``mem_out`` into channels [mem_channel, +Cv) and ``S`` into
        channels [s_channel, +2*topl) of the caller's fp32 buffer ``feats`` (B*N, C, H, W), which is either contiguous
        (NCHW, the reference layout) or channels-last (NHWC; the kernels then write pixel-major).  The reference
        layout is ``matching_features``; inference engines that split the fusion conv use a narrower buffer."""
        banks = self._banks()
        if not banks:
            raise RuntimeError('matching() before any memorize(): memory is empty')
        return self._readout_launch(_f32c(qk.detach(), 'qk'), [_f32c(b['kappa'].detach(), 'kappa') for b in banks],
                                    [_f32c(b['nu'].detach(), 'nu') for b in banks], feats, mem_channel, s_channel,
                                    first_bank=(banks[0]['kappa'], banks[0]['nu']),
                                    update_bank=(banks[1]['kappa'], banks[1]['nu']) if len(banks) == 2 else None, drop_mask=drop_mask)

    def _readout_launch(self, qk, kap, nus, feats, mem_channel: int, s_channel: int, first_bank=None, update_bank=None,
                        drop_mask=None) -> torch.Tensor:
        """One ``swem_readout_forward`` call on contiguous fp32 CUDA tensors (kap / nus: one entry per bank).  ``first_bank``:
        the (kappa, nu) tensor OBJECTS of the 'first' bank as the bank holds them, or None when the caller cannot vouch for
        them (training): lets the call reuse that bank's operand images, see below."""
        B, Ck, H, W = qk.shape
        _, N, _, Cv, L = nus[0].shape
        dev = qk.device
        chans = feats.shape[1]
        pixel_major = 0 if feats.is_contiguous() else 1
        if (not feats.is_cuda or feats.dtype != torch.float32
                or not (feats.is_contiguous() or feats.is_contiguous(memory_format=torch.channels_last))
                or feats.shape != (B * N, chans, H, W) or mem_channel < 0 or mem_channel + Cv > chans
                or s_channel < 0 or s_channel + 2 * self.topl > chans):
            raise RuntimeError(f'readout: bad output buffer {tuple(feats.shape)} {feats.dtype} for B*N={B * N}, '
                               f'mem_channel={mem_channel}, s_channel={s_channel}')

        lib = _lib.load()
        dims = _lib.SwemDims(B, N, Ck, Cv, H * W, L, 0, len(nus), self.topl, self.tau)
        path = self.readout_path if drop_mask is None else _lib.PATH_GENERIC     # (memory dropout: generic family only)
        need = lib.swem_readout_workspace_bytes(C.byref(dims), path)
        ws = self._workspace.get(dev, need, 'readout')            # not shared with the EM: the bank images in it survive a memorize
        # SwemReadArgs.bank_images_valid: the 'first' bank does not change while a sequence runs (modules.py:44-60); its
        # operand images stay in this workspace, so from the second readout on only the 'update' bank is converted.  The
        # images are keyed by everything they depend on: the workspace, the shapes, and the bank's tensor objects (held here,
        # so their storage cannot be recycled for other data) with their version counters.
        key = None
        if first_bank is not None and len(nus) == 2 and path != _lib.PATH_GENERIC:
            key = (ws.data_ptr(), ws.numel(), B, N, Ck, Cv, L, first_bank[0]._version, first_bank[1]._version)
        prev = self._image_key
        valid = 1 if (key is not None and prev is not None and prev[0] == key and prev[1] is first_bank[0]
                      and prev[2] is first_bank[1]) else 0
        self._image_key = None if key is None else (key, first_bank[0], first_bank[1])
        src = self._image1_src                              # bank 1: images emitted by the memorize that produced it
        if (src is not None and update_bank is not None and key is not None
                and src[0] == (ws.data_ptr(), ws.numel(), B, N, Ck, Cv, L) and src[1] is update_bank[0] and src[2] is update_bank[1]
                and src[3] == update_bank[0]._version and src[4] == update_bank[1]._version):
            valid |= 2
        args = _lib.SwemReadArgs(dims, qk.data_ptr(),
                                 (C.c_void_p * 2)(*[k.data_ptr() for k in kap] + [None] * (2 - len(kap))),
                                 (C.c_void_p * 2)(*[n.data_ptr() for n in nus] + [None] * (2 - len(nus))),
                                 feats.data_ptr(), chans, mem_channel, s_channel,
                                 ws.data_ptr(), ws.numel(), path, pixel_major, valid, 0, 0.0, 0,
                                 None if drop_mask is None else drop_mask.data_ptr())
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            rc = _invoke('readout', lambda: lib.swem_readout_forward(C.byref(args), stream))
        _lib.check(rc, 'swem_readout_forward')
        self.launches = lib.swem_last_launch_count()
        return feats

    @torch.no_grad()
    def get_affinity(self, qk, mk, mv, n_kernel: int = 0, sigma: float = 7):
        """The reference's ``get_affinity`` (modules.py:232-276) with its signature: ``qk`` (B,Ck,H,W) and ``mk`` (B,N,2,Ck,Lt)
        l2-normalised as ``matching`` passes them (normalising again is the identity to 1e-6), ``mv`` (B,N,2,Cv,Lt) ->
        ``(S (B*N, 2 topl, H, W), mem_out (B,N,Cv,H,W))``.  Inference only (the training path is ``matching``).  ``n_kernel > 0``:
        the kernelised-memory branch (:252-256 with ``gen_kernels`` :210-230) -- off by default in the reference and never switched
        on by its scripts; implemented by the generic kernel family only (one ``swem_readout_forward`` call with
        ``SWEM_PATH_GENERIC``, ``mkm_kernels`` / ``mkm_sigma`` / ``mkm_width``).  In training mode the reference ignores ``n_kernel``
        (:252); so does this."""
        if self.training:
            n_kernel = 0
        qk, mk, mv = _f32c(qk, 'qk'), _f32c(mk, 'mk'), _f32c(mv, 'mv')
        B, Ck, H, W = qk.shape
        _, N, _, Cv, Lt = mv.shape
        topl = self.topl
        feats = torch.empty(B * N, Cv + 2 * topl, H, W, device=qk.device, dtype=torch.float32)
        lib = _lib.load()
        path = _lib.PATH_GENERIC if n_kernel > 0 else self.readout_path
        dims = _lib.SwemDims(B, N, Ck, Cv, H * W, Lt, 0, 1, topl, self.tau)
        ws = self._workspace.get(qk.device, lib.swem_readout_workspace_bytes(C.byref(dims), path), 'affinity')
        args = _lib.SwemReadArgs(dims, qk.data_ptr(), (C.c_void_p * 2)(mk.data_ptr(), None), (C.c_void_p * 2)(mv.data_ptr(), None),
                                 feats.data_ptr(), Cv + 2 * topl, 0, Cv, ws.data_ptr(), ws.numel(), path, 0, 0,
                                 int(n_kernel), float(sigma), W)
        with torch.cuda.device(qk.device):
            stream = torch.cuda.current_stream(qk.device).cuda_stream
            rc = _invoke('readout', lambda: lib.swem_readout_forward(C.byref(args), stream))
        _lib.check(rc, 'swem_readout_forward')
        self.launches = lib.swem_last_launch_count()
        return feats[:, Cv:].contiguous(), feats[:, :Cv].reshape(B, N, Cv, H, W)

    def matching(self, qk, qv):
        feats, n = self.matching_features(qk, qv)
        return self.fusion_layer(feats), n

    def forward(self, qk, qv):
        pass
