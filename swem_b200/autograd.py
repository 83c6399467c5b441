"""Autograd through the memory (training, BASELINE configs[4]; reference contract: SURVEY section 3.4 / 8b).

What is differentiable in the reference (verified on its autograd graph):

* ``swem`` (modules.py:129-168): the E / M / W steps run under ``@torch.no_grad()`` (:93, :112, :122), so kappa,
  zita and the responsibilities z are constants; only ``nu = (zita_ * nu_ + v z) / zita`` (:164-165) carries
  gradient, to the value features ``v`` and to the prior ``nu_``.  Forward = the fused / generic CUDA kernels
  with ``z_last`` saved; backward = ``swem_em_backward`` of the C ABI (two batched GEMMs + a scale kernel).
* ``matching`` (:278-293): gradient flows to the raw query key ``qk`` (through l2norm :282, affinity, exp,
  both the attention ``P`` and the sorted-prefix feature ``S``) and to the memory values ``nu`` of both banks;
  the memory keys are constants.  Forward = the CUDA readout kernels.  Backward = ``swem_readout_backward`` of the C
  ABI (forward recomputed; batched GEMMs + sorted-prefix, softmax and l2norm backward kernels, fp32).  A torch
  re-evaluation (``readout_torch``) is kept as the cross-check of the tests (``ReadoutFunction.native_backward``).

``qv`` and the fusion conv stay in torch autograd (``SWEMCore.matching`` concatenates with ``torch.cat`` when
gradients are needed).
"""
from __future__ import annotations

import ctypes as C
from typing import List

import torch

from . import _lib


# ------------------------------------------------------------------------------------------------
# EM update
# ------------------------------------------------------------------------------------------------
class EMFunction(torch.autograd.Function):
    """(kappa, nu, zita) = swem(x, v, masks | kappa_, nu_, zita_); differentiable in v and nu_ only."""

    @staticmethod
    def forward(ctx, core, x, v, masks, kappa_, nu_, zita_):
        kappa, nu, zita, z = core._em_launch(x, v, masks, kappa_, nu_, zita_, return_z=True)
        ctx.core = core
        ctx.v_shape = v.shape
        ctx.need = (v.requires_grad, nu_.requires_grad)
        ctx.save_for_backward(z, zita_, zita)
        ctx.mark_non_differentiable(kappa, zita)
        return kappa, nu, zita

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, _gk, gnu, _gz):
        z, zita_, zita = ctx.saved_tensors
        need_v, need_p = ctx.need
        if gnu is None or not (need_v or need_p):
            return (None,) * 7
        gnu = gnu.float().contiguous()
        B, N, _, Cv, L = gnu.shape
        HW = z.shape[3]
        dev = gnu.device
        gv = torch.empty(B, N, Cv, HW, device=dev, dtype=torch.float32) if need_v else None
        gp = torch.empty_like(gnu) if need_p else None
        lib = _lib.load()
        dims = _lib.SwemDims(B, N, 0, Cv, HW, L, 0, 0, 0, 1.0)
        from .core import _invoke
        ws = ctx.core._workspace.get(dev, lib.swem_em_backward_workspace_bytes(C.byref(dims)), 'bwd')
        args = _lib.SwemEmBwdArgs(dims, z.data_ptr(), zita_.data_ptr(), zita.data_ptr(), gnu.data_ptr(),
                                  gv.data_ptr() if need_v else None, gp.data_ptr() if need_p else None,
                                  ws.data_ptr(), ws.numel())
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            rc = _invoke('em_backward', lambda: lib.swem_em_backward(C.byref(args), stream))
        _lib.check(rc, 'swem_em_backward')
        return None, None, (gv.view(ctx.v_shape) if need_v else None), None, None, gp, None


# ------------------------------------------------------------------------------------------------
# Readout
# ------------------------------------------------------------------------------------------------
def _unit(t: torch.Tensor, dim: int) -> torch.Tensor:
    return t / (torch.linalg.vector_norm(t, dim=dim, keepdim=True) + 1e-6)


def readout_torch(qk: torch.Tensor, kappas: List[torch.Tensor], nus: List[torch.Tensor], tau: float, topl: int, drop_mask=None):
    """Differentiable torch evaluation of the readout: -> [mem_out | S] (B*N, Cv + 2*topl, H, W).

    qk (B,Ck,H,W) raw; kappas / nus: per bank (B,N,2,Ck,L) / (B,N,2,Cv,L).  Used for the backward only."""
    B, Ck, H, W = qk.shape
    q = _unit(qk, 1).flatten(2)[:, None, None]                                   # B,1,1,Ck,HW
    mk = _unit(torch.cat(kappas, dim=-1), -2)                                    # B,N,2,Ck,Lt
    mv = torch.cat(nus, dim=-1)                                                  # B,N,2,Cv,Lt
    N, Lt, Cv = mk.shape[1], mk.shape[-1], mv.shape[3]
    aff = mk.transpose(-2, -1) @ q                                               # B,N,2,Lt,HW
    e = torch.exp((aff - aff.amax(dim=(2, 3), keepdim=True)) / tau)
    if drop_mask is not None:                                                    # memory dropout (modules.py:258-263): (B,N,Lt) of 0 / 1
        ed = e * drop_mask[:, :, None, :, None]
        p = ed / (ed.sum(dim=(2, 3), keepdim=True) + 1e-6)
    else:
        p = e / e.sum(dim=(2, 3), keepdim=True)
    mem_out = (mv.transpose(2, 3).flatten(-2) @ p.flatten(2, 3)).reshape(B * N, Cv, H, W)
    run = torch.topk(e, k=topl, dim=3)[0].cumsum(dim=3)                          # sorted descending -> running sums over rank
    f = (run[:, :, 0] / (run[:, :, 0] + run[:, :, 1])).reshape(B * N, topl, H, W)
    return torch.cat([mem_out, f, 1 - f], dim=1)


class ReadoutFunction(torch.autograd.Function):
    """[mem_out | S] = readout(qk | banks); differentiable in qk and in the nu of every bank."""

    @staticmethod
    def forward(ctx, core, qk, n_banks, drop_mask, *bank_tensors):
        kappas, nus = list(bank_tensors[:n_banks]), list(bank_tensors[n_banks:])
        B, _, H, W = qk.shape
        N, Cv = nus[0].shape[1], nus[0].shape[3]
        out = torch.empty(B * N, Cv + 2 * core.topl, H, W, device=qk.device, dtype=torch.float32)
        core._readout_launch(qk, kappas, nus, out, 0, Cv, drop_mask=drop_mask)
        ctx.tau, ctx.topl, ctx.n_banks = core.tau, core.topl, n_banks
        ctx.core = core
        ctx.drop_mask = drop_mask
        ctx.save_for_backward(qk, *kappas, *nus)
        return out

    #: True: `swem_readout_backward` of the C ABI (hand-written kernels).  False: differentiate `readout_torch` (kept as the
    #: cross-check the tests compare against).
    native_backward = True

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gout):
        qk, *rest = ctx.saved_tensors
        nb = ctx.n_banks
        kappas, nus = rest[:nb], rest[nb:]
        need_q = ctx.needs_input_grad[1]
        need_nu = [ctx.needs_input_grad[4 + nb + k] for k in range(nb)]
        if not (need_q or any(need_nu)):
            return (None,) * (4 + 2 * nb)
        if ReadoutFunction.native_backward:
            gq, gn = _readout_backward_native(ctx, qk, kappas, nus, gout.float().contiguous(), need_q, need_nu)
        else:
            with torch.enable_grad():
                q_ = qk.detach().requires_grad_(need_q)
                nus_ = [n.detach().requires_grad_(need_nu[k]) for k, n in enumerate(nus)]
                out = readout_torch(q_, [k.detach() for k in kappas], nus_, ctx.tau, ctx.topl, ctx.drop_mask)
                wrt = ([q_] if need_q else []) + [n for k, n in enumerate(nus_) if need_nu[k]]
                grads = list(torch.autograd.grad(out, wrt, gout))
            gq = grads.pop(0) if need_q else None
            gn = [grads.pop(0) if need_nu[k] else None for k in range(nb)]
        return (None, gq, None, None) + (None,) * nb + tuple(gn)


def _readout_backward_native(ctx, qk, kappas, nus, gout, need_q, need_nu):
    from .core import _invoke
    nb = ctx.n_banks
    B, Ck, H, W = qk.shape
    _, N, _, Cv, L = nus[0].shape
    dev = qk.device
    chans = gout.shape[1]
    gq = torch.empty_like(qk) if need_q else None
    gn = [torch.empty_like(nus[k]) if need_nu[k] else None for k in range(nb)]
    lib = _lib.load()
    dims = _lib.SwemDims(B, N, Ck, Cv, H * W, L, 0, nb, ctx.topl, ctx.tau)
    ws = ctx.core._workspace.get(dev, lib.swem_readout_backward_workspace_bytes(C.byref(dims)), 'bwd')
    ptr = lambda t: None if t is None else t.data_ptr()
    pad = [None] * (2 - nb)
    args = _lib.SwemReadBwdArgs(dims, qk.data_ptr(),
                                (C.c_void_p * 2)(*[ptr(k) for k in kappas] + pad), (C.c_void_p * 2)(*[ptr(n) for n in nus] + pad),
                                gout.data_ptr(), chans, 0, Cv, ptr(gq), (C.c_void_p * 2)(*[ptr(g) for g in gn] + pad),
                                ws.data_ptr(), ws.numel(), ptr(ctx.drop_mask))
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        rc = _invoke('readout_backward', lambda: lib.swem_readout_backward(C.byref(args), stream))
    _lib.check(rc, 'swem_readout_backward')
    return gq, gn
