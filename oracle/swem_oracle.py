"""CPU oracle for the SWEM memory hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A restatement, in plain torch-on-CPU arithmetic (fp32 like the reference, or fp64 as a
higher-precision yardstick), of the sequential weighted EM memory of lmm077/SWEM:

    reference file                        what is restated here
    ------------------------------------  -----------------------------------------
    methods/SWEM/modules.py:7-9           l2norm
    methods/SWEM/modules.py:93-110        w_step      (sww_step)
    methods/SWEM/modules.py:112-120       e_step      (swe_step)
    methods/SWEM/modules.py:122-127       m_step      (swm_step)
    methods/SWEM/modules.py:129-168       em_memorize (swem) incl. the nu update
    methods/SWEM/modules.py:170-178       random_init
    methods/SWEM/modules.py:29-60,183-193 MemoryBanks (MemoryBank x2 + memorize bookkeeping)
    methods/SWEM/modules.py:198-208       perm_inv_feat
    methods/SWEM/modules.py:232-276       readout     (get_affinity, default branch only)
    methods/SWEM/modules.py:278-306       OracleSWEMCore.matching / get_mem
    methods/SWEM/swem.py:69-86            build_em_masks (mask prep of SWEM.memorize)
    methods/SWEM/swem.py:92-116           aggregate / decode tail
    methods/SWEM/swem_evaluator.py:59-148 run_davis_sequence / run_ytvos_sequence

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu-baseline / ``--impl reference``
legs may import this module.  The product package ``swem_b200`` never does: its hot path is the
CUDA library and it raises when that library is missing.

Parity pinning: the reference ships no tests or golden vectors for this path, so
the oracle is pinned against outputs of the reference itself: ``tests/golden/make_golden.py``
imports the unmodified reference from /root/reference (authoring container only), runs it on
seeded inputs and stores inputs + outputs as fixtures under ``tests/golden/``;
``tests/test_oracle_golden.py`` replays them through this file and demands bit-equality in fp32.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Bases = Dict[str, torch.Tensor]          # {'kappa': (B,N,2,Ck,L), 'nu': (B,N,2,Cv,L), 'zita': (B,N,2,1,L)}

EPS_NORM = 1e-6                            # modules.py:8
ZITA_INIT = 1e-6                           # modules.py:176


def l2norm(t: torch.Tensor, dim: int) -> torch.Tensor:
    """t / (||t||_2 + 1e-6) along ``dim`` -- epsilon added to the norm (modules.py:7-9)."""
    return t / (torch.linalg.norm(t, dim=dim, keepdim=True) + EPS_NORM)


def random_init(B: int, N: int, Ck: int, L: int, Cv: int, dtype=torch.float32,
                generator: Optional[torch.Generator] = None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """kappa ~ N(0, sqrt(2/L)) then l2-normalised over Ck; nu = 0; zita = 1e-6 (modules.py:170-178)."""
    kappa = torch.zeros(B, N, 2, Ck, L, dtype=dtype)
    kappa.normal_(0, math.sqrt(2.0 / L), generator=generator)
    kappa = l2norm(kappa, dim=-2)
    nu = torch.zeros(B, N, 2, Cv, L, dtype=dtype)
    zita = torch.zeros(B, N, 2, 1, L, dtype=dtype) + ZITA_INIT
    return kappa, nu, zita


# ------------------------------------------------------------------------------------------
# The three EM steps.  Shapes: x_t (B,1,1,HW,Ck)  x (B,1,1,Ck,HW)  kappa (B,N,2,Ck,L)
#                              weights/masks (B,N,2,HW,1)  z (B,N,2,HW,L)
# ------------------------------------------------------------------------------------------
def e_step(x_t, kappa, weights, tau):
    """Responsibilities: per-side softmax over bases of raw-key . unit-basis / tau, times the pixel weight."""
    logits = torch.matmul(x_t, l2norm(kappa, dim=-2))
    logits = logits - logits.max(dim=-1, keepdim=True)[0]
    return F.softmax(logits / tau, dim=-1) * weights


def m_step(z, x, kappa_prior, zita_prior):
    """Weight-normalised re-estimate of the key bases, always starting from the prior."""
    zita = zita_prior + z.sum(dim=-2, keepdim=True)
    kappa = (zita_prior * kappa_prior + torch.matmul(x, z)) / zita
    return kappa, zita


def w_step(kappa, x_t, masks, tau):
    """New pixel weights: mask x (1 - probability mass the pixel already gives to its own side)."""
    sim = torch.matmul(l2norm(x_t, dim=-1), l2norm(kappa, dim=-2))          # cosine, (B,N,2,HW,L)
    peak = sim.max(dim=-1, keepdim=True)[0].max(dim=2, keepdim=True)[0]      # over bases and sides
    mass = torch.exp((sim - peak) / tau).sum(dim=-1, keepdim=True)           # (B,N,2,HW,1)
    share = mass / mass.sum(dim=2, keepdim=True)
    return masks * (1 - share)


def em_memorize(x: torch.Tensor, v: torch.Tensor, masks: torch.Tensor, prior: Optional[Bases],
                n_bases: int, n_iters: int, tau: float,
                generator: Optional[torch.Generator] = None, trace: Optional[dict] = None) -> Bases:
    """One sequential weighted EM update (modules.py:129-168).

    x (B,Ck,H,W) raw key, v (B,N,Cv,H,W) values, masks (B,N,2,H,W) [bg, fg].  ``prior`` None or a
    bases dict that may hold fewer objects than ``masks`` (late-appearing objects get a fresh
    random init appended).  ``trace`` (optional dict) receives per-iteration z / weights / kappa.
    """
    B, Ck, H, W = x.shape
    N = masks.shape[1]
    Cv = v.shape[2]
    if prior is None:
        kappa_p, nu_p, zita_p = random_init(B, N, Ck, n_bases, Cv, x.dtype, generator)
    else:
        kappa_p, nu_p, zita_p = prior['kappa'], prior['nu'], prior['zita']
    n_new = N - kappa_p.shape[1]
    if n_new > 0:
        k2, n2, z2 = random_init(B, n_new, Ck, n_bases, Cv, x.dtype, generator)
        kappa_p = torch.cat([kappa_p, k2], dim=1)
        nu_p = torch.cat([nu_p, n2], dim=1)
        zita_p = torch.cat([zita_p, z2], dim=1)

    xf = x.flatten(start_dim=-2)[:, None, None]                # B,1,1,Ck,HW
    x_t = xf.transpose(-2, -1)                                 # B,1,1,HW,Ck
    m = masks.flatten(start_dim=-2).unsqueeze(-1)              # B,N,2,HW,1
    weights = m.clone()
    kappa = kappa_p.clone()
    if trace is not None:
        trace.update(kappa_prior=kappa_p, nu_prior=nu_p, zita_prior=zita_p, z=[], weights=[m.clone()], kappa=[])
    z = zita = None
    for it in range(n_iters):
        z = e_step(x_t, kappa, weights, tau)
        kappa, zita = m_step(z, xf, kappa_p, zita_p)
        if it < n_iters - 1:
            weights = w_step(kappa, x_t, m, tau)
        if trace is not None:
            trace['z'].append(z)
            trace['kappa'].append(kappa)
            if it < n_iters - 1:
                trace['weights'].append(weights)
    vf = v.flatten(start_dim=-2).unsqueeze(2)                  # B,N,1,Cv,HW  (both sides share v)
    nu = (zita_p * nu_p + torch.matmul(vf, z)) / zita
    return {'kappa': kappa, 'nu': nu, 'zita': zita}


# ------------------------------------------------------------------------------------------
# Memory banks
# ------------------------------------------------------------------------------------------
class MemoryBanks:
    """'first' bank (fixed: only appends objects it has not seen) + 'update' bank (replaced every call).

    Mirrors MemoryBank (modules.py:29-60) and the bookkeeping of SWEMCore.memorize (:183-193).
    """

    def __init__(self):
        self.first: Optional[Bases] = None
        self.update: Optional[Bases] = None
        self.first_n = 0

    def clear(self):
        self.first = self.update = None
        self.first_n = 0

    def prior(self) -> Optional[Bases]:
        return self.update if self.update is not None else self.first

    def commit(self, bases: Bases):
        if self.first is None:
            self.first, self.first_n = bases, bases['kappa'].shape[1]
            return                                              # first call fills only 'first'
        n = bases['kappa'].shape[1]
        if n > self.first_n:
            self.first = {k: torch.cat([self.first[k], bases[k][:, self.first_n:]], dim=1) for k in bases}
        self.first_n = n
        self.update = bases

    def read(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """Concatenate the non-empty banks along the basis axis, order [first | update] (:295-306)."""
        banks = [b for b in (self.first, self.update) if b is not None]
        return (torch.cat([b['kappa'] for b in banks], dim=-1),
                torch.cat([b['nu'] for b in banks], dim=-1))


# ------------------------------------------------------------------------------------------
# Readout
# ------------------------------------------------------------------------------------------
def perm_inv_feat(e: torch.Tensor, topl: int) -> torch.Tensor:
    """e (BN,2,Lt,H,W) un-normalised exp-affinities -> (BN, 2*topl, H, W).

    Per pixel and side: the topl largest values in descending order, running sums over rank,
    then the background share of the running sum at each rank, and its complement (:198-208).
    The running sum is accumulated sequentially in rank order like the reference's loop.
    """
    top = torch.topk(e, k=topl, dim=2)[0]
    run = torch.zeros_like(top)
    run[:, :, 0] = top[:, :, 0]
    for r in range(1, topl):
        run[:, :, r] = run[:, :, r - 1] + top[:, :, r]
    f = run[:, 0] / (run[:, 0] + run[:, 1])
    return torch.cat([f, 1 - f], dim=1)


def kernel_weights(aff: torch.Tensor, H: int, W: int, n_kernel: int, sigma: float, tau: float) -> torch.Tensor:
    """The reference's kernelised-memory weights (`gen_kernels`, :210-230; inference only, off by default): every basis
    puts a Gaussian of width sigma on the n_kernel pixels it matches best, a pixel keeps the largest of them:
    aff (B,N,2,Lt,HW) raw affinities -> exp(-min_k d^2(p, p_k) / (2 sigma^2) / tau) of the same shape."""
    idx = torch.topk(aff, k=n_kernel, dim=-1)[1]                             # B,N,2,Lt,k  pixel indices
    xk, yk = (idx % W).unsqueeze(-2), ((idx // W) % H).unsqueeze(-2)         # B,N,2,Lt,1,k
    yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing='ij')
    xx = xx.reshape(1, 1, 1, 1, H * W, 1).to(aff.dtype)
    yy = yy.reshape(1, 1, 1, 1, H * W, 1).to(aff.dtype)
    g = -((xx - xk) ** 2 + (yy - yk) ** 2) / (2 * sigma ** 2)                # B,N,2,Lt,HW,k
    return torch.exp(g.max(dim=-1)[0] / tau)


def readout(qk_unit: torch.Tensor, mk_unit: torch.Tensor, mv: torch.Tensor, tau: float, topl: int,
            trace: Optional[dict] = None, n_kernel: int = 0, sigma: float = 7,
            drop_mask: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """qk_unit (B,Ck,H,W) and mk_unit (B,N,2,Ck,Lt) already l2-normalised; mv (B,N,2,Cv,Lt).

    Returns S (BN, 2*topl, H, W) and mem_out (B,N,Cv,H,W) (:232-276; n_kernel > 0: the kernelised branch :252-256, in which the
    attention -- not S -- is weighted by `kernel_weights` and normalised with + 1e-8; drop_mask (B,N,1,Lt,1) of 0 / 1: the memory
    dropout of the training branch :258-263 -- the reference draws it as torch.rand(B,N,1,Lt,1) > p_drop -- normalised with + 1e-6).
    """
    B, Ck, H, W = qk_unit.shape
    N, Lt = mk_unit.shape[1], mk_unit.shape[-1]
    q = qk_unit.flatten(start_dim=-2)[:, None, None]                       # B,1,1,Ck,HW
    aff = torch.matmul(mk_unit.transpose(-2, -1), q)                        # B,N,2,Lt,HW
    peak = aff.max(dim=2, keepdim=True)[0].max(dim=3, keepdim=True)[0]      # B,N,1,1,HW
    e = torch.exp((aff - peak) / tau)
    if n_kernel > 0:
        eg = (e * kernel_weights(aff, H, W, n_kernel, sigma, tau)).flatten(start_dim=2, end_dim=3)
        p = eg / (eg.sum(dim=2, keepdim=True) + 1e-8)
    elif drop_mask is not None:
        ed = e * drop_mask
        p = (ed / (ed.sum(dim=[2, 3], keepdim=True) + 1e-6)).flatten(start_dim=2, end_dim=3)
    else:
        p = (e / e.sum(dim=[2, 3], keepdim=True)).flatten(start_dim=2, end_dim=3)   # B,N,2Lt,HW
    S = perm_inv_feat(e.view(B * N, 2, Lt, H, W), topl)
    vals = mv.transpose(2, 3).flatten(start_dim=-2)                         # B,N,Cv,2Lt  (column = side*Lt + j)
    mem_out = torch.matmul(vals, p).view(B, N, -1, H, W)
    if trace is not None:
        trace.update(aff=aff, e=e, p=p)
    return S, mem_out


class OracleSWEMCore:
    """Same surface as the reference SWEMCore minus the fusion conv (which is a torch module
    owned by the caller): empty / memorize / get_mem / matching_features."""

    def __init__(self, n_bases=256, valdim=512, n_iters=4, tau=0.05, topl=64):
        assert tau > 0
        self.n_bases, self.valdim, self.n_iters, self.tau = n_bases, valdim, n_iters, tau
        self.topl = int(min(n_bases, topl))
        self.banks = MemoryBanks()
        self.generator: Optional[torch.Generator] = None      # None -> global torch RNG, like the reference

    def empty(self):
        self.banks.clear()

    def memorize(self, qk, qv, masks):
        bases = em_memorize(qk, qv, masks, self.banks.prior(), self.n_bases, self.n_iters, self.tau,
                            generator=self.generator)
        self.banks.commit(bases)

    def get_mem(self):
        return self.banks.read()

    def matching_features(self, qk, qv):
        """-> concat [mem_out | qv | S] of shape (B*N, 2*Cv + 2*topl, H, W), and N (:278-291)."""
        mk, mv = self.get_mem()
        S, mem_out = readout(l2norm(qk, dim=1), l2norm(mk, dim=-2), mv, self.tau, self.topl)
        qv_e = qv.unsqueeze(1).expand_as(mem_out)
        return torch.cat([mem_out.flatten(end_dim=1), qv_e.flatten(end_dim=1), S], dim=1), mk.shape[1]


# ------------------------------------------------------------------------------------------
# Glue around the core (swem.py) and the per-sequence loops (swem_evaluator.py)
# ------------------------------------------------------------------------------------------
def build_em_masks(masks_hard: torch.Tensor, masks_soft: torch.Tensor, h16: int, w16: int) -> torch.Tensor:
    """(B,N+1,Hm,Wm) hard + soft masks -> (B,N,2,h16,w16) [bg, fg] EM weights (swem.py:80-84)."""
    hard = F.interpolate(masks_hard[:, 1:].float(), size=(h16, w16), mode='nearest')
    soft = F.interpolate(masks_soft[:, 1:], size=(h16, w16), mode='bilinear')
    return torch.stack([(1 - hard) * (1 - soft), hard * soft], dim=2)


def aggregate(prob: torch.Tensor) -> torch.Tensor:
    """Soft aggregation of per-object probabilities into (N+1)-way logits (swem.py:110-116)."""
    allp = torch.cat([torch.prod(1 - prob, dim=1, keepdim=True), prob], dim=1).clamp(1e-7, 1 - 1e-7)
    return torch.log(allp / (1 - allp))


class OracleSWEM:
    """Whole-model CPU path: the torch networks of ``swem_b200.networks`` (device-agnostic torch
    modules, not hot-path code) around OracleSWEMCore.  Used for free-running mask parity and as
    the CPU baseline of bench.py.  ``nets`` is any object exposing key_encoder, value_encoder,
    key_proj, key_comp, decoder, single_object and fusion_layer (or swem_core.fusion_layer) (e.g. a swem_b200.model.SWEM on CPU).
    """

    def __init__(self, nets, n_bases, n_iters, tau, topl, valdim=512):
        self.nets = nets
        self.fusion_layer = getattr(nets, 'fusion_layer', None) or nets.swem_core.fusion_layer
        self.core = OracleSWEMCore(n_bases, valdim, n_iters, tau, topl)

    def encode_key(self, frame):
        s16, s8, s4 = self.nets.key_encoder(frame)
        return self.nets.key_proj(s16), self.nets.key_comp(s16), s16, s8, s4

    def encode_value(self, frame, masks, s16):
        N = masks.shape[1] - 1
        others = 1 - masks - masks[:, 0:1]
        fg = masks[:, 1:].flatten(end_dim=1).unsqueeze(1)
        ot = others[:, 1:].flatten(end_dim=1).unsqueeze(1)
        fr = frame.unsqueeze(1).expand(-1, N, -1, -1, -1).flatten(end_dim=1)
        s = s16.unsqueeze(1).expand(-1, N, -1, -1, -1).flatten(end_dim=1)
        mv = self.nets.value_encoder(fr, s, fg) if self.nets.single_object else self.nets.value_encoder(fr, s, fg, ot)
        return mv.view(-1, N, *mv.shape[1:])

    def memorize(self, qk16, mv16, hard, soft):
        self.core.memorize(qk16, mv16, build_em_masks(hard, soft, *qk16.shape[-2:]))

    def init(self, qk16, mv16, mask):
        self.core.empty()
        self.memorize(qk16, mv16, mask, mask.float())

    def match(self, qk16, qv16):
        feats, n = self.core.matching_features(qk16, qv16)
        return self.fusion_layer(feats), n

    def segment(self, n, context, s8, s4, out_size):
        s8e = s8.unsqueeze(1).expand(-1, n, -1, -1, -1).flatten(end_dim=1)
        s4e = s4.unsqueeze(1).expand(-1, n, -1, -1, -1).flatten(end_dim=1)
        prob = torch.sigmoid(self.nets.decoder(context, s8e, s4e, out_size))
        logits = aggregate(prob.view(-1, n, *prob.shape[-2:]))
        return logits, F.softmax(logits, dim=1)


def one_hot_from_argmax(pred_mask: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """(B,N+1,H,W) scores -> (argmax (B,1,H,W), one-hot int64 (B,N+1,H,W)) (swem_evaluator.py:83-87)."""
    pred = torch.argmax(pred_mask, dim=1, keepdim=True)
    idx = torch.arange(pred_mask.shape[1], dtype=pred.dtype, device=pred.device).view(1, -1, 1, 1)
    return pred, (pred == idx).type_as(pred)


@torch.no_grad()
def run_davis_sequence(model: OracleSWEM, frames: torch.Tensor, init_mask: torch.Tensor, out_size,
                       timers: Optional[dict] = None) -> List[torch.Tensor]:
    """frames (1,T,3,h,w); init_mask (1,N+1,H,W) one-hot float -> list of T-1 argmax masks (1,H,W).

    Restates SWEMEvaluator.evaluate_davis_seq (swem_evaluator.py:59-102).
    """
    import time
    b, t, c, h, w = frames.shape
    clock = (lambda: time.perf_counter()) if timers is not None else None

    def lap(name, t0):
        if timers is not None:
            timers[name] = timers.get(name, 0.0) + (clock() - t0)

    t0 = clock() if clock else 0
    mk16, _, s16, _, _ = model.encode_key(frames[:, 0])
    m0 = F.interpolate(init_mask, size=(h, w), mode='nearest')
    mv16 = model.encode_value(frames[:, 0], m0.float(), s16)
    model.init(mk16, mv16, init_mask)
    lap('init', t0)
    preds = []
    for i in range(1, t):
        t0 = clock() if clock else 0
        qk16, qv16, s16, s8, s4 = model.encode_key(frames[:, i]); lap('encode_key', t0)
        t0 = clock() if clock else 0
        context, n = model.match(qk16, qv16); lap('match', t0)
        t0 = clock() if clock else 0
        _, pred_mask = model.segment(n, context, s8, s4, out_size); lap('segment', t0)
        pred, hard = one_hot_from_argmax(pred_mask)
        if i < t - 1:
            t0 = clock() if clock else 0
            soft = F.interpolate(pred_mask, size=(h, w), mode='bilinear', align_corners=False)
            mv16 = model.encode_value(frames[:, i], soft, s16); lap('encode_value', t0)
            t0 = clock() if clock else 0
            model.memorize(qk16, mv16, hard, soft); lap('memorize', t0)
        preds.append(pred[:, 0])
    return preds


@torch.no_grad()
def run_ytvos_sequence(model: OracleSWEM, frames: torch.Tensor, init_masks: List[Optional[torch.Tensor]],
                       out_size) -> List[torch.Tensor]:
    """Like run_davis_sequence but objects may appear later: ``init_masks[i]`` is None or a
    (1,N'+1,H,W) one-hot mask holding the NEW objects of frame i (swem_evaluator.py:104-148)."""
    b, t, c, h, w = frames.shape
    mk16, _, s16, _, _ = model.encode_key(frames[:, 0])
    m0 = F.interpolate(init_masks[0], size=(h, w), mode='nearest')
    mv16 = model.encode_value(frames[:, 0], m0.float(), s16)
    model.init(mk16, mv16, init_masks[0])
    preds = []
    for i in range(1, t):
        qk16, qv16, s16, s8, s4 = model.encode_key(frames[:, i])
        context, n = model.match(qk16, qv16)
        _, pred_mask = model.segment(n, context, s8, s4, out_size)
        if init_masks[i] is not None:
            new = init_masks[i][:, 1:].sum(dim=1, keepdim=True).expand_as(pred_mask)
            pred_mask[new > 0] = 0
            pred_mask = torch.cat([pred_mask, init_masks[i][:, 1:]], dim=1)
        pred, hard = one_hot_from_argmax(pred_mask)
        if i < t - 1:
            soft = F.interpolate(pred_mask, size=(h, w), mode='bilinear', align_corners=False)
            mv16 = model.encode_value(frames[:, i], soft, s16)
            model.memorize(qk16, mv16, hard, soft)
        preds.append(pred[:, 0])
    return preds
