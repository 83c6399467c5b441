"""Seeded synthetic inputs shaped like the reference's datasets (no dataset ships offline).

Shapes follow SURVEY section 8d: DAVIS frames are 480x854 bicubic-resized to 480x864 before the
model (basic_evaluator.py:160); YouTube-VOS frames are short-side-480 (480x848) or native 720p.
"""
from __future__ import annotations

import zlib
from typing import List, Optional, Tuple

import torch
import torch.nn.functional as F


def smooth_video(t: int, h: int, w: int, seed: int = 1, noise: float = 0.02) -> torch.Tensor:
    """(1,T,3,h,w) in [0,1]: low-resolution noise upsampled bicubically, drifting (4,6) px/frame."""
    g = torch.Generator().manual_seed(seed)
    base = torch.rand(1, 3, h // 16 + 2, w // 16 + 2, generator=g)
    base = F.interpolate(base, size=(h, w), mode='bicubic', align_corners=False).clamp(0, 1)[0]
    frames = []
    for i in range(t):
        f = torch.roll(base, shifts=(4 * i, 6 * i), dims=(1, 2))
        frames.append((f + noise * torch.rand(3, h, w, generator=g)).clamp(0, 1))
    return torch.stack(frames)[None]


def rectangle_masks(n_obj: int, h: int, w: int, seed: int = 1) -> torch.Tensor:
    """(1,N+1,h,w) float one-hot first-frame annotation: N non-overlapping rectangles on a grid."""
    g = torch.Generator().manual_seed(seed + 1000)
    labels = torch.zeros(h, w, dtype=torch.int64)
    cols = max(1, min(n_obj, 3))
    rows = (n_obj + cols - 1) // cols
    for k in range(n_obj):
        r, c = divmod(k, cols)
        y0, y1 = int(h * (r + 0.15) / rows), int(h * (r + 0.85) / rows)
        x0, x1 = int(w * (c + 0.15) / cols), int(w * (c + 0.85) / cols)
        jy, jx = (int(v) for v in torch.randint(-8, 9, (2,), generator=g))
        labels[max(0, y0 + jy):min(h, y1 + jy), max(0, x0 + jx):min(w, x1 + jx)] = k + 1
    return F.one_hot(labels, n_obj + 1).permute(2, 0, 1)[None].float()


def davis_sequence(t: int = 50, n_obj: int = 5, seed: int = 1, size: Tuple[int, int] = (480, 864)):
    """DAVIS-2017-shaped sequence (BASELINE config 2): frames (1,T,3,480,864), init mask (1,N+1,480,864)."""
    h, w = size
    return smooth_video(t, h, w, seed), rectangle_masks(n_obj, h, w, seed)


def ytvos_sequences(n_seq: int, seed: int = 7) -> List[dict]:
    """YouTube-VOS-shaped batch (BASELINE config 3): mixed sizes, 1-6 objects, some appearing late."""
    g = torch.Generator().manual_seed(seed)
    sizes = [(480, 848), (480, 864), (720, 1280)]
    out = []
    for k in range(n_seq):
        h, w = sizes[int(torch.randint(0, 3, (1,), generator=g))]
        t = int(torch.randint(20, 37, (1,), generator=g))
        n = int(torch.randint(1, 7, (1,), generator=g))
        n_late = int(torch.randint(0, 2, (1,), generator=g)) if n > 1 else 0
        out.append(dict(h=h, w=w, t=t, n_obj=n, n_late=n_late, late_frame=t // 2, seed=seed * 1000 + k))
    return out


def ytvos_materialise(spec: dict):
    """spec from :func:`ytvos_sequences` -> frames (1,T,3,h,w) and the init_masks list."""
    frames = smooth_video(spec['t'], spec['h'], spec['w'], spec['seed'])
    full = rectangle_masks(spec['n_obj'], spec['h'], spec['w'], spec['seed'])      # (1,N+1,h,w)
    n0 = spec['n_obj'] - spec['n_late']
    first = full[:, :n0 + 1].clone()
    first[:, 0] = 1 - first[:, 1:].sum(dim=1)
    init_masks: List[Optional[torch.Tensor]] = [None] * spec['t']
    init_masks[0] = first
    if spec['n_late'] > 0:
        late = torch.cat([torch.zeros_like(full[:, :1]), full[:, n0 + 1:]], dim=1)
        late[:, 0] = 1 - late[:, 1:].sum(dim=1)
        init_masks[spec['late_frame']] = late
    return frames, init_masks


def em_inputs(B: int, N: int, Ck: int, Cv: int, H: int, W: int, seed: int = 0, fg_prob: float = 0.3):
    """Micro-benchmark inputs with the statistics measured on random-init encoders (SURVEY section 8d,
    config 4): x ~ N(0, 2.3^2), v ~ N(0, 1.8^2), hard fg masks ~ Bernoulli(fg_prob), soft = hard."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, Ck, H, W, generator=g) * 2.3
    v = torch.randn(B, N, Cv, H, W, generator=g) * 1.8
    fg = (torch.rand(B, N, H, W, generator=g) < fg_prob).float()
    masks = torch.stack([1 - fg, fg], dim=2)
    return x, v, masks


def fill_deterministic_(module: torch.nn.Module, seed: int = 0) -> None:
    """Overwrite every parameter/buffer with values that depend only on (name, shape, seed), so two
    differently-constructed but key-compatible models (reference vs. this repo) get equal weights."""
    sd = module.state_dict()
    for name in sorted(sd):
        t = sd[name]
        if not t.is_floating_point():
            continue
        g = torch.Generator().manual_seed((zlib.crc32(name.encode()) + seed) & 0x7FFFFFFF)
        if name.endswith('running_var'):
            val = 1.0 + 0.1 * torch.rand(t.shape, generator=g)
        elif name.endswith('running_mean') or name in ('mean', 'std') or name.endswith('.mean') or name.endswith('.std'):
            if name.endswith('running_mean'):
                val = 0.05 * torch.randn(t.shape, generator=g)
            else:
                continue                                    # keep the ImageNet normalisation buffers
        elif t.dim() >= 2:
            fan_in = t[0].numel()
            val = torch.randn(t.shape, generator=g) * (1.0 / fan_in) ** 0.5
        elif name.endswith('weight'):                       # BN scale
            val = 1.0 + 0.1 * torch.randn(t.shape, generator=g)
        else:                                               # biases
            val = 0.02 * torch.randn(t.shape, generator=g)
        t.copy_(val.to(t.dtype))


def clustered_em_inputs(B: int, N: int, Ck: int, Cv: int, H: int, W: int, seed: int = 0, n_clusters: int = 24,
                        noise: float = 0.35, fg_prob: float = 0.3):
    """EM inputs with encoder-like key statistics: every pixel's key is one of `n_clusters` centres
    (entries ~ N(0, 2.3^2), so ||x|| ~ 18 at Ck = 64) plus small noise, centres assigned in spatial patches; values follow
    the same patches (centre + noise, std ~ 1.8).  Assignments of such keys are near-hard and well separated, which makes
    the multi-iteration EM well conditioned (unlike i.i.d. Gaussian keys, where rounding noise is amplified ~100x per
    iteration and even the fp32 reference sits 1e-2 .. 1e-1 from exact arithmetic)."""
    g = torch.Generator().manual_seed(seed)
    ck_centres = torch.randn(B, n_clusters, Ck, generator=g) * 2.3
    cv_centres = torch.randn(B, N, n_clusters, Cv, generator=g) * 1.6
    ph, pw = max(1, H // 6), max(1, W // 6)
    coarse = torch.randint(0, n_clusters, (B, (H + ph - 1) // ph, (W + pw - 1) // pw), generator=g)
    assign = coarse.repeat_interleave(ph, 1).repeat_interleave(pw, 2)[:, :H, :W]            # (B,H,W)
    idx = assign.reshape(B, H * W)
    x = torch.gather(ck_centres, 1, idx[:, :, None].expand(-1, -1, Ck)).transpose(1, 2).reshape(B, Ck, H, W)
    x = x + noise * torch.randn(B, Ck, H, W, generator=g)
    v = torch.gather(cv_centres, 2, idx[:, None, :, None].expand(-1, N, -1, Cv)).transpose(2, 3).reshape(B, N, Cv, H, W)
    v = v + 0.8 * torch.randn(B, N, Cv, H, W, generator=g)
    fg = (torch.rand(B, N, (H + ph - 1) // ph, (W + pw - 1) // pw, generator=g) < fg_prob).float()
    fg = fg.repeat_interleave(ph, 2).repeat_interleave(pw, 3)[:, :, :H, :W]
    soft = (fg * 0.9 + 0.05)                                                               # soft masks, like a decoder's
    masks = torch.stack([(1 - fg) * (1 - soft), fg * soft], dim=2)
    return x, v, masks
