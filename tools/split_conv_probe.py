"""Accuracy of the conv forms on a B200 against fp64: cuDNN IEEE fp32, TF32, three TF32 convs over hi/lo splits summed in
fp32, and main + cross-term convs (FrameEngine(split_tf32=True)), and ONE TF32 conv over [hi|hi|lo] x [wh;wl;wh]; heuristic and autotuned algorithms."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from swem_b200.engine import FrameEngine

dev = 'cuda:0'
sp = FrameEngine._tf32_split
cl = lambda t: t.contiguous(memory_format=torch.channels_last)
rel = lambda a, b: ((a.double() - b).abs().max() / b.abs().max()).item()
g = torch.Generator().manual_seed(0)
for (n, ci, co, h, w, k) in ((5, 256, 256, 120, 216, 3), (1, 1024, 256, 30, 54, 1), (5, 64, 64, 120, 216, 3), (5, 640, 1024, 30, 54, 3)):
    x = cl(torch.relu(torch.randn(n, ci, h, w, generator=g)).to(dev))
    wt = cl((torch.randn(co, ci, k, k, generator=g) / (ci * k * k) ** 0.5).to(dev))
    torch.backends.cudnn.allow_tf32 = False
    want = F.conv2d(x[:1].double(), wt.double(), padding=k // 2)
    for bench in (False, True):
        torch.backends.cudnn.benchmark = bench
        torch.backends.cudnn.allow_tf32 = False
        e_fp32 = rel(F.conv2d(x, wt, padding=k // 2)[:1], want)
        torch.backends.cudnn.allow_tf32 = True
        e_tf32 = rel(F.conv2d(x, wt, padding=k // 2)[:1], want)
        xh, xl = sp(x); wh, wl = sp(wt)
        y3 = F.conv2d(xh, cl(wh), padding=k // 2) + F.conv2d(xh, cl(wl), padding=k // 2) + F.conv2d(cl(xl), cl(wh), padding=k // 2)
        e_3 = rel(y3[:1], want)
        y1 = F.conv2d(cl(torch.cat([xh, xh, xl], 1)), cl(torch.cat([wh, wl, wh], 1)), padding=k // 2)
        e_1 = rel(y1[:1], want)
        y1b = F.conv2d(cl(torch.cat([xl, xh, xh], 1)), cl(torch.cat([wh, wl, wh], 1)), padding=k // 2)   # small terms first
        e_1b = rel(y1b[:1], want)
        y2 = F.conv2d(xh, cl(wh), padding=k // 2) + F.conv2d(cl(torch.cat([xh, xl], 1)), cl(torch.cat([wl, wh], 1)), padding=k // 2)
        e_2 = rel(y2[:1], want)
        yb = F.conv2d(xh, cl(wh), padding=k // 2) + F.conv2d(cl(torch.cat([x, xl], 1).bfloat16()), cl(torch.cat([wl, wh], 1)).bfloat16(), padding=k // 2).float()
        e_b = rel(yb[:1], want)
        e_hh = rel(F.conv2d(xh, cl(wh), padding=k // 2)[:1], want)
        print(f'N={n} {ci}->{co} {k}x{k} {h}x{w} autotune={int(bench)}: fp32 {e_fp32:.1e}  tf32 {e_tf32:.1e}  hi.hi only {e_hh:.1e}  '
              f'3 convs {e_3:.1e}  main + cross {e_2:.1e}  main + bf16 cross {e_b:.1e}  stacked [hi|hi|lo] {e_1:.1e}  stacked [lo|hi|hi] {e_1b:.1e}', flush=True)
