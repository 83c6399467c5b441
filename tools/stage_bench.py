#!/usr/bin/env python
"""Stage-kernel micro-benchmark (stages.cu): the NHWC glue kernels of FrameEngine at the DAVIS-17 shape
(480x864, 5 objects), CUDA-event timing with a 256 MB L2 flush before every call, against the HBM roofline:
achieved GB/s = algorithmic bytes (each input read once, each output written once) / time.

    python tools/stage_bench.py [--reps 20]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from swem_b200 import SWEM, make_config  # noqa: E402
from swem_b200.engine import FrameEngine  # noqa: E402


def timed(fn, reps, flush):
    fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        flush.add_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / reps * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--reps', type=int, default=20)
    args = ap.parse_args()
    dev = torch.device('cuda', 0)
    torch.manual_seed(0)
    eng = FrameEngine(SWEM(make_config()).eval().to(dev))
    eng.refresh()
    peak = 6549.0
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')))['hbm_gbs'])
    except Exception:
        pass
    flush = torch.zeros(64 << 20, device=dev)
    cl = lambda *s: torch.randn(*s, device=dev).contiguous(memory_format=torch.channels_last)
    n = 5
    rows = []

    def row(name, fn, nbytes):
        us = timed(fn, args.reps, flush)
        rows.append((name, us, nbytes / 1e6, nbytes / us / 1e3, nbytes / us / 1e3 / peak))

    mb = lambda *ts: sum(t.numel() * 4 for t in ts)
    # decoder: 1/16 -> 1/8 (512 channels) and 1/8 -> 1/4 (256 channels)
    for (c, h, w) in ((512, 30, 54), (256, 60, 108)):
        lo_a, lo_b, skip, bias = cl(n, c, h, w), cl(n, c, h, w), cl(1, c, 2 * h, 2 * w), torch.randn(c, device=dev)
        out = 2 * n * c * 4 * h * w * 4
        row(f'upsample_add C={c} {h}x{w}->{2 * h}x{2 * w}', lambda: eng._upsample_add(lo_a, lo_b, bias, skip, n), mb(lo_a, lo_b, skip) + out)
    a, b, bias = cl(n, 256, 120, 216), cl(n, 256, 120, 216), torch.randn(256, device=dev)
    row('resblock_tail_pred C=256 120x216', lambda: eng._tail_pred(a, b, bias, n), mb(a, b) + n * 120 * 216 * 4)
    row('bias_add_act C=256 120x216', lambda: eng._add_act(a, b, None, bias, n, relu=True), 3 * mb(a))
    t = cl(n, 64, 240, 432)
    row('maxpool3x3s2 C=64 240x432', lambda: eng._maxpool(t), mb(t) * 1.25)
    y, sh, gb = cl(n, 1024, 30, 54), cl(1, 1024, 30, 54), torch.randn(1024, device=dev)
    row('glu_gate C=2x512 30x54', lambda: eng._glu(y, sh, gb, n), mb(y) * 1.5 + mb(sh))
    print(f'# stage kernels at the DAVIS-17 shape (5 objects), L2 flushed before every call, HBM peak {peak:.0f} GB/s (measured)')
    print(f'# {"kernel":44s} {"us":>8s} {"MB":>8s} {"GB/s":>8s} {"frac":>6s}')
    for name, us, mbs, gbs, frac in rows:
        print(f'  {name:44s} {us:8.1f} {mbs:8.1f} {gbs:8.0f} {frac:6.3f}')


if __name__ == '__main__':
    main()
