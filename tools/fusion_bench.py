"""GLU fusion layer at the BASELINE shape (5 objects, 30 x 54, [mem_out | S] = 640 channels -> 2 x 512): swem_fusion_conv_glu (operand
images + tcgen05 implicit GEMM with the gate in the epilogue) against the engine's cuDNN paths -- parity mode (TF32 main + bf16
cross-term convolutions + swem_glu_gate) and plain TF32 (one convolution + swem_glu_gate).  CUDA events, cuDNN autotuned.

    python tools/fusion_bench.py [--reps 50] [--objects 5]
"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from swem_b200 import SWEM, make_config
from swem_b200.engine import FrameEngine


def timed(fn, reps):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--reps', type=int, default=50)
    ap.add_argument('--objects', type=int, default=5)
    args = ap.parse_args()
    dev = torch.device('cuda:0')
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(0)
    model = SWEM(make_config()).eval().to(dev)
    n, h, w = args.objects, 30, 54
    cv, tl = model.swem_core.valdim, model.swem_core.topl
    feats = torch.randn(n, cv + 2 * tl, h, w, device=dev).contiguous(memory_format=torch.channels_last)
    g = torch.randn(1, 2 * cv, h, w, device=dev).contiguous(memory_format=torch.channels_last)
    flops = 2.0 * n * h * w * (cv + 2 * tl) * 9 * 2 * cv
    with torch.no_grad():
        rows = []
        ref = None
        for name, kw in (('tcgen05 kernel (fp16 hi/lo x 3, fp32-accurate)', dict(split_tf32=True, cross_bf16=True, fusion_kernel=True)),
                         ('cuDNN TF32 main + bf16 cross convs + glu_gate (parity mode)', dict(split_tf32=True, cross_bf16=True, fusion_kernel=False)),
                         ('cuDNN TF32 conv + glu_gate (fails the mask gate)', dict(split_tf32=False, fusion_kernel=False))):
            eng = FrameEngine(model, **kw)
            eng.refresh()
            if kw.get('fusion_kernel'):
                fn = lambda: eng._fusion_conv_glu(feats, g, n, n, h, w)
            else:
                fn = lambda: eng._glu(eng._conv(feats, eng.g_obj), g, eng.g_bias, n)
            out = fn()
            if ref is None:
                x64 = feats.double()
                y = torch.nn.functional.conv2d(x64, eng.g_obj[0].double(), None, padding=1) + g.double() + eng.g_bias.double().view(1, -1, 1, 1)
                ref = y[:, :cv] * torch.sigmoid(y[:, cv:])
            err = float((out.double() - ref).abs().max() / ref.abs().max())
            t = timed(fn, args.reps)
            rows.append((name, t, err))
        print(f'# GLU fusion layer, {n} objects x {h} x {w}, {cv + 2 * tl} -> 2 x {cv} channels, 3 x 3: {flops / 1e9:.1f} GFLOP algorithmic')
        for name, t, err in rows:
            print(f'{name:62s} {t:8.1f} us   {flops / t / 1e6:7.1f} TFLOP/s algorithmic   max-rel err vs fp64 {err:.2e}')


if __name__ == '__main__':
    main()
