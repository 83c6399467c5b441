#!/usr/bin/env python
"""EM / readout kernel micro-benchmark sweep (BASELINE configs[3]):
HW in {1620, 6480} x L in {64, 128, 256, 512} x EM iterations 1..4, Ck = 64, Cv = 512, N = 5, B = 1.

Inputs follow the measured random-init statistics (synthetic.em_inputs); the prior is the output of one previous
call.  CUDA-event timing over `--reps` back-to-back calls, once with the working set left in L2 and once with a
256 MB L2 flush before every call (timed per call with event pairs).  Reports which kernel family ran (AUTO dispatch:
fused tcgen05 where the shape is covered, generic fp32 otherwise), us per call and the fraction of the measured
tensor peak for the algorithmic FLOPs F_mem = 4 N HW L [Ck (3I - 1) + Cv], F_read = 4 N HW Lt (Ck + Cv).

    python tools/microbench.py [--reps 100] [--quick] [--eager]

--eager adds the reference algorithm in eager PyTorch on the same device (the oracle core on CUDA tensors: what the unmodified
modules.py does on a GPU) as two more columns -- the bar these kernels have to beat (SURVEY top of file).  Measurement tool:
imports the oracle, like tools/precision_study.py; the product package never does.
"""
import argparse
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from swem_b200 import SWEMCore, _lib  # noqa: E402
from swem_b200.synthetic import em_inputs  # noqa: E402

CK, CV, N = 64, 512, 5      # CK is overridden by --ck


def peak_tflops():
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')
    return json.load(open(p))['bf16_tflops_sustained'] if os.path.isfile(p) else 1400.0


def timed(fn, reps, flush=None):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    if flush is None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e3
    tot = 0.0
    for _ in range(reps):
        flush.add_(1.0)                                   # 256 MB read-modify-write > 126 MB L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--reps', type=int, default=100)
    ap.add_argument('--quick', action='store_true', help='HW=1620 only, I in {1,4}')
    ap.add_argument('--ck', type=int, default=64, help='key channels (64 = BASELINE, 128 = reference CLI default)')
    ap.add_argument('--eager', action='store_true', help='also time the oracle core on CUDA tensors (eager PyTorch reference)')
    args = ap.parse_args()
    global CK
    CK = args.ck
    dev = torch.device('cuda:0')
    lib = _lib.load()
    peak = peak_tflops()
    flush = torch.zeros(64 << 20, device=dev)
    print(f'# EM / readout micro-benchmark, B=1 N={N} Ck={CK} Cv={CV}, {args.reps} reps, tensor peak {peak:.0f} TFLOP/s (measured bf16 sustained)')
    print('#   HW     L  I  family(em/read)     em_us  em_us(L2 flushed)  em_TF/s  em_frac   read_us  read_us(flushed)  read_TF/s  read_frac' + ('  eager_em_us  eager_read_us  speedup(em/read)' if args.eager else ''))
    shapes = [(30, 54)] if args.quick else [(30, 54), (60, 108)]
    for (H, W) in shapes:
        HW = H * W
        x, v, masks = (t.to(dev) for t in em_inputs(1, N, CK, CV, H, W, seed=0))
        for L in (64, 128, 256, 512):
            for I in ((1, 4) if args.quick else (1, 2, 3, 4)):
                core = SWEMCore(n_bases=L, valdim=CV, n_iters=I, tau=0.05, topl=64).to(dev).eval()
                dims = _lib.SwemDims(1, N, CK, CV, HW, L, I, 2, min(L, 64), 0.05)
                fam = ('fused' if lib.swem_em_fused_supported(C.byref(dims)) else 'generic',
                       'fused' if lib.swem_readout_fused_supported(C.byref(dims)) else 'generic')
                with torch.no_grad():
                    torch.manual_seed(1)
                    core.memorize(x, v, masks)
                    core.memorize(x, v, masks)
                    prior = core.memories['update'].bases
                    em = lambda: core.swem(x, v, masks, prior)
                    rd = lambda: core.matching_features(x, v[:, 0])
                    t_em, t_em_f = timed(em, args.reps), timed(em, max(10, args.reps // 4), flush)
                    feats = torch.empty(N, 2 * CV + 2 * min(L, 64), H, W, device=dev).contiguous(memory_format=torch.channels_last)
                    rd = lambda: core.readout_into(x, feats, 0, 2 * CV)          # kernels only, the engine's pixel-major layout
                    t_rd, t_rd_f = (timed(rd, args.reps), timed(rd, max(10, args.reps // 4), flush)) if I == 4 or args.quick else (float('nan'),) * 2
                    t_eem = t_erd = float('nan')
                    if args.eager:
                        from oracle import swem_oracle as O
                        ref = O.OracleSWEMCore(n_bases=L, valdim=CV, n_iters=I, tau=0.05, topl=64)
                        ref.banks.first = {k: t.clone() for k, t in core.memories['first'].bases.items()}
                        ref.banks.update = {k: t.clone() for k, t in prior.items()}
                        ref.banks.first_n = N
                        eprior = {k: t.reshape(1, N, 2, -1, L) if k == 'zita' else t for k, t in prior.items()}
                        t_eem = timed(lambda: O.em_memorize(x, v, masks, eprior, L, I, 0.05), max(5, args.reps // 10))
                        if I == 4 or args.quick:
                            t_erd = timed(lambda: ref.matching_features(x, v[:, 0]), max(5, args.reps // 10))
                f_mem = 4 * N * HW * L * (CK * (3 * I - 1) + CV)
                f_read = 4 * N * HW * 2 * L * (CK + CV)
                tf_em, tf_rd = f_mem / t_em / 1e6, f_read / t_rd / 1e6
                print(f'  {HW:5d} {L:5d} {I:2d}  {fam[0] + "/" + fam[1]:16s} {t_em:9.1f} {t_em_f:14.1f} {tf_em:12.1f} {tf_em / peak:8.4f} '
                      f'{t_rd:9.1f} {t_rd_f:14.1f} {tf_rd:12.1f} {tf_rd / peak:9.4f}'
                      + (f' {t_eem:12.1f} {t_erd:14.1f}   {t_eem / t_em:6.1f}x / {t_erd / t_rd:5.1f}x' if args.eager else ''), flush=True)
                del core


if __name__ == '__main__':
    main()
