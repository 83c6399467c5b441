"""Phase-level timing of the fused EM kernel (CTA 0 time stamps) + CUDA-event timing of both entry points."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from swem_b200 import SWEMCore, _lib
from swem_b200.synthetic import em_inputs

def main(N=5, H=30, W=54, I=4, reps=50):
    dev = torch.device('cuda:0')
    lib = _lib.load()
    Ck, Cv, L = 64, 512, 128
    x, v, masks = (t.to(dev) for t in em_inputs(1, N, Ck, Cv, H, W, seed=0))
    core = SWEMCore(n_bases=L, valdim=Cv, n_iters=I, tau=0.05, topl=64).to(dev).eval()
    with torch.no_grad():
        core.memorize(x, v, masks)
        core.memorize(x, v, masks)
        prior = core.memories['update'].bases
        for _ in range(3):
            core.swem(x, v, masks, prior)
        buf = torch.zeros(256, dtype=torch.int64, device=dev)
        _lib.check(lib.swem_set_profile_buffer(buf.data_ptr(), buf.numel() * 8), 'set_profile')
        core.swem(x, v, masks, prior)
        torch.cuda.synchronize()
        lib.swem_set_profile_buffer(None, 0)
        st = buf.cpu().tolist()
        n = st[0]
        if n < 0:                         # labelled stamps of em_res_kernel: (ns << 8) | label
            LABELS = {0: 'start', 1: 'setup', 2: 'logits GEMM', 3: 'epilogue', 4: 'M GEMM', 5: 'reduce-add issued', 6: 'fence + arrive',
                      7: 'wait tiles', 8: 'finalize', 9: 'nu GEMM own side', 10: 'nu GEMM peer side', 11: 'nu drain', 12: 'wait nu',
                      13: 'nu slice', 20: '  staged + bulk issued', 21: '  bulk reduce complete', 22: '  arrived', 23: '  totals loaded', 24: '  CTA released', 30: '  clearing issued', 35: '  set-up loads issued', 36: '  TMEM allocated', 37: '  zero buffer ready', 31: '  prior khat staged', 32: '  X staged', 33: '  V round 0 stored', 34: '  clearing complete', 25: '  W statistics sent', 26: '  W statistics received'}
            n = -n
            raw = st[1:1 + n]
            t = [r >> 8 for r in raw]
            ids = [r & 255 for r in raw]
            print(f'N={N} HW={H*W} I={I}: em_res_kernel, {n} stamps, total {(t[-1]-t[0])/1e3:.1f} us')
            for k in range(1, n):
                print(f'  {LABELS.get(ids[k], "?"):24s} {(t[k]-t[k-1])/1e3:8.2f} us')
            n = 0
        t = st[1:1 + n]
        if n:
            print(f'N={N} HW={H*W} I={I}: {n} stamps, total {(t[-1]-t[0])/1e3:.1f} us')
        names = ['setup']
        for it in range(I):
            names += [f'it{it} logits GEMM', f'it{it} epilogue', f'it{it} M GEMM']
            if it == I - 1:
                names += ['nu GEMM', 'nu drain']
            names += [f'it{it} reduce-add', f'it{it} wait tiles', f'it{it} finalize']
        names += ['nu slice', 'exit']
        for k in range(1, n):
            print(f'  {names[k-1] if k-1 < len(names) else "?":24s} {(t[k]-t[k-1])/1e3:8.2f} us')
        _lib.check(lib.swem_set_profile_buffer(buf.data_ptr(), buf.numel() * 8), 'set_profile')
        core.matching_features(x, v[:, 0])
        torch.cuda.synchronize()
        lib.swem_set_profile_buffer(None, 0)
        st = buf.cpu().tolist()
        n = st[128]
        t = st[129:129 + n]
        rn = ['setup', 'scores GEMM', 'max pass', 'exp pass', 'PV issue (+ top-l sort)', 'PV drain wait', 'store']
        print(f'readout_fused CTA0: total {(t[-1]-t[0])/1e3:.1f} us')
        for k in range(1, n):
            print(f'  {rn[k-1] if k-1 < len(rn) else "?":24s} {(t[k]-t[k-1])/1e3:8.2f} us')
        feats_cl = torch.empty(N, 2 * Cv + 128, H, W, device=dev).contiguous(memory_format=torch.channels_last)
        _lib.check(lib.swem_set_profile_buffer(buf.data_ptr(), buf.numel() * 8), 'set_profile')
        core.readout_into(x, feats_cl, 0, 2 * Cv)
        torch.cuda.synchronize()
        lib.swem_set_profile_buffer(None, 0)
        st = buf.cpu().tolist()
        n = st[128]
        t = st[129:129 + n]
        print(f'readout_fused CTA0, pixel-major output: total {(t[-1]-t[0])/1e3:.1f} us')
        for k in range(1, n):
            print(f'  {rn[k-1] if k-1 < len(rn) else "?":24s} {(t[k]-t[k-1])/1e3:8.2f} us')
        for name, fn in (('memorize(EM)', lambda: core.swem(x, v, masks, prior)),
                         ('readout', lambda: core.matching_features(x, v[:, 0])),
                         ('readout (pixel-major out)', lambda: core.readout_into(x, feats_cl, 0, 2 * Cv))):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for _ in range(5):
                fn()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            print(f'{name}: {e0.elapsed_time(e1) / reps * 1e3:.1f} us per call (events, {reps} reps, launches/call {core.launches})')

if __name__ == '__main__':
    main()
    main(N=1)
    main(N=5, I=1)
