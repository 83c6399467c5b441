"""Run-to-run spread of the free-running mask agreement vs the fp32 CPU oracle (fused kernels: L2-atomic reduction order),
with the torch convolutions in fp32, in TF32 (torch's cuDNN default, what bench.py runs) and in TF32 with cuDNN autotuning."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import swem_oracle as O
from swem_b200 import SWEM, make_config, _lib
from swem_b200.engine import FrameEngine
from swem_b200.evaluator import evaluate_davis_seq
from swem_b200.synthetic import davis_sequence
DEV = 'cuda:0'
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
T, N, h, w = 8, 5, 480, 864
torch.manual_seed(0)
cfg = make_config(keydim=64, n_bases=128, n_iters=4, topl=64)
nets_cpu = SWEM(cfg).eval()
model = SWEM(cfg).eval(); model.load_state_dict(nets_cpu.state_dict()); model = model.to(DEV)
frames, init = davis_sequence(T, N, seed=1, size=(h, w))
prior = dict(zip(('kappa', 'nu', 'zita'), O.random_init(1, N, 64, 128, 512, generator=torch.Generator().manual_seed(4))))
oracle = O.OracleSWEM(nets_cpu, 128, 4, 0.05, 64)
model.swem_core.random_init = lambda size, norm_dim=-2, dtype=None, device=None: tuple(t.to(device) for t in (prior['kappa'], prior['nu'], prior['zita']))
real = O.random_init
O.random_init = lambda *a, **k: (prior['kappa'], prior['nu'], prior['zita'])
want = torch.stack(O.run_davis_sequence(oracle, frames, init, (h, w)))
O.random_init = real
fr, im = frames.to(DEV), init.to(DEV)
torch.backends.cudnn.benchmark = True
print('== split-TF32 convs (main + cross-term TF32 convs over hi/lo splits, FrameEngine(split_tf32=True)) + cudnn.benchmark', flush=True)
for label, kw in (('bf16 cross terms, fusion layer on swem_fusion_conv_glu (bench.py)', dict(cross_bf16=True, fusion_kernel=True)),
                  ('bf16 cross terms, fusion layer on cuDNN', dict(cross_bf16=True, fusion_kernel=False)),
                  ('TF32 cross terms, fusion layer on swem_fusion_conv_glu', dict(cross_bf16=False, fusion_kernel=True))):
    eng = FrameEngine(model, split_tf32=True, **kw)
    print(f'-- {label}', flush=True)
    for rep in range(3):
        with torch.no_grad():
            got, _ = evaluate_davis_seq(eng, fr, [im] + [None] * (T - 1), (h, w))
        got = torch.stack(got).cpu()
        dis = 1 - (got == want).flatten(1).float().mean(dim=1)
        print(f'engine+fused     rep {rep}: per-frame disagreement ' + ' '.join(f'{d:.1e}' for d in dis.tolist()) + f' | pooled {dis.mean():.1e} max {dis.max():.1e}', flush=True)
if '--split-only' in sys.argv:
    sys.exit(0)
torch.backends.cudnn.benchmark = False
modes = (('fp32 convs', False, False), ('fp32 convs + cudnn.benchmark', False, True),
         ('tf32 convs (torch default)', True, False), ('tf32 convs + cudnn.benchmark', True, True))
for conv_name, tf32, bench in modes:
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.benchmark = bench
    print(f'== {conv_name}', flush=True)
    for name, stages, path in (('engine+fused', FrameEngine(model), _lib.PATH_AUTO), ('modules+fused', model, _lib.PATH_AUTO),
                               ('modules+generic', model, _lib.PATH_GENERIC), ('engine+generic', FrameEngine(model), _lib.PATH_GENERIC)):
        if bench and name != 'engine+fused':
            continue
        model.swem_core.em_path = model.swem_core.readout_path = path
        for rep in range(2):
            with torch.no_grad():
                got, _ = evaluate_davis_seq(stages, fr, [im] + [None] * (T - 1), (h, w))
            got = torch.stack(got).cpu()
            dis = 1 - (got == want).flatten(1).float().mean(dim=1)
            print(f'{name:16s} rep {rep}: per-frame disagreement ' + ' '.join(f'{d:.1e}' for d in dis.tolist()) + f' | pooled {dis.mean():.1e} max {dis.max():.1e}', flush=True)
