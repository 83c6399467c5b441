"""Pairwise per-frame disagreement of the four runners (deterministic generic kernels, fp32 convs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from swem_b200 import SWEM, make_config, _lib
from swem_b200.engine import FrameEngine
from swem_b200.evaluator import GraphedSequenceRunner, PipelinedSequenceRunner, SequenceRunner
from swem_b200.synthetic import davis_sequence
DEV = 'cuda:0'
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.manual_seed(0)
model = SWEM(make_config(keydim=64, n_bases=128, n_iters=4, topl=64)).eval().to(DEV)
fam = sys.argv[1] if len(sys.argv) > 1 else 'generic'
model.swem_core.em_path = model.swem_core.readout_path = _lib.PATH_GENERIC if fam == 'generic' else _lib.PATH_AUTO
eng = FrameEngine(model)
T, N, h, w = 8, 5, 480, 864
frames, init = davis_sequence(T, N, seed=2, size=(h, w))
frames, init = frames.to(DEV), init.to(DEV)
def run(kind):
    torch.manual_seed(5)
    model.swem_core.static_banks = False
    if kind in ('seq', 'graph'):
        r = (SequenceRunner if kind == 'seq' else GraphedSequenceRunner)(eng, (h, w))
        r.start(frames[:, 0], init)
        out = [r.step(frames[:, i]).clone() for i in range(1, T)]
    else:
        r = PipelinedSequenceRunner(eng, (h, w), use_graph=(kind == 'pipe_graph'))
        r.start(frames[:, 0], init); r.prime(frames[:, 1])
        out = [r.step(frames[:, i + 1] if i + 1 < T else None).clone() for i in range(1, T)]
    model.swem_core.static_banks = False
    return torch.stack(out).cpu()
with torch.no_grad():
    res = {k: run(k) for k in ('seq', 'seq2', 'graph', 'pipe_eager', 'pipe_graph')} if False else None
    res = {}
    for k in ('seq', 'graph', 'pipe_eager', 'pipe_graph'):
        res[k] = run(k)
    res['seq_again'] = run('seq')
for k in ('seq_again', 'graph', 'pipe_eager', 'pipe_graph'):
    d = 1 - (res['seq'] == res[k]).flatten(1).float().mean(dim=1)
    print(f'{fam} seq vs {k:11s}: ' + ' '.join(f'{x:.1e}' for x in d.tolist()))
d = 1 - (res['graph'] == res['pipe_graph']).flatten(1).float().mean(dim=1)
print(f'{fam} graph vs pipe_graph: ' + ' '.join(f'{x:.1e}' for x in d.tolist()))
