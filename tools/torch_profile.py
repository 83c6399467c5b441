"""torch.profiler view of one eager frame step (FrameEngine): which aten ops / kernels take the time, with shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
from swem_b200 import SWEM, make_config
from swem_b200.engine import FrameEngine
from swem_b200.evaluator import SequenceRunner
from swem_b200.synthetic import davis_sequence

dev = torch.device('cuda:0')
torch.manual_seed(0)
model = SWEM(make_config(keydim=64, n_bases=128, n_iters=4, topl=64)).eval().to(dev).to(memory_format=torch.channels_last)
split = os.environ.get('SWEM_SPLIT_TF32', '0') == '1'           # parity mode: fp32-accurate convs as two TF32 convs
torch.backends.cudnn.benchmark = os.environ.get('SWEM_CUDNN_BENCHMARK', '1') == '1'
if split:
    torch.backends.cudnn.allow_tf32 = False
eng = FrameEngine(model, split_tf32=split)
frames, init = davis_sequence(8, 5, seed=1)
frames, init = frames.to(dev), init.to(dev)
runner = SequenceRunner(eng, (480, 864))
with torch.no_grad():
    runner.start(frames[:, 0], init)
    for i in range(1, 5):
        runner.step(frames[:, i])
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
        for i in range(5, 8):
            runner.step(frames[:, i])
        torch.cuda.synchronize()
rows = [e for e in prof.key_averages(group_by_input_shape=True) if e.key.startswith('aten::') and e.self_device_time_total > 0]
rows.sort(key=lambda e: -e.self_device_time_total)
print('self CUDA us per frame | calls per frame | op | input shapes')
for e in rows[:70]:
    print(f'{e.self_device_time_total / 3:10.1f} {e.count / 3:6.1f}  {e.key:34s} {str(e.input_shapes)[:150]}')
print('total self CUDA us per frame', sum(e.self_device_time_total for e in rows) / 3)

ks = [e for e in prof.key_averages() if 'swem::' in e.key]
ks.sort(key=lambda e: -e.self_device_time_total)
print('library kernels: self CUDA us per frame | calls per frame | kernel')
for e in ks:
    print(f'{e.self_device_time_total / 3:10.1f} {e.count / 3:6.1f}  {e.key[:110]}')
