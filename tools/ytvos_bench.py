#!/usr/bin/env python
"""YouTube-VOS-shaped batch sharded by sequence (BASELINE configs[2]):
mixed 480x848 / 480x864 / 720x1280 sequences of 20-36 frames, 1-6 objects, some appearing mid-sequence; sequences are
assigned longest-first to ranks (swem_b200/sharding.py), every rank runs its own with no collective on the hot path,
one final gather of per-sequence results.  Aggregate frames/s = all frames / max-over-ranks device time.

    python tools/ytvos_bench.py --sequences 16                       # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \\
        tools/ytvos_bench.py --sequences 16
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from swem_b200 import SWEM, make_config  # noqa: E402
from swem_b200.engine import FrameEngine  # noqa: E402
from swem_b200.evaluator import evaluate_ytvos_seq  # noqa: E402
from swem_b200.sharding import assign_sequences, gather_results, merge_by_index  # noqa: E402
from swem_b200.synthetic import ytvos_materialise, ytvos_sequences  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--sequences', type=int, default=16)
    ap.add_argument('--max-frames', type=int, default=0, help='truncate every sequence (0 = full length)')
    args = ap.parse_args()
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    torch.manual_seed(0)
    model = SWEM(make_config(keydim=64, n_bases=128, n_iters=4, topl=64)).eval().to(dev).to(memory_format=torch.channels_last)
    engine = FrameEngine(model)
    specs = ytvos_sequences(args.sequences)
    if args.max_frames:
        for s in specs:
            s['t'] = min(s['t'], args.max_frames)
            s['late_frame'] = min(s['late_frame'], s['t'] // 2)
    costs = [s['t'] * s['n_obj'] * (s['h'] // 16) * (s['w'] // 16) for s in specs]
    mine = assign_sequences(costs, world)[rank]
    data = {i: ytvos_materialise(specs[i]) for i in mine}              # host tensors, built before the clock starts

    def run(i, max_t=None):
        frames, init = data[i]
        if max_t is not None:
            frames, init = frames[:, :max_t], init[:max_t]
        init = [None if m is None else m.to(dev) for m in init]
        torch.manual_seed(100 + i)
        preds = evaluate_ytvos_seq(engine, frames.to(dev), init, (specs[i]['h'], specs[i]['w']))
        return {'frames': len(preds), 'checksum': int(sum(int(p.sum()) for p in preds))}

    # cuDNN times its candidate algorithms once per conv shape (frame size x object count): the warm-up below meets every
    # shape of this rank's sequences (first frames + the frames around a late object) before the clock starts
    torch.backends.cudnn.benchmark = os.environ.get('SWEM_CUDNN_BENCHMARK', '1') == '1'
    with torch.no_grad():
        for i in mine:                                                 # warm-up: cuDNN algorithms, allocator, lazy module load
            run(i, max_t=min(specs[i]['t'], specs[i]['late_frame'] + 3))
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        results = {i: run(i) for i in mine}
        t1.record()
        torch.cuda.synchronize(dev)
    ms = torch.tensor([t0.elapsed_time(t1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    merged = merge_by_index(gather_results(results))
    if rank == 0:
        assert sorted(merged) == list(range(len(specs)))
        frames = sum(r['frames'] for r in merged.values())
        print(json.dumps({'workload': f'ytvos_synthetic_{len(specs)}seq', 'n_gpus': world, 'frames': frames,
                          'ms': ms.item(), 'frames_per_s': frames / (ms.item() / 1e3),
                          'objects_per_seq': [s['n_obj'] for s in specs],
                          'sizes': sorted({(s['h'], s['w']) for s in specs}), 'scaling': 'strong (fixed batch of sequences)',
                          'checksum': sum(r['checksum'] for r in merged.values())}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
