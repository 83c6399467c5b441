#!/usr/bin/env python
"""Training-step timing (BASELINE configs[4]): 3-frame clips, batch 8 per GPU, 2 objects
(the second one sometimes empty), 384x384 crops (HW = 576), forward + backward + AdamW through the CUDA memory
(`swem_em_forward` / `swem_em_backward` / `swem_readout_forward` + autograd.py), DDP when launched under torchrun.
The frame loop restates `SWEMTrainer.one_step` (swem_trainer.py:59-108); the loss is plain cross-entropy (the
reference's bootstrapped loss is out of scope).  BN layers stay in eval mode like the reference (:37-39).

    python tools/train_step_bench.py [--steps 10] [--batch 8]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/train_step_bench.py
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from swem_b200 import SWEM, make_config  # noqa: E402
from swem_b200.synthetic import rectangle_masks, smooth_video  # noqa: E402


def clip_batch(batch, t, n_obj, size, seed):
    frames, masks = [], []
    for b in range(batch):
        frames.append(smooth_video(t, size, size, seed * 100 + b))
        m = rectangle_masks(n_obj, size, size, seed * 100 + b)
        if b % 3 == 2:                                   # padded sample: the last object is empty (video_dataset.py:334-335)
            m[:, 0] += m[:, -1]
            m[:, -1] = 0
        masks.append(m)
    return torch.cat(frames), torch.cat(masks)


def one_step(model, frames, init):
    B, T = frames.shape[:2]
    h, w = frames.shape[-2:]
    n = init.shape[1] - 1
    label = init.argmax(dim=1)
    mk16, _, s16, _, _ = model('encode_key', frames[:, 0])
    mv16 = model('encode_value', frames[:, 0], init, s16)
    model('init', mk16, mv16, init.long())
    loss = 0
    for i in range(1, T):
        qk16, qv16, s16, s8, s4 = model('encode_key', frames[:, i])
        ctx, n = model('match', qk16, qv16)
        logits, prob = model('segment', n, ctx, s8, s4, None, (h, w))
        loss = loss + F.cross_entropy(logits, label)
        if i < T - 1:
            hard = F.one_hot(prob.argmax(1), n + 1).permute(0, 3, 1, 2)
            mv16 = model('encode_value', frames[:, i], prob, s16)
            model('memorize', qk16, mv16, hard, prob)
    return loss / (T - 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--batch', type=int, default=8)
    ap.add_argument('--size', type=int, default=384)
    args = ap.parse_args()
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    torch.manual_seed(0)
    model = SWEM(make_config(keydim=64, n_bases=128, n_iters=4, topl=64)).to(dev).train()
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.eval()
    if world > 1:
        model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], output_device=local,
                                                          broadcast_buffers=False)      # as swem_trainer.py:41-43
    opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-5)
    frames, init = clip_batch(args.batch, 3, 2, args.size, seed=1 + rank)
    frames, init = frames.to(dev), init.to(dev)
    losses = []

    def step():
        opt.zero_grad(set_to_none=True)
        loss = one_step(model, frames, init)
        loss.backward()
        opt.step()
        return loss.detach()

    for _ in range(args.warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        losses.append(step())
    t1.record()
    torch.cuda.synchronize(dev)
    ms = torch.tensor([t0.elapsed_time(t1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({'workload': f'train_step_b{args.batch}_t3_n2_{args.size}x{args.size}', 'n_gpus': world, 'steps': args.steps,
                          'ms_per_step': ms.item() / args.steps, 'clips_per_s': world * args.batch * args.steps / (ms.item() / 1e3),
                          'loss_first': float(losses[0]), 'loss_last': float(losses[-1]),
                          'memory': 'fused EM + readout kernels forward, swem_em_backward + swem_readout_backward'}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
