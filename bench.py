#!/usr/bin/env python
"""Headline benchmark: 480p frames/sec of the SWEM per-frame loop plus the roofline of the EM + readout kernels.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

N = 1 (BASELINE.json configs[1]): one synthetic DAVIS-2017-shaped sequence (854x480 -> 864x480, 5 objects, ResNet-50 key
encoder, Ck = 64, L = 128 bases, 4 EM iterations).  A step = one frame of the reference's hot loop
(swem_evaluator.py:73-93): encode_key -> match (readout kernel) -> segment -> encode_value -> memorize (EM kernel).
N > 1 (BASELINE.json configs[2]): a fixed YouTube-VOS-shaped batch of sequences (mixed 480x848 / 480x864 / 720x1280,
1-6 objects, some appearing mid-sequence) sharded by sequence over the ranks, longest first, no collective on the data
path, one final gather (strong scaling).  One process per GPU (torchrun), timing = CUDA events bracketed by barrier +
synchronize, max over ranks.

The timed arithmetic is the one that PASSES the north-star mask gate: convolutions at fp32 accuracy on the tensor cores
(FrameEngine(split_tf32=True, cross_bf16=True)), EM / readout on the tcgen05 kernels.  The gate is measured inside this
run (`parity.min_frame_agreement`, N = 1): the same frames through the CPU oracle and through the exact timed
configuration.  torch's default TF32 convolutions (faster, fail the gate on random-init weights) are a side key.

Prints ONE JSON line (rank 0).  `value`: frames resident in HBM.  `e2e`: each step's frame is copied from pinned host
memory and its mask read back to the host inside the timed region.  `roofline`: algorithmic FLOPs of the EM kernel /
its CUDA-event time.  `cpu_baseline` / `--impl reference`: the reference algorithm's CPU port (oracle/) on host cores.
`gpu_eager_baseline`: the same reference algorithm as plain eager PyTorch on the B200.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

CFG = dict(keydim=64, valdim=512, n_bases=128, n_iters=4, tau=0.05, topl=64, single_obj=False, backbone='resnet50')
H, W = 480, 864
# dram__bytes_read.sum + dram__bytes_write.sum of one EM kernel launch at this workload, from the committed
# `ncu --set full` capture under profiles/ (see PROFILE_SOURCE); algorithmic bytes: 23.0 MB
NCU_TRAFFIC = {'fused-tcgen05': None}
PROFILE_SOURCE = 'profiles/r2z_em_res_kernel_ncu_full.txt'
_traffic_file = os.path.join(ROOT, 'profiles', 'r2_em_traffic.json')
if os.path.isfile(_traffic_file):
    NCU_TRAFFIC['fused-tcgen05'] = json.load(open(_traffic_file)).get('dram_bytes_per_launch')
METRIC = '480p frames/sec'
UNIT = 'frames/s'
GATE = 0.999
POOL = 48                    # distinct synthetic frames kept in pinned host memory; longer runs cycle through them
MIN_SECONDS = 2.0            # the K-step window is repeated until the timed windows add up to this much


def env_flag(name, default='1'):
    return os.environ.get(name, default) == '1'


def workload_config(n_obj, extra=None):
    c = {'workload': f'davis17_synthetic_{H}x{W}_{n_obj}obj', 'frame': [H, W], 'objects': n_obj,
         'key_dim': CFG['keydim'], 'value_dim': CFG['valdim'], 'bases': CFG['n_bases'], 'em_iters': CFG['n_iters'],
         'tau': CFG['tau'], 'topl': CFG['topl'], 'backbone': CFG['backbone'], 'weights': 'random-init seed 0',
         'sharding': 'one sequence per rank'}
    c.update(extra or {})
    return c


def hot_path_flops(n_obj, hw, lt):
    """Algorithmic FLOPs per frame of memorize + readout (SURVEY section 8a / BASELINE.md section 3)."""
    ck, cv, L, it = CFG['keydim'], CFG['valdim'], CFG['n_bases'], CFG['n_iters']
    f_mem = 4 * n_obj * hw * L * (ck * (3 * it - 1) + cv)
    f_read = 4 * n_obj * hw * lt * (ck + cv)
    return f_mem, f_read


def hot_path_bytes(n_obj, hw, lt):
    ck, cv, L, tl = CFG['keydim'], CFG['valdim'], CFG['n_bases'], CFG['topl']
    b_mem = 4 * (ck * hw + n_obj * cv * hw + 2 * n_obj * hw + 4 * n_obj * L * (ck + cv + 1))
    b_read = 4 * (ck * hw + 2 * n_obj * lt * (ck + cv) + n_obj * cv * hw + 2 * n_obj * tl * hw)
    return b_mem, b_read


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(tflops=d['bf16_tflops_sustained'], tflops_burst=d['bf16_tflops'], hbm=d['hbm_gbs'], source='measured')
    return dict(tflops=1400.0, tflops_burst=1590.0, hbm=6650.0, source='fallback')


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms during the timed region."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def __enter__(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix='.csv')
            os.close(fd)
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        if self.proc is None or not self.path:
            return out
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            f = [x.strip() for x in line.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        try:
            os.unlink(self.path)
        except OSError:
            pass
        return out


def build_model(device):
    from swem_b200 import SWEM, make_config
    torch.manual_seed(0)
    model = SWEM(make_config(**CFG)).eval().to(device)
    if env_flag('SWEM_CHANNELS_LAST'):          # torch-side layout choice (cuDNN picks NHWC kernels anyway)
        model = model.to(memory_format=torch.channels_last)
    return model


def make_sequence(n_frames, n_obj, seed):
    from swem_b200.synthetic import davis_sequence
    return davis_sequence(n_frames, n_obj, seed=seed, size=(H, W))


def fixed_prior(n_obj, seed=4):
    """The same initial bases for the oracle and the CUDA path (random_init draws from the device RNG, so a CPU and a GPU
    run never see the same draw otherwise)."""
    from oracle import swem_oracle as O
    return O.random_init(1, n_obj, CFG['keydim'], CFG['n_bases'], CFG['valdim'], generator=torch.Generator().manual_seed(seed))


# --------------------------------------------------------------------------------------------
# The reference algorithm as plain PyTorch (oracle/ core + the same torch networks): on host cores it is the
# cpu_baseline leg / --impl reference and the checker of the parity leg; on the device it is gpu_eager_baseline.
# --------------------------------------------------------------------------------------------
def reference_run(n_obj, steps, warmup, device='cpu', seed=1, prior=None, keep_masks=False):
    """frames/s of the reference algorithm over `steps` frames after `warmup` frames -> dict(fps, seconds, masks, ...)."""
    from oracle import swem_oracle as O
    from swem_b200 import SWEM, make_config
    from torch.nn import functional as F
    dev = torch.device(device)
    on_gpu = dev.type == 'cuda'
    if not on_gpu:
        torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    nets = SWEM(make_config(**CFG)).eval().to(dev)
    model = O.OracleSWEM(nets, CFG['n_bases'], CFG['n_iters'], CFG['tau'], CFG['topl'], CFG['valdim'])
    frames, init = make_sequence(1 + warmup + steps + 1, n_obj, seed)
    frames, init = frames.to(dev), init.to(dev)
    if prior is None:
        prior = fixed_prior(n_obj)
    prior = tuple(t.to(dev) for t in prior)
    real_init = O.random_init
    O.random_init = lambda *a, **k: prior                        # (also puts the draw on the right device)
    masks, hot = [], {'memorize': [], 'readout': []}

    def clock():
        if on_gpu:
            torch.cuda.synchronize(dev)
        return time.perf_counter()

    def timed(bucket, fn, *a):
        if not on_gpu:
            return fn(*a)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn(*a)
        e1.record()
        hot[bucket].append((e0, e1))
        return out

    try:
        with torch.no_grad():
            mk16, _, s16, _, _ = model.encode_key(frames[:, 0])
            mv16 = model.encode_value(frames[:, 0], F.interpolate(init, size=(H, W), mode='nearest'), s16)
            model.init(mk16, mv16, init)
            t0 = None
            for i in range(1, 1 + warmup + steps):
                if i == 1 + warmup:
                    for b in hot.values():
                        b.clear()
                    t0 = clock()
                qk16, qv16, s16, s8, s4 = model.encode_key(frames[:, i])
                feats, n = timed('readout', model.core.matching_features, qk16, qv16)
                ctx = model.fusion_layer(feats)
                _, pm = model.segment(n, ctx, s8, s4, (H, W))
                pred, hard = O.one_hot_from_argmax(pm)
                if keep_masks:
                    masks.append(pred[0, 0].to('cpu', torch.uint8))          # (H, W)
                soft = F.interpolate(pm, size=(H, W), mode='bilinear', align_corners=False)
                mv16 = model.encode_value(frames[:, i], soft, s16)
                em_masks = O.build_em_masks(hard, soft, *qk16.shape[-2:])
                timed('memorize', model.core.memorize, qk16, mv16, em_masks)
            dt = clock() - t0
    finally:
        O.random_init = real_init
    out = {'fps': steps / dt, 'seconds': dt, 'masks': masks, 'threads': torch.get_num_threads()}
    if on_gpu:
        out['memorize_us'] = 1e3 * statistics.mean(a.elapsed_time(b) for a, b in hot['memorize'])
        out['readout_us'] = 1e3 * statistics.mean(a.elapsed_time(b) for a, b in hot['readout'])
    return out


def run_reference(args, rank, world):
    if rank != 0:
        return
    r = reference_run(args.objects, args.steps, args.warmup)
    fps, dt, cores = r['fps'], r['seconds'], r['threads']
    sample = f'{args.steps} frames after {args.warmup} warm-up frames of the same workload, fp32, torch CPU'
    line = {'impl': 'reference', 'metric': METRIC, 'value': fps, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(args.objects, {'device': 'host CPU'}),
            'cpu_baseline': {'value': fps, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': fps, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------
class Bench:
    def __init__(self, args, rank, world, local_rank):
        import torch.distributed as dist
        from swem_b200 import _lib
        assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU fallback)'
        self.args, self.rank, self.world, self.local_rank, self.dist = args, rank, world, local_rank, dist
        torch.cuda.set_device(local_rank)
        self.dev = torch.device('cuda', local_rank)
        self.lib = _lib.load()
        _lib.check(self.lib.swem_device_check(local_rank), 'swem_device_check')
        if world > 1:
            dist.init_process_group('nccl', device_id=self.dev)
        # cuDNN picks its conv kernels by heuristic unless asked to time the candidates (once per shape, during the eager
        # warm-up frames, before the step graph is captured)
        self.cudnn_autotune = env_flag('SWEM_CUDNN_BENCHMARK')
        torch.backends.cudnn.benchmark = self.cudnn_autotune
        if 'SWEM_CUDNN_BENCH_LIMIT' in os.environ:
            torch.backends.cudnn.benchmark_limit = int(os.environ['SWEM_CUDNN_BENCH_LIMIT'])
        self.use_graph = env_flag('SWEM_CUDA_GRAPH')
        self.use_pipe = self.use_graph and env_flag('SWEM_PIPELINE')     # key encoder one frame ahead on a side stream
        self.K, self.n_obj = args.steps, args.objects
        self.Wm = max(args.warmup, 3) if self.use_graph else args.warmup  # frames 1-2 eager, frame 3 captures the graph
        self.model = build_model(self.dev)
        self.core = self.model.swem_core
        self.hw = (H // 16) * (W // 16)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize(self.dev)

    def set_conv_mode(self, mode):
        """'parity': convolutions at fp32 accuracy on the tensor cores (x = hi + lo, w = hi + lo on the TF32 grid; main term
        as a cuDNN TF32 conv, cross terms as one bf16 conv) -- the arithmetic that passes the mask gate.
        'tf32': torch's default cuDNN TF32 convolutions.  'fp32': cuDNN IEEE fp32 (no tensor cores on sm_100)."""
        from swem_b200.engine import FrameEngine
        tf32 = mode == 'tf32'
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        if not env_flag('SWEM_ENGINE'):
            return self.model
        return FrameEngine(self.model, channels_last=env_flag('SWEM_CHANNELS_LAST'), fused_conv=env_flag('SWEM_FUSED_CONV'),
                           split_tf32=mode == 'parity', cross_bf16=mode == 'parity' and env_flag('SWEM_CROSS_BF16'))

    def run_windows(self, stages, frames_pinned, init_dev, host_io, graphed, prior=None, min_seconds=MIN_SECONDS,
                    max_windows=64, keep_masks=0, seed=1234):
        """start on frame 0, Wm warm-up steps, then windows of exactly K timed steps (each bracketed by barrier +
        synchronize, CUDA events, max over ranks) until they add up to `min_seconds` -> (window ms list, clocks, masks).
        Step k segments and memorizes frame 1 + Wm + k of the (cyclic) pool; with the pipelined runner it also encodes the
        key of the following frame meanwhile."""
        from swem_b200.evaluator import FrameUploader, GraphedSequenceRunner, PipelinedSequenceRunner, SequenceRunner
        dev, K, Wm = self.dev, self.K, self.Wm
        pipe = graphed and self.use_pipe
        runner = (PipelinedSequenceRunner if pipe else GraphedSequenceRunner if graphed else SequenceRunner)(stages, (H, W))
        self.core.static_banks = False
        real_init = self.core.random_init
        if prior is not None:
            self.core.random_init = lambda size, norm_dim=-2, dtype=None, device=None: tuple(t.to(device) for t in prior)
        P = frames_pinned.shape[0]
        frame = lambda i: frames_pinned[i % P:i % P + 1]
        resident = None if host_io else frames_pinned.to(dev)
        masks = []
        try:
            torch.manual_seed(seed + self.rank)
            runner.start(frame(0).to(dev), init_dev)
            la = 1 if pipe else 0                                # frame a step consumes = the one it segments + la
            if pipe:
                runner.prime(frame(1).to(dev))
            for i in range(1, 1 + Wm):
                pred = runner.step(frame(i + la).to(dev))
                if len(masks) < keep_masks:
                    masks.append(pred[0].to('cpu', torch.uint8))
            mask_host = torch.empty(K, H, W, dtype=torch.uint8).pin_memory()
            windows, nxt = [], 1 + Wm + la
            with ClockSampler(self.local_rank) as clk:
                while len(windows) < max_windows and (not windows or sum(windows) < 1e3 * min_seconds):
                    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    self.barrier()
                    t0.record()
                    if host_io:                                  # every frame is uploaded inside the timed region (side stream:
                        up = FrameUploader((1, 3, H, W), dev)    # the next frame travels while the current one is processed)
                        up.submit(0, frame(nxt))
                    for k in range(K):
                        if host_io:
                            if k + 1 < K:
                                up.submit(k + 1, frame(nxt + k + 1))
                            pred = runner.step(up.get(k))
                            up.release(k)
                            mask_host[k].copy_(pred[0].to(torch.uint8), non_blocking=True)
                        else:
                            pred = runner.step(resident[(nxt + k) % P:(nxt + k) % P + 1])
                        if len(masks) < keep_masks:
                            masks.append(pred[0].to('cpu', torch.uint8))
                    t1.record()
                    self.barrier()
                    ms = t0.elapsed_time(t1)
                    if self.world > 1:
                        t = torch.tensor([ms], device=dev)
                        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
                        ms = t.item()
                    windows.append(ms)
                    nxt += K
                    if keep_masks and len(masks) >= keep_masks:
                        break
            return windows, clk.summary(), masks
        finally:
            self.core.random_init = real_init
            self.core.static_banks = False

    def mask_agreement(self, stages, frames_pinned, init_dev, prior, want):
        """Per-frame argmax agreement of the TIMED configuration (same engine, same runner class, graph capture and replay
        included) with the oracle's masks `want` (frames 1 .. len(want)), both started from the initial bases `prior`."""
        n = want.shape[0]
        saveK, saveW = self.K, self.Wm
        self.K = max(1, n - saveW)
        try:
            _, _, got = self.run_windows(stages, frames_pinned, init_dev, host_io=False, graphed=self.use_graph, prior=prior,
                                         min_seconds=0.0, keep_masks=n)
        finally:
            self.K, self.Wm = saveK, saveW
        got = torch.stack(got[:n])
        return (got == want).flatten(1).float().mean(dim=1)

    # ---- the YouTube-VOS-shaped batch (BASELINE configs[2]) -------------------------------------------------------
    def run_ytvos_batch(self, stages, n_seq):
        """A fixed batch of `n_seq` sequences sharded over the ranks (longest-processing-time first, swem_b200/sharding.py),
        no collective on the data path, one final all_gather_object of the per-sequence results.  Strong scaling: the batch
        does not grow with the rank count.  -> dict (rank 0) with aggregate frames/s = all frames / max-over-ranks device time."""
        from swem_b200.evaluator import evaluate_ytvos_seq
        from swem_b200.sharding import assign_sequences, gather_results, merge_by_index
        from swem_b200.synthetic import ytvos_materialise, ytvos_sequences
        dev, world, rank = self.dev, self.world, self.rank
        specs = ytvos_sequences(n_seq)
        costs = [s['t'] * (1 + s['n_obj']) * (s['h'] // 16) * (s['w'] // 16) for s in specs]
        mine = assign_sequences(costs, world)[rank]
        shapes = {}
        for i in mine:                                             # one representative per (size, object-count trajectory)
            s = specs[i]
            shapes.setdefault((s['h'], s['w'], s['n_obj'], s['n_late']), i)

        def run(i, max_t=None):
            frames, init = ytvos_materialise(specs[i])             # host tensors (synthesised outside the clock: see below)
            if max_t is not None:
                frames, init = frames[:, :max_t], init[:max_t]
            return frames, init

        def process(frames, init, i):
            init = [None if m is None else m.to(dev, non_blocking=True) for m in init]
            torch.manual_seed(100 + i)
            preds = evaluate_ytvos_seq(stages, frames.to(dev, non_blocking=True), init, (specs[i]['h'], specs[i]['w']))
            return {'frames': len(preds), 'checksum': int(sum(int(p.sum()) for p in preds)),
                    'n_obj': specs[i]['n_obj'], 'size': [specs[i]['h'], specs[i]['w']]}

        with torch.no_grad():
            # warm-up: cuDNN times its candidate algorithms once per conv shape (frame size x object count); every shape of
            # this rank's shard is met here (first frames + the frames around a late object) before the clock starts
            warm_frames = 0
            for i in shapes.values():
                f, m = run(i, max_t=min(specs[i]['t'], specs[i]['late_frame'] + 3))
                warm_frames += process(f, m, i)['frames']
            data = {i: tuple(run(i)) for i in mine}
            data = {i: (f.pin_memory(), m) for i, (f, m) in data.items()}
            self.barrier()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            launches0 = self.lib.swem_total_launch_count()
            with ClockSampler(self.local_rank) as clk:
                t0.record()
                results = {i: process(*data[i], i) for i in mine}
                t1.record()
                self.barrier()
            launches = self.lib.swem_total_launch_count() - launches0
        my_ms = t0.elapsed_time(t1)
        ms = torch.tensor([my_ms], device=dev)
        if world > 1:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        per_rank = gather_results({'rank': rank, 'ms': my_ms, 'frames': sum(r['frames'] for r in results.values()),
                                   'sequences': len(mine)})
        merged = merge_by_index(gather_results(results))
        if rank != 0:
            return None
        assert sorted(merged) == list(range(len(specs)))
        frames = sum(r['frames'] for r in merged.values())
        rank_ms = [p['ms'] for p in sorted(per_rank, key=lambda p: p['rank'])]
        return {'workload': f'ytvos_synthetic_{len(specs)}seq_mixed_480x848_480x864_720x1280_1to6obj', 'n_gpus': world,
                'sequences': len(specs), 'frames': frames, 'ms': ms.item(), 'value': frames / (ms.item() / 1e3), 'unit': UNIT,
                'scaling': 'strong (fixed batch of sequences)', 'sharding': 'whole sequences, longest-processing-time first; '
                'no data-path collective, one final all_gather_object',
                'per_rank_ms': rank_ms, 'imbalance': max(rank_ms) / (sum(rank_ms) / len(rank_ms)),
                'per_rank_frames': [p['frames'] for p in sorted(per_rank, key=lambda p: p['rank'])],
                'objects_per_seq': [s['n_obj'] for s in specs], 'late_objects': sum(s['n_late'] for s in specs),
                'frame_step': 'eager (object count and frame size change between and inside sequences)',
                'h2d': 'each sequence is uploaded from pinned host memory inside the timed region',
                'checksum': sum(r['checksum'] for r in merged.values()), 'clocks': clk.summary(),
                'warmup_frames': warm_frames, 'launches_rank0': launches}


def run_b200(args, rank, world, local_rank):
    import ctypes as C
    import swem_b200.core as core_mod
    from swem_b200 import _lib
    bn = Bench(args, rank, world, local_rank)
    dev, K, Wm, n_obj, lib = bn.dev, bn.K, bn.Wm, bn.n_obj, bn.lib
    stages = bn.set_conv_mode('parity')
    peaks = measured_peaks()
    dims = _lib.SwemDims(1, n_obj, CFG['keydim'], CFG['valdim'], bn.hw, CFG['n_bases'], CFG['n_iters'], 2, CFG['topl'], CFG['tau'])
    family = {'em': 'fused-tcgen05' if lib.swem_em_fused_supported(C.byref(dims)) else 'unsupported',
              'readout': 'fused-tcgen05' if lib.swem_readout_fused_supported(C.byref(dims)) else 'unsupported'}
    conv_note = ('FrameEngine(split_tf32=True, cross_bf16=' + os.environ.get('SWEM_CROSS_BF16', '1') + '): x = hi + lo, w = hi + lo on the '
                 'TF32 grid, conv(x, w) = cuDNN TF32 conv(xh, wh) + cross terms conv([x|xl], [wl;wh]) in bf16 -- fp32-accurate '
                 '(2e-6 .. 1e-5 of fp64 per conv, like cuDNN fp32), on the tensor cores'
                 + (', autotuned (cudnn.benchmark)' if bn.cudnn_autotune else ', heuristic algos')
                 + ('; the GLU fusion layer (per-object channels) on swem_fusion_conv_glu: tcgen05 implicit GEMM, fp16 hi/lo x 3 products, '
                    'gate in the epilogue' if os.environ.get('SWEM_FUSION_KERNEL', '1') == '1' else ''))

    if world > 1:
        # ---- BASELINE configs[2]: the sharded YouTube-VOS-shaped batch is the multi-GPU line --------------------------
        batch = bn.run_ytvos_batch(stages, args.sequences)
        # the per-rank DAVIS replica (configs[1] on every rank, weak scaling) stays as a side key for continuity with round 1
        frames, init = make_sequence(POOL, n_obj, seed=1 + rank)
        wins, clocks, _ = bn.run_windows(stages, frames[0].pin_memory(), init.to(dev), host_io=False, graphed=bn.use_graph,
                                         min_seconds=0.0)
        if rank == 0:
            line = {'metric': METRIC, 'value': batch['value'], 'unit': UNIT, 'n_gpus': world, 'steps': batch['frames'],
                    'warmup': batch['warmup_frames'], 'ms_per_step': batch['ms'] / batch['frames'] * world,
                    'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
                    'dtype': 'f32 I/O; EM/readout contractions f16 (hi+lo split on the key GEMMs), f32 accumulate', 'data': 'synthetic',
                    'config': {'workload': batch['workload'], 'sequences': batch['sequences'], 'frames': batch['frames'],
                               'objects_per_seq': batch['objects_per_seq'], 'late_objects': batch['late_objects'],
                               'sharding': batch['sharding'], 'per_rank_ms': batch['per_rank_ms'], 'imbalance': batch['imbalance'],
                               'per_rank_frames': batch['per_rank_frames'], 'frame_step': batch['frame_step'],
                               'steps_note': 'a step = one frame of the batch; steps = all frames of the batch (fixed, not --steps); warmup = '
                               'frames rank 0 ran before the clock (every frame size x object count of its shard); --steps / --warmup '
                               f'({K} / {args.warmup}) drive the davis_replicas side measurement',
                               'key_dim': CFG['keydim'], 'bases': CFG['n_bases'], 'em_iters': CFG['n_iters'], 'kernel_family': family,
                               'torch_convs': conv_note, 'l2': 'every frame reads > 230 MB of fp32 weights + activations (> 126 MB L2)'},
                    'e2e': {'value': batch['value'], 'unit': UNIT, 'h2d_bytes_per_step': 3 * 4 * 480 * 864, 'd2h_bytes_per_step': 8,
                            'note': 'the batch leg IS end to end: every sequence travels pinned host -> device inside the timed region; '
                                    'h2d bytes are per frame of the smallest size (720p frames: 11 MB); masks are reduced to a checksum'},
                    'gpu_launches': batch['launches_rank0'], 'clocks': {k: batch['clocks'][k] for k in ('sm_mhz', 'sm_max_mhz', 'reasons')},
                    'davis_replicas': {'value': world * K / (statistics.median(wins) / 1e3), 'unit': UNIT, 'scaling': 'weak',
                                       'workload': workload_config(n_obj)['workload'], 'ms_per_step': statistics.median(wins) / K},
                    'checksum': batch['checksum']}
            print(json.dumps(line), flush=True)
        bn.dist.destroy_process_group()
        return

    # ---- N = 1: BASELINE configs[1] ------------------------------------------------------------------------------------
    frames, init = make_sequence(POOL, n_obj, seed=1 + rank)
    frames_pinned, init_dev = frames[0].pin_memory(), init.to(dev)
    prior = fixed_prior(n_obj)

    # pass 1 (eager, instrumented): CUDA events on the launching stream right around the two C-ABI calls
    # (swem_em_forward / swem_readout_forward), and the library's own launch counter -> roofline
    ev = {'em': [], 'readout': []}
    plain_invoke = core_mod._invoke

    def timed_invoke(name, call):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = call()
        e1.record()
        ev.setdefault(name, []).append((e0, e1))
        return rc

    core_mod._invoke = timed_invoke
    launches0 = [None]
    orig_barrier = bn.barrier

    def marking_barrier():                      # the eager pass has one window: the first barrier marks the start of the timed frames
        orig_barrier()
        if launches0[0] is None:
            for b in ev.values():
                b.clear()
            launches0[0] = lib.swem_total_launch_count()
    bn.barrier = marking_barrier
    wins_eager, _, _ = bn.run_windows(stages, frames_pinned, init_dev, host_io=False, graphed=False, min_seconds=0.0)
    bn.barrier = orig_barrier
    core_mod._invoke = plain_invoke
    n_launch = lib.swem_total_launch_count() - launches0[0]
    em_ms = statistics.mean(a.elapsed_time(b) for a, b in ev['em'])
    read_ms = statistics.mean(a.elapsed_time(b) for a, b in ev['readout'])
    ms_eager = wins_eager[0]

    # pass 2: `value` (frames resident in HBM); pass 3: `e2e` (frame H2D + mask D2H inside the timed region)
    wins_res, clocks, _ = bn.run_windows(stages, frames_pinned, init_dev, host_io=False, graphed=bn.use_graph)
    wins_e2e, clocks_e2e, _ = bn.run_windows(stages, frames_pinned, init_dev, host_io=True, graphed=bn.use_graph)
    ms_res, ms_e2e = statistics.median(wins_res), statistics.median(wins_e2e)
    fps, fps_e2e = K / (ms_res / 1e3), K / (ms_e2e / 1e3)

    # parity leg: the exact timed configuration (same engine, same runner, graph replay included) on the frames the CPU
    # oracle sees, from the same initial bases -> per-frame argmax agreement (the north-star gate, >= 99.9 %)
    parity, cpu_baseline, tf32_side, gpu_eager = None, None, None, None
    if not args.no_cpu_baseline:
        n_par = args.cpu_steps
        ref = reference_run(n_obj, n_par, 1, device='cpu', seed=1, prior=prior, keep_masks=True)
        cpu_baseline = {'value': ref['fps'], 'unit': UNIT, 'cores': ref['threads'], 'kind': 'port',
                        'sample': f'{n_par} frames (after 1 warm-up frame) of the same workload, fp32 torch CPU, {ref["seconds"]:.1f} s'}
        want = torch.stack(ref['masks'])                            # frames 1 .. 1 + n_par
        pf = bn.mask_agreement(stages, frames_pinned, init_dev, prior, want)
        parity = {'min_frame_agreement': pf.min().item(), 'mean_frame_agreement': pf.mean().item(), 'frames': int(pf.numel()),
                  'gate': GATE, 'passes': bool(pf.min().item() >= GATE),
                  'against': 'fp32 CPU oracle (oracle/swem_oracle.py + the same torch networks), same frames, same initial bases',
                  'configuration': 'the timed one: ' + ('pipelined CUDA-graph runner' if bn.use_pipe else 'graph runner' if bn.use_graph else 'eager')
                                   + ', FrameEngine parity convolutions, tcgen05 EM / readout'}
        # side key: torch-default TF32 convolutions (what the reference itself would run on a GPU) -- faster, fails the gate
        st_tf32 = bn.set_conv_mode('tf32')
        w_t, _, _ = bn.run_windows(st_tf32, frames_pinned, init_dev, host_io=False, graphed=bn.use_graph, min_seconds=0.5)
        w_te, _, _ = bn.run_windows(st_tf32, frames_pinned, init_dev, host_io=True, graphed=bn.use_graph, min_seconds=0.5)
        pf_t = bn.mask_agreement(st_tf32, frames_pinned, init_dev, prior, want)
        tf32_side = {'value': K / (statistics.median(w_t) / 1e3), 'unit': UNIT, 'e2e': K / (statistics.median(w_te) / 1e3),
                     'min_frame_agreement': pf_t.min().item(), 'passes_gate': bool(pf_t.min().item() >= GATE),
                     'convs': 'cuDNN TF32 (torch default allow_tf32), same engine and kernels otherwise'}
        # gpu_eager_baseline: the reference algorithm as plain eager PyTorch on this GPU (oracle core on CUDA tensors + plain
        # nn.Modules), torch defaults (cudnn.allow_tf32 = True, matmul fp32) and with IEEE-fp32 convolutions
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = True, False
        ge = reference_run(n_obj, K, 3, device=str(dev), seed=1, prior=prior)
        torch.backends.cudnn.allow_tf32 = False
        ge32 = reference_run(n_obj, min(K, 6), 2, device=str(dev), seed=1, prior=prior)
        gpu_eager = {'value': ge['fps'], 'unit': UNIT, 'steps': K, 'memorize_us': ge['memorize_us'], 'readout_us': ge['readout_us'],
                     'what': 'reference algorithm in eager PyTorch on the same B200: oracle core on CUDA tensors + plain nn.Modules, fp32 '
                             'tensors, torch defaults (cudnn.allow_tf32=True, matmul.allow_tf32=False, cudnn.benchmark as this run); '
                             'memorize_us / readout_us = CUDA events around SWEMCore.memorize / the readout up to the concat',
                     'ieee_fp32_convs': {'value': ge32['fps'], 'memorize_us': ge32['memorize_us'], 'readout_us': ge32['readout_us']},
                     'kernel_speedup': {'memorize': ge['memorize_us'] / (em_ms * 1e3), 'readout': ge['readout_us'] / (read_ms * 1e3)}}
        stages = bn.set_conv_mode('parity')

    # the YouTube-VOS-shaped batch at N = 1 (the strong-scaling base point of the multi-GPU runs)
    batch = None
    if args.sequences > 0 and not args.no_batch:
        batch = bn.run_ytvos_batch(stages, args.sequences)

    f_mem, f_read = hot_path_flops(n_obj, bn.hw, 2 * CFG['n_bases'])
    b_mem, b_read = hot_path_bytes(n_obj, bn.hw, 2 * CFG['n_bases'])
    hot_s = (em_ms + read_ms) / 1e3
    achieved = f_mem / (em_ms / 1e3) / 1e12                     # dominant kernel: the EM kernel (one launch per frame)
    cfg_extra = {'kernel_family': family,
                 'frame_step': ('CUDA graph replay, key encoder of the next frame on a second stream' if bn.use_pipe
                                else 'CUDA graph replay' if bn.use_graph else 'eager'),
                 'eager_ms_per_step': ms_eager / K,
                 'l2': 'every step reads a new 5 MB frame and > 230 MB of fp32 weights + activations (> 126 MB L2); no explicit flush',
                 'torch_convs': conv_note,
                 'timed_windows': {'steps_per_window': K, 'windows': len(wins_res), 'min_ms': min(wins_res), 'median_ms': ms_res,
                                   'max_ms': max(wins_res), 'e2e_windows': len(wins_e2e), 'e2e_min_ms': min(wins_e2e),
                                   'e2e_median_ms': ms_e2e, 'clock_samples': clocks.get('samples')},
                 'torch_stages': ('FrameEngine (BN folded, fused conv+bias+relu, object-independent conv halves computed once per frame)'
                                  if env_flag('SWEM_ENGINE') else 'plain nn.Modules')}
    if parity is not None:
        cfg_extra['mask_agreement_vs_fp32_cpu_oracle'] = parity['min_frame_agreement']
    if batch is not None:
        cfg_extra['ytvos_batch_1gpu'] = {k: batch[k] for k in ('workload', 'value', 'frames', 'ms')}
    line = {
        'metric': METRIC, 'value': fps, 'unit': UNIT, 'n_gpus': world, 'steps': K, 'warmup': Wm,
        'ms_per_step': ms_res / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32 I/O; EM/readout contractions f16 (hi+lo split on the key GEMMs), f32 accumulate; convs TF32+bf16 split (fp32-accurate)',
        'data': 'synthetic',
        'config': workload_config(n_obj, cfg_extra),
        'e2e': {'value': fps_e2e, 'unit': UNIT, 'h2d_bytes_per_step': 3 * H * W * 4, 'd2h_bytes_per_step': H * W,
                'ms_per_step': ms_e2e / K, 'upload': 'pinned host -> device on a side stream, double-buffered (FrameUploader)'},
        'gpu_launches': n_launch,
        'clocks': {k: clocks[k] for k in ('sm_mhz', 'sm_max_mhz', 'reasons')},
        'roofline': {'bound': 'tensor', 'achieved': achieved, 'peak': peaks['tflops'], 'unit': 'TFLOP/s',
                     'frac': achieved / peaks['tflops'], 'traffic': NCU_TRAFFIC.get(family['em']),
                     'kernel': 'EM kernel via swem_em_forward (one launch per frame)',
                     'peak_source': peaks['source'] + ' bf16 sustained (kernel timed with CUDA events inside an eager pass of the same K frames)',
                     'algorithmic_flops_per_launch': f_mem, 'algorithmic_bytes_per_launch': b_mem,
                     'em_us': em_ms * 1e3, 'readout_us': read_ms * 1e3,
                     'readout': {'achieved': f_read / (read_ms / 1e3) / 1e12, 'frac': f_read / (read_ms / 1e3) / 1e12 / peaks['tflops'],
                                 'algorithmic_flops_per_call': f_read, 'algorithmic_bytes_per_call': b_read},
                     'hot_path': {'achieved': (f_mem + f_read) / hot_s / 1e12, 'frac': (f_mem + f_read) / hot_s / 1e12 / peaks['tflops'],
                                  'hbm_gbs': (b_mem + b_read) / hot_s / 1e9, 'hbm_frac': (b_mem + b_read) / hot_s / 1e9 / peaks['hbm'],
                                  'share_of_eager_step': (em_ms + read_ms) / (ms_eager / K)}},
    }
    if parity is not None:
        line['parity'] = parity
        line['cpu_baseline'] = cpu_baseline
        line['tf32_convs'] = tf32_side
        line['gpu_eager_baseline'] = gpu_eager
    if batch is not None:
        line['ytvos_batch'] = batch
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--objects', type=int, default=5)
    ap.add_argument('--cpu-steps', type=int, default=8, help='frames of the bounded cpu_baseline sample = frames of the parity leg')
    ap.add_argument('--sequences', type=int, default=64, help='sequences of the YouTube-VOS-shaped batch (configs[2])')
    ap.add_argument('--no-cpu-baseline', action='store_true', help='skip the CPU oracle (parity leg, cpu_baseline) and the side keys')
    ap.add_argument('--no-batch', action='store_true', help='N = 1: skip the YouTube-VOS-shaped batch')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, world, local_rank)


if __name__ == '__main__':
    main()
