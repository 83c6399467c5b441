"""Per-sequence inference loops on tensors -- the host-side mirror of the reference's
``SWEMEvaluator.evaluate_davis_seq`` / ``evaluate_davis_seq_ms`` / ``evaluate_ytvos_seq``
(methods/SWEM/swem_evaluator.py:34-148).

Dataset loading, PNG dumping and J&F scoring of the reference's ``BasicEvaluator`` are out of
scope; these loops take frames and first-frame masks as tensors
and return the per-frame argmax masks, calling the model only through ``model(mode, ...)``.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn.functional as F


def _to_frame_size(pred_mask: torch.Tensor, h: int, w: int) -> torch.Tensor:
    """The reference always calls F.interpolate(pred_mask, (h, w), bilinear, align_corners=False)
    (swem_evaluator.py:91); for equal sizes that map is the identity (source index = destination index,
    weight 1), so the launch is skipped -- bit-identical result."""
    if pred_mask.shape[-2:] == (h, w):
        return pred_mask
    return F.interpolate(pred_mask, size=(h, w), mode='bilinear', align_corners=False)


def hard_masks_from_scores(pred_mask: torch.Tensor):
    """(B,N+1,H,W) scores -> argmax (B,1,H,W) and its int64 one-hot (B,N+1,H,W).  Probabilities that come out of the fused decode
    tail (``SWEM.decode_from_lowres`` -> ``swem_decode_tail_masks``) carry both already."""
    fused = getattr(pred_mask, 'swem_hard_masks', None)
    if fused is not None:
        return fused
    pred = torch.argmax(pred_mask, dim=1, keepdim=True)
    idx = torch.arange(pred_mask.shape[1], dtype=pred.dtype, device=pred.device).view(1, -1, 1, 1)
    return pred, (pred == idx).type_as(pred)


@torch.no_grad()
def evaluate_davis_seq(model, frames: torch.Tensor, init_masks: List[Optional[torch.Tensor]], out_size,
                       on_frame=None):
    """frames (B,T,3,h,w); init_masks[0] (B,N+1,H,W) float one-hot.  Returns (preds, pred_scores):
    lists of T-1 tensors (B,H,W) int64 and (B,N+1,H,W) float.  ``on_frame(i, pred)`` is an optional
    hook called after frame i's mask is available (used by bench.py for the D2H read)."""
    b, t, c, h, w = frames.shape
    mk16, _, s16, _, _ = model('encode_key', frames[:, 0])
    init_mask = F.interpolate(init_masks[0], size=(h, w), mode='nearest')
    mv16 = model('encode_value', frames[:, 0], init_mask.float(), s16)
    model('init', mk16, mv16, init_masks[0])
    preds, scores = [], []
    for i in range(1, t):
        qk16, qv16, s16, s8, s4 = model('encode_key', frames[:, i])
        context, n = model('match', qk16, qv16)
        _, pred_mask = model('segment', n, context, s8, s4, None, out_size)
        scores.append(pred_mask.clone())
        pred, hard = hard_masks_from_scores(pred_mask)
        if i < t - 1:
            soft = _to_frame_size(pred_mask, h, w)
            mv16 = model('encode_value', frames[:, i], soft, s16)
            model('memorize', qk16, mv16, hard, soft)
        preds.append(pred[:, 0])
        if on_frame is not None:
            on_frame(i, pred[:, 0])
    return preds, scores


@torch.no_grad()
def evaluate_davis_seq_ms(model, frames: torch.Tensor, init_masks: List[Optional[torch.Tensor]], out_size,
                          scales=(480,), is_flip: bool = False):
    """Multi-scale / horizontal-flip test-time augmentation (``SWEMEvaluator.evaluate_davis_seq_ms``,
    swem_evaluator.py:34-57): the sequence is re-run at every scale (bicubic resize to (s, s/480*864)) and, with
    ``is_flip``, once more mirrored; the per-frame scores are averaged (flip pair first, then over scales) and the
    argmax of the average is returned as a list of T-1 tensors (B,H_o,W_o).  Scale 960 gives HW = 6480 key pixels."""
    assert len(scales) > 0
    total = [0 for _ in range(frames.shape[1] - 1)]
    for scale in scales:
        h, w = scale, int((scale / 480) * 864)
        resized = F.interpolate(frames[0], size=(h, w), mode='bicubic', align_corners=False).unsqueeze(0)
        _, scores = evaluate_davis_seq(model, resized, init_masks, out_size)
        if is_flip:
            masks_flip = [torch.flip(m, dims=[-1]) for m in init_masks if m is not None]
            _, scores_flip = evaluate_davis_seq(model, torch.flip(resized, dims=[-1]), masks_flip, out_size)
            scores = [(a + torch.flip(b, dims=[-1])) / 2 for a, b in zip(scores, scores_flip)]
        total = [acc + sc / len(scales) for acc, sc in zip(total, scores)]
    return [torch.argmax(sc, dim=1, keepdim=False) for sc in total]


@torch.no_grad()
def evaluate_ytvos_seq(model, frames: torch.Tensor, init_masks: List[Optional[torch.Tensor]], out_size):
    """Like :func:`evaluate_davis_seq`, but ``init_masks[i]`` (None or (B,N'+1,H,W)) introduces the
    objects that first appear in frame i; their scores are spliced in before the argmax and the
    memory grows by N' objects at the next memorize."""
    b, t, c, h, w = frames.shape
    mk16, _, s16, _, _ = model('encode_key', frames[:, 0])
    init_mask = F.interpolate(init_masks[0], size=(h, w), mode='nearest')
    mv16 = model('encode_value', frames[:, 0], init_mask.float(), s16)
    model('init', mk16, mv16, init_masks[0])
    preds = []
    for i in range(1, t):
        qk16, qv16, s16, s8, s4 = model('encode_key', frames[:, i])
        context, n = model('match', qk16, qv16)
        _, pred_mask = model('segment', n, context, s8, s4, None, out_size)
        if init_masks[i] is not None:
            fresh = init_masks[i][:, 1:]
            taken = fresh.sum(dim=1, keepdim=True).expand_as(pred_mask)
            pred_mask[taken > 0] = 0
            pred_mask = torch.cat([pred_mask, fresh], dim=1)
        pred, hard = hard_masks_from_scores(pred_mask)
        if i < t - 1:
            soft = _to_frame_size(pred_mask, h, w)
            mv16 = model('encode_value', frames[:, i], soft, s16)
            model('memorize', qk16, mv16, hard, soft)
        preds.append(pred[:, 0])
    return preds


class SequenceRunner:
    """Frame-at-a-time form of :func:`evaluate_davis_seq` (same calls in the same order), for callers
    that stream frames -- bench.py uploads one frame per step and reads one mask back per step."""

    def __init__(self, model, out_size):
        self.model, self.out_size = model, out_size

    @torch.no_grad()
    def start(self, frame0: torch.Tensor, init_mask: torch.Tensor) -> None:
        """frame0 (B,3,h,w); init_mask (B,N+1,H,W) float one-hot: encode + 'init' (frame 0 of the loop)."""
        h, w = frame0.shape[-2:]
        mk16, _, s16, _, _ = self.model('encode_key', frame0)
        m0 = F.interpolate(init_mask, size=(h, w), mode='nearest')
        mv16 = self.model('encode_value', frame0, m0.float(), s16)
        self.model('init', mk16, mv16, init_mask)

    @torch.no_grad()
    def step(self, frame: torch.Tensor, memorize: bool = True) -> torch.Tensor:
        """One frame: encode_key -> match -> segment -> (encode_value -> memorize).  Returns (B,H,W) int64."""
        h, w = frame.shape[-2:]
        qk16, qv16, s16, s8, s4 = self.model('encode_key', frame)
        context, n = self.model('match', qk16, qv16)
        _, pred_mask = self.model('segment', n, context, s8, s4, None, self.out_size)
        pred, hard = hard_masks_from_scores(pred_mask)
        if memorize:
            soft = _to_frame_size(pred_mask, h, w)
            mv16 = self.model('encode_value', frame, soft, s16)
            self.model('memorize', qk16, mv16, hard, soft)
        return pred[:, 0]


def _graph_key(core, frame, memorize: bool):
    """What a captured frame step is specific to: the memorize flag, the frame shape, the object count -- and both banks
    must exist (the 'update' bank is created by the second memorize; a graph captured before that would bake in tensors
    that are then replaced).  None = not capturable yet."""
    first, upd = core.memories['first'].bases, core.memories['update'].bases
    if first is None or upd is None:
        return None
    return (bool(memorize), tuple(frame.shape), int(first['kappa'].shape[1]), int(upd['kappa'].shape[1]))


class GraphedSequenceRunner(SequenceRunner):
    """:class:`SequenceRunner` whose steady-state step is captured once into a CUDA graph and replayed.

    A frame of the loop is ~450 kernel launches (torch encoders/decoder + our kernels); replaying them as
    one graph removes the CPU launch cost.  The first ``eager_steps`` frames run eagerly (the memory reaches
    its steady shape -- both banks present -- after the second memorize), then one step is captured with a
    static input frame; the 'update' bank is switched to in-place updates (``SWEMCore.static_banks``) so the
    addresses baked into the graph stay valid.  ``step`` returns a static tensor that the next step
    overwrites.  Sequences whose object count changes must use the eager runner.
    """

    def __init__(self, model, out_size, eager_steps: int = 2):
        super().__init__(model, out_size)
        self.eager_steps = max(2, eager_steps)
        self._seen = 0
        self._graph = None
        self._key = None
        self._frame = None
        self._pred = None

    @torch.no_grad()
    def start(self, frame0, init_mask):
        self._seen, self._graph, self._key = 0, None, None
        self.model.swem_core.static_banks = False
        super().start(frame0, init_mask)

    @torch.no_grad()
    def step(self, frame, memorize: bool = True):
        key = _graph_key(self.model.swem_core, frame, memorize)
        if self._graph is not None and key != self._key:         # another step than the captured one: drop the graph
            self._graph, self._key = None, None
            self.model.swem_core.static_banks = False
        if self._graph is None:
            if self._seen < self.eager_steps or key is None:     # (key None: a bank is still missing -> not capturable)
                self._seen += 1
                return super().step(frame, memorize)
            core = self.model.swem_core                          # (FrameEngine forwards .swem_core to its model)
            core.static_banks = True
            self._frame = frame.clone()
            torch.cuda.synchronize(frame.device)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self._pred = super().step(self._frame, memorize)
            self._graph, self._key = graph, key
            # capture does not execute: run the captured step once for this frame
        self._frame.copy_(frame, non_blocking=True)
        self._graph.replay()
        return self._pred


class FrameUploader:
    """Double-buffered host -> device upload of frames on a side stream, so that the copy of frame k+1 from pinned host
    memory overlaps the processing of frame k (what a streaming caller of the runners does; bench.py's end-to-end leg).

        up = FrameUploader((1, 3, h, w), device)
        up.submit(0, host[0])
        for k in range(n):
            if k + 1 < n: up.submit(k + 1, host[k + 1])
            mask = runner.step(up.get(k))
            up.release(k)
    """

    def __init__(self, shape, device, depth: int = 2):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(self.device)
        self.bufs = [torch.empty(shape, device=self.device) for _ in range(depth)]
        self.ready = [torch.cuda.Event() for _ in range(depth)]
        self.free = [None] * depth
        self.stream.wait_stream(torch.cuda.current_stream(self.device))     # nothing is copied before the caller's "now"

    def submit(self, k: int, host_frame: torch.Tensor) -> None:
        i = k % len(self.bufs)
        with torch.cuda.stream(self.stream):
            if self.free[i] is not None:
                self.stream.wait_event(self.free[i])                        # the consumer of the previous occupant is done
            self.bufs[i].copy_(host_frame, non_blocking=True)
            self.ready[i].record(self.stream)

    def get(self, k: int) -> torch.Tensor:
        i = k % len(self.bufs)
        torch.cuda.current_stream(self.device).wait_event(self.ready[i])
        return self.bufs[i]

    def release(self, k: int) -> None:
        i = k % len(self.bufs)
        self.free[i] = torch.cuda.Event()
        self.free[i].record(torch.cuda.current_stream(self.device))


class PipelinedSequenceRunner(SequenceRunner):
    """Frame loop with the key encoder running one frame ahead on a second stream.

    In the reference loop (swem_evaluator.py:73-93) ``encode_key(frame i+1)`` depends on nothing computed for frame i,
    but it is ~60 small single-image kernels (a ResNet-50 at batch 1) that leave most of the GPU idle when they run on
    their own.  Here step i runs two branches that join at its end:

        main stream : match -> segment -> encode_value -> memorize          of frame i   (features from the previous step)
        side stream : encode_key                                             of frame i+1

    so the small kernels fill the gaps of the large ones; with a ``FrameEngine`` the side branch also runs the frame's
    object-independent convolutions (``FrameEngine.shared_parts``).  Results are those of :class:`SequenceRunner` (same
    calls, same order per frame); only the schedule differs.  With ``use_graph`` the whole two-branch step is captured once into a
    CUDA graph (fork / join inside the capture) and replayed, like :class:`GraphedSequenceRunner`.

        runner.start(frame0, init_mask); runner.prime(frame1)
        for i in range(1, T):
            mask_i = runner.step(frames[i + 1] if i + 1 < T else None)     # None: no look-ahead (last frame)
    """

    def __init__(self, model, out_size, use_graph: bool = True, eager_steps: int = 2):
        super().__init__(model, out_size)
        self.use_graph, self.eager_steps = use_graph, max(2, eager_steps)
        self._side = None
        self._reset()

    def _reset(self):
        self._seen, self._graph, self._cur, self._cur_frame, self._next_in, self._pred = 0, None, None, None, None, None
        self._cur_shared = None
        self._key = None

    def _encode(self, frame):
        """Key features of a frame + (FrameEngine only) its object-independent convolutions, as a flat tensor list."""
        feats = list(self.model('encode_key', frame))
        if hasattr(self.model, 'shared_parts'):
            sh = self.model.shared_parts(feats[1], feats[2], feats[3], feats[4])
            feats += [sh['g'], sh['f_c1']] + list(sh['skip']) + ([sh['f_dn']] if 'f_dn' in sh else [])
        return feats

    def _shared_kw(self):
        ex = self._cur[5:]                                   # [g, f_c1, skip8, skip4 (, f_dn)]
        if not ex:
            return {}
        sh = {'g': ex[0], 'f_c1': ex[1], 'skip': ex[2:4]}
        if len(ex) > 4:
            sh['f_dn'] = ex[4]
        return {'shared': sh}

    @torch.no_grad()
    def start(self, frame0, init_mask):
        self._reset()
        self.model.swem_core.static_banks = False
        if self._side is None:
            self._side = torch.cuda.Stream(frame0.device)
        super().start(frame0, init_mask)

    @torch.no_grad()
    def prime(self, frame):
        """Encode the first frame to be segmented (what the side branch does for every later frame)."""
        self._cur = [t.clone(memory_format=torch.preserve_format) for t in self._encode(frame)]
        self._cur_frame = frame.clone()

    def _heavy(self, memorize: bool):
        qk16, qv16, s16, s8, s4 = self._cur[:5]
        kw = self._shared_kw()
        h, w = self._cur_frame.shape[-2:]
        context, n = self.model('match', qk16, qv16, **kw)
        _, pred_mask = self.model('segment', n, context, s8, s4, None, self.out_size, **kw)
        pred, hard = hard_masks_from_scores(pred_mask)
        if memorize:
            soft = _to_frame_size(pred_mask, h, w)
            mv16 = self.model('encode_value', self._cur_frame, soft, s16, **kw)
            self.model('memorize', qk16, mv16, hard, soft)
        return pred[:, 0]

    def _two_branch_step(self, next_frame, memorize: bool):
        main = torch.cuda.current_stream(next_frame.device)
        self._side.wait_stream(main)
        with torch.cuda.stream(self._side):
            nxt = self._encode(next_frame)
        pred = self._heavy(memorize)
        main.wait_stream(self._side)
        capturing = torch.cuda.is_current_stream_capturing()
        for cur, new in zip(self._cur, nxt):                 # rotate: the look-ahead features become the current ones
            if not capturing:
                new.record_stream(main)                      # allocated on the side stream, read here on the main one
            cur.copy_(new)
        self._cur_frame.copy_(next_frame)
        return pred

    @torch.no_grad()
    def step(self, next_frame, memorize: bool = True):
        """Segment (and memorize) the current frame; ``next_frame`` is encoded meanwhile.  Returns (B,H,W) int64."""
        if self._cur is None:
            raise RuntimeError('PipelinedSequenceRunner.step() before prime()')
        if next_frame is None:
            return self._heavy(memorize)
        key = _graph_key(self.model.swem_core, next_frame, memorize)
        if self._graph is not None and key != self._key:         # another step than the captured one: drop the graph
            self._graph, self._key = None, None
            self.model.swem_core.static_banks = False
        if not self.use_graph or self._seen < self.eager_steps or key is None:
            self._seen += 1
            return self._two_branch_step(next_frame, memorize)
        if self._graph is None:
            self.model.swem_core.static_banks = True
            self._next_in = next_frame.clone()
            torch.cuda.synchronize(next_frame.device)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self._pred = self._two_branch_step(self._next_in, memorize)
            self._graph, self._key = graph, key
        self._next_in.copy_(next_frame, non_blocking=True)
        self._graph.replay()
        return self._pred
