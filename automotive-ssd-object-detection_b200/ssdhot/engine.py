"""Pre-planned execution of the hot path: static buffers, raw C-ABI calls, optional CUDA graph.

`HotPathStep` is the allocation-free form of the post-backbone part of the reference's eval step
(SSD_test_step, SSD_trainer.py:214-256): targets + both losses, then `predict`.  All outputs and
workspaces are allocated once; `run()` issues exactly the library's kernels on the current
stream (or replays them from a captured CUDA graph), with no host synchronisation.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import _lib
from . import dist as _dist
from .api import METRICS, HeadSet, PackedTargets  # noqa: F401  (HeadSet: the head-direct launches below)
from .priors import PriorSet


class HotPathStep:
    def __init__(self, priors: PriorSet, batch: int, n_classes: int, iou_thresh: float = 0.5,
                 neg_pos_ratio: float = 3.0, score_thresh: float = 0.01, nms_thresh: float = 0.45,
                 max_per_img: int = 200, class_agnostic: bool = False, metric: str = "diou",
                 norm_wh=(300.0, 300.0), train_half: bool = True, infer_half: bool = True, max_gt: int = 64,
                 concurrent: bool = True, group=None):
        self.ps, self.B, self.C = priors, int(batch), int(n_classes)
        self.iou_thresh, self.ratio = float(iou_thresh), float(neg_pos_ratio)
        self.score_thresh, self.nms_thresh = float(score_thresh), float(nms_thresh)
        self.max_per_img, self.agnostic, self.metric = int(max_per_img), bool(class_agnostic), METRICS[metric]
        self.norm_wh = (float(norm_wh[0]), float(norm_wh[1]))
        self.train_half, self.infer_half = train_half, infer_half
        self.concurrent, self._fork = bool(concurrent), None
        # with a PeerSums the exchange kernel folds the per-image partials itself: loss kernel -> exchange, no finalize launch
        self._fold_in_peer = isinstance(group, _dist.PeerSums)
        self.group = group          # ranks that share the batch: a torch.distributed group (True = default) or a dist.PeerSums;
                                    # the three sums are all-reduced inside run()
        dev = priors.device
        L = _lib.lib()
        self.sums = torch.zeros((3,), dtype=torch.float64, device=dev)
        self.max_gt = int(max_gt)
        self.loss_work = torch.empty((int(L.ssdhot_loss_workspace_bytes(self.B, priors.P, self.max_gt)),), dtype=torch.uint8, device=dev)
        self.pred_work = torch.empty((int(L.ssdhot_predict_workspace_bytes(self.B, priors.P, self.C)),),
                                     dtype=torch.uint8, device=dev)
        m = self.max_per_img
        self.labels = torch.zeros((self.B, m), dtype=torch.int64, device=dev)
        self.scores = torch.zeros((self.B, m), dtype=torch.float32, device=dev)
        self.boxes = torch.zeros((self.B, m, 4), dtype=torch.float32, device=dev)
        self.count = torch.zeros((self.B,), dtype=torch.int32, device=dev)
        self.n_pos = torch.zeros((self.B,), dtype=torch.int32, device=dev)
        # key hand-off between the halves when one conf_all feeds both (ssdhot.h: ssdhot_share_bytes)
        self.share = torch.empty((int(L.ssdhot_share_bytes(self.B, priors.P)),), dtype=torch.uint8, device=dev)
        self.share_keys = True
        self._graphs: Dict[tuple, tuple] = {}      # key -> (graph, tensors kept alive), in LRU order

    # -- raw launches -------------------------------------------------------------------------
    def launch_loss(self, loc: torch.Tensor, conf: torch.Tensor, gt: PackedTargets, stream: int, share: bool = False) -> None:
        ps = self.ps
        if gt.max_gt > self.max_gt:
            raise _lib.SsdhotError(f"HotPathStep was planned for at most {self.max_gt} boxes per image, got {gt.max_gt}")
        rc = _lib.lib().ssdhot_multibox_loss_fwd(
            ps.priors.data_ptr(), ps.priors_xyxy.data_ptr(), ps.aux.data_ptr(), ps.P, ps.layout,
            gt.boxes.data_ptr(), gt.labels.data_ptr(), gt.offsets.data_ptr(), self.B, gt.max_gt,
            self.norm_wh[0], self.norm_wh[1], loc.data_ptr(), conf.data_ptr(), self.C,
            self.iou_thresh, ps.variances[0], ps.variances[1], self.ratio,
            None if self._fold_in_peer else self.sums.data_ptr(), self.loss_work.data_ptr(), None, None, self.n_pos.data_ptr(), None,
            self.share.data_ptr() if share else None, stream)
        _lib.check(rc, "ssdhot_multibox_loss_fwd")

    def launch_loss_heads(self, heads: HeadSet, gt: PackedTargets, stream: int, share: bool = False) -> None:
        """launch_loss reading the head outputs directly (train_image_kernel with a per-level source)."""
        ps = self.ps
        if gt.max_gt > self.max_gt:
            raise _lib.SsdhotError(f"HotPathStep was planned for at most {self.max_gt} boxes per image, got {gt.max_gt}")
        rc = _lib.lib().ssdhot_multibox_loss_heads_fwd(
            ps.priors.data_ptr(), ps.priors_xyxy.data_ptr(), ps.aux.data_ptr(), ps.layout,
            gt.boxes.data_ptr(), gt.labels.data_ptr(), gt.offsets.data_ptr(), self.B, gt.max_gt,
            self.norm_wh[0], self.norm_wh[1], heads.loc_ptr, heads.conf_ptr, heads.layout, self.C,
            self.iou_thresh, ps.variances[0], ps.variances[1], self.ratio,
            None if self._fold_in_peer else self.sums.data_ptr(), self.loss_work.data_ptr(), None, None, self.n_pos.data_ptr(), None,
            self.share.data_ptr() if share else None, stream)
        _lib.check(rc, "ssdhot_multibox_loss_heads_fwd")

    def launch_predict_heads(self, heads: HeadSet, stream: int, stages: int = 3, share: bool = False) -> None:
        """launch_predict reading the head outputs directly (score_kernel / nms_image_kernel with a per-level source)."""
        ps = self.ps
        rc = _lib.lib().ssdhot_predict_heads(
            ps.priors.data_ptr(), heads.loc_ptr, heads.conf_ptr, heads.layout, self.B, self.C,
            self.score_thresh, self.nms_thresh, self.max_per_img, 1 if self.agnostic else 0, self.metric,
            ps.variances[0], ps.variances[1], float(ps.img_w), float(ps.img_h),
            self.labels.data_ptr(), self.scores.data_ptr(), self.boxes.data_ptr(), None,
            self.count.data_ptr(), self.pred_work.data_ptr(), int(stages), self.share.data_ptr() if share else None, stream)
        _lib.check(rc, "ssdhot_predict_heads")

    def launch_predict(self, loc: torch.Tensor, conf: torch.Tensor, stream: int, stages: int = 3, share: bool = False) -> None:
        """stages: 1 = score_kernel only (fills the candidate lists), 2 = nms_image_kernel only, 3 = both."""
        ps = self.ps
        rc = _lib.lib().ssdhot_predict_stages(
            ps.priors.data_ptr(), ps.P, loc.data_ptr(), conf.data_ptr(), self.B, self.C,
            self.score_thresh, self.nms_thresh, self.max_per_img, 1 if self.agnostic else 0, self.metric,
            ps.variances[0], ps.variances[1], float(ps.img_w), float(ps.img_h),
            self.labels.data_ptr(), self.scores.data_ptr(), self.boxes.data_ptr(), None,
            self.count.data_ptr(), self.pred_work.data_ptr(), int(stages), self.share.data_ptr() if share else None, stream)
        _lib.check(rc, "ssdhot_predict_stages")

    def run(self, loc: torch.Tensor, conf_train: torch.Tensor, conf_infer: torch.Tensor, gt: PackedTargets,
            use_graph: bool = False) -> None:
        """Both halves, issued from the current stream (the predict half forks onto a second stream when
        `concurrent`, since the halves share nothing but read-only inputs).  With use_graph the launches
        are captured once per distinct set of input buffers and replayed afterwards."""
        dev = self.ps.device
        if not use_graph:
            cur = torch.cuda.current_stream(dev)
            if self.train_half and self.infer_half and self.concurrent:
                # the two halves are independent: fork the predict half onto a second stream and join
                if self._fork is None:
                    self._fork = torch.cuda.Stream(dev)
                # ONE conf_all for both halves (SSD_test_step): the loss kernel's stream hands predict its row keys
                share = self.share_keys and conf_train.data_ptr() == conf_infer.data_ptr()
                if share:
                    _lib.check(_lib.lib().ssdhot_share_reset(self.share.data_ptr(), self.B, cur.cuda_stream), "ssdhot_share_reset")
                self._fork.wait_stream(cur)
                self.launch_loss(loc, conf_train, gt, cur.cuda_stream, share=share)
                self.launch_predict(loc, conf_infer, self._fork.cuda_stream, share=share)
                self._reduce()                                   # overlaps the predict half
                cur.wait_stream(self._fork)
                return
            share = self.share_keys and self.train_half and self.infer_half and conf_train.data_ptr() == conf_infer.data_ptr()
            if share:
                _lib.check(_lib.lib().ssdhot_share_reset(self.share.data_ptr(), self.B, cur.cuda_stream), "ssdhot_share_reset")
            if self.train_half:
                self.launch_loss(loc, conf_train, gt, cur.cuda_stream, share=share)
                self._reduce()
            if self.infer_half:
                self.launch_predict(loc, conf_infer, cur.cuda_stream, share=share)
            return
        tensors = (loc, conf_train, conf_infer, gt.boxes, gt.labels, gt.offsets)
        g = self._graph_for(("packed",) + tuple(t.data_ptr() for t in tensors) + (gt.max_gt, self.share_keys), tensors,
                            lambda: self.run(loc, conf_train, conf_infer, gt))
        g.replay()

    GRAPH_CACHE = 16

    def _graph_for(self, key, keep_alive, issue) -> "torch.cuda.CUDAGraph":
        """The captured step for this exact set of input buffers.  The key holds EVERY captured pointer and the cache entry
        keeps the tensors alive next to the graph, so an address cannot be recycled by the caching allocator while a graph
        still points at it; the cache is bounded (least recently used entry dropped)."""
        hit = self._graphs.pop(key, None)
        if hit is None:
            dev = self.ps.device
            g = torch.cuda.CUDAGraph()
            # the capture stream carries the loss branch, whose last node is the exchange: high priority lets its CTAs and the
            # small exchange kernel ahead of the predict branch's (N = 2, lag 0: 81.9 -> 78.8 us per step; no effect at N = 1)
            side = torch.cuda.Stream(dev, priority=-1)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                issue()                                              # warm (sets the shared-memory opt-ins)
                side.synchronize()
                with torch.cuda.graph(g, stream=side):
                    issue()
            torch.cuda.current_stream(dev).wait_stream(side)
            hit = (g, keep_alive)
            while len(self._graphs) >= self.GRAPH_CACHE:
                self._graphs.pop(next(iter(self._graphs)))
        self._graphs[key] = hit                                      # (re-inserted last = most recently used)
        return hit[0]

    def run_heads(self, train_heads: HeadSet, infer_heads: HeadSet, gt: PackedTargets, use_graph: bool = False) -> None:
        """run() fed by the head outputs (HeadSet: NCHW or channels_last) instead of loc_all / conf_all: both halves read the
        six tensors of each branch directly, the predict half forked onto a second stream; use_graph captures the launches
        once per distinct pair of HeadSets."""
        dev = self.ps.device
        if not use_graph:
            cur = torch.cuda.current_stream(dev)
            if self._fork is None:
                self._fork = torch.cuda.Stream(dev)
            share = self.share_keys and train_heads is infer_heads
            if share:
                _lib.check(_lib.lib().ssdhot_share_reset(self.share.data_ptr(), self.B, cur.cuda_stream), "ssdhot_share_reset")
            self._fork.wait_stream(cur)
            self.launch_loss_heads(train_heads, gt, cur.cuda_stream, share=share)
            self.launch_predict_heads(infer_heads, self._fork.cuda_stream, share=share)
            self._reduce()
            cur.wait_stream(self._fork)
            return
        tensors = tuple(train_heads.tensors) + tuple(infer_heads.tensors) + (gt.boxes, gt.labels, gt.offsets)
        g = self._graph_for(("heads", train_heads.layout, infer_heads.layout) + tuple(t.data_ptr() for t in tensors) + (gt.max_gt, self.share_keys),
                            tensors + (train_heads, infer_heads),    # (the HeadSets own the host pointer arrays the launches read)
                            lambda: self.run_heads(train_heads, infer_heads, gt))
        g.replay()

    def _reduce(self) -> None:
        """The sharded path's only exchange: all-reduce [sum smooth-L1, sum CE, sum positives] in place."""
        if self._fold_in_peer:                          # one kernel, part of the step's CUDA graph: fold the partials + exchange
            self.group.allreduce_partials(self.loss_work, self.B, self.n_pos, self.sums)
        elif self.group is not None:
            _dist.reduce_sums(self.sums, self.group)

    def losses(self, group=None):
        """(loc_loss, conf_loss) of the last run; all-reduces the three sums first when sharded."""
        sums = self.sums.clone()
        if group is not None:
            _dist.reduce_sums(sums, group)
        return _dist.losses_from_sums(sums)


class StepPipeline:
    """Several eval steps in flight.  Consecutive batches of an evaluation pass (SSD_test_step, SSD_trainer.py:208-256) do not
    depend on each other, so the logit stream of batch i+1 (HBM-bound) may run beside the NMS tail of batch i (a chain of short
    phases that leaves most issue slots idle): `depth` HotPathStep objects -- each with its own outputs, workspaces, share
    buffer, CUDA graphs and stream -- take the batches round-robin.  One B200, B = 256: 63.6 us per step with one step in
    flight, 49.3 us with two (three: no further gain -- two CTA slots per SM are then always taken).

    submit() returns the step object that will hold the batch's results; wait(step) makes the current stream wait for them.
    A step object is reused `depth` submits later: read (or copy) its outputs before that.  Not for a TRAINING loop, whose
    steps depend on each other through the optimizer."""

    def __init__(self, make_step, depth: int = 2):
        if depth < 1:
            raise ValueError("depth must be at least 1")
        self.steps = [make_step(k) for k in range(depth)]
        dev = self.steps[0].ps.device
        self.streams = [torch.cuda.Stream(dev) for _ in range(depth)]
        self._dev, self._next = dev, 0

    def submit(self, loc: torch.Tensor, conf_train: torch.Tensor, conf_infer: torch.Tensor, gt: PackedTargets,
               use_graph: bool = True) -> HotPathStep:
        k = self._next % len(self.steps)
        self._next += 1
        step, st = self.steps[k], self.streams[k]
        st.wait_stream(torch.cuda.current_stream(self._dev))        # the inputs are ready where the caller produced them
        with torch.cuda.stream(st):
            step.run(loc, conf_train, conf_infer, gt, use_graph=use_graph)
        return step

    def wait(self, step: Optional[HotPathStep] = None) -> None:
        """The current stream waits for `step` (default: for every step in flight)."""
        cur = torch.cuda.current_stream(self._dev)
        for s, st in zip(self.steps, self.streams):
            if step is None or s is step:
                cur.wait_stream(st)
