"""Drop-in Python surface of the hot path: the same call signatures as the reference's
`build_targets` / `CELoss_w_neg_mining` (SSD_trainer.py:491, :551) and `mySSD.encode_ssd` /
`decode_ssd` / `iou_nms` / `predict` (SSD_from_scratch.py:697, :776, :664, :338), each a thin shim
(argument validation identical to the reference, ground-truth packing, output allocation, ONE
C-ABI call per stage).  Tensors must live on a CUDA device; there is no CPU path.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.utils.data

from . import _lib
from . import dist as _dist
from .priors import PriorSet

METRICS = {"diou": 0, "ciou": 1, "iou": 2}


def _stream(dev: torch.device) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _need_cuda(*tensors: torch.Tensor) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise _lib.SsdhotError("ssdhot runs on CUDA tensors only (no CPU fallback)")
        dev = t.device if dev is None else dev
        if t.device != dev:
            raise _lib.SsdhotError("all tensors must be on the same CUDA device")
    return dev


class PackedTargets:
    """Ragged ground truth of a batch (hard part 9): boxes [sumG,4] f32 pixel xyxy, labels [sumG] i64,
    offsets [B+1] i32 -- three views of ONE byte buffer laid out [offsets | pad to 16 B | boxes | labels], so that a batch's
    ground truth crosses PCIe as a single copy (`to(device)`) instead of the reference's three tiny copies per image
    (SSD_trainer.py:66-69).  max_gt and the per-image counts are known on the host."""

    def __init__(self, boxes: torch.Tensor, labels: torch.Tensor, offsets: torch.Tensor, max_gt: int, n_img: int,
                 buffer: Optional[torch.Tensor] = None, counts: Optional[Sequence[int]] = None):
        self.boxes, self.labels, self.offsets, self.max_gt, self.n_img = boxes, labels, offsets, int(max_gt), int(n_img)
        self.buffer, self.counts = buffer, (list(counts) if counts is not None else None)

    @staticmethod
    def layout(n_img: int, total: int) -> Tuple[int, int, int]:
        """-> (byte offset of boxes, byte offset of labels, buffer bytes); at least one (unused) row is always present."""
        rows = max(int(total), 1)
        box_off = (4 * (n_img + 1) + 15) & ~15
        lab_off = box_off + 16 * rows
        return box_off, lab_off, lab_off + 8 * rows

    @classmethod
    def from_buffer(cls, buffer: torch.Tensor, n_img: int, total: int, max_gt: int, counts=None) -> "PackedTargets":
        box_off, lab_off, nbytes = cls.layout(n_img, total)
        rows = max(int(total), 1)
        offsets = buffer[: 4 * (n_img + 1)].view(torch.int32)
        boxes = buffer[box_off:lab_off].view(torch.float32).view(rows, 4)
        labels = buffer[lab_off:nbytes].view(torch.int64)
        return cls(boxes, labels, offsets, max_gt, n_img, buffer, counts)

    @property
    def device(self) -> torch.device:
        return self.boxes.device

    @property
    def total(self) -> int:
        return int(sum(self.counts)) if self.counts is not None else int(self.boxes.shape[0])

    def to(self, device, non_blocking: bool = True) -> "PackedTargets":
        """The whole batch's ground truth in ONE copy (cudaMemcpyAsync from pinned memory when the buffer is pinned)."""
        device = torch.device(device)
        if self.boxes.device == device or (device.type == "cuda" and device.index is None and self.boxes.is_cuda):
            return self
        if self.buffer is None:
            return PackedTargets(self.boxes.to(device, non_blocking=non_blocking), self.labels.to(device, non_blocking=non_blocking),
                                 self.offsets.to(device, non_blocking=non_blocking), self.max_gt, self.n_img, None, self.counts)
        return PackedTargets.from_buffer(self.buffer.to(device, non_blocking=non_blocking), self.n_img, self.total, self.max_gt, self.counts)

    def pin_memory(self) -> "PackedTargets":            # (torch.utils.data.DataLoader(pin_memory=True) calls this on custom batches)
        if self.buffer is None or self.buffer.is_pinned() or self.buffer.is_cuda:
            return self
        return PackedTargets.from_buffer(self.buffer.pin_memory(), self.n_img, self.total, self.max_gt, self.counts)

    def record_stream(self, stream) -> None:
        for t in ((self.buffer,) if self.buffer is not None else (self.boxes, self.labels, self.offsets)):
            if t.is_cuda:
                t.record_stream(stream)

    def as_list(self) -> List[Dict[str, torch.Tensor]]:
        """The reference's List[Dict] form (views, no copies): what collate_detection's second output used to be."""
        counts = self.counts if self.counts is not None else (self.offsets[1:] - self.offsets[:-1]).tolist()
        out, at = [], 0
        for c in counts:
            out.append({"boxes": self.boxes[at:at + c], "labels": self.labels[at:at + c]})
            at += c
        return out


def _pack_host(targets: Sequence[Dict[str, torch.Tensor]], pinned: bool) -> PackedTargets:
    counts = [int(t["boxes"].shape[0]) if t["boxes"].numel() else 0 for t in targets]
    n_img, total = len(counts), sum(counts)
    _, _, nbytes = PackedTargets.layout(n_img, total)
    buf = torch.zeros((nbytes,), dtype=torch.uint8, pin_memory=pinned)
    packed = PackedTargets.from_buffer(buf, n_img, total, max(counts) if counts else 0, counts)
    offs = [0]
    for c in counts:
        offs.append(offs[-1] + c)
    packed.offsets.copy_(torch.tensor(offs, dtype=torch.int32))
    live = [i for i, c in enumerate(counts) if c > 0]
    if live:
        torch.cat([targets[i]["boxes"].as_subclass(torch.Tensor).reshape(-1, 4).to(torch.float32) for i in live], 0, out=packed.boxes[:total])
        torch.cat([targets[i]["labels"].as_subclass(torch.Tensor).reshape(-1).to(torch.int64) for i in live], 0, out=packed.labels[:total])
    return packed


def pack_targets(targets: Sequence[Dict[str, torch.Tensor]], device) -> PackedTargets:
    """List[Dict] (collate_detection, SSD_trainer.py:806-813) -> PackedTargets on `device`.  CPU member tensors are packed
    into one pinned host buffer and cross to the device in ONE copy; member tensors that already live on a CUDA device are
    concatenated there (only the B+1 offsets come from the host).  A PackedTargets is moved (one copy) or returned as is."""
    device = torch.device(device)
    if isinstance(targets, PackedTargets):
        return targets.to(device)
    if any(t["boxes"].is_cuda for t in targets):
        counts = [int(t["boxes"].shape[0]) if t["boxes"].numel() else 0 for t in targets]
        offs = [0]
        for c in counts:
            offs.append(offs[-1] + c)
        offsets = torch.tensor(offs, dtype=torch.int32).to(device, non_blocking=True)
        live = [i for i, c in enumerate(counts) if c > 0]
        if not live:
            boxes = torch.zeros((1, 4), dtype=torch.float32, device=device)   # never read (max_gt = 0)
            labels = torch.zeros((1,), dtype=torch.int64, device=device)
        else:
            boxes = torch.cat([targets[i]["boxes"].as_subclass(torch.Tensor).reshape(-1, 4).to(device=device, dtype=torch.float32) for i in live], 0)
            labels = torch.cat([targets[i]["labels"].reshape(-1).to(device=device, dtype=torch.int64) for i in live], 0)
        return PackedTargets(boxes.contiguous(), labels.contiguous(), offsets, max(counts) if counts else 0, len(counts), None, counts)
    return _pack_host(targets, pinned=device.type == "cuda" and torch.cuda.is_available()).to(device)


def collate_detection(batch):
    """Drop-in for SSD_trainer.collate_detection (SSD_trainer.py:806-813): list of (img, target) -> (images [B,C,H,W],
    PackedTargets).  The ground truth of the whole batch is packed ONCE, here, into a single (pinned, when this process may
    pin) host buffer [offsets | boxes | labels]; `targets.to(device)` -- which the ssdhot step functions, build_targets and
    multibox_loss do themselves -- is then one cudaMemcpyAsync instead of the reference's 3 copies per image
    (SSD_trainer.py:66-69).  `targets.as_list()` gives the reference's List[Dict] back."""
    imgs = [img for img, _ in batch]
    tgts = [tgt for _, tgt in batch]
    in_worker = torch.utils.data.get_worker_info() is not None       # (a DataLoader worker must not touch CUDA; the loader's
    pinned = torch.cuda.is_available() and not in_worker             #  pin_memory thread pins the buffer through pin_memory())
    return torch.stack(imgs, dim=0), _pack_host(tgts, pinned)


# ------------------------------------------------------------------------------------------------
# match + encode
# ------------------------------------------------------------------------------------------------
def match_encode_batch(priors: PriorSet, packed: PackedTargets, iou_thresh: float, norm_wh: Tuple[float, float],
                       want_loc: str = "none", want_cls: bool = True, want_pos: bool = True,
                       want_matched_idx: bool = False, want_matched_box: bool = False):
    """One ssdhot_match_encode launch.  want_loc: 'none' | 'positives' | 'all'.  Returns a dict."""
    dev = priors.device
    B, P = packed.n_img, priors.P
    out = {}
    out["n_pos"] = torch.empty((B,), dtype=torch.int32, device=dev)
    loc = None
    if want_loc != "none":
        loc = torch.empty((B, P, 4), dtype=torch.float32, device=dev)
    out["loc_t"] = loc
    out["cls_t"] = torch.empty((B, P), dtype=torch.int64, device=dev) if want_cls else None
    out["pos_mask"] = torch.empty((B, P), dtype=torch.bool, device=dev) if want_pos else None
    out["matched_gt"] = torch.empty((B, P), dtype=torch.int32, device=dev) if want_matched_idx else None
    out["matched_cxcywh"] = torch.empty((B, P, 4), dtype=torch.float32, device=dev) if want_matched_box else None
    work = _workspace("match", dev, _lib.lib().ssdhot_match_workspace_bytes(B, packed.max_gt))
    with torch.cuda.device(dev):
        rc = _lib.lib().ssdhot_match_encode(
            priors.priors.data_ptr(), priors.priors_xyxy.data_ptr(), priors.aux.data_ptr(), P, priors.layout,
            packed.boxes.data_ptr(), packed.labels.data_ptr(), packed.offsets.data_ptr(), B, packed.max_gt,
            float(norm_wh[0]), float(norm_wh[1]), float(iou_thresh), priors.variances[0], priors.variances[1],
            _ptr(loc), 1 if want_loc == "positives" else 0, _ptr(out["cls_t"]), _ptr(out["pos_mask"]),
            _ptr(out["matched_gt"]), _ptr(out["matched_cxcywh"]), out["n_pos"].data_ptr(), None, work.data_ptr(),
            _stream(dev))
    _lib.check(rc, "ssdhot_match_encode")
    return out


def build_targets(model, targets: List[Dict], H: int = 300, W: int = 300, iou_thresh: float = 0.50,
                  device="cuda") -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Drop-in for SSD_trainer.build_targets (SSD_trainer.py:491-547):
    -> (pos_mask [B,P] bool, loc_t[pos_mask] [N_pos,4] f32, cls_t [B,P] i64)."""
    if not (0.0 < iou_thresh < 1.0):
        raise ValueError(f"Score threshold should be greater than 0 and less than 1, recieved {iou_thresh}.")
    priors = PriorSet.of(model)
    if torch.device(device).type != "cuda":
        raise _lib.SsdhotError("ssdhot.build_targets needs device='cuda' (no CPU fallback)")
    packed = pack_targets(targets, priors.device)
    r = match_encode_batch(priors, packed, iou_thresh, (W, H), want_loc="positives")
    total = int(r["n_pos"].sum().item())            # the reference's return shape needs N_pos on the host
    loc_pm = torch.empty((total, 4), dtype=torch.float32, device=priors.device)
    if total > 0:
        with torch.cuda.device(priors.device):
            rc = _lib.lib().ssdhot_compact_rows(r["loc_t"].data_ptr(), r["pos_mask"].data_ptr(), r["n_pos"].data_ptr(),
                                                packed.n_img, priors.P, loc_pm.data_ptr(), _stream(priors.device))
        _lib.check(rc, "ssdhot_compact_rows")
    return r["pos_mask"], loc_pm, r["cls_t"]


def encode_ssd(model, gt_boxes_xyxy: torch.Tensor, gt_labels: torch.Tensor, iou_thresh: float = 0.5,
               background_class: int = 0):
    """Drop-in for mySSD.encode_ssd (SSD_from_scratch.py:697-773), one image, normalised boxes:
    -> (loc_target [P,4], cls_target [P], pos_mask [P] bool, matched_gt_cxcywh [P,4])."""
    if background_class != 0:
        raise ValueError(f"Background should be 0, recieved {background_class}.")
    priors = PriorSet.of(model)
    _need_cuda(gt_boxes_xyxy, gt_labels)
    packed = pack_targets([{"boxes": gt_boxes_xyxy, "labels": gt_labels}], priors.device)
    r = match_encode_batch(priors, packed, iou_thresh, (1.0, 1.0), want_loc="all", want_matched_box=True)
    cls = r["cls_t"][0]
    if gt_labels.dtype != torch.int64:
        cls = cls.to(gt_labels.dtype)
    return r["loc_t"][0], cls, r["pos_mask"][0], r["matched_cxcywh"][0]


# ------------------------------------------------------------------------------------------------
# losses
# ------------------------------------------------------------------------------------------------
_work_cache: Dict[Tuple, torch.Tensor] = {}


def _workspace(tag: str, dev: torch.device, nbytes: int) -> torch.Tensor:
    key = (tag, dev)
    buf = _work_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty((max(int(nbytes), 256),), dtype=torch.uint8, device=dev)
        _work_cache[key] = buf
    return buf


class _MinedCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, conf_all, cls_t, pos_mask, total_pos, ratio):
        dev = _need_cuda(conf_all, cls_t, pos_mask)
        B, P, C = conf_all.shape
        conf = conf_all.detach().to(torch.float32).contiguous()
        cls = cls_t.to(torch.int64).contiguous()
        pos = pos_mask.to(torch.bool).contiguous()
        sums = torch.empty((3,), dtype=torch.float64, device=dev)
        sel = torch.empty((B, P), dtype=torch.int8, device=dev) if conf_all.requires_grad else None
        work = _workspace("loss", dev, _lib.lib().ssdhot_loss_workspace_bytes(B, P, 0))
        with torch.cuda.device(dev):
            rc = _lib.lib().ssdhot_mined_ce_fwd(conf.data_ptr(), cls.data_ptr(), pos.data_ptr(), B, P, C, float(ratio),
                                                sums.data_ptr(), work.data_ptr(), _ptr(sel), _stream(dev))
        _lib.check(rc, "ssdhot_mined_ce_fwd")
        total = torch.as_tensor(total_pos, device=dev).to(torch.float64)
        ctx.save_for_backward(conf, sel, total)
        return (sums[1] / total).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        conf, sel, total = ctx.saved_tensors
        dev = conf.device
        B, P, C = conf.shape
        scales = torch.stack((torch.zeros((), dtype=torch.float64, device=dev), g.to(torch.float64) / total))
        grad = torch.empty_like(conf)
        with torch.cuda.device(dev):
            rc = _lib.lib().ssdhot_multibox_loss_bwd(None, P, None, None, B, 1.0, 1.0, None, conf.data_ptr(), C, 0.1, 0.2,
                                                     sel.data_ptr(), None, scales.data_ptr(), None, grad.data_ptr(),
                                                     _stream(dev))
        _lib.check(rc, "ssdhot_multibox_loss_bwd")
        return grad, None, None, None, None


def CELoss_w_neg_mining(conf_all: torch.Tensor, cls_t: torch.Tensor, pos_mask: torch.Tensor,
                        num_pos_per_img: torch.Tensor, total_pos, neg_pos_ratio: float = 3.0) -> torch.Tensor:
    """Drop-in for SSD_trainer.CELoss_w_neg_mining (SSD_trainer.py:551-600); differentiable w.r.t.
    conf_all.  `num_pos_per_img` is accepted for signature parity; the kernel recounts positives
    from pos_mask on the device instead of syncing once per image as the reference does (:585)."""
    return _MinedCE.apply(conf_all, cls_t, pos_mask, total_pos, float(neg_pos_ratio))


def _check_group_for_grad(group, need_grad: bool) -> None:
    """A lagged PeerSums delivers the PREVIOUS call's reduced sums: fine for logging an eval step, wrong for a training
    step, whose backward is scaled by 1 / (positives of THIS batch) (SSD_trainer.py:105-130)."""
    if need_grad and isinstance(group, _dist.PeerSums) and group.lag != 0:
        raise _lib.SsdhotError("a differentiable loss needs the sums of its own batch: use PeerSums(lag=0) (lag=1 delivers the "
                               "previous call's sums and would mis-scale the gradients)")


class _FusedLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, loc_all, conf_all, priors: PriorSet, packed: PackedTargets, iou_thresh, ratio, norm_wh, group, share=None):
        dev = _need_cuda(loc_all, conf_all)
        B, P, C = conf_all.shape
        loc = loc_all.detach().to(torch.float32).contiguous()
        conf = conf_all.detach().to(torch.float32).contiguous()
        need_grad = loc_all.requires_grad or conf_all.requires_grad
        sums = torch.empty((3,), dtype=torch.float64, device=dev)
        sel = torch.empty((B, P), dtype=torch.int8, device=dev) if need_grad else None
        matched = torch.empty((B, P), dtype=torch.int16, device=dev) if need_grad else None
        work = _workspace("loss", dev, _lib.lib().ssdhot_loss_workspace_bytes(B, P, packed.max_gt))
        with torch.cuda.device(dev):
            rc = _lib.lib().ssdhot_multibox_loss_fwd(
                priors.priors.data_ptr(), priors.priors_xyxy.data_ptr(), priors.aux.data_ptr(), P, priors.layout,
                packed.boxes.data_ptr(), packed.labels.data_ptr(), packed.offsets.data_ptr(), B, packed.max_gt,
                float(norm_wh[0]), float(norm_wh[1]), loc.data_ptr(), conf.data_ptr(), C,
                float(iou_thresh), priors.variances[0], priors.variances[1], float(ratio),
                sums.data_ptr(), work.data_ptr(), _ptr(sel), _ptr(matched), None, None, _ptr(share), _stream(dev))
        _lib.check(rc, "ssdhot_multibox_loss_fwd")
        if group is not None:
            # the only exchange of the sharded path: [sum smooth-L1, sum CE, sum positives]
            _check_group_for_grad(group, need_grad)
            _dist.reduce_sums(sums, group)
        total = sums[2].clamp_min(1.0)
        if need_grad:
            ctx.save_for_backward(loc, conf, sel, matched, total, packed.boxes, packed.offsets)
            ctx.meta = (priors, norm_wh)
        losses = (sums[:2] / total).to(torch.float32)
        ctx.mark_non_differentiable(sums)
        return losses[0], losses[1], sums

    @staticmethod
    def backward(ctx, g_loc, g_conf, _g_sums):
        loc, conf, sel, matched, total, gt_boxes, gt_offsets = ctx.saved_tensors
        priors, norm_wh = ctx.meta
        dev = conf.device
        B, P, C = conf.shape
        scales = torch.stack((g_loc.to(torch.float64) / total, g_conf.to(torch.float64) / total))
        d_loc = torch.empty_like(loc) if ctx.needs_input_grad[0] else None
        d_conf = torch.empty_like(conf) if ctx.needs_input_grad[1] else None
        with torch.cuda.device(dev):
            rc = _lib.lib().ssdhot_multibox_loss_bwd(
                priors.priors.data_ptr(), P, gt_boxes.data_ptr(), gt_offsets.data_ptr(), B,
                float(norm_wh[0]), float(norm_wh[1]), loc.data_ptr(), conf.data_ptr(), C,
                priors.variances[0], priors.variances[1], sel.data_ptr(), matched.data_ptr(), scales.data_ptr(),
                _ptr(d_loc), _ptr(d_conf), _stream(dev))
        _lib.check(rc, "ssdhot_multibox_loss_bwd")
        return d_loc, d_conf, None, None, None, None, None, None, None


def multibox_loss(model, loc_all: torch.Tensor, conf_all: torch.Tensor, targets, iou_thresh: float = 0.5,
                  neg_pos_ratio: float = 3.0, H: int = 300, W: int = 300, group=None, return_sums: bool = False,
                  _share: Optional[torch.Tensor] = None):
    """The whole post-backbone training step of SSD_train_step (SSD_trainer.py:92-117) in one
    launch: -> (batch_loc_loss, batch_conf_loss), 0-d fp32, differentiable w.r.t. loc_all/conf_all.
    With `group` (a torch.distributed process group, or True for the default group) the batch is
    this rank's shard: the three partial sums are all-reduced so every rank returns the
    global-batch losses."""
    if not (0.0 < iou_thresh < 1.0):
        raise ValueError(f"Score threshold should be greater than 0 and less than 1, recieved {iou_thresh}.")
    priors = PriorSet.of(model)
    packed = pack_targets(targets, priors.device)
    l_loc, l_conf, sums = _FusedLoss.apply(loc_all, conf_all, priors, packed, float(iou_thresh), float(neg_pos_ratio),
                                           (W, H), group, _share)
    if return_sums:
        return l_loc, l_conf, sums
    return l_loc, l_conf


def smooth_l1_positive_loss(loc_all: torch.Tensor, pos_mask: torch.Tensor, loc_t_pm: torch.Tensor, total_pos) -> torch.Tensor:
    """SSD_trainer.py:108 as the reference writes it (kept in torch: it is one gather + one library
    loss on N_pos rows; the fused path above never materialises it)."""
    return torch.nn.functional.smooth_l1_loss(loc_all[pos_mask], loc_t_pm, reduction="sum") / total_pos


# ------------------------------------------------------------------------------------------------
# decode / NMS / predict
# ------------------------------------------------------------------------------------------------
def decode_ssd(loc: torch.Tensor, priors: torch.Tensor, variances: Tuple[float, float]) -> torch.Tensor:
    """Drop-in for mySSD.decode_ssd (SSD_from_scratch.py:776-800)."""
    dev = _need_cuda(loc, priors)
    loc_c = loc.to(torch.float32).contiguous()
    pri_c = priors.to(torch.float32).contiguous()
    out = torch.empty_like(loc_c)
    with torch.cuda.device(dev):
        rc = _lib.lib().ssdhot_decode(loc_c.data_ptr(), pri_c.data_ptr(), int(loc_c.shape[0]), float(variances[0]),
                                      float(variances[1]), out.data_ptr(), _stream(dev))
    _lib.check(rc, "ssdhot_decode")
    return out


def nms_sets(boxes: torch.Tensor, scores: torch.Tensor, set_sizes: Sequence[int], thresh: float, metric: str = "diou",
             max_keep: int = 0):
    """Greedy NMS of several independent box sets stored back to back -> (keep [total] i64 with set
    s's survivors at offset sum(set_sizes[:s]), keep_count [n_sets] i32)."""
    dev = _need_cuda(boxes, scores)
    offs = [0]
    for n in set_sizes:
        offs.append(offs[-1] + int(n))
    total = offs[-1]
    set_offsets = torch.tensor(offs, dtype=torch.int32).to(dev)
    bx = boxes.to(torch.float32).contiguous()
    sc = scores.to(torch.float32).contiguous()
    keep = torch.empty((max(total, 1),), dtype=torch.int64, device=dev)
    count = torch.empty((len(set_sizes),), dtype=torch.int32, device=dev)
    work = _workspace("nms", dev, _lib.lib().ssdhot_nms_workspace_bytes(total))
    with torch.cuda.device(dev):
        rc = _lib.lib().ssdhot_nms(bx.data_ptr(), sc.data_ptr(), set_offsets.data_ptr(), len(set_sizes), total,
                                   max(set_sizes) if len(set_sizes) else 0, float(thresh), METRICS[metric], int(max_keep),
                                   keep.data_ptr(), count.data_ptr(), work.data_ptr(), _stream(dev))
    _lib.check(rc, "ssdhot_nms")
    return keep[:total], count


def iou_nms(boxes: torch.Tensor, scores: torch.Tensor, iou_threshold: float, metric: str = "diou") -> torch.Tensor:
    """Drop-in for mySSD.iou_nms (SSD_from_scratch.py:664-692): LongTensor of kept indices,
    score-descending (equal scores: ascending index)."""
    if boxes.numel() == 0:
        return boxes.new_zeros((0,), dtype=torch.long)
    keep, count = nms_sets(boxes, scores, [int(boxes.shape[0])], iou_threshold, metric)
    return keep[: int(count.item())]


def predict_padded(model, loc_all: torch.Tensor, conf_all: torch.Tensor, score_thresh: float = 0.2,
                   nms_thresh: float = 0.5, max_per_img: int = 100, class_agnostic: bool = False,
                   metric: str = "diou", want_cand: bool = False, _share: Optional[torch.Tensor] = None):
    """Device-resident form of predict: padded outputs, no host synchronisation.
    -> (labels [B,max] i64, scores [B,max] f32, boxes [B,max,4] f32 px xyxy, count [B] i32[, cand])."""
    if not (0.0 <= score_thresh < 1.0):
        raise ValueError(f"Score threshold should be greater than 0 and less than 1, recieved {score_thresh}.")
    if not (0.0 < nms_thresh < 1.0):
        raise ValueError(f"NMS threshold should be greater than 0 and less than 1, recieved {nms_thresh}.")
    priors = PriorSet.of(model)
    dev = _need_cuda(loc_all, conf_all)
    B, P, C = conf_all.shape
    assert P == priors.P
    assert C >= 2
    loc = loc_all.detach().to(torch.float32).contiguous()
    conf = conf_all.detach().to(torch.float32).contiguous()
    labels = torch.empty((B, max_per_img), dtype=torch.int64, device=dev)
    scores = torch.empty((B, max_per_img), dtype=torch.float32, device=dev)
    boxes = torch.empty((B, max_per_img, 4), dtype=torch.float32, device=dev)
    cand = torch.empty((B, max_per_img), dtype=torch.int32, device=dev) if want_cand else None
    count = torch.empty((B,), dtype=torch.int32, device=dev)
    work = _workspace("predict", dev, _lib.lib().ssdhot_predict_workspace_bytes(B, P, C))
    with torch.cuda.device(dev):
        rc = _lib.lib().ssdhot_predict_stages(priors.priors.data_ptr(), P, loc.data_ptr(), conf.data_ptr(), B, C,
                                              float(score_thresh), float(nms_thresh), int(max_per_img), 1 if class_agnostic else 0,
                                              METRICS[metric], priors.variances[0], priors.variances[1],
                                              float(priors.img_w), float(priors.img_h),
                                              labels.data_ptr(), scores.data_ptr(), boxes.data_ptr(), _ptr(cand),
                                              count.data_ptr(), work.data_ptr(), 3, _ptr(_share), _stream(dev))
    _lib.check(rc, "ssdhot_predict_stages")
    if want_cand:
        return labels, scores, boxes, count, cand
    return labels, scores, boxes, count


@torch.no_grad()
def predict(model, images: Optional[torch.Tensor], score_thresh: float = 0.2, nms_thresh: float = 0.5,
            max_per_img: int = 100, class_agnostic: bool = False, pre_loc_all: Optional[torch.Tensor] = None,
            pre_conf_all: Optional[torch.Tensor] = None, metric: str = "diou") -> List[Dict[str, torch.Tensor]]:
    """Drop-in for mySSD.predict (SSD_from_scratch.py:338-476): List[Dict] with 'labels' i64[K],
    'scores' f32[K], 'boxes' f32[K,4] pixel xyxy, K <= max_per_img, score-descending."""
    if not (0.0 <= score_thresh < 1.0):
        raise ValueError(f"Score threshold should be greater than 0 and less than 1, recieved {score_thresh}.")
    if not (0.0 < nms_thresh < 1.0):
        raise ValueError(f"NMS threshold should be greater than 0 and less than 1, recieved {nms_thresh}.")
    if hasattr(model, "eval"):
        model.eval()                                   # side effect of the reference (:375)
    if (pre_loc_all is not None) and (pre_conf_all is not None):
        loc_all, conf_all = pre_loc_all, pre_conf_all
    elif _has_ssd_heads(model):
        # the model's own backbone and heads, but no permute / cat tail: the kernels read the head outputs directly
        loc_heads, conf_heads = forward_heads(model, images)
        return predict_heads(model, loc_heads, conf_heads, score_thresh, nms_thresh, max_per_img, class_agnostic, metric)
    else:
        loc_all, conf_all = model(images)
    if hasattr(model, "num_classes"):
        assert conf_all.shape[-1] == model.num_classes and conf_all.shape[-1] >= 2
    labels, scores, boxes, count = predict_padded(model, loc_all, conf_all, score_thresh, nms_thresh, max_per_img,
                                                  class_agnostic, metric)
    ks = count.tolist()                                # the one D2H of the call
    return [{"labels": l[:k], "scores": s[:k], "boxes": b[:k]}
            for l, s, b, k in zip(labels.unbind(0), scores.unbind(0), boxes.unbind(0), ks)]


_eval_fork: Dict[torch.device, "torch.cuda.Stream"] = {}


@torch.no_grad()
def eval_step(model, loc_all: torch.Tensor, conf_all: torch.Tensor, targets, iou_thresh: float = 0.5, neg_pos_ratio: float = 3.0,
              score_thresh: float = 0.05, nms_thresh: float = 0.5, max_per_img: int = 100, class_agnostic: bool = False,
              H: int = 300, W: int = 300, group=None, metric: str = "diou"):
    """The post-backbone part of one SSD_test_step batch (SSD_trainer.py:214-256) on ONE pair of head outputs: targets + both
    losses and predict, one kernel each on two streams; the loss kernel's logit stream hands predict its row keys, so conf_all
    is read from HBM once.  -> (loc_loss, conf_loss, labels [B,max] i64, scores [B,max], boxes [B,max,4], count [B] i32), all on the
    device, no host synchronisation."""
    dev = _need_cuda(loc_all, conf_all)
    cur = torch.cuda.current_stream(dev)
    side = _eval_fork.get(dev)
    if side is None:
        side = _eval_fork[dev] = torch.cuda.Stream(dev)
    priors = PriorSet.of(model)
    packed = pack_targets(targets, dev)
    # one conf_all feeds both halves: the loss kernel's logit stream leaves predict's row keys in `share` (ssdhot.h)
    B, P, _ = conf_all.shape
    share = _workspace("share", dev, _lib.lib().ssdhot_share_bytes(B, P))
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().ssdhot_share_reset(share.data_ptr(), B, _stream(dev)), "ssdhot_share_reset")
    side.wait_stream(cur)
    l_loc, l_conf = multibox_loss(priors, loc_all, conf_all, packed, iou_thresh, neg_pos_ratio, H, W, group, _share=share)
    with torch.cuda.stream(side):
        labels, scores, boxes, count = predict_padded(priors, loc_all, conf_all, score_thresh, nms_thresh, max_per_img,
                                                      class_agnostic, metric, _share=share)
    cur.wait_stream(side)
    for t in (loc_all, conf_all):
        t.record_stream(side)
    for t in (labels, scores, boxes, count):
        t.record_stream(cur)
    return l_loc, l_conf, labels, scores, boxes, count


# ------------------------------------------------------------------------------------------------
# head-output packing (the tail of mySSD.forward)
# ------------------------------------------------------------------------------------------------
_LEVELS = ((38, 4), (19, 6), (10, 6), (5, 6), (3, 4), (1, 4))


def pack_heads(loc_heads: Sequence[torch.Tensor], conf_heads: Sequence[torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor]:
    """Drop-in for the tail of mySSD.forward (SSD_from_scratch.py:249-269): the six NCHW outputs of the box heads
    [B, A*4, H, W] and of the class heads [B, A*C, H, W] -> (loc_all [B,8732,4], conf_all [B,8732,C]), one launch per
    branch instead of 12 permute().contiguous() + 2 cat."""
    import ctypes
    outs = []
    for heads, what in ((loc_heads, "loc"), (conf_heads, "conf")):
        if len(heads) != len(_LEVELS):
            raise ValueError(f"expected {len(_LEVELS)} {what} head outputs, got {len(heads)}")
        dev = _need_cuda(*heads)
        B = int(heads[0].shape[0])
        D = int(heads[0].shape[1]) // _LEVELS[0][1]
        hs = []
        for h, (side, shapes) in zip(heads, _LEVELS):
            if tuple(h.shape) != (B, shapes * D, side, side):
                raise ValueError(f"{what} head of the {side}x{side} level has shape {tuple(h.shape)}, expected {(B, shapes * D, side, side)}")
            hs.append(h.detach().to(torch.float32).contiguous())
        out = torch.empty((B, 8732, D), dtype=torch.float32, device=dev)
        ptrs = (ctypes.c_void_p * len(hs))(*[h.data_ptr() for h in hs])
        with torch.cuda.device(dev):
            rc = _lib.lib().ssdhot_pack_heads(ctypes.cast(ptrs, ctypes.c_void_p), B, D, out.data_ptr(), _stream(dev))
        _lib.check(rc, "ssdhot_pack_heads")
        outs.append(out)
    return outs[0], outs[1]


# ------------------------------------------------------------------------------------------------
# the hot path straight from the head outputs (no permute / cat / pack pass at all)
# ------------------------------------------------------------------------------------------------
HEADS_NCHW, HEADS_NHWC = 0, 1
_SSD_TRUNK = ("VGG16_UpTo_conv4_3", "VGG16_extras", "extra_conv6", "extra_conv7", "extra_conv8_2", "extra_conv9_2",
              "extra_conv10_2", "extra_conv11_2", "box_head", "cls_head")


def _has_ssd_heads(model) -> bool:
    return all(hasattr(model, name) for name in _SSD_TRUNK)


def forward_heads(model, x: torch.Tensor) -> Tuple[List[torch.Tensor], List[torch.Tensor]]:
    """mySSD.forward up to the head convolutions (SFS:236-262), through the model's own modules, WITHOUT the
    permute(0,2,3,1).contiguous() x 12 + cat x 2 tail (SFS:249-269): -> (six box-head outputs [B, A*4, H, W], six
    class-head outputs [B, A*C, H, W]) for predict_heads / multibox_loss_heads.  model(x) == pack_heads(*forward_heads(model, x)).
    The backbone and the heads stay PyTorch modules (out of this path's scope)."""
    f0 = model.VGG16_UpTo_conv4_3(x)
    f1 = model.extra_conv7(model.extra_conv6(model.VGG16_extras(f0)))
    f2 = model.extra_conv8_2(f1)
    f3 = model.extra_conv9_2(f2)
    f4 = model.extra_conv10_2(f3)
    f5 = model.extra_conv11_2(f4)
    feats = (f0, f1, f2, f3, f4, f5)
    return [h(f) for h, f in zip(model.box_head, feats)], [h(f) for h, f in zip(model.cls_head, feats)]


class HeadSet:
    """The 2 x 6 head outputs of a batch ([B, A*D, side, side] each, SFS:249-262), validated, with the two HOST arrays of
    DEVICE pointers the head-direct entry points take (ssdhot_predict_heads, ssdhot_multibox_loss_heads_fwd/_bwd).
    Heads that are channels_last in memory are read as NHWC rows (layout HEADS_NHWC), everything else as NCHW planes
    (HEADS_NCHW); a head that is neither (or is not 16-byte aligned) is made contiguous first.  The object keeps the
    tensors alive; prepare it once when the same buffers are launched repeatedly (HotPathStep.launch_*_heads)."""

    def __init__(self, loc_heads: Sequence[torch.Tensor], conf_heads: Sequence[torch.Tensor]):
        import ctypes
        if len(loc_heads) != len(_LEVELS) or len(conf_heads) != len(_LEVELS):
            raise ValueError(f"expected {len(_LEVELS)} loc and conf head outputs, got {len(loc_heads)} and {len(conf_heads)}")
        self.device = _need_cuda(*loc_heads, *conf_heads)
        self.B = int(conf_heads[0].shape[0])
        self.C = int(conf_heads[0].shape[1]) // _LEVELS[0][1]
        for heads, D, what in ((loc_heads, 4, "loc"), (conf_heads, self.C, "conf")):
            for h, (side, shapes) in zip(heads, _LEVELS):
                if tuple(h.shape) != (self.B, shapes * D, side, side):
                    raise ValueError(f"{what} head of the {side}x{side} level has shape {tuple(h.shape)}, "
                                     f"expected {(self.B, shapes * D, side, side)}")
        every = list(loc_heads) + list(conf_heads)
        nhwc = all(h.is_contiguous(memory_format=torch.channels_last) for h in every) and not all(h.is_contiguous() for h in every)
        fmt = torch.channels_last if nhwc else torch.contiguous_format
        self.layout = HEADS_NHWC if nhwc else HEADS_NCHW
        self.tensors = []
        for h in every:
            t = h.detach().to(torch.float32).contiguous(memory_format=fmt)
            if t.data_ptr() % 16:
                t = t.clone(memory_format=fmt)
            self.tensors.append(t)
        self._loc_array = (ctypes.c_void_p * 6)(*[t.data_ptr() for t in self.tensors[:6]])
        self._conf_array = (ctypes.c_void_p * 6)(*[t.data_ptr() for t in self.tensors[6:]])
        self.loc_ptr = ctypes.cast(self._loc_array, ctypes.c_void_p)
        self.conf_ptr = ctypes.cast(self._conf_array, ctypes.c_void_p)

    def used_on(self, stream: "torch.cuda.Stream") -> None:
        """Tell the caching allocator that `stream` reads the (possibly re-laid-out) tensors."""
        for t in self.tensors:
            t.record_stream(stream)


def predict_heads_padded(model, loc_heads: Sequence[torch.Tensor], conf_heads: Sequence[torch.Tensor], score_thresh: float = 0.2,
                         nms_thresh: float = 0.5, max_per_img: int = 100, class_agnostic: bool = False,
                         metric: str = "diou", want_cand: bool = False):
    """predict_padded from the six head outputs of each branch instead of (loc_all, conf_all): the tail of mySSD.forward
    (12 permute().contiguous() + 2 cat, SFS:249-269) never runs -- score_kernel / nms_image_kernel address the heads
    directly.  Same results as predict_padded(model, *pack_heads(loc_heads, conf_heads), ...), bit for bit.
    Class counts other than 6 are packed first (ssdhot_pack_heads) and take the packed entry point."""
    if not (0.0 <= score_thresh < 1.0):
        raise ValueError(f"Score threshold should be greater than 0 and less than 1, recieved {score_thresh}.")
    if not (0.0 < nms_thresh < 1.0):
        raise ValueError(f"NMS threshold should be greater than 0 and less than 1, recieved {nms_thresh}.")
    hs = loc_heads if isinstance(loc_heads, HeadSet) else HeadSet(loc_heads, conf_heads)
    dev, B, C = hs.device, hs.B, hs.C
    priors = PriorSet.of(model)
    assert priors.P == 8732
    assert C >= 2
    if C != 6:
        return predict_padded(model, *pack_heads(hs.tensors[:6], hs.tensors[6:]), score_thresh, nms_thresh, max_per_img,
                              class_agnostic, metric, want_cand)
    labels = torch.empty((B, max_per_img), dtype=torch.int64, device=dev)
    scores = torch.empty((B, max_per_img), dtype=torch.float32, device=dev)
    boxes = torch.empty((B, max_per_img, 4), dtype=torch.float32, device=dev)
    cand = torch.empty((B, max_per_img), dtype=torch.int32, device=dev) if want_cand else None
    count = torch.empty((B,), dtype=torch.int32, device=dev)
    work = _workspace("predict", dev, _lib.lib().ssdhot_predict_workspace_bytes(B, 8732, C))
    with torch.cuda.device(dev):
        rc = _lib.lib().ssdhot_predict_heads(priors.priors.data_ptr(), hs.loc_ptr, hs.conf_ptr, hs.layout, B, C,
                                             float(score_thresh), float(nms_thresh), int(max_per_img), 1 if class_agnostic else 0,
                                             METRICS[metric], priors.variances[0], priors.variances[1],
                                             float(priors.img_w), float(priors.img_h),
                                             labels.data_ptr(), scores.data_ptr(), boxes.data_ptr(), _ptr(cand),
                                             count.data_ptr(), work.data_ptr(), 3, None, _stream(dev))
    _lib.check(rc, "ssdhot_predict_heads")
    hs.used_on(torch.cuda.current_stream(dev))
    if want_cand:
        return labels, scores, boxes, count, cand
    return labels, scores, boxes, count


@torch.no_grad()
def predict_heads(model, loc_heads: Sequence[torch.Tensor], conf_heads: Sequence[torch.Tensor], score_thresh: float = 0.2,
                  nms_thresh: float = 0.5, max_per_img: int = 100, class_agnostic: bool = False,
                  metric: str = "diou") -> List[Dict[str, torch.Tensor]]:
    """mySSD.predict (SFS:338-476) fed with the head outputs: List[Dict] exactly as predict returns."""
    labels, scores, boxes, count = predict_heads_padded(model, loc_heads, conf_heads, score_thresh, nms_thresh, max_per_img,
                                                        class_agnostic, metric)
    ks = count.tolist()
    return [{"labels": l[:k], "scores": s[:k], "boxes": b[:k]}
            for l, s, b, k in zip(labels.unbind(0), scores.unbind(0), boxes.unbind(0), ks)]


class _FusedLossHeads(torch.autograd.Function):
    """_FusedLoss on the head layouts: forward = ssdhot_multibox_loss_heads_fwd, backward = ssdhot_multibox_loss_heads_bwd,
    which writes the gradients as twelve tensors laid out like the inputs."""

    @staticmethod
    def forward(ctx, priors: PriorSet, packed: PackedTargets, iou_thresh, ratio, norm_wh, group, *heads):
        hs = HeadSet(heads[:6], heads[6:])
        dev, B, C, layout, keep = hs.device, hs.B, hs.C, hs.layout, hs.tensors
        need_grad = any(h.requires_grad for h in heads)
        sums = torch.empty((3,), dtype=torch.float64, device=dev)
        sel = torch.empty((B, 8732), dtype=torch.int8, device=dev) if need_grad else None
        matched = torch.empty((B, 8732), dtype=torch.int16, device=dev) if need_grad else None
        work = _workspace("loss", dev, _lib.lib().ssdhot_loss_workspace_bytes(B, 8732, packed.max_gt))
        with torch.cuda.device(dev):
            rc = _lib.lib().ssdhot_multibox_loss_heads_fwd(
                priors.priors.data_ptr(), priors.priors_xyxy.data_ptr(), priors.aux.data_ptr(), priors.layout,
                packed.boxes.data_ptr(), packed.labels.data_ptr(), packed.offsets.data_ptr(), B, packed.max_gt,
                float(norm_wh[0]), float(norm_wh[1]), hs.loc_ptr, hs.conf_ptr, layout, C,
                float(iou_thresh), priors.variances[0], priors.variances[1], float(ratio),
                sums.data_ptr(), work.data_ptr(), _ptr(sel), _ptr(matched), None, None, None, _stream(dev))
        _lib.check(rc, "ssdhot_multibox_loss_heads_fwd")
        hs.used_on(torch.cuda.current_stream(dev))
        if group is not None:
            _check_group_for_grad(group, need_grad)
            _dist.reduce_sums(sums, group)
        total = sums[2].clamp_min(1.0)
        if need_grad:
            ctx.save_for_backward(sel, matched, total, packed.boxes, packed.offsets, *keep)
            ctx.meta = (priors, norm_wh, layout, B, C)
        losses = (sums[:2] / total).to(torch.float32)
        ctx.mark_non_differentiable(sums)
        return losses[0], losses[1], sums

    @staticmethod
    def backward(ctx, g_loc, g_conf, _g_sums):
        import ctypes
        sel, matched, total, gt_boxes, gt_offsets, *keep = ctx.saved_tensors
        priors, norm_wh, layout, B, C = ctx.meta
        dev = sel.device
        scales = torch.stack((g_loc.to(torch.float64) / total, g_conf.to(torch.float64) / total))
        want_loc, want_conf = any(ctx.needs_input_grad[6:12]), any(ctx.needs_input_grad[12:18])
        d_loc = [torch.empty_like(t) for t in keep[:6]] if want_loc else None
        d_conf = [torch.empty_like(t) for t in keep[6:]] if want_conf else None

        def arr(ts):
            if ts is None:
                return None, None
            a = (ctypes.c_void_p * 6)(*[t.data_ptr() for t in ts])
            return a, ctypes.cast(a, ctypes.c_void_p)
        (a0, lp), (a1, cp), (a2, glp), (a3, gcp) = arr(keep[:6]), arr(keep[6:]), arr(d_loc), arr(d_conf)
        with torch.cuda.device(dev):
            rc = _lib.lib().ssdhot_multibox_loss_heads_bwd(
                priors.priors.data_ptr(), gt_boxes.data_ptr(), gt_offsets.data_ptr(), B, float(norm_wh[0]), float(norm_wh[1]),
                lp, cp, layout, C, priors.variances[0], priors.variances[1], sel.data_ptr(), matched.data_ptr(),
                scales.data_ptr(), glp, gcp, _stream(dev))
        _lib.check(rc, "ssdhot_multibox_loss_heads_bwd")
        grads = (d_loc or [None] * 6) + (d_conf or [None] * 6)
        return (None,) * 6 + tuple(grads)


def multibox_loss_heads(model, loc_heads: Sequence[torch.Tensor], conf_heads: Sequence[torch.Tensor], targets,
                        iou_thresh: float = 0.5, neg_pos_ratio: float = 3.0, H: int = 300, W: int = 300, group=None,
                        return_sums: bool = False):
    """multibox_loss from the six head outputs of each branch: the post-backbone training step (SSD_trainer.py:92-117)
    without ever forming loc_all / conf_all -- differentiable w.r.t. the twelve head tensors, whose gradients come back in
    the heads' own memory layout.  Same sums as multibox_loss on the packed tensors, bit for bit.  Inputs the head kernel
    does not cover (C != 6, more than 64 boxes per image, non-SSD300 priors) go through the forward's own permute + cat
    (differentiable) and the packed entry point."""
    if not (0.0 < iou_thresh < 1.0):
        raise ValueError(f"Score threshold should be greater than 0 and less than 1, recieved {iou_thresh}.")
    priors = PriorSet.of(model)
    packed = pack_targets(targets, priors.device)
    if len(loc_heads) != len(_LEVELS) or len(conf_heads) != len(_LEVELS):
        raise ValueError(f"expected {len(_LEVELS)} loc and conf head outputs, got {len(loc_heads)} and {len(conf_heads)}")
    C = int(conf_heads[0].shape[1]) // _LEVELS[0][1]
    if C != 6 or priors.layout != 1 or packed.max_gt > 64:           # (1 = SSDHOT_LAYOUT_SSD300)
        B = int(conf_heads[0].shape[0])
        loc_all = torch.cat([h.permute(0, 2, 3, 1).contiguous().view(B, -1, 4) for h in loc_heads], 1)       # SFS:249-269
        conf_all = torch.cat([h.permute(0, 2, 3, 1).contiguous().view(B, -1, C) for h in conf_heads], 1)
        return multibox_loss(model, loc_all, conf_all, packed, iou_thresh, neg_pos_ratio, H, W, group, return_sums)
    l_loc, l_conf, sums = _FusedLossHeads.apply(priors, packed, float(iou_thresh), float(neg_pos_ratio), (W, H), group,
                                                *loc_heads, *conf_heads)
    if return_sums:
        return l_loc, l_conf, sums
    return l_loc, l_conf


# ------------------------------------------------------------------------------------------------
# patching the reference in place
# ------------------------------------------------------------------------------------------------
def patch(model=None, trainer_module=None, steps: bool = True):
    """Route a reference `mySSD` instance and/or the imported `SSD_trainer` module through ssdhot:
    model.encode_ssd / decode_ssd / iou_nms / predict and trainer.build_targets / CELoss_w_neg_mining /
    collate_detection keep their signatures.  With `steps` the module's SSD_train_step /
    SSD_test_step are replaced by the fused forms (ssdhot/trainer.py: one launch for targets + both losses, straight
    from the head outputs); the reference's own step functions also run unchanged on the patched callables."""
    import types
    if model is not None:
        model.encode_ssd = types.MethodType(lambda self, *a, **k: encode_ssd(self, *a, **k), model)
        model.predict = types.MethodType(lambda self, *a, **k: predict(self, *a, **k), model)
        model.decode_ssd = decode_ssd
        model.iou_nms = iou_nms
    if trainer_module is not None:
        from . import trainer as _trainer
        trainer_module.build_targets = build_targets
        trainer_module.CELoss_w_neg_mining = CELoss_w_neg_mining
        trainer_module.collate_detection = collate_detection
        if steps:
            trainer_module.SSD_train_step = _trainer.SSD_train_step
            trainer_module.SSD_test_step = _trainer.make_test_step(trainer_module if hasattr(trainer_module, "MeanAveragePrecision") else None)
    return model
