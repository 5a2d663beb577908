"""ORACLE -- test infrastructure only, never the product path.

A restatement of the SSD300 multibox post-backbone hot path of
ElliotBlackstone/automotive-ssd-object-detection, written against plain torch tensor
ops so that it reproduces the reference's fp32 arithmetic op for op.  It is imported only
by ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs, and only as the checker (or the timed CPU baseline) -- the
shipped package ``ssdhot`` never imports it and has no CPU fallback.

Parity pin: the reference ships no tests or golden vectors of its own, so
the oracle is pinned against outputs of the *reference itself*, imported unmodified from
/root/reference in the build container by ``tests/golden/make_golden.py``; the resulting
fixtures live in ``tests/golden/*.npz`` and ``tests/test_oracle_golden.py`` asserts
bit-for-bit equality of this file's outputs with them (CPU).

The box arithmetic of the reference lives in a third-party dependency that is not vendored
under /root/reference: torchvision (pinned 0.20.1 in the reference's requirements.txt:83,
0.24.1 in app_files/requirements.txt:26; 0.26.0 in this image), functions
``box_convert`` (ops/_box_convert.py:5-47), ``box_area`` / ``_box_inter_union`` /
``box_iou`` (ops/boxes.py:273-371), ``_box_diou_iou`` (ops/boxes.py:462-480),
``distance_box_iou`` (:437-459) and ``complete_box_iou`` (:404-434).  Their published
algorithm is restated below, one separately-rounded fp32 op per line, and
``tests/test_oracle_golden.py`` additionally checks the restatement bit-for-bit against the
image's torchvision.

Every function is device-agnostic: run on CPU tensors it is the CPU oracle; run on CUDA
tensors (GPU box only) the very same code is the *device-matched* oracle -- eager torch
CUDA kernels use IEEE +,-,*,/ and CUDA libdevice atanf/expf/logf, which is what the
ssdhot kernels are written to reproduce bit for bit.

Reference line numbers cite /root/reference/SSD_from_scratch.py (``SFS``) and
/root/reference/SSD_trainer.py (``TR``).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

NUM_PRIORS = 8732
IMG_SIZE = 300
FMAP_SIDES = (38, 19, 10, 5, 3, 1)
EXTRA_RATIOS = ((2,), (2, 3), (2, 3), (2, 3), (2,), (2,))
PRIORS_SHA256 = "b4314e383aa627d8707e5e5fdb65712f96024e2b0b14d45dbe14deb3d1ed0b34"
PRIORS_XYXY_SHA256 = "308e341f2bdeed01516e5f802eeb2de321414055e7e187bf25415ddfb4b4bd6c"


# --------------------------------------------------------------------------------------
# f.3 head-output packing (the tail of mySSD.forward)                          SFS:249-269
# --------------------------------------------------------------------------------------
def pack_heads(loc_heads: Sequence[torch.Tensor], conf_heads: Sequence[torch.Tensor], num_classes: int):
    """Six NCHW head outputs per branch -> (loc_all [B,8732,4], conf_all [B,8732,C]) exactly as the reference writes it:
    permute(0,2,3,1).contiguous() per level (SFS:249-262), flatten + cat (SFS:265-266), view (SFS:268-269)."""
    loc_list = [o.permute(0, 2, 3, 1).contiguous() for o in loc_heads]
    cls_list = [o.permute(0, 2, 3, 1).contiguous() for o in conf_heads]
    loc_output = torch.cat([o.view(o.size(0), -1) for o in loc_list], 1)
    cls_output = torch.cat([o.view(o.size(0), -1) for o in cls_list], 1)
    return loc_output.view(loc_output.size(0), -1, 4), cls_output.view(cls_output.size(0), -1, num_classes)


# --------------------------------------------------------------------------------------
# a1. default boxes                                                     SFS:275-331, :32-35
# --------------------------------------------------------------------------------------
def default_boxes(s_min: float = 0.2, s_max: float = 0.9) -> torch.Tensor:
    """8732 x (cx, cy, w, h), built in Python float64 and cast once to fp32 (SFS:325).

    Level l has scale s_l = s_min + (s_max - s_min) * l / 5; per cell the shapes are
    [s_l square, sqrt(s_l s_{l+1}) square, then for each extra ratio a: (s sqrt a, s / sqrt a)
    and its transpose] (SFS:304-315); cells are visited row-major (SFS:318-323);
    centres are clamped to [0,1] and sizes to [1e-6,1] after the fp32 cast (SFS:326-330).
    """
    n_levels = len(FMAP_SIDES)
    scales = [s_min + (s_max - s_min) * (l / (n_levels - 1)) for l in range(n_levels)] + [1.0]
    rows: List[List[float]] = []
    for l, side in enumerate(FMAP_SIDES):
        shapes = [(scales[l], scales[l])]
        mid = float(math.sqrt(scales[l] * scales[l + 1]))
        shapes.append((mid, mid))
        for a in EXTRA_RATIOS[l]:
            r = float(math.sqrt(a))
            shapes.append((scales[l] * r, scales[l] / r))
            shapes.append((scales[l] / r, scales[l] * r))
        for iy in range(side):
            cy = (iy + 0.5) / side
            for ix in range(side):
                cx = (ix + 0.5) / side
                rows.extend([cx, cy, w, h] for (w, h) in shapes)
    out = torch.tensor(rows, dtype=torch.float32)
    out[:, :2].clamp_(0.0, 1.0)
    out[:, 2:].clamp_(1e-6, 1.0)
    return out


def cxcywh_to_xyxy(b: torch.Tensor) -> torch.Tensor:
    """torchvision ops/_box_convert.py:5-24 -- x1 = cx - 0.5*w (a multiply, then a subtract)."""
    cx, cy, w, h = b.unbind(-1)
    hw = 0.5 * w
    hh = 0.5 * h
    return torch.stack((cx - hw, cy - hh, cx + hw, cy + hh), dim=-1)


def xyxy_to_cxcywh(b: torch.Tensor) -> torch.Tensor:
    """torchvision ops/_box_convert.py:27-47 -- cx = (x1 + x2) / 2, w = x2 - x1."""
    x1, y1, x2, y2 = b.unbind(-1)
    return torch.stack(((x1 + x2) / 2, (y1 + y2) / 2, x2 - x1, y2 - y1), dim=-1)


def prior_tables(device="cpu") -> Tuple[torch.Tensor, torch.Tensor]:
    """(priors cxcywh, priors xyxy clamped to [0,1]) as registered at SFS:32-35."""
    pri = default_boxes()
    pri_xyxy = cxcywh_to_xyxy(pri).clamp(0, 1)
    return pri.to(device), pri_xyxy.to(device)


# --------------------------------------------------------------------------------------
# a2'. pairwise IoU / DIoU / CIoU                          torchvision ops/boxes.py:273-480
# --------------------------------------------------------------------------------------
def _area(b: torch.Tensor) -> torch.Tensor:
    return (b[..., 2] - b[..., 0]) * (b[..., 3] - b[..., 1])          # boxes.py:296


def pairwise_iou(rows: torch.Tensor, cols: torch.Tensor) -> torch.Tensor:
    """[N,4] x [M,4] -> [N,M]; boxes.py:308-371."""
    a_r = _area(rows)
    a_c = _area(cols)
    lo = torch.max(rows[:, None, :2], cols[None, :, :2])              # :321
    hi = torch.min(rows[:, None, 2:], cols[None, :, 2:])              # :322
    ext = (hi - lo).clamp(min=0)                                      # :336
    inter = ext[..., 0] * ext[..., 1]                                 # :337
    union = a_r[:, None] + a_c[None, :] - inter                       # :339
    return inter / union                                              # :370


def pairwise_diou(rows: torch.Tensor, cols: torch.Tensor, eps: float = 1e-7):
    """-> (diou, iou), both [N,M]; boxes.py:462-480."""
    iou = pairwise_iou(rows, cols)
    lo = torch.min(rows[:, None, :2], cols[None, :, :2])              # :465
    hi = torch.max(rows[:, None, 2:], cols[None, :, 2:])              # :466
    hull = (hi - lo).clamp(min=0)                                     # :467
    diag2 = (hull[..., 0] ** 2) + (hull[..., 1] ** 2) + eps           # :468
    xr = (rows[:, 0] + rows[:, 2]) / 2                                # :470-473
    yr = (rows[:, 1] + rows[:, 3]) / 2
    xc = (cols[:, 0] + cols[:, 2]) / 2
    yc = (cols[:, 1] + cols[:, 3]) / 2
    dist2 = ((xr[:, None] - xc[None, :]) ** 2) + ((yr[:, None] - yc[None, :]) ** 2)   # :475-477
    return iou - (dist2 / diag2), iou                                 # :480


def pairwise_ciou(rows: torch.Tensor, cols: torch.Tensor, eps: float = 1e-7) -> torch.Tensor:
    """[N,M] complete IoU; boxes.py:404-434.  atan is taken on [N,1] / [1,M] before the
    broadcast, i.e. it is a per-row / per-column constant."""
    diou, iou = pairwise_diou(rows, cols, eps)
    w_r = rows[:, None, 2] - rows[:, None, 0]
    h_r = rows[:, None, 3] - rows[:, None, 1]
    w_c = cols[None, :, 2] - cols[None, :, 0]
    h_c = cols[None, :, 3] - cols[None, :, 1]
    v = (4 / (torch.pi ** 2)) * torch.pow(torch.atan(w_r / h_r) - torch.atan(w_c / h_c), 2)   # :430
    alpha = v / (1 - iou + v + eps)                                   # :432
    return diou - alpha * v                                           # :433


# --------------------------------------------------------------------------------------
# a2. match + encode one image                                              SFS:697-773
# --------------------------------------------------------------------------------------
def match_encode(priors: torch.Tensor, priors_xyxy: torch.Tensor, gt_xyxy: torch.Tensor,
                 gt_labels: torch.Tensor, iou_thresh: float = 0.5,
                 variances: Tuple[float, float] = (0.1, 0.2), background_class: int = 0,
                 return_match: bool = False):
    """-> (loc_target [P,4], cls_target [P], pos_mask [P] bool, matched_gt_cxcywh [P,4])
    (+ best_gt [P] i64, best_val [P] f32 when ``return_match``)."""
    if background_class != 0:                                         # SFS:719-720
        raise ValueError(f"Background should be 0, recieved {background_class}.")
    n_pri = priors.shape[0]
    n_gt = gt_xyxy.shape[0]
    dev = priors.device
    if n_gt == 0:                                                     # SFS:731-736
        res = (torch.zeros((n_pri, 4), dtype=priors.dtype, device=dev),
               torch.full((n_pri,), background_class, dtype=gt_labels.dtype, device=dev),
               torch.zeros((n_pri,), dtype=torch.bool, device=dev),
               torch.zeros((n_pri, 4), dtype=priors.dtype, device=dev))
        if return_match:
            res = res + (torch.zeros((n_pri,), dtype=torch.int64, device=dev),
                         torch.zeros((n_pri,), dtype=priors.dtype, device=dev))
        return res
    q = pairwise_ciou(priors_xyxy, gt_xyxy)                           # SFS:744  [P,G]
    champion = q.argmax(dim=0)                                        # SFS:746  best prior per GT
    q[champion, torch.arange(n_gt, device=dev)] = 2.0                 # SFS:747  forced match
    best_gt = q.argmax(dim=1)                                         # SFS:749
    best_val = q.gather(1, best_gt.view(-1, 1)).squeeze(1)            # SFS:750
    pos = best_val >= iou_thresh                                      # SFS:751
    g = xyxy_to_cxcywh(gt_xyxy)[best_gt]                              # SFS:754-755
    v_c, v_s = variances
    t_xy = (g[:, :2] - priors[:, :2]) / priors[:, 2:] / v_c           # SFS:759
    t_wh = torch.log((g[:, 2:] / priors[:, 2:]).clamp(min=1e-12)) / v_s   # SFS:760-762
    loc_t = torch.cat((t_xy, t_wh), dim=1)                            # SFS:764-766
    lab = gt_labels[best_gt]
    cls_t = torch.full((n_pri,), background_class, dtype=lab.dtype, device=dev)
    cls_t[pos] = lab[pos] + 1                                         # SFS:769-771
    if return_match:
        return loc_t, cls_t, pos, g, best_gt, best_val
    return loc_t, cls_t, pos, g


# --------------------------------------------------------------------------------------
# a3. batch targets                                                          TR:491-547
# --------------------------------------------------------------------------------------
def batch_targets(priors: torch.Tensor, priors_xyxy: torch.Tensor, targets: Sequence[Dict],
                  H: int = 300, W: int = 300, iou_thresh: float = 0.5,
                  variances: Tuple[float, float] = (0.1, 0.2), dense: bool = False):
    """-> (pos_mask [B,P] bool, loc_t[pos_mask] [N_pos,4], cls_t [B,P] i64)
    (``dense`` additionally returns the un-compacted loc_t [B,P,4])."""
    if not (0.0 < iou_thresh < 1.0):                                  # TR:516-517
        raise ValueError(f"Score threshold should be greater than 0 and less than 1, recieved {iou_thresh}.")
    dev = priors.device
    scale = torch.tensor([W, H, W, H], device=dev, dtype=torch.float32)   # TR:519
    loc_l, cls_l, pos_l = [], [], []
    for t in targets:                                                 # TR:525
        px = t["boxes"]
        if px.numel() == 0:                                           # TR:529-530
            unit = px.new_zeros((0, 4))
        else:
            unit = px / scale                                         # TR:532 (tensor / tensor)
        loc_t, cls_t, pos, _ = match_encode(priors, priors_xyxy, unit, t["labels"], iou_thresh, variances)
        loc_l.append(loc_t)
        cls_l.append(cls_t)
        pos_l.append(pos)
    loc_t = torch.stack(loc_l, 0)
    cls_t = torch.stack(cls_l, 0)
    pos = torch.stack(pos_l, 0)
    if dense:
        return pos, loc_t[pos], cls_t, loc_t
    return pos, loc_t[pos], cls_t                                     # TR:547


# --------------------------------------------------------------------------------------
# a4 / a5. losses                                              TR:104-108, TR:551-600
# --------------------------------------------------------------------------------------
def loc_loss(loc_all: torch.Tensor, pos: torch.Tensor, loc_t_pm: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """-> (smooth-L1 sum / total_pos, num_pos_per_img [B], total_pos 0-d f32)."""
    n_img = pos.sum(dim=1)                                            # TR:104
    total = n_img.sum().clamp_min(1).float()                          # TR:105
    val = F.smooth_l1_loss(loc_all[pos], loc_t_pm, reduction="sum") / total   # TR:108
    return val, n_img, total


def mined_negative_budget(n_pos: int, n_neg: int, ratio: float) -> int:
    """How many negatives one image contributes (TR:585-596)."""
    want = int(ratio) if n_pos == 0 else int(ratio * n_pos)
    if n_neg == 0 or want == 0:
        return 0
    return min(want, n_neg)


def mined_ce_loss(conf_all: torch.Tensor, cls_t: torch.Tensor, pos: torch.Tensor,
                  n_img: torch.Tensor, total, ratio: float = 3.0,
                  return_parts: bool = False):
    """Softmax cross-entropy over positives + the `ratio`:1 hardest negatives per image."""
    n_b, n_p, n_c = conf_all.shape
    ce = F.cross_entropy(conf_all.view(-1, n_c), cls_t.view(-1), reduction="none").view(n_b, n_p)   # TR:577
    acc_pos = ce[pos].sum()                                           # TR:580
    acc_neg = torch.zeros((), device=conf_all.device)
    for b in range(n_b):                                              # TR:584
        neg = ce[b].masked_select(~pos[b])                            # TR:592
        k = mined_negative_budget(int(n_img[b].item()), neg.numel(), ratio)
        if k == 0:
            continue
        vals, _ = torch.topk(neg, k, largest=True, sorted=False)      # TR:597
        acc_neg += vals.sum()                                         # TR:598
    out = (acc_pos + acc_neg) / total                                 # TR:600
    if return_parts:
        return out, ce, acc_pos, acc_neg
    return out


def train_half(priors, priors_xyxy, loc_all, conf_all, targets, iou_thresh=0.5, ratio=3.0,
               H=300, W=300, variances=(0.1, 0.2)):
    """The post-backbone part of one training step (TR:92-117): -> (loc_loss, conf_loss)."""
    pos, loc_t_pm, cls_t = batch_targets(priors, priors_xyxy, targets, H, W, iou_thresh, variances)
    l_loc, n_img, total = loc_loss(loc_all, pos, loc_t_pm)
    l_conf = mined_ce_loss(conf_all, cls_t, pos, n_img, total, ratio)
    return l_loc, l_conf


# --------------------------------------------------------------------------------------
# a6. decode                                                                SFS:776-800
# --------------------------------------------------------------------------------------
def decode(loc: torch.Tensor, priors: torch.Tensor, variances: Tuple[float, float]) -> torch.Tensor:
    v_c, v_s = variances
    cx = loc[:, 0] * v_c * priors[:, 2] + priors[:, 0]                # SFS:793
    cy = loc[:, 1] * v_c * priors[:, 3] + priors[:, 1]                # SFS:794
    w = priors[:, 2] * torch.exp(loc[:, 2] * v_s)                     # SFS:796
    h = priors[:, 3] * torch.exp(loc[:, 3] * v_s)                     # SFS:797
    return torch.stack((cx, cy, w, h), dim=1)


# --------------------------------------------------------------------------------------
# a8. greedy DIoU NMS                                                       SFS:664-692
# --------------------------------------------------------------------------------------
def greedy_nms(boxes: torch.Tensor, scores: torch.Tensor, thr: float, metric: str = "diou",
               limit: Optional[int] = None) -> torch.Tensor:
    """Indices of survivors, score-descending.  A box survives a kept box iff
    ``metric(kept, box) <= thr`` (NaN therefore suppresses).  The reference's argsort is
    unstable (SFS:677); here equal scores are ordered by ascending input index, which is the
    documented ssdhot tie rule.  ``limit`` stops after that many survivors (lossless for the
    post-NMS top-k of SFS:465, used only to bound oracle run time)."""
    if boxes.numel() == 0:                                            # SFS:674-675
        return boxes.new_zeros((0,), dtype=torch.long)
    queue = scores.argsort(descending=True, stable=True)              # SFS:677
    kept = []
    while queue.numel() > 0:                                          # SFS:680
        top = queue[0]
        kept.append(top)
        if queue.numel() == 1 or (limit is not None and len(kept) >= limit):
            break
        rest = queue[1:]
        if metric == "diou":
            sim = pairwise_diou(boxes[top].unsqueeze(0), boxes[rest])[0].squeeze(0)   # SFS:688
        elif metric == "ciou":
            sim = pairwise_ciou(boxes[top].unsqueeze(0), boxes[rest]).squeeze(0)
        else:
            raise ValueError(metric)
        queue = rest[sim <= thr]                                      # SFS:690
    return torch.stack(kept)


# --------------------------------------------------------------------------------------
# a7. post-process                                                          SFS:338-476
# --------------------------------------------------------------------------------------
def postprocess(priors: torch.Tensor, loc_all: torch.Tensor, conf_all: torch.Tensor,
                score_thresh: float = 0.2, nms_thresh: float = 0.5, max_per_img: int = 100,
                class_agnostic: bool = False, variances: Tuple[float, float] = (0.1, 0.2),
                H: int = IMG_SIZE, W: int = IMG_SIZE, metric: str = "diou",
                with_index: bool = False, nms_limit: bool = False) -> List[Dict[str, torch.Tensor]]:
    """-> per image {'labels' i64[K], 'scores' f32[K], 'boxes' f32[K,4] px xyxy}
    (+ 'cand' i64[K]: flat candidate id prior*(C-1)+class when ``with_index``).

    ``nms_limit=False`` runs every NMS to exhaustion exactly as the reference does (this is
    what the timed CPU baseline uses); ``nms_limit=True`` stops each NMS after ``max_per_img``
    survivors, which cannot change the output (SFS:465 keeps at most that many, in score
    order) and only shortens test run time."""
    cap = max_per_img if nms_limit else None
    if not (0.0 <= score_thresh < 1.0):                               # SFS:369-370
        raise ValueError(f"Score threshold should be greater than 0 and less than 1, recieved {score_thresh}.")
    if not (0.0 < nms_thresh < 1.0):                                  # SFS:372-373
        raise ValueError(f"NMS threshold should be greater than 0 and less than 1, recieved {nms_thresh}.")
    n_b, n_p, n_c = conf_all.shape
    assert n_p == priors.shape[0] and n_c >= 2                        # SFS:384-385
    dev = conf_all.device
    fg = conf_all.softmax(dim=-1)[..., 1:]                            # SFS:388
    results = []
    for b in range(n_b):                                              # SFS:397
        sc = fg[b]
        live = sc > score_thresh                                      # SFS:402 (strict)
        if not live.any():                                            # SFS:403-409
            item = {"labels": torch.empty(0, dtype=torch.int64, device=dev),
                    "scores": torch.empty(0, dtype=torch.float32, device=dev),
                    "boxes": priors.new_zeros((0, 4))}
            if with_index:
                item["cand"] = torch.empty(0, dtype=torch.int64, device=dev)
            results.append(item)
            continue
        pri_i, cls_i = live.nonzero(as_tuple=True)                    # SFS:412 (prior-major)
        box = decode(loc_all[b, pri_i], priors[pri_i], variances)     # SFS:415-419
        cx, cy, w, h = box.unbind(dim=1)
        x1 = (cx - 0.5 * w).clamp(0, 1) * W                           # SFS:422-425
        y1 = (cy - 0.5 * h).clamp(0, 1) * H
        x2 = (cx + 0.5 * w).clamp(0, 1) * W
        y2 = (cy + 0.5 * h).clamp(0, 1) * H
        xyxy = torch.stack((x1, y1, x2, y2), dim=1)
        s = sc[pri_i, cls_i]
        if class_agnostic:                                            # SFS:433-436
            keep = greedy_nms(xyxy, s, nms_thresh, metric, limit=cap)
        else:                                                         # SFS:439-463
            parts = []
            for c in range(n_c - 1):
                members = (cls_i == c).nonzero(as_tuple=True)[0]
                if members.numel() == 0:
                    continue
                local = greedy_nms(xyxy[members], s[members], nms_thresh, metric, limit=cap)
                parts.append(members[local])
            keep = torch.cat(parts)
            keep = keep[s[keep].argsort(descending=True, stable=True)]     # SFS:463
            # stable + ascending candidate id inside equal scores: `parts` is class-major, so
            # re-impose candidate order among exact score ties (documented tie rule).
            if keep.numel() > 1:
                sk = s[keep]
                if bool((sk[1:] == sk[:-1]).any()):
                    order = torch.tensor(sorted(range(keep.numel()),
                                                key=lambda i: (-float(sk[i]), int(keep[i]))), device=dev)
                    keep = keep[order]
        keep = keep[:max_per_img]                                     # SFS:465
        item = {"labels": cls_i[keep], "scores": s[keep], "boxes": xyxy[keep]}   # SFS:468-474
        if with_index:
            item["cand"] = pri_i[keep] * (n_c - 1) + cls_i[keep]
        results.append(item)
    return results
