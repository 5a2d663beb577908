"""Seeded synthetic inputs for the five BASELINE.json configurations.

Everything is generated on the CPU with ``torch.Generator().manual_seed(seed)`` so that the
CUDA path, the oracle and the reference arm all see the same bits; callers move the tensors
to the GPU themselves.  Shapes follow the reference: ``loc_all [B,8732,4]`` and
``conf_all [B,8732,C]`` fp32 as produced by ``mySSD.forward`` (SSD_from_scratch.py:265-271),
targets a ``List[Dict]`` with ``'boxes' [G,4]`` fp32 pixel xyxy and ``'labels' [G]`` int64 as
produced by ``collate_detection`` (SSD_trainer.py:806-813).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch

P = 8732
C = 6          # 5 foreground classes + background (SSD_from_scratch.py:25, ssd_demo_app.py:26)
IMG = 300.0


def make_targets(n_img: int, g_lo: int, g_hi: int, gen: torch.Generator, n_fg: int = C - 1) -> List[Dict[str, torch.Tensor]]:
    """Ragged ground truth: G_i ~ U{g_lo..g_hi} boxes per image; centres U(0,300)^2, sides
    300*exp(U(ln .03, ln .6)) (log-uniform: the data set's boxes are small), clipped to the
    image with every side >= 1 px; labels U{0..n_fg-1}."""
    out = []
    counts = torch.randint(g_lo, g_hi + 1, (n_img,), generator=gen).tolist() if g_hi > g_lo else [g_lo] * n_img
    for g in counts:
        if g == 0:
            out.append({"boxes": torch.zeros((0, 4), dtype=torch.float32), "labels": torch.zeros((0,), dtype=torch.int64)})
            continue
        ctr = torch.rand((g, 2), generator=gen) * IMG
        u = torch.rand((g, 2), generator=gen)
        side = IMG * torch.exp(math.log(0.03) + u * (math.log(0.6) - math.log(0.03)))
        lo = (ctr - 0.5 * side).clamp(0.0, IMG - 1.0)
        hi = torch.minimum(torch.maximum(ctr + 0.5 * side, lo + 1.0), torch.tensor(IMG))
        boxes = torch.cat((lo, hi), dim=1).to(torch.float32).contiguous()
        labels = torch.randint(0, n_fg, (g,), generator=gen, dtype=torch.int64)
        out.append({"boxes": boxes, "labels": labels})
    return out


def make_heads(n_img: int, gen: torch.Generator, bg_bias: float = 0.0, n_cls: int = C) -> Tuple[torch.Tensor, torch.Tensor]:
    """Random head outputs: loc ~ N(0,1) [B,P,4], conf ~ N(0,1) [B,P,C] (+bg_bias on class 0)."""
    loc = torch.randn((n_img, P, 4), generator=gen, dtype=torch.float32)
    conf = torch.randn((n_img, P, n_cls), generator=gen, dtype=torch.float32)
    if bg_bias != 0.0:
        conf[..., 0] += bg_bias
    return loc, conf


def dedup_scores(conf: torch.Tensor, max_rounds: int = 64) -> torch.Tensor:
    """Nudge logits (1 ulp at a time) until, per image, all foreground softmax scores are
    distinct.  The reference sorts scores with an *unstable* argsort
    (SSD_from_scratch.py:439,463,677), so only score-distinct inputs have a unique answer."""
    conf = conf.clone()
    for b in range(conf.shape[0]):
        for rnd in range(max_rounds):
            s = conf[b].softmax(-1)[:, 1:].reshape(-1)
            order = torch.argsort(s, stable=True)
            ss = s[order]
            dup = (ss[1:] == ss[:-1]).nonzero(as_tuple=True)[0]
            if dup.numel() == 0:
                break
            flat = order[dup + 1]
            pri = flat // (conf.shape[-1] - 1)
            cls = flat % (conf.shape[-1] - 1) + 1
            cur = conf[b, pri, cls]
            for _ in range(rnd + 1):   # a 1-ulp logit step may round to the same score: widen per round
                cur = torch.nextafter(cur, torch.full_like(cur, float("inf")))
            conf[b, pri, cls] = cur
        else:
            raise RuntimeError("could not de-duplicate scores")
    return conf


CONFIGS = {
    # idx: (seed, B, (g_lo, g_hi), bg_bias, iou_thresh, ratio, score_thresh, nms_thresh, max_per_img)
    1: dict(seed=0, batch=1, g=(5, 5), bg_bias=6.0, iou_thresh=0.5, ratio=3.0, score_thresh=0.01, nms_thresh=0.45, max_per_img=200),
    2: dict(seed=1, batch=32, g=(1, 20), bg_bias=6.0, iou_thresh=0.5, ratio=3.0, score_thresh=0.01, nms_thresh=0.45, max_per_img=200),
    3: dict(seed=2, batch=256, g=(1, 20), bg_bias=6.0, iou_thresh=0.5, ratio=3.0, score_thresh=0.01, nms_thresh=0.45, max_per_img=200),
    4: dict(seed=3, batch=4096, g=(1, 20), bg_bias=6.0, iou_thresh=0.5, ratio=3.0, score_thresh=0.01, nms_thresh=0.45, max_per_img=200),
    5: dict(seed=4, batch=1024, g=(64, 64), bg_bias=0.0, iou_thresh=0.5, ratio=3.0, score_thresh=0.0, nms_thresh=0.45, max_per_img=200),
}


def config(idx: int, batch: Optional[int] = None, dedup: bool = False, seed_offset: int = 0) -> Dict:
    """Materialise configuration ``idx`` (1..5), optionally at a reduced batch size.

    Training half uses ``conf_train`` (no background bias); the inference half uses
    ``conf_infer = conf_train`` with ``bg_bias`` added to the background logit, as in
    SURVEY.md section 8d.  ``seed_offset`` derives independent shards (rank r of a sharded
    run uses seed_offset=r)."""
    spec = dict(CONFIGS[idx])
    n_img = batch if batch is not None else spec["batch"]
    gen = torch.Generator().manual_seed(spec["seed"] + 1000 * seed_offset)
    targets = make_targets(n_img, spec["g"][0], spec["g"][1], gen)
    loc, conf = make_heads(n_img, gen)
    conf_infer = conf.clone()
    if spec["bg_bias"] != 0.0:
        conf_infer[..., 0] += spec["bg_bias"]
    if dedup:
        conf_infer = dedup_scores(conf_infer)
    spec.update(batch=n_img, targets=targets, loc_all=loc, conf_train=conf, conf_infer=conf_infer)
    return spec


def pack_targets(targets: List[Dict[str, torch.Tensor]]) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Ragged list -> (gt_boxes [sumG,4] f32, gt_labels [sumG] i64, gt_offsets [B+1] i32), CPU."""
    counts = [int(t["boxes"].shape[0]) for t in targets]
    offs = torch.zeros(len(targets) + 1, dtype=torch.int32)
    if counts:
        offs[1:] = torch.tensor(counts, dtype=torch.int32).cumsum(0)
    if sum(counts) == 0:
        return torch.zeros((0, 4), dtype=torch.float32), torch.zeros((0,), dtype=torch.int64), offs
    boxes = torch.cat([t["boxes"].reshape(-1, 4).to(torch.float32) for t in targets], 0)
    labels = torch.cat([t["labels"].reshape(-1).to(torch.int64) for t in targets], 0)
    return boxes.contiguous(), labels.contiguous(), offs


LEVELS = ((38, 4), (19, 6), (10, 6), (5, 6), (3, 4), (1, 4))      # (side, shapes per cell) of the six SSD300 feature maps


def heads_from_packed(all_rows: torch.Tensor, channels_last: bool = False) -> List[torch.Tensor]:
    """Synthetic head outputs: the six tensors [B, A*D, H, W] whose reference packing (permute(0,2,3,1) + view + cat,
    SSD_from_scratch.py:249-269) is `all_rows` [B, 8732, D] -- pure data movement.  channels_last=True returns them in
    channels_last memory format (what a channels_last conv head produces)."""
    B, _, D = all_rows.shape
    out, off = [], 0
    for side, shapes in LEVELS:
        n = side * side * shapes
        h = all_rows[:, off:off + n, :].reshape(B, side, side, shapes * D).permute(0, 3, 1, 2)
        out.append(h.contiguous(memory_format=torch.channels_last) if channels_last else h.contiguous())
        off += n
    assert off == 8732
    return out
