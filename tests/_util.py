"""Shared helpers for the parity tests: golden loading and input regeneration."""
from __future__ import annotations

import hashlib
import os
from typing import Dict, List

import numpy as np
import torch

from ssdhot import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sha(t: torch.Tensor) -> str:
    return hashlib.sha256(t.detach().cpu().contiguous().numpy().tobytes()).hexdigest()


def load(name: str) -> Dict[str, np.ndarray]:
    with np.load(os.path.join(GOLDEN, name)) as z:
        return {k: z[k] for k in z.files}


def unpack_targets(g) -> List[Dict[str, torch.Tensor]]:
    offs = g["gt_offsets"]
    out = []
    for i in range(len(offs) - 1):
        a, b = int(offs[i]), int(offs[i + 1])
        out.append({"boxes": torch.from_numpy(g["gt_boxes"][a:b].copy()).reshape(-1, 4),
                    "labels": torch.from_numpy(g["gt_labels"][a:b].copy())})
    return out


def train_inputs(g):
    """(targets, loc_all, conf) of a train_* fixture, regenerated from the seed and checked
    against the fixture's sha256 (a mismatch means the CPU random generator diverged between
    the machine that made the fixture and this one -- not a parity failure)."""
    n = int(g["batch"])
    cfg_idx = int(g["cfg"])
    full = synth.config(cfg_idx, batch=None if cfg_idx == 2 else n)
    loc, conf = full["loc_all"][:n], full["conf_train"][:n]
    if n != full["batch"] and cfg_idx == 2 and not (sha(loc) == str(g["loc_all_sha"])):
        full = synth.config(cfg_idx, batch=n)
        loc, conf = full["loc_all"], full["conf_train"]
    assert sha(loc) == str(g["loc_all_sha"]) and sha(conf) == str(g["conf_sha"]), \
        "seeded input generator diverged from the one that produced the golden fixture"
    return unpack_targets(g), loc.contiguous(), conf.contiguous()


def predict_inputs(g):
    n = int(g["batch"])
    cfg = synth.config(int(g["cfg"]), batch=n)
    loc, conf = cfg["loc_all"], cfg["conf_infer"].clone()
    idx = torch.from_numpy(g["conf_patch_idx"].astype(np.int64))
    if idx.numel():
        conf[idx[:, 0], idx[:, 1], idx[:, 2]] = torch.from_numpy(g["conf_patch_val"])
    assert sha(loc) == str(g["loc_all_sha"]) and sha(conf) == str(g["conf_sha"]), \
        "seeded input generator diverged from the one that produced the golden fixture"
    return loc, conf


def _patched(t: torch.Tensor, idx: np.ndarray, val: np.ndarray) -> torch.Tensor:
    t = t.clone()
    i = torch.from_numpy(idx.astype(np.int64))
    if i.numel():
        t[i[:, 0], i[:, 1], i[:, 2]] = torch.from_numpy(val.copy())
    return t


def nonfinite_predict_inputs(g):
    """(loc, conf) of nonfinite.npz's predict half: cfg 3, two images, de-duplicated scores, then the NaN / Inf patches."""
    cfg = synth.config(3, batch=2)
    conf = _patched(cfg["conf_infer"], g["pred_dedup_idx"], g["pred_dedup_val"])
    return _patched(cfg["loc_all"], g["pred_loc_idx"], g["pred_loc_val"]), _patched(conf, g["pred_conf_idx"], g["pred_conf_val"])


def nonfinite_loss_cases(g):
    """Yields (name, loc, conf, targets, cfg, (want_loc_loss, want_conf_loss)) for the loss half of nonfinite.npz (cfg 1)."""
    cfg = synth.config(1)
    assert sha(cfg["loc_all"]) == str(g["loss_loc_sha"]) and sha(cfg["conf_train"]) == str(g["loss_conf_sha"])
    starts = g["loss_case_start"]
    for k, name in enumerate(g["loss_case_names"]):
        idx, val = g["loss_idx"][starts[k]:starts[k + 1]], g["loss_val"][starts[k]:starts[k + 1]]
        which = str(g["loss_case_which"][k])
        loc = _patched(cfg["loc_all"], idx, val) if which == "loc" else cfg["loc_all"]
        conf = _patched(cfg["conf_train"], idx, val) if which == "conf" else cfg["conf_train"]
        yield str(name), loc, conf, cfg["targets"], cfg, (float(g["loss_want"][k][0]), float(g["loss_want"][k][1]))


def same_float(a: float, b: float, rel: float = 1e-5) -> bool:
    """Equality of two loss values that may be NaN or infinite."""
    if a != a or b != b:
        return a != a and b != b
    if a in (float("inf"), float("-inf")) or b in (float("inf"), float("-inf")):
        return a == b
    return abs(a - b) <= rel * max(abs(a), abs(b), 1e-30)


def equal_nan(a: torch.Tensor, b: torch.Tensor) -> bool:
    """torch.equal that treats NaN == NaN (same shape, same dtype, same values)."""
    return a.shape == b.shape and a.dtype == b.dtype and bool(torch.isclose(a, b, rtol=0.0, atol=0.0, equal_nan=True).all())


def split_rows(counts, labels, scores, boxes):
    ends = np.cumsum(counts)
    out = []
    for i, e in enumerate(ends):
        s = e - counts[i]
        out.append({"labels": torch.from_numpy(labels[s:e].astype(np.int64)), "scores": torch.from_numpy(scores[s:e].copy()),
                    "boxes": torch.from_numpy(boxes[s:e].copy()).reshape(-1, 4)})
    return out


def split_predictions(g):
    counts = g["counts"]
    ends = np.cumsum(counts)
    out = []
    for i, e in enumerate(ends):
        s = e - counts[i]
        out.append({"labels": torch.from_numpy(g["labels"][s:e].astype(np.int64)),
                    "scores": torch.from_numpy(g["scores"][s:e].copy()),
                    "boxes": torch.from_numpy(g["boxes"][s:e].copy()).reshape(-1, 4)})
    return out


def unpack_bits(bits: np.ndarray, n: int) -> torch.Tensor:
    return torch.from_numpy(np.unpackbits(bits, axis=-1)[..., :n].astype(bool))


LEVELS = synth.LEVELS
unpack_heads = synth.heads_from_packed      # [B, 8732, D] -> the six head outputs whose reference packing it is


def _pack_heads(loc_heads, conf_heads, num_classes):
    from oracle import ssd_oracle as O
    return O.pack_heads(loc_heads, conf_heads, num_classes)


class TinySSD(torch.nn.Module):
    """A stand-in with mySSD's module names and feature-map sizes (SFS:120-231) but 8-channel trunks, so that the head
    plumbing can be checked on the CPU without the VGG weights."""

    def __init__(self, num_classes=6):
        super().__init__()
        nn = torch.nn
        self.num_classes = num_classes

        def down(cin, cout, size):
            return nn.Sequential(nn.Conv2d(cin, cout, 1), nn.AdaptiveAvgPool2d(size), nn.ReLU())
        self.VGG16_UpTo_conv4_3 = down(3, 8, 38)
        self.VGG16_extras = down(8, 8, 19)
        self.extra_conv6 = nn.Sequential(nn.Conv2d(8, 8, 1), nn.ReLU())
        self.extra_conv7 = nn.Sequential(nn.Conv2d(8, 8, 1), nn.ReLU())
        self.extra_conv8_2 = down(8, 8, 10)
        self.extra_conv9_2 = down(8, 8, 5)
        self.extra_conv10_2 = down(8, 8, 3)
        self.extra_conv11_2 = down(8, 8, 1)
        shapes = (4, 6, 6, 6, 4, 4)
        self.box_head = nn.ModuleList([nn.Conv2d(8, a * 4, 3, padding=1) for a in shapes])
        self.cls_head = nn.ModuleList([nn.Conv2d(8, a * num_classes, 3, padding=1) for a in shapes])

    def forward(self, x):                                   # the reference's forward, tail included (SFS:234-271)
        f0 = self.VGG16_UpTo_conv4_3(x)
        f1 = self.extra_conv7(self.extra_conv6(self.VGG16_extras(f0)))
        f2 = self.extra_conv8_2(f1); f3 = self.extra_conv9_2(f2); f4 = self.extra_conv10_2(f3); f5 = self.extra_conv11_2(f4)
        feats = (f0, f1, f2, f3, f4, f5)
        return _pack_heads([h(f) for h, f in zip(self.box_head, feats)], [h(f) for h, f in zip(self.cls_head, feats)],
                            self.num_classes)
