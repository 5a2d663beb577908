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
