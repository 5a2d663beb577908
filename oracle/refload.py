"""ORACLE -- test infrastructure only.  Imports the UNMODIFIED reference staged under oracle/_ref/ (stage_reference.py).

`load()` returns the two reference modules (SSD_from_scratch, SSD_trainer) exactly as they are on disk.  The two
third-party modules SSD_trainer.py imports at top level but never touches on the hot path, and which this image lacks
(torchmetrics -- SSD_trainer.py:4, matplotlib -- SSD_trainer.py:12), are stubbed in sys.modules; nothing else is altered.
`train_half` / `predict_half` call the reference's own functions in the order SSD_train_step / SSD_test_step do
(SSD_trainer.py:92-117, :214-256), on whatever device the tensors live on: CPU = the timed baseline, CUDA = the
device-matched checker.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types
from typing import Dict, List, Optional, Tuple

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
CLASSES = {"car": 0, "truck": 1, "pedestrian": 2, "bicyclist": 3, "light": 4}     # 5 foreground classes (ssd_demo_app.py:26)

_modules = None


class _StubMAP:
    """Stand-in for torchmetrics' MeanAveragePrecision (absent from this image): records the updates it is given."""

    def __init__(self, *a, **k):
        self.preds, self.targets = [], []

    def to(self, *_a, **_k):
        return self

    def reset(self):
        self.preds, self.targets = [], []

    def update(self, preds, target):
        self.preds.append(preds)
        self.targets.append(target)

    def compute(self):
        return {"map_50": torch.tensor(-1.0)}


def available() -> bool:
    return all(os.path.isfile(os.path.join(REF_DIR, f)) for f in ("SSD_from_scratch.py", "SSD_trainer.py"))


def load():
    """-> (SSD_from_scratch module, SSD_trainer module), imported once from oracle/_ref."""
    global _modules
    if _modules is not None:
        return _modules
    if not available():
        raise FileNotFoundError(f"{REF_DIR} holds no staged reference: run `python oracle/stage_reference.py` in the build "
                                "container (it needs /root/reference)")
    for name in ("torchmetrics", "torchmetrics.detection", "torchmetrics.detection.mean_ap", "matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:                                   # noqa: BLE001 -- absent: stub it
                sys.modules[name] = types.ModuleType(name)
    mp = sys.modules["torchmetrics.detection.mean_ap"]
    if not hasattr(mp, "MeanAveragePrecision"):
        mp.MeanAveragePrecision = _StubMAP
    mods = []
    for name in ("SSD_from_scratch", "SSD_trainer"):
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF_DIR, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod                                  # SSD_trainer does `from SSD_from_scratch import mySSD`
        spec.loader.exec_module(mod)
        mods.append(mod)
    _modules = tuple(mods)
    return _modules


_models: Dict[str, torch.nn.Module] = {}


def model(device="cpu"):
    """A reference mySSD (random weights; the hot path only uses its prior buffers, variances and class count)."""
    key = str(torch.device(device))
    if key not in _models:
        sfs, _ = load()
        gen_state = torch.random.get_rng_state()
        _models[key] = sfs.mySSD(class_to_idx_dict=dict(CLASSES)).to(device).eval()
        torch.random.set_rng_state(gen_state)
    return _models[key]


def to_device(targets: List[Dict[str, torch.Tensor]], device) -> List[Dict[str, torch.Tensor]]:
    return [{k: v.to(device) for k, v in t.items()} for t in targets]


def train_half(loc_all: torch.Tensor, conf_all: torch.Tensor, targets, iou_thresh: float, ratio: float,
               mdl=None) -> Tuple[torch.Tensor, torch.Tensor, Tuple]:
    """SSD_trainer.py:92-117 after the forward pass, verbatim calls: -> (loc_loss, conf_loss, (pos_mask, loc_t_pm, cls_t))."""
    _, tr = load()
    dev = loc_all.device
    mdl = mdl if mdl is not None else model(dev)
    pos_mask, loc_t_pm, cls_t = tr.build_targets(model=mdl, targets=targets, H=300, W=300, iou_thresh=iou_thresh, device=dev)
    num_pos_per_img = pos_mask.sum(dim=1)
    total_pos = num_pos_per_img.sum().clamp_min(1).float()
    l_loc = torch.nn.functional.smooth_l1_loss(loc_all[pos_mask], loc_t_pm, reduction="sum") / total_pos
    l_conf = tr.CELoss_w_neg_mining(conf_all=conf_all, cls_t=cls_t, pos_mask=pos_mask, num_pos_per_img=num_pos_per_img,
                                    total_pos=total_pos, neg_pos_ratio=ratio)
    return l_loc, l_conf, (pos_mask, loc_t_pm, cls_t)


def predict_half(loc_all: torch.Tensor, conf_all: torch.Tensor, score_thresh: float, nms_thresh: float, max_per_img: int,
                 class_agnostic: bool = False, mdl=None):
    """mySSD.predict on precomputed head outputs (SSD_trainer.py:240-246): -> List[Dict]."""
    mdl = mdl if mdl is not None else model(loc_all.device)
    return mdl.predict(images=None, score_thresh=score_thresh, nms_thresh=nms_thresh, max_per_img=max_per_img,
                       class_agnostic=class_agnostic, pre_loc_all=loc_all, pre_conf_all=conf_all)
