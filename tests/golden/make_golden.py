"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

Run in the build container only (it is the one place /root/reference exists):

    python tests/golden/make_golden.py

The reference (/root/reference/SSD_from_scratch.py, SSD_trainer.py) is imported as is; the
two modules it imports but never uses on the hot path and that are absent from this image
(torchmetrics, matplotlib -- SSD_trainer.py:4,12) are stubbed in ``sys.modules``.  Inputs
come from ``ssdhot.synth`` (seeded, CPU); every fixture stores the sha256 of the generated
head tensors so a divergent random generator is detected instead of mis-reported as a
parity failure.  Outputs are stored in compact dtypes (bit masks, int8 classes) without loss.
"""
from __future__ import annotations

import hashlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "automotive-ssd-object-detection_b200"))

from ssdhot import synth  # noqa: E402


def import_reference():
    for name in ("torchmetrics", "torchmetrics.detection", "torchmetrics.detection.mean_ap",
                 "matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["torchmetrics.detection.mean_ap"].MeanAveragePrecision = object
    sys.path.insert(0, "/root/reference")
    import SSD_from_scratch as sfs
    import SSD_trainer as tr
    return sfs, tr


def sha(t: torch.Tensor) -> str:
    return hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()


def pack_targets_np(targets):
    boxes, labels, offs = synth.pack_targets(targets)
    return boxes.numpy(), labels.numpy(), offs.numpy()


def ref_train(model, tr, targets, loc_all, conf_all, iou_thresh, ratio):
    pos, loc_t_pm, cls_t = tr.build_targets(model=model, targets=targets, H=300, W=300,
                                            iou_thresh=iou_thresh, device="cpu")
    n_img = pos.sum(dim=1)
    total = n_img.sum().clamp_min(1).float()
    l_loc = torch.nn.functional.smooth_l1_loss(loc_all[pos], loc_t_pm, reduction="sum") / total
    l_conf = tr.CELoss_w_neg_mining(conf_all=conf_all, cls_t=cls_t, pos_mask=pos,
                                    num_pos_per_img=n_img, total_pos=total, neg_pos_ratio=ratio)
    return pos, loc_t_pm, cls_t, n_img, l_loc, l_conf


def save(name, **arrays):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def train_fixture(model, tr, name, cfg_idx, batch=None, extra_targets=None, iou_thresh=None, ratio=None):
    cfg = synth.config(cfg_idx, batch=batch)
    targets = cfg["targets"] if extra_targets is None else extra_targets
    n = len(targets)
    loc_all, conf = cfg["loc_all"][:n], cfg["conf_train"][:n]
    thr = cfg["iou_thresh"] if iou_thresh is None else iou_thresh
    rat = cfg["ratio"] if ratio is None else ratio
    pos, loc_t_pm, cls_t, n_img, l_loc, l_conf = ref_train(model, tr, targets, loc_all, conf, thr, rat)
    # per-image full encode_ssd outputs for image 0 (covers negatives' loc targets too)
    norm = torch.tensor([300, 300, 300, 300], dtype=torch.float32)
    g0 = targets[0]["boxes"] / norm if targets[0]["boxes"].numel() else targets[0]["boxes"].new_zeros((0, 4))
    e_loc, e_cls, e_pos, e_match = model.encode_ssd(g0, targets[0]["labels"], iou_thresh=thr)
    gb, gl, go = pack_targets_np(targets)
    assert int(cls_t.max()) < 127
    save(name,
         cfg=np.int64(cfg_idx), batch=np.int64(n), iou_thresh=np.float64(thr), ratio=np.float64(rat),
         gt_boxes=gb, gt_labels=gl, gt_offsets=go,
         loc_all_sha=np.array(sha(loc_all)), conf_sha=np.array(sha(conf)),
         pos_bits=np.packbits(pos.numpy(), axis=1), cls_t=cls_t.numpy().astype(np.int8),
         loc_t_pm=loc_t_pm.numpy(), n_pos=n_img.numpy(),
         loc_loss=np.float32(l_loc.item()), conf_loss=np.float32(l_conf.item()),
         enc0_loc=e_loc.numpy(), enc0_cls=e_cls.numpy().astype(np.int8),
         enc0_pos=np.packbits(e_pos.numpy()), enc0_match=e_match.numpy())


def predict_fixture(model, name, cfg_idx, batch, score_thresh=None, nms_thresh=None, max_per_img=None,
                    class_agnostic=False, dedup=True):
    cfg = synth.config(cfg_idx, batch=batch, dedup=dedup)
    st = cfg["score_thresh"] if score_thresh is None else score_thresh
    nt = cfg["nms_thresh"] if nms_thresh is None else nms_thresh
    mx = cfg["max_per_img"] if max_per_img is None else max_per_img
    loc_all, conf = cfg["loc_all"], cfg["conf_infer"]
    out = model.predict(None, score_thresh=st, nms_thresh=nt, max_per_img=mx,
                        class_agnostic=class_agnostic, pre_loc_all=loc_all, pre_conf_all=conf)
    counts = np.array([o["labels"].numel() for o in out], dtype=np.int64)
    # the de-duplicated conf differs from the seeded one only at a few logits: store the patch
    raw = synth.config(cfg_idx, batch=batch)["conf_infer"]
    diff = (raw != conf).nonzero()
    save(name,
         cfg=np.int64(cfg_idx), batch=np.int64(batch), score_thresh=np.float64(st), nms_thresh=np.float64(nt),
         max_per_img=np.int64(mx), class_agnostic=np.bool_(class_agnostic),
         loc_all_sha=np.array(sha(loc_all)), conf_sha=np.array(sha(conf)),
         conf_patch_idx=diff.numpy().astype(np.int32), conf_patch_val=conf[raw != conf].numpy(),
         counts=counts,
         labels=torch.cat([o["labels"] for o in out]).numpy().astype(np.int8),
         scores=torch.cat([o["scores"] for o in out]).numpy(),
         boxes=torch.cat([o["boxes"] for o in out]).numpy())


def edge_targets():
    """Hand-made ground truth that exercises the tie / degenerate rules of SURVEY.md section 8a."""
    f = lambda rows: torch.tensor(rows, dtype=torch.float32).reshape(-1, 4)
    i = lambda rows: torch.tensor(rows, dtype=torch.int64)
    return [
        {"boxes": f([]), "labels": i([])},                                           # no GT at all
        {"boxes": f([[60, 60, 120, 120], [60, 60, 120, 120]]), "labels": i([1, 3])},   # duplicate GT: lower index wins
        {"boxes": f([[100, 100, 100, 100], [30, 40, 200, 220]]), "labels": i([2, 0])}, # zero-size GT -> NaN column
        {"boxes": f([[0, 0, 300, 300]]), "labels": i([4])},                          # whole image
        {"boxes": f([[10, 10, 11, 11], [290, 290, 300, 300], [149, 149, 151, 151]]), "labels": i([0, 1, 2])},  # tiny boxes
        {"boxes": f([[20, 20, 140, 140], [22, 22, 142, 142], [24, 20, 144, 140], [20, 24, 140, 144]]),
         "labels": i([0, 1, 2, 3])},                                                 # GTs competing for one prior
        {"boxes": f([[150, 0, 150, 300]]), "labels": i([1])},                         # zero width, full height
    ]


def main():
    torch.manual_seed(0)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    sfs, tr = import_reference()
    model = sfs.mySSD(class_to_idx_dict={"biker": 0, "car": 1, "pedestrian": 2, "trafficLight": 3, "truck": 4})
    model.eval()
    save("priors.npz", priors_sha=np.array(sha(model.priors)), priors_xyxy_sha=np.array(sha(model.priors_xyxy)),
         priors_head=model.priors[:8].numpy(), priors_tail=model.priors[-8:].numpy())

    train_fixture(model, tr, "train_cfg1.npz", 1)
    train_fixture(model, tr, "train_cfg2.npz", 2)
    train_fixture(model, tr, "train_cfg2_thr04.npz", 2, batch=8, iou_thresh=0.4, ratio=2.5)
    train_fixture(model, tr, "train_cfg5_b2.npz", 5, batch=2)
    train_fixture(model, tr, "train_edges.npz", 2, extra_targets=edge_targets())

    predict_fixture(model, "predict_cfg1.npz", 1, 1)
    predict_fixture(model, "predict_cfg3_b4.npz", 3, 4)
    predict_fixture(model, "predict_cfg3_b2_notebook.npz", 3, 2, score_thresh=0.2, nms_thresh=0.3, max_per_img=100)
    predict_fixture(model, "predict_cfg3_b2_agnostic.npz", 3, 2, class_agnostic=True)
    predict_fixture(model, "predict_cfg3_b2_empty.npz", 3, 2, score_thresh=0.999)
    predict_fixture(model, "predict_cfg5_b1.npz", 5, 1)

    # stand-alone static methods: decode_ssd and iou_nms
    gen = torch.Generator().manual_seed(77)
    n = 600
    ctr = torch.rand((n, 2), generator=gen) * 300
    wh = torch.rand((n, 2), generator=gen) * 80 + 4
    boxes = torch.cat((ctr - wh / 2, ctr + wh / 2), 1).clamp(0, 300)
    scores = torch.rand((n,), generator=gen)
    assert scores.unique().numel() == n
    keep = sfs.mySSD.iou_nms(boxes, scores, 0.45)
    keep30 = sfs.mySSD.iou_nms(boxes, scores, 0.30)
    loc = torch.randn((P_ := 8732, 4), generator=gen)
    dec = sfs.mySSD.decode_ssd(loc, model.priors, (0.1, 0.2))
    save("static_methods.npz", boxes=boxes.numpy(), scores=scores.numpy(), keep45=keep.numpy(),
         keep30=keep30.numpy(), loc_sha=np.array(sha(loc)), decoded=dec.numpy())


if __name__ == "__main__":
    main()
