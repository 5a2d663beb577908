"""The two callers of the hot path, SSD_train_step and SSD_test_step (SSD_trainer.py:21-150, :155-293), with the same
signatures and return dictionaries, routed through the fused kernels: the post-backbone part of a step is ONE launch for
targets + both losses (train_image_kernel reading the head outputs where the conv heads left them -- no permute / cat
tail, no [B,P] target tensors in HBM) and, in the eval step, the two launches of predict.

`ssdhot.patch(model, trainer_module)` installs them over the reference's functions; the reference's own step functions keep
working after a patch too (they then call the drop-in build_targets / CELoss_w_neg_mining / predict one by one).

What is kept from the reference on purpose: the per-batch `.item()` of the three losses (its return value is a running
sum of Python floats), zero_grad(set_to_none=True) -> backward -> step -> scheduler.step() order, `model.train()` /
`model.eval()` + `torch.inference_mode()`, the timing dictionary's keys, and the in-place move of the caller's target dicts
to the device in the eval step (the mAP metric consumes them there).
"""
from __future__ import annotations

import time
from typing import Dict

import torch

from . import api


def _losses(model, images, targets, iou_thresh, neg_pos_ratio, device):
    """forward + fused targets/losses -> (loc_loss, conf_loss, loc_all, conf_all, head outputs or None)."""
    H, W = images.shape[-2], images.shape[-1]
    packed = api.pack_targets(targets, device)
    if api._has_ssd_heads(model) and getattr(model, "num_classes", 6) == 6:
        loc_heads, conf_heads = api.forward_heads(model, images)
        l_loc, l_conf = api.multibox_loss_heads(model, loc_heads, conf_heads, packed, iou_thresh, neg_pos_ratio, H, W)
        return l_loc, l_conf, (loc_heads, conf_heads), None
    loc_all, conf_all = model(images)
    l_loc, l_conf = api.multibox_loss(model, loc_all, conf_all, packed, iou_thresh, neg_pos_ratio, H, W)
    return l_loc, l_conf, None, (loc_all, conf_all)


def SSD_train_step(model, dataloader, optimizer, iou_thresh: float = 0.5, neg_pos_ratio: float = 3.0, device="cpu",
                   timing: bool = False, scheduler=None) -> Dict:
    """Drop-in for SSD_trainer.SSD_train_step (SSD_trainer.py:21-150).  Batches may carry the reference's List[Dict]
    targets or the PackedTargets of ssdhot.collate_detection."""
    if torch.device(device).type != "cuda":
        raise api._lib.SsdhotError("ssdhot.SSD_train_step needs device='cuda' (no CPU fallback)")
    model.train()
    train_loss = loc_loss = conf_loss = 0.0
    batch_count = 0
    time_device = time_forward = 0.0
    for _batch, (images, targets) in enumerate(dataloader):
        t0 = time.perf_counter()
        images = images.to(device, non_blocking=True)
        packed = api.pack_targets(targets, device)                  # one pinned buffer, one H2D copy (SSD_trainer.py:66-69: 3 per image)
        t1 = time.perf_counter()
        l_loc, l_conf, _, _ = _losses(model, images, packed, iou_thresh, neg_pos_ratio, device)
        batch_loss = l_loc + l_conf
        t2 = time.perf_counter()
        ll, lc, lt = torch.stack((l_loc.detach(), l_conf.detach(), batch_loss.detach())).tolist()      # one sync, not three
        loc_loss += ll
        conf_loss += lc
        train_loss += lt
        optimizer.zero_grad(set_to_none=True)
        batch_loss.backward()
        optimizer.step()
        if scheduler is not None:
            scheduler.step()
        batch_count += 1
        time_device += t1 - t0
        time_forward += t2 - t1
    n = len(dataloader)
    time_dict = {"to device": time_device / max(batch_count, 1), "model forward": time_forward / max(batch_count, 1),
                 "build targets": 0.0}        # (targets are built inside the fused loss kernel: part of "model forward")
    return {"training loss": train_loss / n, "localization loss": loc_loss / n, "classification loss": conf_loss / n,
            "timing": time_dict}


def make_test_step(metric_source=None):
    """SSD_test_step bound to where its mAP metric class comes from: a module whose `MeanAveragePrecision` attribute is looked
    up at call time (the patched SSD_trainer module), a callable, or None (torchmetrics itself)."""

    def SSD_test_step(model, dataloader, iou_thresh: float = 0.5, neg_pos_ratio: float = 3.0, score_thresh: float = 0.05,
                      nms_thresh: float = 0.5, max_detections_per_img: int = 100, device="cpu", timing: bool = False):
        """Drop-in for SSD_trainer.SSD_test_step (SSD_trainer.py:155-293)."""
        if torch.device(device).type != "cuda":
            raise api._lib.SsdhotError("ssdhot.SSD_test_step needs device='cuda' (no CPU fallback)")
        model.eval()
        conf_loss = loc_loss = test_loss = 0.0
        batch_count = 0
        time_pred = time_map = 0.0
        if metric_source is None:
            from torchmetrics.detection.mean_ap import MeanAveragePrecision as factory      # noqa: N813
        else:
            factory = metric_source if callable(metric_source) else getattr(metric_source, "MeanAveragePrecision")
        map_metric = factory(box_format="xyxy", iou_type="bbox", iou_thresholds=[0.50], class_metrics=True).to(device)
        map_metric.reset()
        with torch.inference_mode():
            for _batch, (images, targets) in enumerate(dataloader):
                images = images.to(device, non_blocking=True)
                if isinstance(targets, api.PackedTargets):
                    packed = targets.to(device)
                    targets = packed.as_list()
                else:
                    packed = api.pack_targets(targets, device)          # one packed copy feeds the kernels ...
                    for i in range(len(targets)):                       # ... the metric wants the reference's dicts on the device
                        for key in targets[i]:
                            targets[i][key] = targets[i][key].to(device=device, non_blocking=True)
                l_loc, l_conf, heads, packed_out = _losses(model, images, packed, iou_thresh, neg_pos_ratio, device)
                ll, lc = torch.stack((l_loc, l_conf)).tolist()
                loc_loss += ll
                conf_loss += lc
                test_loss += ll + lc
                t0 = time.perf_counter()
                if heads is not None:
                    preds = api.predict_heads(model, heads[0], heads[1], score_thresh, nms_thresh, max_detections_per_img, False)
                else:
                    preds = api.predict(model, None, score_thresh, nms_thresh, max_detections_per_img, False,
                                        pre_loc_all=packed_out[0], pre_conf_all=packed_out[1])
                time_pred += time.perf_counter() - t0
                map_metric.update(preds=preds, target=targets)
                batch_count += 1
        n = len(dataloader)
        t0 = time.perf_counter()
        mAP = map_metric.compute()
        time_map += time.perf_counter() - t0
        time_dict = {"model prediction": time_pred / max(batch_count, 1), "mAP time": time_map, "build targets": 0.0}
        return {"testing loss": test_loss / n, "localization loss": loc_loss / n, "classification loss": conf_loss / n,
                "mAP": mAP, "timing": time_dict}

    return SSD_test_step


SSD_test_step = make_test_step(None)
