"""Golden fixture for NaN / +-Inf head outputs (SURVEY.md section 7 hard part 3, section 8a "torch.maximum/minimum/clamp
propagate NaN"), produced by the UNMODIFIED reference like the other fixtures:

    python tests/golden/make_golden_nonfinite.py            (build container only: needs /root/reference)

predict (SSD_from_scratch.py:388-465): a NaN / +Inf logit makes the whole softmax row NaN, and NaN > thresh is False, so
the row yields no candidate; -Inf in the background logit renormalises the others; a NaN box offset survives decode and
`clamp(0, 1)` (torch.clamp propagates NaN, :422-425), and a NaN box that is its class's best keeps its place and suppresses
every later box of the class (`d <= thr` is False for NaN, :690) while a NaN box further down is suppressed by whoever
is kept before it.  losses (SSD_trainer.py:104-108, :577-600): a NaN cross-entropy of a negative prior is the largest
element for torch.topk and therefore always mined -> the loss is NaN; CE = +Inf likewise; non-finite offsets matter only
at positive priors.  The fixture stores the patches applied to the seeded inputs and the reference's outputs.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402
from ssdhot import synth  # noqa: E402

NAN, INF = float("nan"), float("inf")


def apply(t: torch.Tensor, patch):
    t = t.clone()
    for (b, p, j, v) in patch:
        t[b, p, j] = v
    return t


def patch_arrays(patch):
    return (np.array([[b, p, j] for (b, p, j, _) in patch], dtype=np.int32).reshape(-1, 3),
            np.array([v for (*_, v) in patch], dtype=np.float32))


def main():
    torch.manual_seed(0)
    sfs, tr = MG.import_reference()
    model = sfs.mySSD(class_to_idx_dict={"biker": 0, "car": 1, "pedestrian": 2, "trafficLight": 3, "truck": 4}).eval()
    out = {}

    # ---- predict: cfg 3, two images -----------------------------------------------------------------------------
    cfg = synth.config(3, batch=2, dedup=True)
    st, nt, mx = cfg["score_thresh"], cfg["nms_thresh"], cfg["max_per_img"]
    loc, conf = cfg["loc_all"], cfg["conf_infer"]
    scores = conf.softmax(-1)[..., 1:]                                  # [2, P, 5]
    clean = model.predict(None, st, nt, mx, False, loc, conf)

    def prior_of(b, rank):                                              # (prior, class) of the rank-th clean detection
        s = clean[b]["scores"][rank]
        hit = (scores[b] == s).nonzero()
        assert hit.shape[0] == 1
        return int(hit[0, 0]), int(hit[0, 1])

    used = set()

    def fresh(b, rank):
        while True:
            p, c = prior_of(b, rank)
            if (b, p) not in used:
                used.add((b, p))
                return p, c, rank
            rank += 1

    loc_patch, conf_patch = [], []
    p, c, _ = fresh(0, 0)                       # the best box of its class becomes all-NaN: kept, kills the rest of the class
    loc_patch += [(0, p, j, NAN) for j in range(4)]
    p, c, _ = fresh(0, 1)                       # infinite width: clamps to the full image width
    loc_patch += [(0, p, 2, INF)]
    p, c, _ = fresh(0, 40)                      # a NaN coordinate further down: suppressed by whoever is kept before it
    loc_patch += [(0, p, 0, NAN)]
    p, c, _ = fresh(0, 3)                       # NaN logit: the whole row stops being a candidate
    conf_patch += [(0, p, 2, NAN)]
    p, c, _ = fresh(0, 5)                       # +Inf logit: inf - inf = NaN in the softmax, same effect
    conf_patch += [(0, p, c + 1, INF)]
    p, c, _ = fresh(0, 7)                       # -Inf background logit: the foreground scores renormalise upwards
    conf_patch += [(0, p, 0, -INF)]
    p, c, _ = fresh(1, 0)                       # cx = inf and w = inf: x1 = inf - inf = NaN, x2 clamps to 1
    loc_patch += [(1, p, 0, INF), (1, p, 2, INF)]
    p, c, _ = fresh(1, 2)                       # w = pw * exp(-inf) = 0: an empty box
    loc_patch += [(1, p, 2, -INF)]
    p, c, _ = fresh(1, 4)                       # every logit -Inf: max = -inf, x - max = NaN
    conf_patch += [(1, p, j, -INF) for j in range(6)]
    loc_p, conf_p = apply(loc, loc_patch), apply(conf, conf_patch)
    got = model.predict(None, st, nt, mx, False, loc_p, conf_p)
    agn = model.predict(None, st, nt, mx, True, loc_p, conf_p)
    li, lv = patch_arrays(loc_patch)
    ci, cv = patch_arrays(conf_patch)
    raw = synth.config(3, batch=2)["conf_infer"]
    diff = (raw != conf).nonzero()
    out.update(pred_loc_idx=li, pred_loc_val=lv, pred_conf_idx=ci, pred_conf_val=cv,
               pred_dedup_idx=diff.numpy().astype(np.int32), pred_dedup_val=conf[raw != conf].numpy(),
               pred_counts=np.array([o["labels"].numel() for o in got], dtype=np.int64),
               pred_labels=torch.cat([o["labels"] for o in got]).numpy().astype(np.int8),
               pred_scores=torch.cat([o["scores"] for o in got]).numpy(),
               pred_boxes=torch.cat([o["boxes"] for o in got]).numpy(),
               agn_counts=np.array([o["labels"].numel() for o in agn], dtype=np.int64),
               agn_labels=torch.cat([o["labels"] for o in agn]).numpy().astype(np.int8),
               agn_scores=torch.cat([o["scores"] for o in agn]).numpy(),
               agn_boxes=torch.cat([o["boxes"] for o in agn]).numpy(),
               clean_counts=np.array([o["labels"].numel() for o in clean], dtype=np.int64))
    n_nan = int(torch.isnan(torch.cat([o["boxes"] for o in got])).any(dim=1).sum())
    print(f"predict: {out['pred_counts'].tolist()} detections (clean {out['clean_counts'].tolist()}), {n_nan} with a NaN box")
    assert n_nan >= 2

    # ---- losses: cfg 1 (one image, 5 boxes), one patch per case ----------------------------------------------------
    c1 = synth.config(1)
    loc1, conf1, targets = c1["loc_all"], c1["conf_train"], c1["targets"]
    pos, loc_t_pm, cls_t, n_img, l_loc, l_conf = MG.ref_train(model, tr, targets, loc1, conf1, c1["iou_thresh"], c1["ratio"])
    pos_idx = pos[0].nonzero().flatten()
    p_pos = int(pos_idx[0])
    p_neg = int((~pos[0]).nonzero().flatten()[100])
    cases = [
        ("neg_nan_logit", "conf", [(0, p_neg, 2, NAN)]),            # CE NaN at a negative: topk takes it first -> NaN
        ("neg_bg_minus_inf", "conf", [(0, p_neg, 0, -INF)]),        # CE(background) = +Inf at a negative -> +Inf
        ("pos_nan_logit", "conf", [(0, p_pos, 1, NAN)]),
        ("pos_nan_offset", "loc", [(0, p_pos, 1, NAN)]),            # smooth-L1 NaN; the CE side stays finite
        ("neg_nan_offsets", "loc", [(0, p_neg, j, NAN) for j in range(4)]),      # never read: both losses unchanged
        ("neg_plus_inf_logit", "conf", [(0, p_neg, 3, INF)]),
        ("neg_fg_minus_inf", "conf", [(0, p_neg, 1, -INF)]),        # finite: that class just drops out of the softmax
        ("pos_inf_offset", "loc", [(0, p_pos, 2, INF)]),
    ]
    names, which, idx_all, val_all, starts, want = [], [], [], [], [0], []
    for name, w, patch in cases:
        lp = apply(loc1, patch) if w == "loc" else loc1
        cp = apply(conf1, patch) if w == "conf" else conf1
        r = MG.ref_train(model, tr, targets, lp, cp, c1["iou_thresh"], c1["ratio"])
        assert torch.equal(r[0], pos)
        i, v = patch_arrays(patch)
        names.append(name); which.append(w); idx_all.append(i); val_all.append(v); starts.append(starts[-1] + len(patch))
        want.append([np.float32(r[4].item()), np.float32(r[5].item())])
        print(f"{name:20s} loc_loss={r[4].item():.6g} conf_loss={r[5].item():.6g}")
    out.update(loss_case_names=np.array(names), loss_case_which=np.array(which), loss_case_start=np.array(starts, dtype=np.int32),
               loss_idx=np.concatenate(idx_all, 0), loss_val=np.concatenate(val_all, 0),
               loss_want=np.array(want, dtype=np.float32), loss_clean=np.array([l_loc.item(), l_conf.item()], dtype=np.float32),
               loss_loc_sha=np.array(MG.sha(loc1)), loss_conf_sha=np.array(MG.sha(conf1)))
    MG.save("nonfinite.npz", **out)


if __name__ == "__main__":
    main()
