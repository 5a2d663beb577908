"""CPU: pin the oracle (oracle/ssd_oracle.py) bit-for-bit to the golden fixtures that
tests/golden/make_golden.py produced by running the unmodified reference, and to the image's
torchvision for the restated box arithmetic."""
import numpy as np
import pytest
import torch

import _util as U
from oracle import ssd_oracle as O

TRAIN = ["train_cfg1.npz", "train_cfg2.npz", "train_cfg2_thr04.npz", "train_cfg5_b2.npz", "train_edges.npz"]
PREDICT = ["predict_cfg1.npz", "predict_cfg3_b4.npz", "predict_cfg3_b2_notebook.npz",
           "predict_cfg3_b2_agnostic.npz", "predict_cfg3_b2_empty.npz", "predict_cfg5_b1.npz"]


def test_priors_match_reference():
    g = U.load("priors.npz")
    pri, pri_xyxy = O.prior_tables()
    assert pri.shape == (O.NUM_PRIORS, 4)
    assert U.sha(pri) == str(g["priors_sha"]) == O.PRIORS_SHA256
    assert U.sha(pri_xyxy) == str(g["priors_xyxy_sha"]) == O.PRIORS_XYXY_SHA256
    assert np.array_equal(pri[:8].numpy(), g["priors_head"]) and np.array_equal(pri[-8:].numpy(), g["priors_tail"])


def test_box_math_matches_torchvision_bitwise():
    from torchvision.ops import box_convert, box_iou, complete_box_iou, distance_box_iou
    gen = torch.Generator().manual_seed(5)
    a = torch.rand((700, 2), generator=gen)
    rows = torch.cat((a, a + torch.rand((700, 2), generator=gen) * 0.5 + 1e-3), 1)
    b = torch.rand((90, 2), generator=gen)
    cols = torch.cat((b, b + torch.rand((90, 2), generator=gen) * 0.5 + 1e-3), 1)
    cols[3] = cols[2]                       # duplicate column
    cols[7, 2:] = cols[7, :2]               # zero-size box -> NaN in CIoU
    assert torch.equal(O.pairwise_iou(rows, cols), box_iou(rows, cols))
    assert torch.equal(O.pairwise_diou(rows, cols)[0], distance_box_iou(rows, cols))
    x, y = O.pairwise_ciou(rows, cols), complete_box_iou(rows, cols)
    assert torch.equal(torch.isnan(x), torch.isnan(y)) and torch.isnan(x).any()
    assert torch.equal(torch.nan_to_num(x, nan=-9.0), torch.nan_to_num(y, nan=-9.0))
    assert torch.equal(O.xyxy_to_cxcywh(rows), box_convert(rows, "xyxy", "cxcywh"))
    assert torch.equal(O.cxcywh_to_xyxy(rows), box_convert(rows, "cxcywh", "xyxy"))
    # DIoU is bitwise symmetric -> a triangular suppression mask is exact
    sq = O.pairwise_diou(rows, rows)[0]
    assert torch.equal(sq, sq.t())


@pytest.mark.parametrize("name", TRAIN)
def test_train_half_bitwise(name):
    g = U.load(name)
    targets, loc_all, conf = U.train_inputs(g)
    pri, pri_xyxy = O.prior_tables()
    thr, ratio = float(g["iou_thresh"]), float(g["ratio"])
    pos, loc_t_pm, cls_t = O.batch_targets(pri, pri_xyxy, targets, 300, 300, thr)
    assert torch.equal(pos, U.unpack_bits(g["pos_bits"], O.NUM_PRIORS))
    assert np.array_equal(cls_t.numpy(), g["cls_t"].astype(np.int64))
    assert np.array_equal(loc_t_pm.numpy(), g["loc_t_pm"])
    l_loc, n_img, total = O.loc_loss(loc_all, pos, loc_t_pm)
    assert np.array_equal(n_img.numpy(), g["n_pos"])
    l_conf = O.mined_ce_loss(conf, cls_t, pos, n_img, total, ratio)
    assert np.float32(l_loc.item()) == g["loc_loss"] and np.float32(l_conf.item()) == g["conf_loss"]
    # full encode_ssd outputs of image 0, negatives included
    t0 = targets[0]
    unit = t0["boxes"] / torch.tensor([300.0] * 4) if t0["boxes"].numel() else t0["boxes"].new_zeros((0, 4))
    e_loc, e_cls, e_pos, e_match = O.match_encode(pri, pri_xyxy, unit, t0["labels"], thr)
    assert np.array_equal(e_loc.numpy(), g["enc0_loc"], equal_nan=True)
    assert np.array_equal(e_cls.numpy(), g["enc0_cls"].astype(np.int64))
    assert torch.equal(e_pos, U.unpack_bits(g["enc0_pos"], O.NUM_PRIORS))
    assert np.array_equal(e_match.numpy(), g["enc0_match"], equal_nan=True)


@pytest.mark.parametrize("name", PREDICT)
def test_postprocess_bitwise(name):
    g = U.load(name)
    loc_all, conf = U.predict_inputs(g)
    pri, _ = O.prior_tables()
    got = O.postprocess(pri, loc_all, conf, float(g["score_thresh"]), float(g["nms_thresh"]),
                        int(g["max_per_img"]), bool(g["class_agnostic"]), nms_limit=True)
    want = U.split_predictions(g)
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert torch.equal(a["labels"], b["labels"])
        assert torch.equal(a["scores"], b["scores"])
        assert torch.equal(a["boxes"], b["boxes"])
        assert a["labels"].dtype == torch.int64 and a["boxes"].shape[1:] == (4,)


def test_static_methods_bitwise():
    g = U.load("static_methods.npz")
    boxes, scores = torch.from_numpy(g["boxes"]), torch.from_numpy(g["scores"])
    assert np.array_equal(O.greedy_nms(boxes, scores, 0.45).numpy(), g["keep45"])
    assert np.array_equal(O.greedy_nms(boxes, scores, 0.30).numpy(), g["keep30"])
    assert O.greedy_nms(boxes[:0], scores[:0], 0.45).shape == (0,)
    gen = torch.Generator().manual_seed(77)
    for _ in range(2):
        torch.rand((600, 2), generator=gen)
    torch.rand((600,), generator=gen)
    loc = torch.randn((8732, 4), generator=gen)
    assert U.sha(loc) == str(g["loc_sha"])
    pri, _ = O.prior_tables()
    assert np.array_equal(O.decode(loc, pri, (0.1, 0.2)).numpy(), g["decoded"])


def test_argument_validation_matches_reference():
    pri, pri_xyxy = O.prior_tables()
    with pytest.raises(ValueError):
        O.batch_targets(pri, pri_xyxy, [], iou_thresh=1.0)
    with pytest.raises(ValueError):
        O.match_encode(pri, pri_xyxy, torch.zeros((1, 4)), torch.zeros((1,), dtype=torch.int64), background_class=1)
    z = torch.zeros((1, 8732, 4)), torch.zeros((1, 8732, 6))
    with pytest.raises(ValueError):
        O.postprocess(pri, z[0], z[1], score_thresh=1.0)
    with pytest.raises(ValueError):
        O.postprocess(pri, z[0], z[1], nms_thresh=0.0)


def test_mined_negative_budget():
    assert O.mined_negative_budget(0, 8732, 3.0) == 3       # "pretend one positive" (SSD_trainer.py:586-588)
    assert O.mined_negative_budget(7, 8725, 2.5) == 17      # int() truncation (:590)
    assert O.mined_negative_budget(5000, 3732, 3.0) == 3732
    assert O.mined_negative_budget(3, 100, 0.2) == 0
    assert O.mined_negative_budget(8732, 0, 3.0) == 0


def test_pack_heads_restatement_layout():
    """oracle.pack_heads (SFS:249-269): element (b, level offset + (y*W + x)*A + a, j) of the packed tensor is element
    (b, a*D + j, y, x) of that level's NCHW head output."""
    gen = torch.Generator().manual_seed(5)
    levels = ((38, 4), (19, 6), (10, 6), (5, 6), (3, 4), (1, 4))
    C = 6
    loc_heads = [torch.randn((2, a * 4, n, n), generator=gen) for n, a in levels]
    conf_heads = [torch.randn((2, a * C, n, n), generator=gen) for n, a in levels]
    loc_all, conf_all = O.pack_heads(loc_heads, conf_heads, C)
    assert loc_all.shape == (2, 8732, 4) and conf_all.shape == (2, 8732, C)
    off = 0
    for (n, a), lh, ch in zip(levels, loc_heads, conf_heads):
        for (y, x, k) in ((0, 0, 0), (n - 1, n // 2, a - 1), (n // 2, n - 1, a // 2)):
            p = off + (y * n + x) * a + k
            assert torch.equal(loc_all[:, p, :], lh[:, k * 4:(k + 1) * 4, y, x])
            assert torch.equal(conf_all[:, p, :], ch[:, k * C:(k + 1) * C, y, x])
        off += n * n * a
    assert off == 8732


def test_nonfinite_head_outputs_bitwise():
    """NaN / +-Inf in loc_all / conf_all (tests/golden/make_golden_nonfinite.py ran the reference): the restatement keeps the
    reference's propagation rules -- torch.clamp / maximum / minimum keep NaN, NaN > thresh is False, `d <= thr` is False
    for a NaN box, torch.topk ranks NaN / Inf cross-entropies first."""
    g = U.load("nonfinite.npz")
    pri, xyxy = O.prior_tables()
    loc, conf = U.nonfinite_predict_inputs(g)
    for agn, pre in ((False, "pred"), (True, "agn")):
        want = U.split_rows(g[pre + "_counts"], g[pre + "_labels"], g[pre + "_scores"], g[pre + "_boxes"])
        got = O.postprocess(pri, loc, conf, 0.01, 0.45, 200, agn)
        assert sum(int(torch.isnan(w["boxes"]).any()) for w in want) >= 1
        for a, b in zip(got, want):
            assert torch.equal(a["labels"], b["labels"]) and torch.equal(a["scores"], b["scores"]) and U.equal_nan(a["boxes"], b["boxes"])
    for name, loc1, conf1, targets, cfg, (w_loc, w_conf) in U.nonfinite_loss_cases(g):
        l_loc, l_conf = O.train_half(pri, xyxy, loc1, conf1, targets, cfg["iou_thresh"], cfg["ratio"])
        assert U.same_float(l_loc.item(), w_loc, 0.0) and U.same_float(l_conf.item(), w_conf, 0.0), (name, l_loc.item(), w_loc, l_conf.item(), w_conf)
