"""GPU parity tests (run on the B200 box with `-m gpu`).  Every call goes through the C ABI of
libssdhot.so (ctypes); the oracle is only the checker.

Two oracles are used:
  * the DEVICE-MATCHED oracle -- oracle/ssd_oracle.py run on CUDA tensors (eager torch-CUDA:
    IEEE +,-,*,/ and libdevice atanf/expf/logf).  ssdhot must equal it BIT FOR BIT: indices,
    masks, keep lists AND every float (targets, boxes, scores);
  * the reference itself (CPU) through the golden fixtures of tests/golden/: indices, masks and
    keep lists must be identical; floats within 1e-5 relative (CPU SLEEF vs CUDA libdevice
    transcendentals differ in the last bit -- SURVEY.md section 7, hard part 2).
"""
import numpy as np
import pytest
import torch

import _util as U
from oracle import ssd_oracle as O

pytestmark = pytest.mark.gpu

RTOL = 1e-5                      # north_star: losses and decoded boxes within 1e-5 relative in fp32
BOX_ATOL = 1e-5 * 300.0          # boxes are pixel coordinates of a 300-px image


@pytest.fixture(scope="module")
def env():
    import ssdhot
    dev = torch.device("cuda:0")
    pri, pri_xyxy = O.prior_tables(dev)
    ps = ssdhot.PriorSet.default(dev)
    return dict(ssdhot=ssdhot, dev=dev, pri=pri, pri_xyxy=pri_xyxy, ps=ps)


def to_dev(targets, dev):
    return [{k: v.to(dev) for k, v in t.items()} for t in targets]


def bit_equal(a, b):
    a, b = a.detach().cpu(), b.detach().cpu()
    if a.shape != b.shape:
        return False
    if a.is_floating_point():
        return bool(((a == b) | (torch.isnan(a) & torch.isnan(b))).all())
    return bool((a == b).all())


def close(a, b, rtol=RTOL, atol=1e-6):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    nan = torch.isnan(a) & torch.isnan(b)
    return a.shape == b.shape and bool((((a - b).abs() <= atol + rtol * b.abs()) | nan).all())


# ------------------------------------------------------------------------------------------------
def test_library_is_native_and_loaded(env):
    s = env["ssdhot"]
    assert s.lib().ssdhot_abi_version() == 2
    before = s.launch_count()
    s.PriorSet.default(env["dev"])
    assert s.launch_count() == before + 1


def test_prior_tables_bitwise(env):
    ps, pri, pri_xyxy = env["ps"], env["pri"], env["pri_xyxy"]
    assert U.sha(ps.priors) == O.PRIORS_SHA256 and U.sha(ps.priors_xyxy) == O.PRIORS_XYXY_SHA256
    w, h = pri_xyxy[:, 2] - pri_xyxy[:, 0], pri_xyxy[:, 3] - pri_xyxy[:, 1]
    assert bit_equal(ps.aux[:, 0], w * h)
    assert bit_equal(ps.aux[:, 1], (pri_xyxy[:, 0] + pri_xyxy[:, 2]) / 2)
    assert bit_equal(ps.aux[:, 2], (pri_xyxy[:, 1] + pri_xyxy[:, 3]) / 2)
    assert bit_equal(ps.aux[:, 3], torch.atan(w / h))
    # a PriorSet built from caller-owned buffers (the reference model's) gives the same tables
    ps2 = env["ssdhot"].PriorSet(pri, pri_xyxy)
    assert bit_equal(ps2.aux, ps.aux)


TRAIN = ["train_cfg1.npz", "train_cfg2.npz", "train_cfg2_thr04.npz", "train_cfg5_b2.npz", "train_edges.npz"]


@pytest.mark.parametrize("name", TRAIN)
def test_match_encode(env, name):
    s, dev, ps = env["ssdhot"], env["dev"], env["ps"]
    g = U.load(name)
    targets, _, _ = U.train_inputs(g)
    thr = float(g["iou_thresh"])
    tg = to_dev(targets, dev)
    # device-matched oracle: everything bit for bit, negatives' targets included
    pos_o, locpm_o, cls_o, locd_o = O.batch_targets(env["pri"], env["pri_xyxy"], tg, 300, 300, thr, dense=True)
    r = s.match_encode_batch(ps, s.pack_targets(targets, dev), thr, (300, 300), want_loc="all",
                             want_matched_idx=True, want_matched_box=True)
    assert bit_equal(r["pos_mask"], pos_o) and bit_equal(r["cls_t"], cls_o) and bit_equal(r["loc_t"], locd_o)
    assert bit_equal(r["n_pos"].long(), pos_o.sum(1))
    for b, t in enumerate(tg):
        unit = t["boxes"] / torch.tensor([300.0] * 4, device=dev) if t["boxes"].numel() else t["boxes"].new_zeros((0, 4))
        e = O.match_encode(env["pri"], env["pri_xyxy"], unit, t["labels"], thr, return_match=True)
        assert bit_equal(r["matched_gt"][b].long(), e[4]), f"matched index, image {b}"
        assert bit_equal(r["matched_cxcywh"][b], e[3])
        if b >= 3:
            break
    # the reference itself (CPU golden): indices identical, offsets within 1e-5
    assert bit_equal(r["pos_mask"], U.unpack_bits(g["pos_bits"], 8732))
    assert bit_equal(r["cls_t"], torch.from_numpy(g["cls_t"].astype(np.int64)))
    assert bit_equal(r["n_pos"].long(), torch.from_numpy(g["n_pos"]))
    assert close(r["loc_t"][0], torch.from_numpy(g["enc0_loc"]), atol=1e-5)
    assert close(r["matched_cxcywh"][0], torch.from_numpy(g["enc0_match"]))
    # drop-in build_targets (SSD_trainer.py:491): same three outputs, same dtypes
    pm, lpm, ct = s.build_targets(ps, tg, 300, 300, thr, "cuda")
    assert pm.dtype == torch.bool and ct.dtype == torch.int64 and lpm.dtype == torch.float32
    assert bit_equal(pm, pos_o) and bit_equal(ct, cls_o) and bit_equal(lpm, locpm_o)
    assert close(lpm, torch.from_numpy(g["loc_t_pm"]), atol=1e-5)


def test_encode_ssd_dropin(env):
    s, dev, ps = env["ssdhot"], env["dev"], env["ps"]
    g = U.load("train_cfg1.npz")
    targets, _, _ = U.train_inputs(g)
    unit = (targets[0]["boxes"] / torch.tensor([300.0] * 4)).to(dev)
    lab = targets[0]["labels"].to(dev)
    got = s.encode_ssd(ps, unit, lab, iou_thresh=0.5)
    want = O.match_encode(env["pri"], env["pri_xyxy"], unit, lab, 0.5)
    for a, b in zip(got, want):
        assert a.dtype == b.dtype and bit_equal(a, b)
    # no ground truth: all-zero outputs (SSD_from_scratch.py:731-736)
    got = s.encode_ssd(ps, unit[:0], lab[:0])
    assert all(int(x.abs().sum()) == 0 for x in got) and got[2].dtype == torch.bool and got[0].shape == (8732, 4)
    with pytest.raises(ValueError):
        s.encode_ssd(ps, unit, lab, background_class=1)


@pytest.mark.parametrize("name", TRAIN)
def test_losses(env, name):
    s, dev, ps = env["ssdhot"], env["dev"], env["ps"]
    g = U.load(name)
    targets, loc_all, conf = U.train_inputs(g)
    thr, ratio = float(g["iou_thresh"]), float(g["ratio"])
    tg, lg, cg = to_dev(targets, dev), loc_all.to(dev), conf.to(dev)
    pos_o, locpm_o, cls_o = O.batch_targets(env["pri"], env["pri_xyxy"], tg, 300, 300, thr)
    o_loc, n_img, total = O.loc_loss(lg, pos_o, locpm_o)
    o_conf = O.mined_ce_loss(cg, cls_o, pos_o, n_img, total, ratio)
    l_loc, l_conf, sums = s.multibox_loss(ps, lg, cg, targets, thr, ratio, return_sums=True)
    assert l_loc.dtype == torch.float32 and l_loc.dim() == 0
    assert float(sums[2]) == float(n_img.sum())
    for got, dev_o, gold in ((l_loc, o_loc, g["loc_loss"]), (l_conf, o_conf, g["conf_loss"])):
        assert abs(got.item() - dev_o.item()) <= RTOL * abs(dev_o.item())
        assert abs(got.item() - float(gold)) <= RTOL * abs(float(gold))
    # drop-in CELoss_w_neg_mining (SSD_trainer.py:551) with the reference's own argument list
    c2 = s.CELoss_w_neg_mining(cg, cls_o, pos_o, n_img, total, ratio)
    assert abs(c2.item() - float(g["conf_loss"])) <= RTOL * abs(float(g["conf_loss"]))
    l2 = s.smooth_l1_positive_loss(lg, pos_o, locpm_o, total)
    assert abs(l2.item() - float(g["loc_loss"])) <= RTOL * abs(float(g["loc_loss"]))


@pytest.mark.parametrize("ratio", [3.0, 0.0, 0.4, 1e6])
def test_mining_budget_edges(env, ratio):
    """int() truncation, the n_pos == 0 rule, k == 0 and k >= #negatives (SSD_trainer.py:585-596)."""
    s, dev, ps = env["ssdhot"], env["dev"], env["ps"]
    g = U.load("train_edges.npz")
    targets, loc_all, conf = U.train_inputs(g)
    tg, lg, cg = to_dev(targets, dev), loc_all.to(dev), conf.to(dev)
    pos_o, locpm_o, cls_o = O.batch_targets(env["pri"], env["pri_xyxy"], tg, 300, 300, 0.5)
    _, n_img, total = O.loc_loss(lg, pos_o, locpm_o)
    want = O.mined_ce_loss(cg, cls_o, pos_o, n_img, total, ratio)
    got = s.CELoss_w_neg_mining(cg, cls_o, pos_o, n_img, total, ratio)
    assert abs(got.item() - want.item()) <= RTOL * max(abs(want.item()), 1e-12)
    _, fused = s.multibox_loss(ps, lg, cg, targets, 0.5, ratio)
    assert abs(fused.item() - want.item()) <= RTOL * max(abs(want.item()), 1e-12)


def test_mining_with_massive_ties(env):
    """All logits equal: every CE value ties; the sum is tie-invariant and the backward mask takes
    the lowest prior indices (documented tie rule)."""
    s, dev, ps = env["ssdhot"], env["dev"], env["ps"]
    g = U.load("train_cfg1.npz")
    targets, loc_all, _ = U.train_inputs(g)
    conf = torch.zeros((1, 8732, 6), device=dev, requires_grad=True)
    lg = loc_all.to(dev)
    tg = to_dev(targets, dev)
    pos_o, locpm_o, cls_o = O.batch_targets(env["pri"], env["pri_xyxy"], tg, 300, 300, 0.5)
    _, n_img, total = O.loc_loss(lg, pos_o, locpm_o)
    want = O.mined_ce_loss(conf.detach(), cls_o, pos_o, n_img, total, 3.0)
    _, got = s.multibox_loss(ps, lg, conf, targets, 0.5, 3.0)
    assert abs(got.item() - want.item()) <= RTOL * abs(want.item())
    got.backward()
    touched = (conf.grad.abs().sum(-1) > 0)[0]
    n_pos = int(n_img[0])
    assert int(touched.sum()) == n_pos + 3 * n_pos
    neg_idx = (~pos_o[0]).nonzero()[:, 0][: 3 * n_pos]
    assert bool(touched[neg_idx].all())


def test_loss_backward_matches_autograd(env):
    s, dev, ps = env["ssdhot"], env["dev"], env["ps"]
    g = U.load("train_cfg2_thr04.npz")
    targets, loc_all, conf = U.train_inputs(g)
    thr, ratio = float(g["iou_thresh"]), float(g["ratio"])
    tg = to_dev(targets, dev)
    la, ca = loc_all.to(dev).requires_grad_(True), conf.to(dev).requires_grad_(True)
    lb, cb = loc_all.to(dev).requires_grad_(True), conf.to(dev).requires_grad_(True)
    l_loc, l_conf = s.multibox_loss(ps, la, ca, targets, thr, ratio)
    (l_loc * 1.5 + l_conf * 0.7).backward()
    pos_o, locpm_o, cls_o = O.batch_targets(env["pri"], env["pri_xyxy"], tg, 300, 300, thr)
    o_loc, n_img, total = O.loc_loss(lb, pos_o, locpm_o)
    o_conf = O.mined_ce_loss(cb, cls_o, pos_o, n_img, total, ratio)
    (o_loc * 1.5 + o_conf * 0.7).backward()
    assert close(la.grad, lb.grad, atol=1e-7) and close(ca.grad, cb.grad, atol=1e-7)
    # drop-in CE alone
    cc = conf.to(dev).requires_grad_(True)
    s.CELoss_w_neg_mining(cc, cls_o, pos_o, n_img, total, ratio).backward()
    cd = conf.to(dev).requires_grad_(True)
    O.mined_ce_loss(cd, cls_o, pos_o, n_img, total, ratio).backward()
    assert close(cc.grad, cd.grad, atol=1e-7)


PREDICT = ["predict_cfg1.npz", "predict_cfg3_b4.npz", "predict_cfg3_b2_notebook.npz",
           "predict_cfg3_b2_agnostic.npz", "predict_cfg3_b2_empty.npz", "predict_cfg5_b1.npz"]


@pytest.mark.parametrize("name", PREDICT)
def test_predict(env, name):
    s, dev, ps = env["ssdhot"], env["dev"], env["ps"]
    g = U.load(name)
    loc_all, conf = U.predict_inputs(g)
    st, nt, mx, ag = float(g["score_thresh"]), float(g["nms_thresh"]), int(g["max_per_img"]), bool(g["class_agnostic"])
    lg, cg = loc_all.to(dev), conf.to(dev)
    want = O.postprocess(env["pri"], lg, cg, st, nt, mx, ag, nms_limit=True, with_index=True)
    gold = U.split_predictions(g)
    got = s.predict(ps, None, st, nt, mx, ag, pre_loc_all=lg, pre_conf_all=cg)
    _, _, _, count, cand = s.predict_padded(ps, lg, cg, st, nt, mx, ag, want_cand=True)
    assert len(got) == len(want) == len(gold)
    for b, (a, w, r) in enumerate(zip(got, want, gold)):
        assert a["labels"].dtype == torch.int64 and a["scores"].dtype == torch.float32 and a["boxes"].shape[1:] == (4,)
        # device-matched: keep list (candidate ids), labels, scores, boxes bit for bit
        assert bit_equal(cand[b, : int(count[b])].long(), w["cand"]), f"keep list, image {b}"
        assert bit_equal(a["labels"], w["labels"]) and bit_equal(a["scores"], w["scores"]) and bit_equal(a["boxes"], w["boxes"])
        # the reference (CPU golden): same detections in the same order
        assert bit_equal(a["labels"], r["labels"])
        assert close(a["scores"], r["scores"]) and close(a["boxes"], r["boxes"], atol=BOX_ATOL)


def test_predict_ties_and_metrics(env):
    """Equal scores everywhere (all logits zero): candidates are ordered by candidate id (prior-major,
    class-minor), the documented resolution of the reference's unstable argsort; and the CIoU / IoU
    NMS variants agree with the oracle's."""
    s, dev, ps = env["ssdhot"], env["dev"], env["ps"]
    gen = torch.Generator().manual_seed(11)
    loc = torch.randn((2, 8732, 4), generator=gen).to(dev)
    conf = torch.zeros((2, 8732, 6), device=dev)
    conf[1] = torch.randn((8732, 6), generator=gen).to(dev)
    for agn in (False, True):
        want = O.postprocess(env["pri"], loc, conf, 0.05, 0.45, 60, agn, nms_limit=True, with_index=True)
        *_, count, cand = s.predict_padded(ps, loc, conf, 0.05, 0.45, 60, agn, want_cand=True)
        for b in range(2):
            assert bit_equal(cand[b, : int(count[b])].long(), want[b]["cand"])
    for metric in ("ciou",):
        want = O.postprocess(env["pri"], loc, conf, 0.05, 0.45, 60, False, metric=metric, nms_limit=True, with_index=True)
        *_, count, cand = s.predict_padded(ps, loc, conf, 0.05, 0.45, 60, False, metric=metric, want_cand=True)
        for b in range(2):
            assert bit_equal(cand[b, : int(count[b])].long(), want[b]["cand"])


def test_static_methods(env):
    s, dev, ps = env["ssdhot"], env["dev"], env["ps"]
    g = U.load("static_methods.npz")
    boxes, scores = torch.from_numpy(g["boxes"]).to(dev), torch.from_numpy(g["scores"]).to(dev)
    for thr, key in ((0.45, "keep45"), (0.30, "keep30")):
        k = s.iou_nms(boxes, scores, thr)
        assert k.dtype == torch.int64 and bit_equal(k, torch.from_numpy(g[key]))
        # idempotence: the survivors survive themselves
        k2 = s.iou_nms(boxes[k], scores[k], thr)
        assert bit_equal(k2, torch.arange(k.numel()))
    assert s.iou_nms(boxes[:0], scores[:0], 0.45).shape == (0,)
    # NaN suppresses (zero-area twins): SURVEY.md section 7 hard part 3
    z = torch.tensor([[5., 5., 5., 5.], [5., 5., 5., 5.], [50., 50., 80., 80.]], device=dev)
    zs = torch.tensor([0.9, 0.8, 0.7], device=dev)
    assert bit_equal(s.iou_nms(z, zs, 0.45), O.greedy_nms(z, zs, 0.45))
    # several sets in one launch + more boxes than one ranking chunk
    gen = torch.Generator().manual_seed(3)
    n = 3000
    ctr = torch.rand((n, 2), generator=gen) * 300
    wh = torch.rand((n, 2), generator=gen) * 60 + 2
    big = torch.cat((ctr - wh / 2, ctr + wh / 2), 1).clamp(0, 300).to(dev)
    bs = torch.rand((n,), generator=gen).to(dev)
    bs[100:140] = bs[100]                      # a run of exactly equal scores
    sizes = [1700, 0, 1300]
    keep, cnt = s.nms_sets(big, bs, sizes, 0.45)
    off = 0
    for i, m in enumerate(sizes):
        want = O.greedy_nms(big[off:off + m], bs[off:off + m], 0.45)
        assert int(cnt[i]) == want.numel() and bit_equal(keep[off: off + want.numel()], want)
        off += m
    for metric in ("ciou",):
        want = O.greedy_nms(big[:1700], bs[:1700], 0.45, metric=metric)
        assert bit_equal(s.iou_nms(big[:1700], bs[:1700], 0.45, metric=metric), want)
    # decode
    gen = torch.Generator().manual_seed(77)
    for _ in range(2):
        torch.rand((600, 2), generator=gen)
    torch.rand((600,), generator=gen)
    loc = torch.randn((8732, 4), generator=gen).to(dev)
    d = s.decode_ssd(loc, ps.priors, (0.1, 0.2))
    assert bit_equal(d, O.decode(loc, env["pri"], (0.1, 0.2))) and close(d, torch.from_numpy(g["decoded"]))


def test_full_size_properties(env):
    """BASELINE configs at full batch (B=256): properties that need no oracle at that size, plus a
    bit-exact spot check of a few images against the device-matched oracle."""
    s, dev, ps = env["ssdhot"], env["dev"], env["ps"]
    from ssdhot import synth
    cfg = synth.config(3)
    lg, ct, ci = cfg["loc_all"].to(dev), cfg["conf_train"].to(dev), cfg["conf_infer"].to(dev)
    targets = cfg["targets"]
    B = cfg["batch"]
    # -- training half: batch invariance + additivity of the un-normalised sums over shards
    _, _, full = s.multibox_loss(ps, lg, ct, targets, 0.5, 3.0, return_sums=True)
    parts = torch.zeros(3, dtype=torch.float64, device=dev)
    for lo in range(0, B, 64):
        _, _, p = s.multibox_loss(ps, lg[lo:lo + 64], ct[lo:lo + 64], targets[lo:lo + 64], 0.5, 3.0, return_sums=True)
        parts += p
    assert float(parts[2]) == float(full[2])
    assert close(parts[:2], full[:2], rtol=1e-9)
    r = s.match_encode_batch(ps, s.pack_targets(targets, dev), 0.5, (300, 300), want_loc="all", want_matched_idx=True)
    pick = [0, 17, 101, 255]
    sub = s.match_encode_batch(ps, s.pack_targets([targets[i] for i in pick], dev), 0.5, (300, 300), want_loc="all",
                               want_matched_idx=True)
    for j, i in enumerate(pick):
        for key in ("pos_mask", "cls_t", "loc_t", "matched_gt"):
            assert bit_equal(r[key][i], sub[key][j])
    tg = to_dev([targets[i] for i in pick], dev)
    pos_o, _, cls_o, locd_o = O.batch_targets(env["pri"], env["pri_xyxy"], tg, 300, 300, 0.5, dense=True)
    assert bit_equal(sub["pos_mask"], pos_o) and bit_equal(sub["cls_t"], cls_o) and bit_equal(sub["loc_t"], locd_o)
    # every ground-truth box got its forced prior: >= 1 positive per distinct best prior, and
    # every positive's class is one of the image's labels + 1
    assert int((r["n_pos"] <= 0).sum()) == 0
    # -- inference half
    labels, scores, boxes, count, cand = s.predict_padded(ps, lg, ci, 0.01, 0.45, 200, want_cand=True)
    cnt = count.tolist()
    assert all(0 <= k <= 200 for k in cnt)
    for i in range(B):
        k = cnt[i]
        sc = scores[i, :k]
        assert bool((sc[1:] <= sc[:-1]).all()) and bool((sc > 0.01).all())
        bx = boxes[i, :k]
        assert bool((bx >= 0).all()) and bool((bx <= 300).all())
        assert bool((labels[i, :k] == (cand[i, :k] % 5)).all())
    want = O.postprocess(env["pri"], lg[pick], ci[pick], 0.01, 0.45, 200, False, nms_limit=True, with_index=True)
    for j, i in enumerate(pick):
        assert bit_equal(cand[i, :cnt[i]].long(), want[j]["cand"])
        assert bit_equal(scores[i, :cnt[i]], want[j]["scores"]) and bit_equal(boxes[i, :cnt[i]], want[j]["boxes"])
    # per-class NMS fixpoint: re-running NMS on one image's detections of one class keeps them all
    i = 5
    k = cnt[i]
    for c in range(5):
        m = labels[i, :k] == c
        if int(m.sum()) > 1:
            kk = s.iou_nms(boxes[i, :k][m], scores[i, :k][m], 0.45)
            assert bit_equal(kk, torch.arange(int(m.sum())))


def test_stress_config_slice(env):
    """cfg 5 (64 GT boxes per image; threshold 0.0 so all 8732 x 5 pairs enter NMS) on a slice."""
    s, dev, ps = env["ssdhot"], env["dev"], env["ps"]
    from ssdhot import synth
    cfg = synth.config(5, batch=3)
    lg, ct, ci = cfg["loc_all"].to(dev), cfg["conf_train"].to(dev), cfg["conf_infer"].to(dev)
    tg = to_dev(cfg["targets"], dev)
    pos_o, locpm_o, cls_o, locd_o = O.batch_targets(env["pri"], env["pri_xyxy"], tg, 300, 300, 0.5, dense=True)
    r = s.match_encode_batch(ps, s.pack_targets(cfg["targets"], dev), 0.5, (300, 300), want_loc="all")
    assert bit_equal(r["pos_mask"], pos_o) and bit_equal(r["cls_t"], cls_o) and bit_equal(r["loc_t"], locd_o)
    want = O.postprocess(env["pri"], lg, ci, 0.0, 0.45, 200, False, nms_limit=True, with_index=True)
    *_, count, cand = s.predict_padded(ps, lg, ci, 0.0, 0.45, 200, want_cand=True)
    for b in range(3):
        assert bit_equal(cand[b, : int(count[b])].long(), want[b]["cand"])
    # no early stop: the stand-alone NMS over one full class (8732 boxes) equals the oracle's full run
    sc = ci[0].softmax(-1)[:, 1]
    box = O.decode(lg[0], env["pri"], (0.1, 0.2))
    xyxy = (O.cxcywh_to_xyxy(box).clamp(0, 1) * 300).contiguous()
    assert bit_equal(s.iou_nms(xyxy, sc, 0.45), O.greedy_nms(xyxy, sc, 0.45))


def test_error_behaviour(env):
    s, dev, ps = env["ssdhot"], env["dev"], env["ps"]
    z_loc, z_conf = torch.zeros((1, 8732, 4), device=dev), torch.zeros((1, 8732, 6), device=dev)
    with pytest.raises(ValueError):
        s.predict(ps, None, score_thresh=1.0, pre_loc_all=z_loc, pre_conf_all=z_conf)
    with pytest.raises(ValueError):
        s.predict(ps, None, nms_thresh=1.0, pre_loc_all=z_loc, pre_conf_all=z_conf)
    with pytest.raises(ValueError):
        s.build_targets(ps, [], iou_thresh=0.0)
    with pytest.raises(s.SsdhotError):
        s.decode_ssd(torch.zeros((4, 4)), torch.zeros((4, 4)), (0.1, 0.2))     # CPU tensors: no fallback
    with pytest.raises(s.SsdhotError):
        s.build_targets(ps, [{"boxes": torch.zeros((0, 4)), "labels": torch.zeros((0,), dtype=torch.int64)}], device="cpu")
    # raw ABI status codes
    L = s.lib()
    assert L.ssdhot_decode(None, None, 4, 0.1, 0.2, None, None) == -1
    assert L.ssdhot_prior_tables(ps.priors.data_ptr(), 20000, ps.priors_xyxy.data_ptr(), ps.aux.data_ptr(), None) == -2
    assert L.ssdhot_decode(ps.priors.data_ptr() + 4, ps.priors.data_ptr(), 4, 0.1, 0.2, ps.aux.data_ptr(), None) == -5


def test_many_classes_fallback_path(env):
    """C = 21 (VOC-sized heads): the dense score array of one image no longer fits in shared memory,
    so predict runs per (image, class) + merge; matching / loss run the generic-C code."""
    s, dev, ps = env["ssdhot"], env["dev"], env["ps"]
    gen = torch.Generator().manual_seed(21)
    loc = torch.randn((2, 8732, 4), generator=gen).to(dev)
    conf = torch.randn((2, 8732, 21), generator=gen)
    conf[..., 0] += 5.0
    conf = conf.to(dev)
    want = O.postprocess(env["pri"], loc, conf, 0.02, 0.45, 100, False, nms_limit=True, with_index=True)
    labels, scores, boxes, count, cand = s.predict_padded(ps, loc, conf, 0.02, 0.45, 100, want_cand=True)
    for b in range(2):
        k = int(count[b])
        assert bit_equal(cand[b, :k].long(), want[b]["cand"]) and bit_equal(scores[b, :k], want[b]["scores"])
        assert bit_equal(boxes[b, :k], want[b]["boxes"]) and bit_equal(labels[b, :k], want[b]["labels"])
    from ssdhot import synth
    targets = synth.make_targets(2, 3, 9, gen, n_fg=20)
    tg = to_dev(targets, dev)
    conf_t = torch.randn((2, 8732, 21), generator=gen).to(dev)
    pos_o, locpm_o, cls_o = O.batch_targets(env["pri"], env["pri_xyxy"], tg, 300, 300, 0.5)
    o_loc, n_img, total = O.loc_loss(loc, pos_o, locpm_o)
    o_conf = O.mined_ce_loss(conf_t, cls_o, pos_o, n_img, total, 3.0)
    l_loc, l_conf = s.multibox_loss(ps, loc, conf_t, targets, 0.5, 3.0)
    assert abs(l_loc.item() - o_loc.item()) <= RTOL * abs(o_loc.item())
    assert abs(l_conf.item() - o_conf.item()) <= RTOL * abs(o_conf.item())


def test_pruned_sweep_equals_exact_sweep(env):
    """The hot path prunes (prior, box) pairs that cannot matter; positives, classes, positives'
    offsets and n_pos must be bit-identical to the exact-everywhere sweep -- including boxes outside
    every prior's reach (fallback), tiny boxes and heavy crowding."""
    s, dev, ps = env["ssdhot"], env["dev"], env["ps"]
    from ssdhot import synth
    gen = torch.Generator().manual_seed(5)
    targets = synth.make_targets(24, 0, 40, gen)
    f = lambda rows: torch.tensor(rows, dtype=torch.float32).reshape(-1, 4)
    targets[1] = {"boxes": f([[310, 310, 330, 340], [-50, -40, -10, -5]]), "labels": torch.tensor([1, 2])}   # outside the image
    targets[2] = {"boxes": f([[0, 0, 1, 1], [299, 299, 300, 300], [150, 150, 150.5, 150.5]]), "labels": torch.tensor([0, 1, 2])}
    targets[3] = {"boxes": f([[100, 100, 100, 100]]), "labels": torch.tensor([3])}                         # NaN column only
    packed = s.pack_targets(targets, dev)
    for thr in (0.5, 0.2, 0.9):
        exact = s.match_encode_batch(ps, packed, thr, (300, 300), want_loc="all", want_matched_idx=True)
        pruned = s.match_encode_batch(ps, packed, thr, (300, 300), want_loc="positives")
        assert bit_equal(pruned["pos_mask"], exact["pos_mask"]) and bit_equal(pruned["cls_t"], exact["cls_t"])
        assert bit_equal(pruned["n_pos"], exact["n_pos"])
        m = exact["pos_mask"]
        assert bit_equal(pruned["loc_t"][m], exact["loc_t"][m])
        tg = to_dev(targets, dev)
        pos_o, _, cls_o = O.batch_targets(env["pri"], env["pri_xyxy"], tg, 300, 300, thr)
        assert bit_equal(exact["pos_mask"], pos_o) and bit_equal(exact["cls_t"], cls_o)


def _raw_loss(s, ps, loc, conf, targets, thr, ratio):
    """ssdhot_multibox_loss_fwd with every optional output -> (sums f64[3], sel i8[B,P], matched i16[B,P], n_pos i32[B])."""
    from ssdhot import _lib
    dev = loc.device
    B, P, C = conf.shape
    packed = s.pack_targets(targets, dev)
    sums = torch.empty((3,), dtype=torch.float64, device=dev)
    sel = torch.empty((B, P), dtype=torch.int8, device=dev)
    matched = torch.empty((B, P), dtype=torch.int16, device=dev)
    n_pos = torch.empty((B,), dtype=torch.int32, device=dev)
    work = torch.empty((int(s.lib().ssdhot_loss_workspace_bytes(B, P, packed.max_gt)),), dtype=torch.uint8, device=dev)
    rc = s.lib().ssdhot_multibox_loss_fwd(
        ps.priors.data_ptr(), ps.priors_xyxy.data_ptr(), ps.aux.data_ptr(), P, ps.layout,
        packed.boxes.data_ptr(), packed.labels.data_ptr(), packed.offsets.data_ptr(), B, packed.max_gt, 300.0, 300.0,
        loc.data_ptr(), conf.data_ptr(), C, float(thr), 0.1, 0.2, float(ratio),
        sums.data_ptr(), work.data_ptr(), sel.data_ptr(), matched.data_ptr(), n_pos.data_ptr(), None, None,
        torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc, "ssdhot_multibox_loss_fwd")
    torch.cuda.synchronize(dev)
    return sums, sel, matched, n_pos


def test_fast_train_path_equals_generic_path(env):
    """The SSD300 fast path (box-centric matching over candidate rectangles, approximate CE +
    error-band refinement for the mining) must pick exactly the positives, matched boxes and hard
    negatives of the layout-agnostic exact kernels, and produce the same sums (same fp32 terms,
    double accumulation in a different order -> 1e-12)."""
    s, dev, ps = env["ssdhot"], env["dev"], env["ps"]
    from ssdhot import synth
    psg = s.PriorSet.default(dev, generic=True)
    assert ps.layout == 1 and psg.layout == 0
    gen = torch.Generator().manual_seed(77)
    f = lambda rows: torch.tensor(rows, dtype=torch.float32).reshape(-1, 4)
    cases = []
    t_a = synth.make_targets(24, 0, 40, gen)
    t_a[1] = {"boxes": f([[310, 310, 330, 340], [-50, -40, -10, -5]]), "labels": torch.tensor([1, 2])}
    t_a[2] = {"boxes": f([[0, 0, 1, 1], [299, 299, 300, 300], [150, 150, 150.5, 150.5]]), "labels": torch.tensor([0, 1, 2])}
    t_a[3] = {"boxes": f([[100, 100, 100, 100]]), "labels": torch.tensor([3])}
    t_a[4] = {"boxes": f([[100, 100, 100, 100], [20, 30, 200, 250], [20, 30, 200, 250]]), "labels": torch.tensor([3, 1, 4])}
    t_a[5] = {"boxes": f([[0, 0, 300, 300], [0, 0, 300, 10], [140, 0, 160, 300]]), "labels": torch.tensor([0, 1, 2])}
    t_a[6] = {"boxes": f([[50, 50, 40, 90], [float("inf"), 0, 10, 10], [10, 10, 60, 60]]), "labels": torch.tensor([0, 1, 2])}
    loc_a, conf_a = synth.make_heads(24, gen)
    cases.append(("edges", t_a, loc_a, conf_a))
    t_b = synth.make_targets(6, 64, 64, gen)
    loc_b, conf_b = synth.make_heads(6, gen)
    cases.append(("g64", t_b, loc_b, conf_b))
    t_c = synth.make_targets(8, 1, 20, gen)
    loc_c, conf_c = synth.make_heads(8, gen)
    conf_big = conf_c * 30.0
    conf_tie = torch.zeros_like(conf_c)
    conf_tie[:, ::7] = conf_c[:, ::7].round()                 # many exact ties, a few distinct values
    conf_bias = conf_c.clone()
    conf_bias[..., 0] += 6.0
    cases += [("big", t_c, loc_c, conf_big), ("ties", t_c, loc_c, conf_tie), ("bias", t_c, loc_c, conf_bias)]
    for name, targets, loc, conf in cases:
        lg, cg = loc.to(dev), conf.to(dev)
        for thr in (0.5, 0.25):
            for ratio in (3.0, 0.4, 0.0, 1e6):
                fast = _raw_loss(s, ps, lg, cg, targets, thr, ratio)
                slow = _raw_loss(s, psg, lg, cg, targets, thr, ratio)
                tag = (name, thr, ratio)
                assert bit_equal(fast[3], slow[3]), tag
                assert bit_equal(fast[2], slow[2]), tag
                assert bit_equal(fast[1], slow[1]), tag
                assert torch.allclose(fast[0], slow[0], rtol=1e-12, atol=1e-9), (tag, fast[0], slow[0])
            # and both agree with the reference semantics (device-matched oracle), exotic boxes included
            pos_o, _, cls_o = O.batch_targets(env["pri"], env["pri_xyxy"], to_dev(targets, dev), 300, 300, thr)
            assert bit_equal(fast[1] > 0, pos_o), (name, thr)
            assert bit_equal(fast[1].long().clamp_min(0)[pos_o], cls_o[pos_o]), (name, thr)


def test_predict_threshold_edge_is_exact(env):
    """score_kernel ranks with approximate softmax scores but must take the candidate SET from the exact
    arithmetic: thresholds placed exactly on, one ulp below and one ulp above real scores flip exactly the
    candidates the strict `score > thresh` of the reference (SFS:402) flips."""
    s, dev, ps, pri = env["ssdhot"], env["dev"], env["ps"], env["pri"]
    gen = torch.Generator().manual_seed(314)
    loc = torch.randn((1, 8732, 4), generator=gen).to(dev)
    conf = torch.randn((1, 8732, 6), generator=gen)
    conf[..., 0] += 3.0
    conf = conf.to(dev)
    scores = conf[0].softmax(-1)[:, 1:].reshape(-1)
    srt = torch.sort(scores, descending=True).values
    for rank in (50, 400, 3000):
        sv = srt[rank]
        for thr in (sv, torch.nextafter(sv, sv.new_tensor(0.0)), torch.nextafter(sv, sv.new_tensor(1.0))):
            thr = float(thr)
            n_ref = int((scores > thr).sum())
            want = O.postprocess(pri, loc, conf, thr, 0.45, 8732 * 5 // 64, False, nms_limit=True, with_index=True)
            *_, count, cand = s.predict_padded(ps, loc, conf, thr, 0.45, 8732 * 5 // 64, want_cand=True)
            assert bit_equal(cand[0, : int(count[0])].long(), want[0]["cand"]), (rank, thr, n_ref)
    # NMS off (threshold close to 1): every candidate survives, so the output count IS the candidate count
    for rank in (20, 150):
        sv = srt[rank]
        for thr, expect in ((float(sv), rank), (float(torch.nextafter(sv, sv.new_tensor(0.0))), rank + 1)):
            dup = int((srt == sv).sum())                     # (exact duplicates of the probe score would shift the count)
            if dup != 1:
                continue
            *_, count = s.predict_padded(ps, loc, conf, thr, 0.999999, 600)
            assert int(count[0]) == expect, (rank, thr, int(count[0]), expect)


def test_predict_multi_round_suppression(env):
    """Heavy suppression: decoded boxes are the priors themselves (loc = 0) and one class takes nearly all the
    score mass, so neighbouring priors of a cell suppress each other and the first pull of ~1.5 max_per_img
    candidates does not yield max_per_img survivors -- the later rounds (tests against earlier survivors,
    per-class lists) must reproduce the reference's greedy order exactly."""
    s, dev, ps, pri = env["ssdhot"], env["dev"], env["ps"], env["pri"]
    gen = torch.Generator().manual_seed(2718)
    loc = (0.05 * torch.randn((3, 8732, 4), generator=gen)).to(dev)
    conf = torch.randn((3, 8732, 6), generator=gen)
    conf[0, :, 2] += 4.0                                     # image 0: almost everything is class 1 (one long class list)
    conf[1, :, 1:] += 1.0                                    # image 1: all classes busy
    conf[2, :, 0] += 8.0                                     # image 2: sparse
    conf = conf.to(dev)
    for thr, nms, keep, agn in ((0.05, 0.3, 200, False), (0.02, 0.45, 400, False), (0.05, 0.3, 150, True)):
        want = O.postprocess(pri, loc, conf, thr, nms, keep, agn, nms_limit=True, with_index=True)
        labels, scores, boxes, count, cand = s.predict_padded(ps, loc, conf, thr, nms, keep, agn, want_cand=True)
        for b in range(3):
            k = int(count[b])
            assert bit_equal(cand[b, :k].long(), want[b]["cand"]), (thr, nms, keep, agn, b, k, want[b]["cand"].numel())
            assert bit_equal(scores[b, :k], want[b]["scores"]) and bit_equal(boxes[b, :k], want[b]["boxes"])
            assert bit_equal(labels[b, :k], want[b]["labels"])


def test_pack_heads_matches_reference_layout(env):
    """ssdhot.pack_heads == the tail of mySSD.forward (permute(0,2,3,1).contiguous() x 12, cat x 2; SFS:249-269),
    bit for bit (pure data movement), for the reference's 6 classes and a VOC-sized head, contiguous or not."""
    s, dev = env["ssdhot"], env["dev"]
    gen = torch.Generator().manual_seed(99)
    levels = ((38, 4), (19, 6), (10, 6), (5, 6), (3, 4), (1, 4))
    for B, C in ((3, 6), (2, 21), (1, 2)):
        loc_heads = [torch.randn((B, a * 4, n, n), generator=gen).to(dev) for n, a in levels]
        conf_heads = [torch.randn((B, a * C, n, n), generator=gen).to(dev) for n, a in levels]
        conf_heads[1] = conf_heads[1].to(memory_format=torch.channels_last)          # a non-contiguous producer
        want_loc, want_conf = O.pack_heads(loc_heads, conf_heads, C)
        got_loc, got_conf = s.pack_heads(loc_heads, conf_heads)
        assert got_loc.shape == (B, 8732, 4) and got_conf.shape == (B, 8732, C)
        assert bit_equal(got_loc, want_loc) and bit_equal(got_conf, want_conf)
    with pytest.raises(ValueError):
        s.pack_heads(loc_heads[:5], conf_heads)


# ------------------------------------------------------------------------------------------------
# head-direct entry points: the kernels read the six head outputs of each branch
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("channels_last", [False, True])
@pytest.mark.parametrize("name", PREDICT)
def test_predict_from_heads(env, name, channels_last):
    """predict_heads(heads) == predict(pack(heads)) == the reference's detections: keep lists, labels, scores and boxes
    bit for bit against the device-matched oracle, for NCHW heads (as the conv heads return them) and channels_last ones."""
    s, dev, ps = env["ssdhot"], env["dev"], env["ps"]
    g = U.load(name)
    loc_all, conf = U.predict_inputs(g)
    st, nt, mx, ag = float(g["score_thresh"]), float(g["nms_thresh"]), int(g["max_per_img"]), bool(g["class_agnostic"])
    lg, cg = loc_all.to(dev), conf.to(dev)
    loc_heads, conf_heads = U.unpack_heads(lg, channels_last), U.unpack_heads(cg, channels_last)
    back_loc, back_conf = O.pack_heads(loc_heads, conf_heads, cg.shape[-1])
    assert bit_equal(back_loc, lg) and bit_equal(back_conf, cg)                 # the helper inverts the reference's packing
    want = O.postprocess(env["pri"], lg, cg, st, nt, mx, ag, nms_limit=True, with_index=True)
    gold = U.split_predictions(g)
    before = s.launch_count()
    got = s.predict_heads(ps, loc_heads, conf_heads, st, nt, mx, ag)
    assert s.launch_count() == before + 1                                       # predict_image_kernel alone: no pack pass, no list in HBM
    labels, scores, boxes, count, cand = s.predict_heads_padded(ps, loc_heads, conf_heads, st, nt, mx, ag, want_cand=True)
    pl, psc, pb, pc, pcand = s.predict_padded(ps, lg, cg, st, nt, mx, ag, want_cand=True)
    assert bit_equal(count, pc)
    for b, (a, w, r) in enumerate(zip(got, want, gold)):
        k = int(count[b])
        assert bit_equal(cand[b, :k], pcand[b, :k]) and bit_equal(scores[b, :k], psc[b, :k]) and bit_equal(boxes[b, :k], pb[b, :k])
        assert bit_equal(cand[b, :k].long(), w["cand"]), f"keep list, image {b}"
        assert bit_equal(a["labels"], w["labels"]) and bit_equal(a["scores"], w["scores"]) and bit_equal(a["boxes"], w["boxes"])
        assert bit_equal(a["labels"], r["labels"])
        assert close(a["scores"], r["scores"]) and close(a["boxes"], r["boxes"], atol=BOX_ATOL)


@pytest.mark.parametrize("channels_last", [False, True])
@pytest.mark.parametrize("name", TRAIN)
def test_losses_from_heads(env, name, channels_last):
    """multibox_loss_heads(heads) == multibox_loss(pack(heads)): the three sums bit for bit, and the losses within 1e-5 of
    the reference's (golden fixtures)."""
    s, dev, ps = env["ssdhot"], env["dev"], env["ps"]
    g = U.load(name)
    targets, loc_all, conf = U.train_inputs(g)
    thr, ratio = float(g["iou_thresh"]), float(g["ratio"])
    lg, cg = loc_all.to(dev), conf.to(dev)
    loc_heads, conf_heads = U.unpack_heads(lg, channels_last), U.unpack_heads(cg, channels_last)
    _, _, sums_packed = s.multibox_loss(ps, lg, cg, targets, thr, ratio, return_sums=True)
    l_loc, l_conf, sums = s.multibox_loss_heads(ps, loc_heads, conf_heads, targets, thr, ratio, return_sums=True)
    assert bit_equal(sums, sums_packed)
    assert abs(l_loc.item() - float(g["loc_loss"])) <= RTOL * abs(float(g["loc_loss"]))
    assert abs(l_conf.item() - float(g["conf_loss"])) <= RTOL * abs(float(g["conf_loss"]))


def test_heads_edge_cases(env):
    """Head-direct paths on the inputs that leave the fast paths: every candidate passes (threshold 0), the exact
    mining tail (budget >= negatives), images without boxes, and a class count the head kernels do not cover."""
    s, dev, ps = env["ssdhot"], env["dev"], env["ps"]
    cfg = s.synth.config(5, batch=2)
    lg, cg = cfg["loc_all"].to(dev), cfg["conf_infer"].to(dev)
    for cl in (False, True):
        lh, ch = U.unpack_heads(lg, cl), U.unpack_heads(cg, cl)
        for st, agn, metric in ((0.0, False, "diou"), (0.0, True, "ciou"), (0.3, False, "iou")):
            a = s.predict_heads_padded(ps, lh, ch, st, 0.45, 200, agn, metric, want_cand=True)
            b = s.predict_padded(ps, lg, cg, st, 0.45, 200, agn, metric, want_cand=True)
            n = b[3]
            assert bit_equal(a[3], n)
            for i in range(lg.shape[0]):
                k = int(n[i])
                assert all(bit_equal(x[i, :k], y[i, :k]) for x, y in zip((a[0], a[1], a[2], a[4]), (b[0], b[1], b[2], b[4])))
        ct = cfg["conf_train"].to(dev)
        cht = U.unpack_heads(ct, cl)
        targets = list(cfg["targets"])
        targets[1] = {"boxes": torch.zeros((0, 4)), "labels": torch.zeros((0,), dtype=torch.int64)}
        for ratio in (3.0, 1e6, 0.0):
            want = s.multibox_loss(ps, lg, ct, targets, 0.5, ratio, return_sums=True)[2]
            got = s.multibox_loss_heads(ps, lh, cht, targets, 0.5, ratio, return_sums=True)[2]
            assert bit_equal(got, want), (cl, ratio)
    # C = 21: packed first, then the packed kernels
    gen = torch.Generator().manual_seed(5)
    loc21 = torch.randn((2, 8732, 4), generator=gen).to(dev)
    conf21 = torch.randn((2, 8732, 21), generator=gen).to(dev)
    a = s.predict_heads_padded(ps, U.unpack_heads(loc21), U.unpack_heads(conf21), 0.2, 0.5, 50)
    b = s.predict_padded(ps, loc21, conf21, 0.2, 0.5, 50)
    assert all(bit_equal(x, y) for x, y in ((a[3], b[3]),)) and int(b[3].sum()) > 0
    for i in range(2):
        k = int(b[3][i])
        assert bit_equal(a[1][i, :k], b[1][i, :k]) and bit_equal(a[2][i, :k], b[2][i, :k])
    with pytest.raises(ValueError):
        s.predict_heads(ps, U.unpack_heads(loc21)[:5], U.unpack_heads(conf21), 0.2, 0.5, 50)


@pytest.mark.parametrize("channels_last", [False, True])
def test_heads_backward_matches_autograd(env, channels_last):
    """Gradients of multibox_loss_heads w.r.t. the twelve head tensors == torch autograd through the reference's own
    permute + cat tail (SFS:249-269) and loss (TR:104-108, :551-600), within 1e-5; delivered in the heads' memory format."""
    s, dev, ps = env["ssdhot"], env["dev"], env["ps"]
    g = U.load("train_cfg2_thr04.npz")
    targets, loc_all, conf = U.train_inputs(g)
    thr, ratio = float(g["iou_thresh"]), float(g["ratio"])
    tg = to_dev(targets, dev)
    mine = [h.requires_grad_(True) for h in U.unpack_heads(loc_all.to(dev), channels_last) + U.unpack_heads(conf.to(dev), channels_last)]
    ref = [h.detach().clone().requires_grad_(True) for h in mine]
    l_loc, l_conf = s.multibox_loss_heads(ps, mine[:6], mine[6:], targets, thr, ratio)
    (l_loc * 1.5 + l_conf * 0.7).backward()
    lb, cb = O.pack_heads(ref[:6], ref[6:], 6)
    pos_o, locpm_o, cls_o = O.batch_targets(env["pri"], env["pri_xyxy"], tg, 300, 300, thr)
    o_loc, n_img, total = O.loc_loss(lb, pos_o, locpm_o)
    o_conf = O.mined_ce_loss(cb, cls_o, pos_o, n_img, total, ratio)
    (o_loc * 1.5 + o_conf * 0.7).backward()
    assert abs(l_loc.item() - o_loc.item()) <= RTOL * abs(o_loc.item()) and abs(l_conf.item() - o_conf.item()) <= RTOL * abs(o_conf.item())
    for a, b in zip(mine, ref):
        assert a.grad is not None and a.grad.shape == a.shape
        assert a.grad.is_contiguous(memory_format=torch.channels_last if channels_last else torch.contiguous_format)
        assert close(a.grad, b.grad, atol=1e-7)
    # only the class branch needs a gradient
    only = [h.detach().clone().requires_grad_(i >= 6) for i, h in enumerate(mine)]
    l_loc, l_conf = s.multibox_loss_heads(ps, only[:6], only[6:], targets, thr, ratio)
    (l_loc + l_conf).backward()
    assert all(h.grad is None for h in only[:6]) and all(h.grad is not None for h in only[6:])


def test_predict_images_goes_through_the_heads(env):
    """predict(model, images) on a model with mySSD's module names runs the model's trunk and heads, then the head-direct
    kernels (no permute / cat / pack): same detections as predict on the model's own (loc_all, conf_all)."""
    s, dev = env["ssdhot"], env["dev"]
    torch.manual_seed(11)
    model = U.TinySSD().to(dev).eval()
    model.register_buffer("priors", s.default_boxes().to(dev))
    x = torch.randn(3, 3, 300, 300, device=dev)
    with torch.no_grad():
        loc_all, conf_all = model(x)
        conf_all = conf_all * 8.0                                   # (the stub's logits are nearly flat)
        for h in model.cls_head:
            h.weight.mul_(8.0); h.bias.mul_(8.0)
    want = s.predict(model, None, 0.15, 0.45, 50, pre_loc_all=loc_all, pre_conf_all=conf_all)
    before = s.launch_count()
    got = s.predict(model, x, 0.15, 0.45, 50)
    assert s.launch_count() == before + 1
    assert sum(int(w["labels"].numel()) for w in want) > 0
    for a, w in zip(got, want):
        assert bit_equal(a["labels"], w["labels"]) and close(a["scores"], w["scores"]) and close(a["boxes"], w["boxes"], atol=BOX_ATOL)
    # channels_last model: the heads come out NHWC and are read as rows
    model_cl = model.to(memory_format=torch.channels_last)
    x_cl = x.contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        loc_cl, conf_cl = model_cl(x_cl)
        lh, ch = s.forward_heads(model_cl, x_cl)
    assert all(h.is_contiguous(memory_format=torch.channels_last) for h in lh + ch)
    want_cl = s.predict(model_cl, None, 0.15, 0.45, 50, pre_loc_all=loc_cl, pre_conf_all=conf_cl)
    got_cl = s.predict(model_cl, x_cl, 0.15, 0.45, 50)
    for a, w in zip(got_cl, want_cl):
        assert bit_equal(a["labels"], w["labels"]) and close(a["scores"], w["scores"]) and close(a["boxes"], w["boxes"], atol=BOX_ATOL)


def test_peer_allreduce_protocol_virtual_ranks(env):
    """ssdhot_allreduce_sums_peer (csrc/peer.cu) with three ranks inside this process, one stream each, sharing their
    mailboxes directly (no IPC): 60 back-to-back steps, eager and replayed from CUDA graphs, always give the sum in rank
    order on every rank -- sequence tags, parity double-buffering and the in-place result all exercised."""
    s, dev = env["ssdhot"], env["dev"]
    from ssdhot.dist import PeerSums
    world, steps = 3, 60
    ranks = PeerSums.virtual(world, dev)
    streams = [torch.cuda.Stream(dev) for _ in range(world)]
    gen = torch.Generator().manual_seed(4)
    vals = torch.rand((steps, world, 3), generator=gen, dtype=torch.float64) * 1e3
    want = vals[:, 0, :].clone()
    for r in range(1, world):
        want = want + vals[:, r, :]                                   # rank order
    dvals = vals.to(dev)
    bufs = [torch.zeros((3,), dtype=torch.float64, device=dev) for _ in range(world)]
    outs = [torch.zeros((steps, 3), dtype=torch.float64, device=dev) for _ in range(world)]
    torch.cuda.synchronize(dev)
    before = s.launch_count()
    for i in range(steps):
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                bufs[r].copy_(dvals[i, r])
                ranks[r].allreduce(bufs[r])
                outs[r][i].copy_(bufs[r])
    torch.cuda.synchronize(dev)
    assert s.launch_count() == before + steps * world
    for r in range(world):
        assert not ranks[r].timed_out()
        assert bit_equal(outs[r], want), r
    # the same from CUDA graphs (one per rank), replayed in a skewed order
    graphs = []
    for r in range(world):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(streams[r]):
            streams[r].synchronize()
            with torch.cuda.graph(g, stream=streams[r]):
                bufs[r].mul_(2.0)
                ranks[r].allreduce(bufs[r])
        graphs.append(g)
    torch.cuda.synchronize(dev)
    for r in range(world):
        bufs[r].copy_(dvals[0, r])
    expect = [want[0].clone()]
    torch.cuda.synchronize(dev)
    for it in range(5):
        for r in (2, 0, 1):
            with torch.cuda.stream(streams[r]):
                graphs[r].replay()
        torch.cuda.synchronize(dev)
        # every rank doubled its buffer, then all-reduced: buf_r(new) = sum_r 2 * buf_r(old); all ranks hold the same value
        if it == 0:
            cur = 2.0 * dvals[0, 0].cpu()
            for r in range(1, world):
                cur = cur + 2.0 * dvals[0, r].cpu()
        else:
            prev = cur
            cur = 2.0 * prev
            for r in range(1, world):
                cur = cur + 2.0 * prev
        for r in range(world):
            assert bit_equal(bufs[r], cur), (it, r)
    assert not any(p.timed_out() for p in ranks)
    # lag = 1: call i delivers the reduced sums of call i - 1 (zeros first); the ranks may run a call apart
    lagged = PeerSums.virtual(world, dev, lag=1)
    outs = [torch.full((steps, 3), -1.0, dtype=torch.float64, device=dev) for _ in range(world)]
    for i in range(steps):
        for r in ((0, 1, 2) if i % 2 == 0 else (2, 1, 0)):
            with torch.cuda.stream(streams[r]):
                bufs[r].copy_(dvals[i, r])
                lagged[r].allreduce(bufs[r])
                outs[r][i].copy_(bufs[r])
    torch.cuda.synchronize(dev)
    shifted = torch.cat((torch.zeros((1, 3), dtype=torch.float64), want[:-1]), 0)
    for r in range(world):
        assert not lagged[r].timed_out()
        assert bit_equal(outs[r], shifted), r
    for p in lagged[1:]:
        p.close()
    lagged[0].close()
    # a lone rank (world 1) is the identity
    solo = PeerSums.virtual(1, dev)[0]
    x = torch.tensor([1.5, 2.5, 3.0], dtype=torch.float64, device=dev)
    solo.allreduce(x)
    assert bit_equal(x, torch.tensor([1.5, 2.5, 3.0], dtype=torch.float64))
    solo.close()
    for p in ranks[1:]:
        p.close()
    ranks[0].close()


def test_hot_path_step_from_heads_equals_packed(env):
    """HotPathStep.run_heads (both halves forked, eager and replayed from a CUDA graph) leaves the same sums and padded
    detections as HotPathStep.run on the packed tensors."""
    s, dev, ps = env["ssdhot"], env["dev"], env["ps"]
    from ssdhot.engine import HeadSet, HotPathStep
    cfg = s.synth.config(2, batch=8)
    loc, ct, ci = cfg["loc_all"].to(dev), cfg["conf_train"].to(dev), cfg["conf_infer"].to(dev)
    gt = s.pack_targets(cfg["targets"], dev)
    step = HotPathStep(ps, 8, 6, cfg["iou_thresh"], cfg["ratio"], cfg["score_thresh"], cfg["nms_thresh"], cfg["max_per_img"])
    step.run(loc, ct, ci, gt)
    torch.cuda.synchronize(dev)
    want = (step.sums.clone(), step.count.clone(), step.labels.clone(), step.scores.clone(), step.boxes.clone())
    for cl in (False, True):
        lh = s.synth.heads_from_packed(loc, cl)
        train, infer = HeadSet(lh, s.synth.heads_from_packed(ct, cl)), HeadSet(lh, s.synth.heads_from_packed(ci, cl))
        for use_graph in (False, True, True):
            for t in (step.sums, step.count, step.scores):
                t.zero_()
            step.run_heads(train, infer, gt, use_graph=use_graph)
            torch.cuda.synchronize(dev)
            assert bit_equal(step.sums, want[0]) and bit_equal(step.count, want[1])
            for b in range(8):
                k = int(want[1][b])
                assert bit_equal(step.labels[b, :k], want[2][b, :k]) and bit_equal(step.scores[b, :k], want[3][b, :k])
                assert bit_equal(step.boxes[b, :k], want[4][b, :k])


def test_nonfinite_head_outputs(env):
    """NaN / +-Inf rows in loc_all / conf_all (tests/golden/nonfinite.npz, produced by the reference; SURVEY.md section 7
    hard part 3): candidates, keep lists and labels bit-exact, NaN boxes stay NaN (torch.clamp propagates them) and suppress
    their class, NaN / Inf cross-entropies are mined first -- packed tensors and both head layouts, fast and generic paths."""
    s, dev, ps = env["ssdhot"], env["dev"], env["ps"]
    g = U.load("nonfinite.npz")
    loc, conf = U.nonfinite_predict_inputs(g)
    lg, cg = loc.to(dev), conf.to(dev)
    for agn, pre in ((False, "pred"), (True, "agn")):
        gold = U.split_rows(g[pre + "_counts"], g[pre + "_labels"], g[pre + "_scores"], g[pre + "_boxes"])
        want = O.postprocess(env["pri"], lg, cg, 0.01, 0.45, 200, agn, nms_limit=True)
        got = s.predict(ps, None, 0.01, 0.45, 200, agn, pre_loc_all=lg, pre_conf_all=cg)
        heads = [s.predict_heads(ps, U.unpack_heads(lg, cl), U.unpack_heads(cg, cl), 0.01, 0.45, 200, agn) for cl in (False, True)]
        n_nan = 0
        for b, (a, w, r) in enumerate(zip(got, want, gold)):
            assert bit_equal(a["labels"], r["labels"]), f"labels / keep list vs the reference, image {b}"
            assert close(a["scores"], r["scores"]) and close(a["boxes"], r["boxes"], atol=BOX_ATOL)
            assert bool((torch.isnan(a["boxes"].cpu()) == torch.isnan(r["boxes"])).all())
            assert bit_equal(a["labels"], w["labels"]) and bit_equal(a["scores"], w["scores"]) and bit_equal(a["boxes"], w["boxes"])
            for h in heads:
                assert bit_equal(h[b]["labels"], a["labels"]) and bit_equal(h[b]["scores"], a["scores"]) and bit_equal(h[b]["boxes"], a["boxes"])
            n_nan += int(torch.isnan(a["boxes"]).any())
        assert n_nan >= 1
    generic = s.PriorSet.default(dev, generic=True)
    for name, loc1, conf1, targets, cfg, (w_loc, w_conf) in U.nonfinite_loss_cases(g):
        l1, c1 = loc1.to(dev), conf1.to(dev)
        for tag, pset in (("fast", ps), ("generic", generic)):
            l_loc, l_conf = s.multibox_loss(pset, l1, c1, targets, cfg["iou_thresh"], cfg["ratio"])
            assert U.same_float(l_loc.item(), w_loc) and U.same_float(l_conf.item(), w_conf), (name, tag, l_loc.item(), w_loc, l_conf.item(), w_conf)
        for cl in (False, True):
            h_loc, h_conf = s.multibox_loss_heads(ps, U.unpack_heads(l1, cl), U.unpack_heads(c1, cl), targets, cfg["iou_thresh"], cfg["ratio"])
            assert U.same_float(h_loc.item(), w_loc) and U.same_float(h_conf.item(), w_conf), (name, cl)
        # the drop-in pieces on the reference's own targets
        pos, loc_pm, cls = s.build_targets(ps, targets, 300, 300, cfg["iou_thresh"], "cuda")
        total = pos.sum().clamp_min(1).float()
        ce = s.CELoss_w_neg_mining(c1, cls, pos, pos.sum(1), total, cfg["ratio"])
        assert U.same_float(ce.item(), w_conf), (name, ce.item(), w_conf)


def test_full_size_stress_config(env):
    """BASELINE cfg 5 at FULL size (B = 1024, 64 boxes per image, score threshold 0.0 so that all 8732 x 5 pairs of every
    image enter NMS): the whole batch runs through the kernels; 8 randomly chosen images are compared with the device-matched
    oracle (matches, class targets, offsets, per-image loss sums, keep lists, scores, boxes), the rest through the
    size-independent properties (sharded sums == whole-batch sums, counts, score order)."""
    s, dev, ps = env["ssdhot"], env["dev"], env["ps"]
    from ssdhot import synth
    B = 1024
    cfg = synth.config(5)
    assert cfg["batch"] == B and all(t["boxes"].shape[0] == 64 for t in cfg["targets"])
    lg, ct = cfg["loc_all"].to(dev), cfg["conf_train"].to(dev)
    packed = s.pack_targets(cfg["targets"], dev)
    pick = sorted(torch.randperm(B, generator=torch.Generator().manual_seed(55))[:8].tolist())
    # ---- match + encode + mined loss ----
    r = s.match_encode_batch(ps, packed, 0.5, (300, 300), want_loc="positives")
    l_loc, l_conf, sums = s.multibox_loss(ps, lg, ct, packed, 0.5, 3.0, return_sums=True)
    assert int(r["n_pos"].sum()) == int(sums[2].item())
    part = torch.zeros((3,), dtype=torch.float64, device=dev)
    for lo in range(0, B, 256):                                        # four shards: the sums add up (cfg 4's property)
        part += s.multibox_loss(ps, lg[lo:lo + 256], ct[lo:lo + 256], cfg["targets"][lo:lo + 256], 0.5, 3.0, return_sums=True)[2]
    assert close(part[:2], sums[:2], rtol=1e-12) and part[2].item() == sums[2].item()
    for b in pick:
        tg = to_dev(cfg["targets"][b:b + 1], dev)
        pos_o, locpm_o, cls_o = O.batch_targets(env["pri"], env["pri_xyxy"], tg, 300, 300, 0.5)
        assert bit_equal(r["pos_mask"][b:b + 1], pos_o) and bit_equal(r["cls_t"][b:b + 1], cls_o), f"image {b}"
        assert bit_equal(r["loc_t"][b][r["pos_mask"][b]], locpm_o)
        o_loc, o_conf = O.train_half(env["pri"], env["pri_xyxy"], lg[b:b + 1], ct[b:b + 1], tg, 0.5, 3.0)
        i_loc, i_conf = s.multibox_loss(ps, lg[b:b + 1], ct[b:b + 1], cfg["targets"][b:b + 1], 0.5, 3.0)
        assert close(i_loc, o_loc) and close(i_conf, o_conf), f"image {b}"
    # ---- decode + threshold 0.0 + DIoU-NMS ----
    ci = cfg["conf_infer"].to(dev)
    labels, scores, boxes, count, cand = s.predict_padded(ps, lg, ci, 0.0, 0.45, 200, want_cand=True)
    assert int(count.min()) == 200 and int(count.max()) == 200         # 43,660 candidates per image: the cap is always reached
    assert bool((scores[:, 1:] <= scores[:, :-1]).all())               # score-descending
    want = O.postprocess(env["pri"], lg[pick], ci[pick], 0.0, 0.45, 200, False, nms_limit=True, with_index=True)
    for k, b in enumerate(pick):
        assert bit_equal(cand[b, :200].long(), want[k]["cand"]), f"keep list, image {b}"
        assert bit_equal(labels[b], want[k]["labels"]) and bit_equal(scores[b], want[k]["scores"]) and bit_equal(boxes[b], want[k]["boxes"])


def test_exchange_kernel_folds_the_partials(env):
    """HotPathStep(group=PeerSums): the loss forward leaves its per-image partials (sums == NULL) and the exchange kernel
    folds them itself (ssdhot_allreduce_partials_peer) -- same sums as loss kernel + finalize_sums_kernel, eager and replayed
    from a CUDA graph, packed tensors and head layouts."""
    s, dev, ps = env["ssdhot"], env["dev"], env["ps"]
    from ssdhot import synth
    from ssdhot.dist import PeerSums
    from ssdhot.engine import HeadSet, HotPathStep
    cfg = synth.config(3, batch=37)
    loc, conf = cfg["loc_all"].to(dev), cfg["conf_infer"].to(dev)
    gt = s.pack_targets(cfg["targets"], dev)
    plain = HotPathStep(ps, 37, 6, 0.5, 3.0, 0.01, 0.45, 200)
    plain.run(loc, conf, conf, gt)
    torch.cuda.synchronize()
    want = plain.sums.clone()
    solo = PeerSums(dev)
    try:
        fused = HotPathStep(ps, 37, 6, 0.5, 3.0, 0.01, 0.45, 200, group=solo)
        for use_graph in (False, True, True):
            fused.sums.zero_()
            fused.run(loc, conf, conf, gt, use_graph=use_graph)
            torch.cuda.synchronize()
            assert torch.allclose(fused.sums, want, rtol=1e-13, atol=0.0) and fused.sums[2].item() == want[2].item(), (fused.sums, want)
        hs = HeadSet(U.unpack_heads(loc), U.unpack_heads(conf))
        fused.sums.zero_()
        fused.run_heads(hs, hs, gt)
        torch.cuda.synchronize()
        assert torch.allclose(fused.sums, want, rtol=1e-13, atol=0.0)
        assert torch.equal(fused.count, plain.count) and not solo.timed_out()
    finally:
        solo.close()


def test_key_handoff_between_the_halves(env):
    """One conf_all for both halves (SSD_test_step): the loss kernel's logit stream leaves predict's row keys in the share buffer
    and predict_image_kernel picks them up instead of streaming the logits again.  Sums and detections are bit-identical to
    the two independent launches: packed tensors and both head layouts, eager and replayed from a CUDA graph, a score
    threshold of 0 (every row a candidate), non-finite logits, and a loss launch that does not deliver (generic path): the
    predict CTAs then stream themselves."""
    s, dev, ps = env["ssdhot"], env["dev"], env["ps"]
    from ssdhot import synth
    from ssdhot.engine import HeadSet, HotPathStep
    B = 37
    cfg = synth.config(3, batch=B)
    loc, conf = cfg["loc_all"].to(dev), cfg["conf_infer"].to(dev).clone()
    conf[3, 100:140, 2] = float("nan")
    conf[5, 17, 0] = float("inf")
    conf[7, 4000:4100] = float("-inf")
    gt = s.pack_targets(cfg["targets"], dev)

    def outputs(step):
        torch.cuda.synchronize()
        return [t.clone() for t in (step.sums, step.n_pos, step.count, step.labels, step.scores, step.boxes)]

    def same(a, b):
        for x, y in zip(a, b):
            assert x.shape == y.shape and bool(((x == y) | ((x != x) & (y != y))).all())

    def picked_up(step):
        return int((step.share[:4 * B].view(torch.int32) == 3).sum())

    for thr in (0.01, 0.0):
        apart = HotPathStep(ps, B, 6, 0.5, 3.0, thr, 0.45, 200)
        apart.share_keys = False
        apart.run(loc, conf, conf, gt)
        want = outputs(apart)
        together = HotPathStep(ps, B, 6, 0.5, 3.0, thr, 0.45, 200)
        together.share.fill_(0xAB)                                  # (whatever the buffer held: the reset clears the flags)
        for use_graph in (False, True, True):
            for t in (together.count, together.labels, together.scores, together.boxes, together.sums):
                t.zero_()
            together.run(loc, conf, conf, gt, use_graph=use_graph)
            same(outputs(together), want)
            assert picked_up(together) >= B - 8, (thr, use_graph, picked_up(together))
        for cl in (False, True):
            hs = HeadSet(U.unpack_heads(loc, cl), U.unpack_heads(conf, cl))
            together.count.zero_()
            together.run_heads(hs, hs, gt)
            same(outputs(together), want)
            assert picked_up(together) >= B - 8, (thr, cl, picked_up(together))
    # more images than the GPU holds CTAs at once (several waves of both kernels, the predict CTAs of late images start before
    # their loss CTAs): every image is still served -- by the hand-off or by its own stream -- with the same results
    Bw = 700
    cw = synth.config(3, batch=Bw, seed_offset=3)
    locw, confw, gtw = cw["loc_all"].to(dev), cw["conf_infer"].to(dev), s.pack_targets(cw["targets"], dev)
    apart = HotPathStep(ps, Bw, 6, 0.5, 3.0, 0.01, 0.45, 200)
    apart.share_keys = False
    apart.run(locw, confw, confw, gtw)
    want = outputs(apart)
    together = HotPathStep(ps, Bw, 6, 0.5, 3.0, 0.01, 0.45, 200)
    for use_graph in (False, True, True):
        together.count.zero_()
        together.run(locw, confw, confw, gtw, use_graph=use_graph)
        same(outputs(together), want)
        assert int((together.share[:4 * Bw].view(torch.int32) == 3).sum()) >= Bw // 2
    # a loss launch that leaves no keys (generic priors -> match_kernel + loss_image_kernel): predict streams itself
    generic = s.PriorSet.default(dev, generic=True)
    apart = HotPathStep(generic, B, 6, 0.5, 3.0, 0.01, 0.45, 200)
    apart.share_keys = False
    apart.run(loc, conf, conf, gt)
    want = outputs(apart)
    together = HotPathStep(generic, B, 6, 0.5, 3.0, 0.01, 0.45, 200)
    together.run(loc, conf, conf, gt)
    same(outputs(together), want)
    assert picked_up(together) == 0
    # the public eval step goes the same way
    l_loc, l_conf, labels, scores, boxes, count = s.eval_step(ps, loc, conf, cfg["targets"], 0.5, 3.0, 0.01, 0.45, 200)
    a_loc, a_conf = s.multibox_loss(ps, loc, conf, cfg["targets"], 0.5, 3.0)
    al, asc, ab, ac = s.predict_padded(ps, loc, conf, 0.01, 0.45, 200)
    torch.cuda.synchronize()
    same([l_loc, l_conf, count], [a_loc, a_conf, ac])
    for b in range(B):
        k = int(ac[b])
        same([labels[b, :k], scores[b, :k], boxes[b, :k]], [al[b, :k], asc[b, :k], ab[b, :k]])


def test_two_steps_in_flight_equal_sequential(env):
    """StepPipeline: two HotPathStep objects on two streams take the batches of an evaluation pass round-robin (the loss
    kernel's stream of batch i+1 beside the NMS tail of batch i).  Every batch's sums and detections equal those of the
    same batch run alone."""
    s, dev, ps = env["ssdhot"], env["dev"], env["ps"]
    from ssdhot import synth
    from ssdhot.engine import HotPathStep, StepPipeline
    B = 24
    batches = []
    for k in range(5):
        cfg = synth.config(3, batch=B, seed_offset=k)
        batches.append((cfg["loc_all"].to(dev), cfg["conf_infer"].to(dev), s.pack_targets(cfg["targets"], dev)))
    mk = lambda k=0: HotPathStep(ps, B, 6, 0.5, 3.0, 0.01, 0.45, 200)
    alone = mk()
    want = []
    for loc, conf, gt in batches:
        alone.run(loc, conf, conf, gt)
        torch.cuda.synchronize()
        want.append([t.clone() for t in (alone.sums, alone.n_pos, alone.count, alone.labels, alone.scores, alone.boxes)])
    pipe = StepPipeline(mk, depth=2)
    for use_graph in (False, True, True):
        got, pending = [], []
        for i, (loc, conf, gt) in enumerate(batches):
            if len(pending) == 2:                                   # the step object about to be reused: collect its batch first
                st = pending.pop(0)
                pipe.wait(st)
                got.append([t.clone() for t in (st.sums, st.n_pos, st.count, st.labels, st.scores, st.boxes)])
            pending.append(pipe.submit(loc, conf, conf, gt, use_graph=use_graph))
        for st in pending:
            pipe.wait(st)
            got.append([t.clone() for t in (st.sums, st.n_pos, st.count, st.labels, st.scores, st.boxes)])
        torch.cuda.synchronize()
        assert len(got) == len(want)
        for g, w in zip(got, want):
            k = w[2].long()
            assert torch.equal(g[0], w[0]) and torch.equal(g[1], w[1]) and torch.equal(g[2], w[2])
            for b in range(B):                                      # (rows beyond count hold what an earlier batch left there)
                n = int(k[b])
                assert torch.equal(g[3][b, :n], w[3][b, :n]) and torch.equal(g[4][b, :n], w[4][b, :n]) and torch.equal(g[5][b, :n], w[5][b, :n])
