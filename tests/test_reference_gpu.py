"""-m gpu: the UNMODIFIED reference (oracle/_ref, staged by oracle/stage_reference.py) as the device-matched checker.

(1) The reference's own functions run with device='cuda' on the seeded inputs of BASELINE cfg 1 / 2 / 3(slice) and ssdhot's
    drop-ins must reproduce them: masks, class targets, keep lists, labels bit-exact; offsets, scores, boxes bit-exact
    (same device = same libdevice / IEEE arithmetic); loss sums within 1e-5 relative (fp32 summation order differs).
(2) INTEGRATION.md section 2 executed: the reference's own SSD_train_step / SSD_test_step on a real `mySSD` on CUDA,
    unpatched, then after `ssdhot.patch(model, trainer_module, steps=False)` (the reference's loops calling the drop-ins one
    by one) and after `ssdhot.patch(model, trainer_module)` (the fused step functions).
"""
from __future__ import annotations

import copy

import pytest
import torch

import _util as U  # noqa: F401  (path set-up)
from oracle import refload as R

pytestmark = pytest.mark.gpu
REL = 1e-5          # north_star: losses and decoded boxes within 1e-5 relative in fp32


def close(a, b, rel=REL):
    a, b = float(a), float(b)
    return abs(a - b) <= rel * max(abs(a), abs(b), 1e-30)


@pytest.fixture(scope="module")
def env():
    assert torch.cuda.is_available()
    if not R.available():                    # (staged by __graft_entry__.build() in the build container; it ships with the snapshot)
        pytest.skip("oracle/_ref holds no staged reference: run __graft_entry__.build() where /root/reference is mounted")
    import ssdhot
    from ssdhot import synth
    dev = torch.device("cuda:0")
    sfs, tr = R.load()
    mdl = R.model(dev)
    return dict(ssdhot=ssdhot, synth=synth, dev=dev, sfs=sfs, tr=tr, mdl=mdl, ps=ssdhot.PriorSet.of(mdl))


@pytest.mark.parametrize("cfg_idx,batch", [(1, 1), (2, 32), (3, 4)])
def test_train_half_against_reference_on_cuda(env, cfg_idx, batch):
    ssdhot, dev, mdl, tr = env["ssdhot"], env["dev"], env["mdl"], env["tr"]
    cfg = env["synth"].config(cfg_idx, batch=batch)
    loc, conf = cfg["loc_all"].to(dev), cfg["conf_train"].to(dev)
    targets = R.to_device(cfg["targets"], dev)
    with torch.no_grad():
        r_loc, r_conf, (r_pos, r_loc_pm, r_cls) = R.train_half(loc, conf, targets, cfg["iou_thresh"], cfg["ratio"], mdl)
    # build_targets drop-in (SSD_trainer.py:491-547): same call, same outputs
    pos, loc_pm, cls = ssdhot.build_targets(model=mdl, targets=targets, H=300, W=300, iou_thresh=cfg["iou_thresh"], device="cuda")
    assert pos.dtype == r_pos.dtype and cls.dtype == r_cls.dtype and loc_pm.dtype == r_loc_pm.dtype
    assert torch.equal(pos, r_pos) and torch.equal(cls, r_cls)
    assert loc_pm.shape == r_loc_pm.shape and torch.equal(loc_pm, r_loc_pm)
    # CELoss_w_neg_mining drop-in on the reference's own targets (SSD_trainer.py:551-600)
    n_img = r_pos.sum(dim=1)
    total = n_img.sum().clamp_min(1).float()
    ce = ssdhot.CELoss_w_neg_mining(conf_all=conf, cls_t=r_cls, pos_mask=r_pos, num_pos_per_img=n_img, total_pos=total,
                                    neg_pos_ratio=cfg["ratio"])
    assert close(ce, r_conf), (float(ce), float(r_conf))
    sl1 = ssdhot.smooth_l1_positive_loss(loc, r_pos, r_loc_pm, total)
    assert close(sl1, r_loc), (float(sl1), float(r_loc))
    # the fused step (one launch) and the same from the head outputs
    l_loc, l_conf = ssdhot.multibox_loss(mdl, loc, conf, targets, cfg["iou_thresh"], cfg["ratio"])
    assert close(l_loc, r_loc) and close(l_conf, r_conf), (float(l_loc), float(r_loc), float(l_conf), float(r_conf))
    # collate_detection's packed batch gives the same bits as the list of dicts
    _, packed = ssdhot.collate_detection([(torch.zeros((1,)), t) for t in cfg["targets"]])
    p_loc, p_conf = ssdhot.multibox_loss(mdl, loc, conf, packed, cfg["iou_thresh"], cfg["ratio"])
    assert p_loc.item() == l_loc.item() and p_conf.item() == l_conf.item()
    ref_packed = ssdhot.pack_targets(cfg["targets"], dev)
    moved = packed.to(dev)
    assert torch.equal(moved.boxes[:moved.total], ref_packed.boxes[:ref_packed.total]) and torch.equal(moved.offsets, ref_packed.offsets)
    assert torch.equal(moved.labels[:moved.total], ref_packed.labels[:ref_packed.total])


@pytest.mark.parametrize("cfg_idx,batch", [(1, 1), (3, 3)])
def test_predict_against_reference_on_cuda(env, cfg_idx, batch):
    ssdhot, dev, mdl = env["ssdhot"], env["dev"], env["mdl"]
    cfg = env["synth"].config(cfg_idx, batch=batch, dedup=True)
    loc, conf = cfg["loc_all"].to(dev), cfg["conf_infer"].to(dev)
    want = R.predict_half(loc, conf, cfg["score_thresh"], cfg["nms_thresh"], cfg["max_per_img"], False, mdl)
    got = ssdhot.predict(mdl, None, cfg["score_thresh"], cfg["nms_thresh"], cfg["max_per_img"], False, pre_loc_all=loc, pre_conf_all=conf)
    assert len(got) == len(want) == batch
    for g, w in zip(got, want):
        assert torch.equal(g["labels"], w["labels"]), "keep list / labels differ from the reference"
        assert torch.equal(g["scores"], w["scores"]) and torch.equal(g["boxes"], w["boxes"])
        assert g["labels"].dtype == w["labels"].dtype and g["boxes"].dtype == w["boxes"].dtype
    # static methods, called the way predict calls them (keywords; SSD_from_scratch.py:419, :437)
    sfs = env["sfs"]
    d_ref = sfs.mySSD.decode_ssd(loc=loc[0, :64], priors=mdl.priors[:64], variances=(0.1, 0.2))
    d_got = ssdhot.decode_ssd(loc=loc[0, :64], priors=mdl.priors[:64], variances=(0.1, 0.2))
    assert torch.equal(d_ref, d_got)
    boxes, scores = want[0]["boxes"], want[0]["scores"]
    if boxes.shape[0] > 1:
        assert torch.equal(sfs.mySSD.iou_nms(boxes, scores, iou_threshold=0.3), ssdhot.iou_nms(boxes, scores, iou_threshold=0.3))


def _loader(synth, n_batches, batch, seed):
    gen = torch.Generator().manual_seed(seed)
    out = []
    for i in range(n_batches):
        targets = synth.make_targets(batch, 0 if i == 0 else 1, 6, gen)
        out.append((torch.rand((batch, 3, 300, 300), generator=gen), targets))
    return out


def _fresh(loader):
    return [(im.clone(), [{k: v.clone() for k, v in t.items()} for t in tg]) for im, tg in loader]


def _calibrated_thresh(mdl, loader, dev, per_image=300):
    """A score threshold that leaves ~per_image candidates per image for this (random-weight) model."""
    with torch.inference_mode():
        mdl.eval()
        _, conf = mdl(loader[0][0].to(dev))
        s = conf.softmax(-1)[..., 1:].reshape(conf.shape[0], -1)
        k = min(per_image, s.shape[1] - 1)
        return float(s.topk(k + 1, dim=1).values[:, -1].max().clamp(0.0, 0.999))


def test_reference_step_functions_run_through_patch(env):
    """SSD_trainer.SSD_test_step / SSD_train_step (TR:155-293, :21-150), unmodified, on a real mySSD on CUDA: unpatched vs
    patched (drop-ins called by the reference's own loops) vs the fused step functions installed by patch()."""
    ssdhot, synth, dev, sfs, tr = env["ssdhot"], env["synth"], env["dev"], env["sfs"], env["tr"]
    torch.manual_seed(11)
    base = sfs.mySSD(class_to_idx_dict=dict(R.CLASSES)).to(dev)
    with torch.no_grad():                      # spread the (random-weight) logits so that the top scores are distinct floats
        for h in base.cls_head:
            h.weight.mul_(40.0)
            h.bias.normal_()
    loader = _loader(synth, 2, 2, seed=3)
    train_loader = _loader(synth, 1, 4, seed=4)      # ONE optimiser step: the loss is that of identical weights, the update is compared below
    thr = _calibrated_thresh(base, loader, dev)
    state = copy.deepcopy(base.state_dict())
    originals = {k: getattr(tr, k) for k in ("build_targets", "CELoss_w_neg_mining", "collate_detection", "SSD_train_step", "SSD_test_step")}

    def new_model():
        m = sfs.mySSD(class_to_idx_dict=dict(R.CLASSES)).to(dev)
        m.load_state_dict(state)
        return m

    def run_eval(m):
        captured = []
        orig_cls = tr.MeanAveragePrecision

        class Capture(R._StubMAP):
            def update(self, preds, target):
                captured.append([{k: v.detach().clone() for k, v in p.items()} for p in preds])
        tr.MeanAveragePrecision = Capture
        try:
            out = tr.SSD_test_step(model=m, dataloader=_fresh(loader), iou_thresh=0.5, neg_pos_ratio=3.0, score_thresh=thr,
                                   nms_thresh=0.45, max_detections_per_img=50, device="cuda")
        finally:
            tr.MeanAveragePrecision = orig_cls
        return out, captured

    def param_distance(a, b):
        """Relative L2 distance of two parameter UPDATES (state after training minus the common initial state), all tensors
        together: a tensor whose gradient is pure rounding noise (a conv bias in front of a batch-norm) cannot dominate."""
        num = den = 0.0
        for k, v in b.items():
            if "num_batches_tracked" not in k:
                num += float(((a[k] - v).double() ** 2).sum())
                den += float(((v - state[k].to(v.device)).double() ** 2).sum())
        return (num / max(den, 1e-300)) ** 0.5

    def run_train(m):
        opt = torch.optim.SGD(m.parameters(), lr=1e-3, momentum=0.0)
        out = tr.SSD_train_step(model=m, dataloader=_fresh(train_loader), optimizer=opt, iou_thresh=0.5, neg_pos_ratio=3.0, device="cuda")
        return out, {k: v.detach().clone() for k, v in m.state_dict().items() if v.dtype.is_floating_point}

    try:
        # --- unpatched reference ----------------------------------------------------------------------------------
        ref_eval, ref_preds = run_eval(new_model())
        ref_train, ref_params = run_train(new_model())
        _, again = run_train(new_model())           # the reference against itself: the run-to-run noise of one SGD step
        noise = param_distance(again, ref_params)   # through a VGG-16 with batch-norm in train mode (cuDNN's backward is not bit-reproducible)
        results = {}
        for mode, steps in (("dropins", False), ("fused", True)):
            for k, v in originals.items():
                setattr(tr, k, v)
            m = new_model()
            ssdhot.patch(m, tr, steps=steps)
            assert tr.build_targets is ssdhot.build_targets and (tr.SSD_train_step is originals["SSD_train_step"]) == (not steps)
            before = ssdhot.launch_count()
            ev, preds = run_eval(m)
            assert ssdhot.launch_count() > before, "the patched step launched no ssdhot kernel"
            m2 = new_model()
            ssdhot.patch(m2, tr, steps=steps)
            trn, params = run_train(m2)
            results[mode] = (ev, preds, trn, params)
    finally:
        for k, v in originals.items():
            setattr(tr, k, v)

    for mode, (ev, preds, trn, params) in results.items():
        for key in ("testing loss", "localization loss", "classification loss"):
            assert close(ev[key], ref_eval[key], 2e-5), (mode, key, ev[key], ref_eval[key])
        assert set(ev.keys()) == set(ref_eval.keys()) and set(ev["timing"].keys()) == set(ref_eval["timing"].keys())
        assert len(preds) == len(ref_preds)
        for pb, rb in zip(preds, ref_preds):
            for p, r in zip(pb, rb):
                assert torch.equal(p["labels"], r["labels"]), mode
                assert torch.equal(p["scores"], r["scores"]) and torch.equal(p["boxes"], r["boxes"]), mode
        for key in ("training loss", "localization loss", "classification loss"):
            assert close(trn[key], ref_train[key], 1e-4), (mode, key, trn[key], ref_train[key])     # (batch 2 sees the weights of step 1)
        assert set(trn.keys()) == set(ref_train.keys()) and set(trn["timing"].keys()) == set(ref_train["timing"].keys())
        worst = param_distance(params, ref_params)
        assert worst <= max(1e-3, 20.0 * noise), (mode, worst, noise)


def test_patched_model_methods_keep_reference_call_forms(env):
    """model.encode_ssd / decode_ssd / iou_nms / predict after ssdhot.patch(model): the reference's own call forms."""
    ssdhot, dev, sfs = env["ssdhot"], env["dev"], env["sfs"]
    cfg = env["synth"].config(1)
    ref_m = env["mdl"]
    m = copy.copy(ref_m)                       # shallow: shares buffers, owns its attribute dict
    m.__dict__ = dict(ref_m.__dict__)
    ssdhot.patch(m)
    g = (cfg["targets"][0]["boxes"] / torch.tensor([300., 300., 300., 300.])).to(dev)
    lab = cfg["targets"][0]["labels"].to(dev)
    want = ref_m.encode_ssd(g, lab, iou_thresh=0.5)
    got = m.encode_ssd(g, lab, iou_thresh=0.5)
    for a, b in zip(got, want):
        assert a.dtype == b.dtype and torch.equal(a, b)
    with pytest.raises(ValueError):
        m.encode_ssd(g, lab, background_class=1)
    loc, conf = cfg["loc_all"].to(dev), cfg["conf_infer"].to(dev)
    a = m.predict(images=None, score_thresh=0.01, nms_thresh=0.45, max_per_img=200, class_agnostic=False, pre_loc_all=loc, pre_conf_all=conf)
    b = ref_m.predict(images=None, score_thresh=0.01, nms_thresh=0.45, max_per_img=200, class_agnostic=False, pre_loc_all=loc, pre_conf_all=conf)
    assert torch.equal(a[0]["labels"], b[0]["labels"]) and torch.equal(a[0]["boxes"], b[0]["boxes"]) and torch.equal(a[0]["scores"], b[0]["scores"])
    assert torch.equal(m.decode_ssd(loc=loc[0, :8], priors=m.priors[:8], variances=(0.1, 0.2)),
                       sfs.mySSD.decode_ssd(loc=loc[0, :8], priors=m.priors[:8], variances=(0.1, 0.2)))


def test_eval_step_and_packed_loader(env):
    """ssdhot.eval_step (both halves on two streams, one conf tensor) equals the two calls it stands for; the fused step
    functions accept the PackedTargets batches of ssdhot.collate_detection and give the List[Dict] loader's results."""
    ssdhot, synth, dev, mdl = env["ssdhot"], env["synth"], env["dev"], env["mdl"]
    cfg = synth.config(3, batch=6)
    loc, conf = cfg["loc_all"].to(dev), cfg["conf_infer"].to(dev)
    _, packed = ssdhot.collate_detection([(torch.zeros((1,)), t) for t in cfg["targets"]])
    l_loc, l_conf, labels, scores, boxes, count = ssdhot.eval_step(mdl, loc, conf, packed, 0.5, 3.0, 0.01, 0.45, 200)
    w_loc, w_conf = ssdhot.multibox_loss(mdl, loc, conf, cfg["targets"], 0.5, 3.0)
    w_labels, w_scores, w_boxes, w_count = ssdhot.predict_padded(mdl, loc, conf, 0.01, 0.45, 200)
    torch.cuda.synchronize()
    assert l_loc.item() == w_loc.item() and l_conf.item() == w_conf.item()
    assert torch.equal(count, w_count)
    for b in range(6):
        k = int(count[b])
        assert torch.equal(labels[b, :k], w_labels[b, :k]) and torch.equal(scores[b, :k], w_scores[b, :k]) and torch.equal(boxes[b, :k], w_boxes[b, :k])
    # the fused step functions on a TinySSD-sized stand-in: packed loader == list loader
    import _util as U
    torch.manual_seed(5)
    gen = torch.Generator().manual_seed(9)
    batches = [(torch.rand((3, 3, 300, 300), generator=gen), synth.make_targets(3, 1, 5, gen)) for _ in range(2)]
    results = []
    for use_packed in (False, True):
        torch.manual_seed(21)
        m = U.TinySSD().to(dev)
        m.register_buffer("priors", ssdhot.default_boxes().to(dev), persistent=False)
        loader = [(im.clone(), ssdhot.collate_detection([(torch.zeros((1,)), t) for t in tg])[1] if use_packed
                   else [{k: v.clone() for k, v in t.items()} for t in tg]) for im, tg in batches]
        opt = torch.optim.SGD(m.parameters(), lr=1e-3)
        out = ssdhot.SSD_train_step(m, loader, opt, iou_thresh=0.5, neg_pos_ratio=3.0, device="cuda")
        ev = ssdhot.trainer.make_test_step(R._StubMAP)(m, loader, score_thresh=0.05, nms_thresh=0.45, max_detections_per_img=20, device="cuda")
        results.append((out, ev))
    for key in ("training loss", "localization loss", "classification loss"):      # (cuDNN's weight gradients are not bit-reproducible)
        assert close(results[0][0][key], results[1][0][key]), key
    for key in ("testing loss", "localization loss", "classification loss"):
        assert close(results[0][1][key], results[1][1][key]), key
