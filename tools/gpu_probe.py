"""Diagnostic run on the GPU box: compares ssdhot with the oracle on CUDA (device-matched) and on
CPU, and with the golden fixtures, printing mismatch statistics instead of stopping at the first."""
import os, sys, time, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "automotive-ssd-object-detection_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch
import ssdhot
from ssdhot import synth
from oracle import ssd_oracle as O
import _util as U

dev = torch.device("cuda")
print(torch.cuda.get_device_name(0), torch.version.cuda)

def section(name):
    print("\n==== " + name, flush=True)

def cmp_f(name, a, b):
    a = a.detach().float().cpu(); b = b.detach().float().cpu()
    if a.shape != b.shape:
        print(f"  {name}: SHAPE {tuple(a.shape)} vs {tuple(b.shape)}"); return
    if a.numel() == 0:
        print(f"  {name}: empty ok"); return
    both_nan = torch.isnan(a) & torch.isnan(b)
    eq = (a == b) | both_nan
    d = (a - b).abs()[~both_nan]
    rel = (d / b[~both_nan].abs().clamp_min(1e-6)).max().item() if d.numel() else 0.0
    print(f"  {name}: bitwise-equal {eq.float().mean().item()*100:.4f}%  max|d|={d.max().item() if d.numel() else 0:.3e} maxrel={rel:.3e}")

def cmp_i(name, a, b):
    a = a.detach().cpu(); b = b.detach().cpu()
    if a.shape != b.shape:
        print(f"  {name}: SHAPE {tuple(a.shape)} vs {tuple(b.shape)}"); return
    ne = (a != b).sum().item()
    print(f"  {name}: mismatches {ne} / {a.numel()}")

def run(fn):
    try:
        fn()
    except Exception:
        traceback.print_exc()

pri_c, pri_xyxy_c = O.prior_tables()
pri_g, pri_xyxy_g = pri_c.to(dev), pri_xyxy_c.to(dev)
PS = ssdhot.PriorSet.default(dev)

def t_priors():
    section("priors")
    cmp_f("priors", PS.priors, pri_c)
    cmp_f("priors_xyxy (kernel) vs oracle", PS.priors_xyxy, pri_xyxy_c)
    w = pri_xyxy_g[:, 2] - pri_xyxy_g[:, 0]; h = pri_xyxy_g[:, 3] - pri_xyxy_g[:, 1]
    cmp_f("aux.area", PS.aux[:, 0], w * h)
    cmp_f("aux.atan vs torch-cuda", PS.aux[:, 3], torch.atan(w / h))
    cmp_f("aux.atan vs torch-cpu", PS.aux[:, 3], torch.atan(w.cpu() / h.cpu()))
run(t_priors)

def t_inputs():
    section("input generator portability")
    for name in ["train_cfg1.npz", "train_cfg2.npz", "predict_cfg3_b4.npz"]:
        g = U.load(name)
        try:
            if name.startswith("train"): U.train_inputs(g)
            else: U.predict_inputs(g)
            print("  ", name, "sha OK")
        except AssertionError as e:
            print("  ", name, "SHA MISMATCH", e)
run(t_inputs)

def t_train(name):
    section("train " + name)
    g = U.load(name)
    targets, loc_all, conf = U.train_inputs(g)
    thr, ratio = float(g["iou_thresh"]), float(g["ratio"])
    tg = [{k: v.to(dev) for k, v in t.items()} for t in targets]
    # oracle on cuda, dense
    pos_o, locpm_o, cls_o, locd_o = O.batch_targets(pri_g, pri_xyxy_g, tg, 300, 300, thr, dense=True)
    packed = ssdhot.pack_targets(targets, dev)
    r = ssdhot.match_encode_batch(PS, packed, thr, (300, 300), want_loc="all", want_matched_idx=True, want_matched_box=True)
    torch.cuda.synchronize()
    cmp_i("pos_mask vs oracle-cuda", r["pos_mask"], pos_o)
    cmp_i("cls_t vs oracle-cuda", r["cls_t"], cls_o)
    cmp_f("loc_t dense vs oracle-cuda", r["loc_t"], locd_o)
    cmp_i("pos_mask vs golden(cpu ref)", r["pos_mask"], U.unpack_bits(g["pos_bits"], 8732))
    cmp_i("cls_t vs golden", r["cls_t"], torch.from_numpy(g["cls_t"].astype(np.int64)))
    cmp_i("n_pos vs golden", r["n_pos"].long(), torch.from_numpy(g["n_pos"]))
    pm, lpm, ct = ssdhot.build_targets(PS, tg, 300, 300, thr, "cuda")
    cmp_f("build_targets loc_t_pm vs golden", lpm, torch.from_numpy(g["loc_t_pm"]))
    cmp_f("build_targets loc_t_pm vs oracle-cuda", lpm, locpm_o)
    # per-image matched idx for image 0 vs oracle
    t0 = tg[0]
    unit = t0["boxes"] / torch.tensor([300.0] * 4, device=dev) if t0["boxes"].numel() else t0["boxes"].new_zeros((0, 4))
    e = O.match_encode(pri_g, pri_xyxy_g, unit, t0["labels"], thr, return_match=True)
    cmp_i("matched_gt[0] vs oracle-cuda", r["matched_gt"][0].long(), e[4])
    cmp_f("matched_box[0] vs oracle-cuda", r["matched_cxcywh"][0], e[3])
    cmp_f("enc0 loc vs golden", r["loc_t"][0], torch.from_numpy(g["enc0_loc"]))
    cmp_f("enc0 matched vs golden", r["matched_cxcywh"][0], torch.from_numpy(g["enc0_match"]))
    # losses
    lg, cg = loc_all.to(dev), conf.to(dev)
    l_loc, l_conf = ssdhot.multibox_loss(PS, lg, cg, targets, thr, ratio)
    o_loc, n_img, total = O.loc_loss(lg, pos_o, locpm_o)
    o_conf = O.mined_ce_loss(cg, cls_o, pos_o, n_img, total, ratio)
    print(f"  loc_loss ssdhot {l_loc.item():.8f} oracle-cuda {o_loc.item():.8f} golden {float(g['loc_loss']):.8f}")
    print(f"  conf_loss ssdhot {l_conf.item():.8f} oracle-cuda {o_conf.item():.8f} golden {float(g['conf_loss']):.8f}")
    c2 = ssdhot.CELoss_w_neg_mining(cg, cls_o, pos_o, n_img, total, ratio)
    print(f"  CELoss_w_neg_mining drop-in {c2.item():.8f}")
for n in ["train_cfg1.npz", "train_cfg2.npz", "train_cfg2_thr04.npz", "train_cfg5_b2.npz", "train_edges.npz"]:
    run(lambda n=n: t_train(n))

def t_predict(name):
    section("predict " + name)
    g = U.load(name)
    loc_all, conf = U.predict_inputs(g)
    st, nt, mx, ag = float(g["score_thresh"]), float(g["nms_thresh"]), int(g["max_per_img"]), bool(g["class_agnostic"])
    lg, cg = loc_all.to(dev), conf.to(dev)
    t0 = time.time()
    want = O.postprocess(pri_g, lg, cg, st, nt, mx, ag, nms_limit=True, with_index=True)
    torch.cuda.synchronize(); t1 = time.time()
    labels, scores, boxes, count, cand = ssdhot.predict_padded(PS, lg, cg, st, nt, mx, ag, want_cand=True)
    torch.cuda.synchronize()
    gold = U.split_predictions(g)
    print(f"  oracle-cuda {t1-t0:.2f}s; counts ssdhot {count.tolist()} oracle {[w['labels'].numel() for w in want]} golden {g['counts'].tolist()}")
    for b in range(len(want)):
        k = int(count[b])
        ko = want[b]["labels"].numel()
        if k == ko:
            cmp_i(f"img{b} cand vs oracle-cuda", cand[b, :k].long(), want[b]["cand"])
            cmp_f(f"img{b} scores vs oracle-cuda", scores[b, :k], want[b]["scores"])
            cmp_f(f"img{b} boxes vs oracle-cuda", boxes[b, :k], want[b]["boxes"])
        if k == gold[b]["labels"].numel():
            cmp_i(f"img{b} labels vs golden", labels[b, :k], gold[b]["labels"])
            cmp_f(f"img{b} scores vs golden", scores[b, :k], gold[b]["scores"])
            cmp_f(f"img{b} boxes vs golden", boxes[b, :k], gold[b]["boxes"])
for n in ["predict_cfg1.npz", "predict_cfg3_b4.npz", "predict_cfg3_b2_notebook.npz", "predict_cfg3_b2_agnostic.npz",
          "predict_cfg3_b2_empty.npz", "predict_cfg5_b1.npz"]:
    run(lambda n=n: t_predict(n))

def t_static():
    section("static methods")
    g = U.load("static_methods.npz")
    boxes, scores = torch.from_numpy(g["boxes"]).to(dev), torch.from_numpy(g["scores"]).to(dev)
    for thr, key in ((0.45, "keep45"), (0.30, "keep30")):
        k = ssdhot.iou_nms(boxes, scores, thr)
        cmp_i(f"iou_nms {thr} vs golden", k, torch.from_numpy(g[key]))
    gen = torch.Generator().manual_seed(77)
    for _ in range(2): torch.rand((600, 2), generator=gen)
    torch.rand((600,), generator=gen)
    loc = torch.randn((8732, 4), generator=gen)
    d = ssdhot.decode_ssd(loc.to(dev), PS.priors, (0.1, 0.2))
    cmp_f("decode vs golden", d, torch.from_numpy(g["decoded"]))
    cmp_f("decode vs oracle-cuda", d, O.decode(loc.to(dev), pri_g, (0.1, 0.2)))
run(t_static)

def t_time():
    section("timing (B=256)")
    cfg = synth.config(3)
    lg, cg_t, cg_i = cfg["loc_all"].to(dev), cfg["conf_train"].to(dev), cfg["conf_infer"].to(dev)
    packed = ssdhot.pack_targets(cfg["targets"], dev)
    def timeit(fn, n=20):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    t = timeit(lambda: ssdhot.multibox_loss(PS, lg, cg_t, packed, 0.5, 3.0))
    print(f"  multibox_loss fwd B=256: {t*1e3:.1f} us  -> {256/t*1e3:.0f} img/s")
    t = timeit(lambda: ssdhot.match_encode_batch(PS, packed, 0.5, (300, 300)))
    print(f"  match_encode B=256: {t*1e3:.1f} us")
    t = timeit(lambda: ssdhot.predict_padded(PS, lg, cg_i, 0.01, 0.45, 200))
    print(f"  predict_padded B=256: {t*1e3:.1f} us  -> {256/t*1e3:.0f} img/s")
    cfg5 = synth.config(5, batch=64)
    l5, c5 = cfg5["loc_all"].to(dev), cfg5["conf_infer"].to(dev)
    t = timeit(lambda: ssdhot.predict_padded(PS, l5, c5, 0.0, 0.45, 200), n=5)
    print(f"  predict_padded stress B=64 thr0: {t*1e3:.1f} us -> {64/t*1e3:.0f} img/s")
    p5 = ssdhot.pack_targets(cfg5["targets"], dev)
    t = timeit(lambda: ssdhot.multibox_loss(PS, l5, cfg5["conf_train"].to(dev), p5, 0.5, 3.0), n=5)
    print(f"  multibox_loss G=64 B=64: {t*1e3:.1f} us -> {64/t*1e3:.0f} img/s")
run(t_time)
print("\nlaunches:", ssdhot.launch_count())
