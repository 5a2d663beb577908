"""Device timing of the two halves fed by the head outputs, CUDA events, rotating inputs > L2:
  packed      the kernels on loc_all / conf_all (what bench.py times; the forward's permute + cat not counted)
  pack+packed ssdhot_pack_heads of both branches (2 launches) followed by the packed kernels
  torch+packed the reference's own tail of forward (12 permute().contiguous() + 2 cat) followed by the packed kernels
  heads NCHW  the kernels reading the six NCHW head outputs directly
  heads NHWC  the kernels reading six channels_last head outputs directly
usage: python tools/time_heads.py [batch=256] [iters=40]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "automotive-ssd-object-detection_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import ssdhot
from ssdhot import synth
from ssdhot.engine import HeadSet, HotPathStep
import _util as U

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 40
dev = torch.device("cuda:0")
n_sets = 3
sets = []
for i in range(n_sets):
    cfg = synth.config(3, batch=batch, seed_offset=i)
    loc, ct, ci = cfg["loc_all"].to(dev), cfg["conf_train"].to(dev), cfg["conf_infer"].to(dev)
    d = dict(loc=loc, ct=ct, ci=ci, gt=ssdhot.pack_targets(cfg["targets"], dev))
    for name, cl in (("nchw", False), ("nhwc", True)):
        lh, cth, cih = U.unpack_heads(loc, cl), U.unpack_heads(ct, cl), U.unpack_heads(ci, cl)
        d[name] = dict(lh=lh, cth=cth, cih=cih, train=HeadSet(lh, cth), infer=HeadSet(lh, cih))
    sets.append(d)
ps = ssdhot.PriorSet.default(dev)
step = HotPathStep(ps, batch, 6, cfg["iou_thresh"], cfg["ratio"], cfg["score_thresh"], cfg["nms_thresh"], cfg["max_per_img"])
st = torch.cuda.current_stream(dev).cuda_stream


def torch_pack(heads, D):       # SFS:249-269
    return torch.cat([h.permute(0, 2, 3, 1).contiguous().view(h.shape[0], -1, D) for h in heads], 1)


def run(mode, train, i):
    s = sets[i % n_sets]
    if mode == "packed":
        step.launch_loss(s["loc"], s["ct"], s["gt"], st) if train else step.launch_predict(s["loc"], s["ci"], st)
    elif mode in ("pack+packed", "torch+packed"):
        h = s["nchw"]
        if mode == "pack+packed":
            loc, conf = ssdhot.pack_heads(h["lh"], h["cth"] if train else h["cih"])
        else:
            loc, conf = torch_pack(h["lh"], 4), torch_pack(h["cth"] if train else h["cih"], 6)
        step.launch_loss(loc, conf, s["gt"], st) if train else step.launch_predict(loc, conf, st)
    else:
        h = s[mode]
        step.launch_loss_heads(h["train"], s["gt"], st) if train else step.launch_predict_heads(h["infer"], st)


out = {"batch": batch, "unit": "us (median of %d, CUDA events)" % iters}
ref = {}
for train in (True, False):
    half = "match_loss" if train else "decode_nms"
    out[half] = {}
    for mode in os.environ.get("MODES", "packed,pack+packed,torch+packed,nchw,nhwc").split(","):
        for i in range(4):
            run(mode, train, i)
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
        for i, (a, b) in enumerate(evs):
            a.record(); run(mode, train, i); b.record()
        torch.cuda.synchronize()
        ts = sorted(a.elapsed_time(b) * 1e3 for a, b in evs)
        out[half][mode] = round(ts[len(ts) // 2], 1)
        if mode == "packed":
            ref[half] = (step.sums.clone(), step.count.clone(), step.scores.clone())
        # all modes end on the same input set: identical results
        if half not in ref:
            continue
        same = bool((step.sums == ref[half][0]).all()) if train else bool((step.count == ref[half][1]).all() and (step.scores == ref[half][2]).all())
        out[half][mode + "_same_result"] = same
print(json.dumps(out))
