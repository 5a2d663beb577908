"""BASELINE cfg 5 at full size: B = 1024, 64 boxes per image, score threshold 0.0 (all 8732 x 5 pairs are candidates).
Checks invariants that need no oracle and times both halves."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "automotive-ssd-object-detection_b200")):
    sys.path.insert(0, p)
import torch, ssdhot
from ssdhot import synth
from ssdhot.engine import HotPathStep
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device("cuda:0")
t0 = time.time(); cfg = synth.config(5, batch=B); print(f"synth {time.time()-t0:.1f}s", flush=True)
ps = ssdhot.PriorSet.default(dev)
loc, ct, ci = cfg["loc_all"].to(dev), cfg["conf_train"].to(dev), cfg["conf_infer"].to(dev)
gt = ssdhot.pack_targets(cfg["targets"], dev)
step = HotPathStep(ps, B, 6, cfg["iou_thresh"], cfg["ratio"], cfg["score_thresh"], cfg["nms_thresh"], cfg["max_per_img"], max_gt=64)
st = torch.cuda.current_stream(dev).cuda_stream
for name, f in (("match+loss", lambda: step.launch_loss(loc, ct, gt, st)), ("predict", lambda: step.launch_predict(loc, ci, st))):
    for _ in range(2): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): f()
    b.record(); torch.cuda.synchronize()
    print(f"cfg5 B={B} {name}: {a.elapsed_time(b)/5*1e3:.0f} us per batch -> {B/(a.elapsed_time(b)/5e3):.0f} img/s", flush=True)
l_loc, l_conf = step.losses()
assert torch.isfinite(l_loc) and torch.isfinite(l_conf)
assert int(step.n_pos.min()) >= 1 and int(step.count.min()) >= 1 and int(step.count.max()) <= cfg["max_per_img"]
k = step.count.long()
s = step.scores
for b in range(0, B, max(1, B // 16)):                      # survivors come out in descending score order
    v = s[b, : int(k[b])]
    assert bool((v[:-1] >= v[1:]).all())
print("ok", l_loc.item(), l_conf.item(), int(step.n_pos.sum()), int(k.sum()))
