"""Quick device timing of each half of the hot path (CUDA events, rotating inputs > L2).
usage: python tools/time_halves.py [batch=256] [iters=50]"""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "automotive-ssd-object-detection_b200")):
    sys.path.insert(0, p)
import torch
import ssdhot
from ssdhot import synth
from ssdhot.engine import HotPathStep

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 50
dev = torch.device("cuda:0")
n_sets = max(2, min(4, (3 * 128 * 2**20) // (batch * 349296) + 1))
sets = []
for i in range(n_sets):
    cfg = synth.config(3, batch=batch, seed_offset=i)
    if os.environ.get("NO_GT"):          # experiment: how much of the train half is matching?
        cfg["targets"] = [{"boxes": torch.zeros((0, 4)), "labels": torch.zeros((0,), dtype=torch.int64)} for _ in cfg["targets"]]
    sets.append((cfg["loc_all"].to(dev), cfg["conf_train"].to(dev), cfg["conf_infer"].to(dev), ssdhot.pack_targets(cfg["targets"], dev)))
ps = ssdhot.PriorSet.default(dev)
step = HotPathStep(ps, batch, 6, cfg["iou_thresh"], cfg["ratio"], cfg["score_thresh"], cfg["nms_thresh"], cfg["max_per_img"])
st = torch.cuda.current_stream(dev).cuda_stream
def run(train, i):
    loc, ct, ci, gt = sets[i % n_sets]
    step.launch_loss(loc, ct, gt, st) if train else step.launch_predict(loc, ci, st)
for name, train in (("match+loss", True), ("predict", False)):
    for i in range(5): run(train, i)
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for i, (a, b) in enumerate(evs):
        a.record(); run(train, i); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) * 1e3 for a, b in evs)
    med = ts[len(ts) // 2]
    print(f"B={batch} {name:11s} median {med:8.1f} us  min {ts[0]:8.1f} us  -> {batch / med:7.3f} img/us  frac_of_hbm {batch * 349296 / (med * 1e-6) / 6549.8e9:5.3f}")
