"""train_image_kernel phase durations (debug timeline, L2-warm single launch) for images with few and with many boxes."""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "automotive-ssd-object-detection_b200")):
    sys.path.insert(0, p)
import torch, ssdhot
from ssdhot import synth
from ssdhot.engine import HotPathStep
batch = 256
dev = torch.device("cuda:0")
cfg = synth.config(3, batch=batch)
ps = ssdhot.PriorSet.default(dev)
loc, ct = cfg["loc_all"].to(dev), cfg["conf_train"].to(dev)
gt = ssdhot.pack_targets(cfg["targets"], dev)
step = HotPathStep(ps, batch, 6, cfg["iou_thresh"], cfg["ratio"])
st = torch.cuda.current_stream(dev).cuda_stream
for _ in range(3): step.launch_loss(loc, ct, gt, st)
tl = torch.zeros((batch, 16), dtype=torch.int64, device=dev)
ssdhot.lib().ssdhot_debug_timeline(tl.data_ptr())
step.launch_loss(loc, ct, gt, st)
torch.cuda.synchronize()
ssdhot.lib().ssdhot_debug_timeline(None)
t = tl.cpu().double() / 1e3
G = [int(x["boxes"].shape[0]) for x in cfg["targets"]]
npos = step.n_pos.cpu().tolist()
order = [0, 1, 11, 12, 13, 14, 3, 4, 5, 6, 7, 8, 9]
names = {0: "start", 1: "cleared", 11: "1a", 12: "1a'+1b", 13: "1c", 14: "1d", 3: "match done", 4: "joined", 5: "claimed", 6: "bracket", 7: "scanned", 8: "lists done", 9: "end"}
for label, sel in (("G<=3", [i for i in range(batch) if G[i] <= 3]), ("8<=G<=12", [i for i in range(batch) if 8 <= G[i] <= 12]), ("G>=18", [i for i in range(batch) if G[i] >= 18])):
    print(label, "images", len(sel), "median positives", statistics.median(npos[i] for i in sel))
    prev = None
    for k in order:
        v = statistics.median((t[i, k] - t[i, 0]).item() for i in sel)
        print(f"   {names[k]:11s} at {v:6.1f} us" + (f"  (+{v - prev:4.1f})" if prev is not None else ""))
        prev = v
