"""Short driver for ncu: a few direct (non-graph) launches of each half at B=256 (BASELINE cfg 3).
4th argument `shared`: one conf_all for both halves, back to back on one stream (the key hand-off of the eval step)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "automotive-ssd-object-detection_b200")):
    sys.path.insert(0, p)
import torch
import ssdhot
from ssdhot import synth
from ssdhot.engine import HotPathStep

cfg_idx = int(sys.argv[1]) if len(sys.argv) > 1 else 3
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 256
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda:0")
cfg = synth.config(cfg_idx, batch=batch)
ps = ssdhot.PriorSet.default(dev)
loc, ct, ci = cfg["loc_all"].to(dev), cfg["conf_train"].to(dev), cfg["conf_infer"].to(dev)
gt = ssdhot.pack_targets(cfg["targets"], dev)
shared = len(sys.argv) > 4 and sys.argv[4] == "shared"
step = HotPathStep(ps, batch, 6, cfg["iou_thresh"], cfg["ratio"], cfg["score_thresh"], cfg["nms_thresh"], cfg["max_per_img"],
                   concurrent=not shared)
for _ in range(iters):
    step.run(loc, ci if shared else ct, ci, gt)
torch.cuda.synchronize()
print("ok", step.losses(), int(step.count.sum()), "keys handed off:", int((step.share[:4 * batch].view(torch.int32) == 3).sum()))
