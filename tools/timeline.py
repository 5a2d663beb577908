"""Per-phase timeline of train_image_kernel (debug hook ssdhot_debug_timeline): median microseconds from kernel start."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "automotive-ssd-object-detection_b200")):
    sys.path.insert(0, p)
import torch, ssdhot
from ssdhot import synth
from ssdhot.engine import HeadSet, HotPathStep
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
source = sys.argv[2] if len(sys.argv) > 2 else "packed"          # packed | nchw | nhwc
dev = torch.device("cuda:0")
cfg = synth.config(3, batch=batch)
ps = ssdhot.PriorSet.default(dev)
loc, ct = cfg["loc_all"].to(dev), cfg["conf_train"].to(dev)
gt = ssdhot.pack_targets(cfg["targets"], dev)
step = HotPathStep(ps, batch, 6, cfg["iou_thresh"], cfg["ratio"])
st = torch.cuda.current_stream(dev).cuda_stream
if source != "packed":
    heads = HeadSet(synth.heads_from_packed(loc, source == "nhwc"), synth.heads_from_packed(ct, source == "nhwc"))
    launch = lambda: step.launch_loss_heads(heads, gt, st)
else:
    launch = lambda: step.launch_loss(loc, ct, gt, st)
for _ in range(3): launch()
COLD = os.environ.get("COLD")            # COLD=1: evict the inputs from L2 (as bench.py's rotating sets do) before the stamped launch
if COLD:
    junk = torch.empty((512 << 20,), dtype=torch.uint8, device=dev); junk.fill_(1); junk.fill_(2); torch.cuda.synchronize()
    print("cold L2")
tl = torch.zeros((batch, 16), dtype=torch.int64, device=dev)
ssdhot.lib().ssdhot_debug_timeline(tl.data_ptr())
launch()
torch.cuda.synchronize()
ssdhot.lib().ssdhot_debug_timeline(None)
print("source:", source)
t = tl.cpu().double()
t0 = t[:, 0].min()
names = {0: "start", 1: "cleared", 11: "1a done", 12: "1a'+1b done", 13: "1c done", 14: "1d done", 3: "match done", 2: "stream done",
         4: "joined", 5: "claimed", 6: "bracket", 7: "scanned", 8: "lists done", 9: "end"}
sm = t[:, 10].long()
shared = torch.tensor([(sm == s).sum().item() > 1 for s in sm])
for k in (0, 1, 11, 12, 13, 14, 3, 2, 4, 5, 6, 7, 8, 9):
    v = (t[:, k] - t0) / 1e3
    print(f"{names[k]:14s} all: med {v.median():6.1f} max {v.max():6.1f} | SM shared: med {v[shared].median():6.1f} | SM alone: med {v[~shared].median() if (~shared).any() else float('nan'):6.1f}")
