import os, sys
ROOT = "/root/repo"
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
for trial in range(2):
    tl = torch.zeros((batch, 16), dtype=torch.int64, device=dev)
    ssdhot.lib().ssdhot_debug_timeline(tl.data_ptr())
    step.launch_loss(loc, ct, gt, st)
    torch.cuda.synchronize()
    ssdhot.lib().ssdhot_debug_timeline(None)
    t = tl.cpu()
    sm = t[:, 10].tolist()
    dur = ((t[:, 9] - t[:, 0]).double() / 1e3).tolist()
    G = [int(x["boxes"].shape[0]) for x in cfg["targets"]]
    same = sum(1 for j in range(108) if sm[j] == sm[j + 148])
    print("trial", trial, "blocks j and j+148 on the same SM:", same, "of 108; distinct SMs among blocks 0..147:", len(set(sm[:148])))
    print("first 20 smids:", sm[:20], "blocks 148..160:", sm[148:160])
    import statistics
    # correlation of duration with G
    byg = {}
    for g, d in zip(G, dur): byg.setdefault(g, []).append(d)
    print("median CTA duration by G:", {g: round(statistics.median(v), 1) for g, v in sorted(byg.items())})
