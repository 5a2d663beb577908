"""Per-phase timeline of nms_image_kernel's first round (debug hook ssdhot_debug_timeline)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "automotive-ssd-object-detection_b200")):
    sys.path.insert(0, p)
import torch, ssdhot
from ssdhot import synth
from ssdhot.engine import HotPathStep
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda:0")
cfg = synth.config(3, batch=batch)
ps = ssdhot.PriorSet.default(dev)
loc, ci = cfg["loc_all"].to(dev), cfg["conf_infer"].to(dev)
step = HotPathStep(ps, batch, 6, cfg["iou_thresh"], cfg["ratio"], cfg["score_thresh"], cfg["nms_thresh"], cfg["max_per_img"])
st = torch.cuda.current_stream(dev).cuda_stream
for _ in range(3): step.launch_predict(loc, ci, st)
COLD = os.environ.get("COLD")            # COLD=1: evict the inputs from L2 (as bench.py's rotating sets do) before the stamped launch
if COLD:
    junk = torch.empty((512 << 20,), dtype=torch.uint8, device=dev); junk.fill_(1); junk.fill_(2); torch.cuda.synchronize()
    print("cold L2")
tl = torch.zeros((batch, 16), dtype=torch.int64, device=dev)
ssdhot.lib().ssdhot_debug_timeline(tl.data_ptr())
step.launch_predict(loc, ci, st)
torch.cuda.synchronize()
ssdhot.lib().ssdhot_debug_timeline(None)
t = tl.cpu().double()
fused = bool((t[:, 14] > 0).any())          # predict_image_kernel stamps its start (before the stream) in slot 14
t0 = t[:, 14].min() if fused else t[:, 0].min()
if fused:
    print(f"fused predict_image_kernel: images that left the fused mould (generic path): {int((t[:, 15] > 0).sum())} of {batch}")
    codes = tl.cpu()[:, 15]
    for code in sorted(set((codes & 255).tolist()) - {0}):
        sel = codes[(codes & 255) == code] >> 8
        print(f"   reason {code} (1 list overflow, 2 no histogram cut, 3 gather overflow, 4 nothing valid, 5 second round): {sel.numel()} images, info "
              f"{[(int(x) & 4095, int(x) >> 12) for x in sel[:6].tolist()]}")
    v = (t[:, 14] - t0) / 1e3
    print(f"{'CTA start':16s} med {v.median():6.1f} max {v.max():6.1f}")
names = {0: "streamed" if fused else "start", 1: "hist built", 2: "cut found", 3: "gathered", 4: "exact keys", 5: "sorted", 6: "decoded+ordered",
         7: "pairs tested", 8: "resolved", 9: "emitted", 10: "end"}
sm = t[:, 11].long()
shared = torch.tensor([(sm == x).sum().item() > 1 for x in sm])
for k in range(11):
    v = (t[:, k] - t0) / 1e3
    print(f"{names[k]:16s} med {v.median():6.1f} max {v.max():6.1f} | SM shared med {v[shared].median():6.1f} | SM alone med {v[~shared].median() if (~shared).any() else float('nan'):6.1f}")
print("candidates per image: med", t[:, 12].median().item(), " pulled in round 1 (K): med", t[:, 13].median().item(), "max", t[:, 13].max().item())
