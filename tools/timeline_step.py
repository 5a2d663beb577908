"""Timeline of ONE forked step (loss kernel on the current stream, predict kernel on a second one, as HotPathStep.run issues
them): when the CTAs of each kernel start and end, and what the key hand-off changes.  SHARE=0 turns the hand-off off."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "automotive-ssd-object-detection_b200")):
    sys.path.insert(0, p)
import torch, ssdhot
from ssdhot import synth
from ssdhot.engine import HotPathStep
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
share = os.environ.get("SHARE", "1") != "0"
dev = torch.device("cuda:0")
cfg = synth.config(3, batch=batch)
ps = ssdhot.PriorSet.default(dev)
loc, conf = cfg["loc_all"].to(dev), cfg["conf_infer"].to(dev)
gt = ssdhot.pack_targets(cfg["targets"], dev)
step = HotPathStep(ps, batch, 6, cfg["iou_thresh"], cfg["ratio"], cfg["score_thresh"], cfg["nms_thresh"], cfg["max_per_img"])
step.share_keys = share
for _ in range(3):
    step.run(loc, conf, conf, gt)
torch.cuda.synchronize()
junk = torch.empty((512 << 20,), dtype=torch.uint8, device=dev); junk.fill_(1); junk.fill_(2); torch.cuda.synchronize()
ta = torch.zeros((batch, 16), dtype=torch.int64, device=dev)
tb = torch.zeros((batch, 16), dtype=torch.int64, device=dev)
L = ssdhot.lib()
cur = torch.cuda.current_stream(dev)
fork = torch.cuda.Stream(dev)
if share:
    L.ssdhot_share_reset(step.share.data_ptr(), batch, cur.cuda_stream)
fork.wait_stream(cur)
L.ssdhot_debug_timeline(ta.data_ptr())
step.launch_loss(loc, conf, gt, cur.cuda_stream, share=share)
L.ssdhot_debug_timeline(tb.data_ptr())
step.launch_predict(loc, conf, fork.cuda_stream, share=share)
L.ssdhot_debug_timeline(None)
cur.wait_stream(fork)
torch.cuda.synchronize()
a, b = ta.cpu().double(), tb.cpu().double()
t0 = min(a[:, 0].min(), b[:, 14].min())
def line(name, v):
    v = (v - t0) / 1e3
    q = torch.quantile(v, torch.tensor([0.0, 0.25, 0.5, 0.75, 1.0], dtype=torch.float64))
    print(f"{name:28s} min {q[0]:6.1f}  q25 {q[1]:6.1f}  med {q[2]:6.1f}  q75 {q[3]:6.1f}  max {q[4]:6.1f}")
print(f"share = {share}; handed off: {int((step.share[:4 * batch].view(torch.int32) == 3).sum())} of {batch}")
line("loss CTA start", a[:, 0])
line("loss stream done", a[:, 2])
line("loss match done", a[:, 3])
line("loss CTA end", a[:, 9])
line("predict CTA start", b[:, 14])
line("predict keys ready", b[:, 0])
line("predict CTA end", b[:, 10])
d = (b[:, 10] - b[:, 14]) / 1e3
print(f"predict CTA duration: med {d.median():.1f} max {d.max():.1f};  loss CTA duration: med {((a[:, 9] - a[:, 0]) / 1e3).median():.1f}")
late = b[:, 14] > a[:, 9].min()
print(f"predict CTAs that started after the first loss CTA ended: {int(late.sum())}")
