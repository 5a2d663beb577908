"""Throughput of the eval step with one and with two steps in flight (two HotPathStep objects replaying their graphs on two
streams, alternating batches): consecutive eval batches are independent, so the loss stream of batch i+1 may overlap the
NMS tail of batch i."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "automotive-ssd-object-detection_b200")):
    sys.path.insert(0, p)
import torch, ssdhot
from ssdhot import synth
from ssdhot.engine import HotPathStep
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda:0")
ps = ssdhot.PriorSet.default(dev)
sets = []
for k in range(4):
    cfg = synth.config(3, batch=batch, seed_offset=k)
    sets.append((cfg["loc_all"].to(dev), cfg["conf_infer"].to(dev), ssdhot.pack_targets(cfg["targets"], dev)))
mk = lambda: HotPathStep(ps, batch, 6, cfg["iou_thresh"], cfg["ratio"], cfg["score_thresh"], cfg["nms_thresh"], cfg["max_per_img"])
def run(depth, n=400):
    steps = [mk() for _ in range(depth)]
    streams = [torch.cuda.Stream(dev) for _ in range(depth)]
    def one(i):
        s = sets[i % 4]
        with torch.cuda.stream(streams[i % depth]):
            steps[i % depth].run(s[0], s[1], s[1], s[2], use_graph=True)
    for i in range(16): one(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cur = torch.cuda.current_stream(dev)
    a.record()
    for st in streams: st.wait_stream(cur)
    for i in range(n): one(i)
    for st in streams: cur.wait_stream(st)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
for d in (1, 2, 3):
    print(f"B={batch} steps in flight {d}: {run(d):.1f} us per step")
