"""CUDA-event time of each half alone and of the forked step (CUDA graph), rotating cold inputs -- what bench.py reports under
roofline.parts and value, in a few seconds (for A/B runs of library variants: SSDHOT_LIB_PATH=...)."""
import os, sys, statistics
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
step = HotPathStep(ps, batch, 6, cfg["iou_thresh"], cfg["ratio"], cfg["score_thresh"], cfg["nms_thresh"], cfg["max_per_img"])
st = torch.cuda.current_stream(dev).cuda_stream
def half(train, iters=48):
    f = (lambda s: step.launch_loss(s[0], s[1], s[2], st)) if train else (lambda s: step.launch_predict(s[0], s[1], st))
    for i in range(4): f(sets[i % 4])
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for i, (a, b) in enumerate(evs):
        a.record(); f(sets[i % 4]); b.record()
    torch.cuda.synchronize()
    return statistics.mean(a.elapsed_time(b) for a, b in evs) * 1e3
def whole(share, n=300):
    step.share_keys = share
    for i in range(12): step.run(sets[i % 4][0], sets[i % 4][1], sets[i % 4][1], sets[i % 4][2], use_graph=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n): step.run(sets[i % 4][0], sets[i % 4][1], sets[i % 4][1], sets[i % 4][2], use_graph=True)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
print(f"{os.environ.get('SSDHOT_LIB_PATH', 'default')}: B={batch} loss {half(True):.1f} us  predict {half(False):.1f} us  step(share) {whole(True):.1f} us  step(no share) {whole(False):.1f} us")
