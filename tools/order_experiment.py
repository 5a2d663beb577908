"""Does the position of an image in the batch (= its CTA's place in the launch order) matter for train_image_kernel?
Times match+loss at B=256 (cold inputs, rotating sets) for several arrangements of the same images by box count."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "automotive-ssd-object-detection_b200")):
    sys.path.insert(0, p)
import torch, ssdhot
from ssdhot import synth
from ssdhot.engine import HotPathStep

batch, iters, n_sets, n_sm = 256, 60, 4, 148
dev = torch.device("cuda:0")
cfgs = [synth.config(3, batch=batch, seed_offset=i) for i in range(n_sets)]
ps = ssdhot.PriorSet.default(dev)
step = HotPathStep(ps, batch, 6, cfgs[0]["iou_thresh"], cfgs[0]["ratio"])
st = torch.cuda.current_stream(dev).cuda_stream


def arrangement(name, counts):
    idx = sorted(range(batch), key=lambda i: -counts[i])          # heaviest first
    if name == "original":
        return list(range(batch))
    if name == "descending":
        return idx
    if name == "ascending":
        return idx[::-1]
    if name == "heavy_alone":
        # blocks 0..147 land on distinct SMs, blocks 148..255 double up on the SMs of blocks 0..107 (simple model):
        # the heaviest 40 images go to blocks 108..147, the rest pair heavy (block j) with light (block 148 + j)
        second = batch - n_sm
        alone = n_sm - second
        out = [None] * batch
        heavy, rest = idx[:alone], idx[alone:]
        for k, i in enumerate(heavy):
            out[second + k] = i
        for j in range(second):
            out[j] = rest[j]
            out[n_sm + j] = rest[len(rest) - 1 - j]
        return out
    raise ValueError(name)


for name in ("original", "descending", "ascending", "heavy_alone", "original"):
    sets = []
    for cfg in cfgs:
        counts = [int(t["boxes"].shape[0]) for t in cfg["targets"]]
        order = arrangement(name, counts)
        sel = torch.tensor(order)
        sets.append((cfg["loc_all"][sel].contiguous().to(dev), cfg["conf_train"][sel].contiguous().to(dev),
                     ssdhot.pack_targets([cfg["targets"][i] for i in order], dev)))
    for i in range(5):
        step.launch_loss(*sets[i % n_sets], st)
    torch.cuda.synchronize()
    sums0 = step.sums.clone()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for i, (a, b) in enumerate(evs):
        a.record(); step.launch_loss(*sets[i % n_sets], st); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) * 1e3 for a, b in evs)
    print(f"{name:12s} median {ts[len(ts) // 2]:6.1f} us  min {ts[0]:6.1f} us   (sum positives {int(step.sums[2].item())})")
