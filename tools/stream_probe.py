"""16-byte read-only loads vs 1-D bulk-async copies (cp.async.bulk + mbarrier) for the one access
pattern of this path -- one CTA per image pulling its 209,568 contiguous bytes of logits.  Both probes (csrc/probe.cu)
consume every byte (a checksum per image); inputs rotate over > L2 worth of data.  usage: python tools/stream_probe.py"""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "automotive-ssd-object-detection_b200")):
    sys.path.insert(0, p)
import torch, ssdhot
from ssdhot import _lib
dev = torch.device("cuda:0")
IMG = 8732 * 6 * 4
st = torch.cuda.current_stream(dev).cuda_stream
for batch in (256, 2048):
    n_sets = max(2, (3 * 126 * 2**20) // (batch * IMG) + 1)
    sets = [torch.randn((batch, 8732, 6), device=dev) for _ in range(n_sets)]
    out = torch.empty((batch,), device=dev)
    ref = [s.sum(dim=(1, 2)) for s in sets]
    for mode, name in ((0, "LDG.128 x3, two row pairs in flight per lane"), (1, "cp.async.bulk ring, 3 x 24 KB per CTA")):
        for i in range(4):
            _lib.check(_lib.lib().ssdhot_debug_stream_probe(sets[i % n_sets].data_ptr(), batch, IMG, mode, out.data_ptr(), st), "probe")
        torch.cuda.synchronize()
        ok = torch.allclose(out, ref[3 % n_sets], rtol=1e-3, atol=1e-2)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(40)]
        for i, (a, b) in enumerate(evs):
            a.record()
            _lib.lib().ssdhot_debug_stream_probe(sets[i % n_sets].data_ptr(), batch, IMG, mode, out.data_ptr(), st)
            b.record()
        torch.cuda.synchronize()
        us = statistics.median(a.elapsed_time(b) for a, b in evs) * 1e3
        print(f"B={batch:5d} mode {mode} ({name}): {us:7.1f} us  {batch * IMG / us / 1e3:7.1f} GB/s  checksum ok: {ok}")
