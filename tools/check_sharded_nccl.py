"""Sharded-path check on real GPUs (torchrun, NCCL): every rank runs the fused match + mined loss on its contiguous
slice of a BASELINE cfg-4 style batch, the three partial sums are all-reduced, and the result must equal the
single-GPU run over the whole batch (1e-12 on the double sums: same fp32 terms, different double summation order).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/check_sharded_nccl.py [images_per_rank]
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "automotive-ssd-object-detection_b200")):
    sys.path.insert(0, p)
import torch
import torch.distributed as dist
import ssdhot
from ssdhot import dist as D, synth

per_rank = int(sys.argv[1]) if len(sys.argv) > 1 else 512
rank, world, local = D.init_from_env()
dev = torch.device("cuda", local)
cfg = synth.config(4, batch=per_rank * world)                 # the same seeded global batch on every rank
lo, hi = D.shard_range(per_rank * world, rank, world)
ps = ssdhot.PriorSet.default(dev)
loc, conf = cfg["loc_all"][lo:hi].to(dev), cfg["conf_train"][lo:hi].to(dev)
l_loc, l_conf, sums = ssdhot.multibox_loss(ps, loc, conf, cfg["targets"][lo:hi], cfg["iou_thresh"], cfg["ratio"],
                                            group=True if world > 1 else None, return_sums=True)
torch.cuda.synchronize(dev)
# the same exchange through the peer-memory kernel (csrc/peer.cu) instead of NCCL, three times in a row
peer = D.PeerSums(dev) if world > 1 else None
peer_sums = []
for _ in range(3):
    if peer is None:
        break
    peer_sums.append(ssdhot.multibox_loss(ps, loc, conf, cfg["targets"][lo:hi], cfg["iou_thresh"], cfg["ratio"], group=peer,
                                          return_sums=True)[2].clone())
torch.cuda.synchronize(dev)
peer_rel = max(((p - sums).abs() / sums.abs().clamp_min(1e-300)).max().item() for p in peer_sums) if peer_sums else 0.0
peer_ok = peer is None or (peer_rel <= 1e-12 and not peer.timed_out())
if peer is not None:                              # lag 1: the second call delivers the first call's reduced sums
    lagged = D.PeerSums(dev, lag=1)
    first = lagged.allreduce(sums.clone() * 0 + (rank + 1.0))
    second = lagged.allreduce(torch.zeros_like(sums))
    torch.cuda.synchronize(dev)
    peer_ok = peer_ok and bool((first == 0).all()) and bool((second == world * (world + 1) / 2.0).all()) and not lagged.timed_out()
    dist.barrier()
    lagged.close()
out = {"world": world, "images": per_rank * world, "loc_loss": l_loc.item(), "conf_loss": l_conf.item(), "sums": sums.tolist()}
if rank == 0:
    f_loc, f_conf, f_sums = ssdhot.multibox_loss(ps, cfg["loc_all"].to(dev), cfg["conf_train"].to(dev), cfg["targets"],
                                                 cfg["iou_thresh"], cfg["ratio"], return_sums=True)
    rel = ((sums - f_sums).abs() / f_sums.abs().clamp_min(1e-300)).max().item()
    out.update({"single_gpu_loc_loss": f_loc.item(), "single_gpu_conf_loss": f_conf.item(), "max_rel_diff_of_sums": rel,
                "peer_allreduce_max_rel_diff_vs_nccl": peer_rel, "peer_allreduce_ok": bool(peer_ok),
                "ok": bool(rel <= 1e-12 and sums[2].item() == f_sums[2].item() and peer_ok)})
    print(json.dumps(out))
    assert out["ok"], out
assert peer_ok, (rank, peer_rel)
if world > 1:
    dist.barrier()
    peer.close()
    dist.destroy_process_group()
