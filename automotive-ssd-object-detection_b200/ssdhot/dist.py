"""Multi-GPU plumbing of the sharded path.

The batch shards by image: every rank runs the whole hot path on its own contiguous slice with
no data-path collective.  The single exchange of a training step is an all-reduce (sum) of three
doubles -- [sum smooth-L1, sum CE, sum positives] -- after which every rank normalises locally
(SSD_trainer.py:105,108,600).  Inference needs no collective.  One process per GPU,
`torch.distributed` (NCCL over NVLink/NVSwitch on the B200 box, gloo in the CPU tests).
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """Join the job described by RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* (torchrun).
    -> (rank, world_size, local_rank); a no-op single-process answer when WORLD_SIZE is unset."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local)
    return rank, world, local


def shard_range(n_img: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous image slice [lo, hi) of rank `rank` (sizes differ by at most one)."""
    base, extra = divmod(n_img, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def combine_sums(sums: torch.Tensor, group=None) -> torch.Tensor:
    """All-reduce the [3] float64 partial sums in place (no-op without a process group)."""
    if dist.is_available() and dist.is_initialized() and (group is not None or dist.get_world_size() > 1):
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    return sums


def losses_from_sums(sums: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """(loc_loss, conf_loss) = sums[0:2] / max(sums[2], 1) as 0-d fp32 (SSD_trainer.py:105-108, :600)."""
    total = sums[2].clamp_min(1.0)
    out = (sums[:2] / total).to(torch.float32)
    return out[0], out[1]


class SumsReducer:
    """Overlaps the sharded path's one exchange with the next step: `submit(sums)` snapshots the three partial sums
    on the current stream and all-reduces the snapshot on a side stream, so the kernels of the next step do not
    wait for the collective's launch latency; `result()` returns the newest reduced sums (the current stream then
    waits for that collective only)."""

    def __init__(self, device, group=None, depth: int = 2):
        self.group, self.depth, self.n = group, int(depth), 0
        self.side = torch.cuda.Stream(device)
        self.bufs = [torch.zeros((3,), dtype=torch.float64, device=device) for _ in range(self.depth)]
        self.copied = [torch.cuda.Event() for _ in range(self.depth)]
        self.reduced = [torch.cuda.Event() for _ in range(self.depth)]

    def submit(self, sums: torch.Tensor) -> None:
        i = self.n % self.depth
        cur = torch.cuda.current_stream(sums.device)
        if self.n >= self.depth:
            cur.wait_event(self.reduced[i])            # the snapshot slot is free again
        self.bufs[i].copy_(sums)
        self.copied[i].record(cur)
        self.side.wait_event(self.copied[i])
        with torch.cuda.stream(self.side):
            combine_sums(self.bufs[i], self.group)
            self.reduced[i].record(self.side)
        self.n += 1

    def result(self) -> torch.Tensor:
        i = (self.n - 1) % self.depth
        torch.cuda.current_stream(self.bufs[i].device).wait_event(self.reduced[i])
        return self.bufs[i]
