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


def reduce_sums(sums: torch.Tensor, group) -> torch.Tensor:
    """The sharded path's exchange for any kind of `group`: a PeerSums (one kernel over NVLink peer memory), True (the
    default torch.distributed group) or a torch.distributed process group."""
    if isinstance(group, PeerSums):
        return group.allreduce(sums)
    return combine_sums(sums, None if group is True else group)


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


class PeerSums:
    """The sharded path's exchange as ONE kernel over NVLink peer memory (csrc/peer.cu, ssdhot_allreduce_sums_peer):
    every rank owns a 1 KB mailbox that the other ranks map through CUDA IPC; `allreduce(sums)` stores this rank's
    three sums into every mailbox, waits for the others' and adds them in rank order.  Unlike a library collective it
    costs no host work beyond one kernel launch and can be captured in the step's CUDA graph (HotPathStep(group=PeerSums)).
    All ranks of `group` must sit on one node (one NVSwitch domain, at most 8 ranks) and call allreduce the same number
    of times.  lag=1: allreduce delivers the reduced sums of the PREVIOUS call (zeros on the first), which were posted a
    step earlier -- no waiting, no per-step re-synchronisation of the ranks; call it once more at the end for the last step."""

    def __init__(self, device, group=None, lag: int = 0, _virtual=None):
        import ctypes
        self.lag = int(lag)
        from . import _lib
        self._lib, self._ct = _lib, ctypes
        L = _lib.lib()
        self.device = torch.device(device)
        self.flags = torch.zeros((1,), dtype=torch.int32, device=self.device)
        self._opened = []
        if _virtual is not None:                       # several ranks inside one process (tests): mailboxes shared directly
            self.rank, self.world, boxes = _virtual
            self._own = None
        else:
            self.world = dist.get_world_size(group) if dist.is_initialized() else 1
            self.rank = dist.get_rank(group) if dist.is_initialized() else 0
            if self.world > 8:
                raise _lib.SsdhotError("PeerSums spans one NVSwitch domain: at most 8 ranks")
            with torch.cuda.device(self.device):
                own = ctypes.c_void_p()
                _lib.check(L.ssdhot_peer_alloc(ctypes.byref(own)), "ssdhot_peer_alloc")
                self._own = own.value
                boxes = [None] * self.world
                boxes[self.rank] = self._own
                if self.world > 1:
                    handle = (ctypes.c_ubyte * 64)()
                    _lib.check(L.ssdhot_peer_export(self._own, handle), "ssdhot_peer_export")
                    mine = torch.tensor(list(handle), dtype=torch.uint8, device=self.device)
                    every = [torch.empty_like(mine) for _ in range(self.world)]
                    dist.all_gather(every, mine, group=group)
                    for r, t in enumerate(every):
                        if r == self.rank:
                            continue
                        raw = (ctypes.c_ubyte * 64)(*t.cpu().tolist())
                        mapped = ctypes.c_void_p()
                        _lib.check(L.ssdhot_peer_open(raw, ctypes.byref(mapped)), "ssdhot_peer_open")
                        boxes[r] = mapped.value
                        self._opened.append(mapped.value)
                    dist.barrier(group=group)
        self._boxes = (ctypes.c_void_p * self.world)(*boxes)
        self._boxes_ptr = ctypes.cast(self._boxes, ctypes.c_void_p)

    @classmethod
    def virtual(cls, world: int, device, lag: int = 0):
        """`world` ranks inside ONE process on one GPU (each to be driven from its own stream): the protocol without IPC."""
        import ctypes
        from . import _lib
        boxes = []
        with torch.cuda.device(device):
            for _ in range(world):
                p = ctypes.c_void_p()
                _lib.check(_lib.lib().ssdhot_peer_alloc(ctypes.byref(p)), "ssdhot_peer_alloc")
                boxes.append(p.value)
        ranks = [cls(device, lag=lag, _virtual=(r, world, boxes)) for r in range(world)]
        ranks[0]._virtual_boxes = boxes                 # freed by ranks[0].close()
        return ranks

    def allreduce(self, sums: torch.Tensor, stream: Optional[int] = None) -> torch.Tensor:
        """All-reduce (sum) the [3] float64 tensor in place on `stream` (default: the current stream of its device)."""
        if sums.dtype != torch.float64 or sums.numel() != 3 or not sums.is_cuda:
            raise self._lib.SsdhotError("PeerSums.allreduce needs a CUDA float64 tensor of 3 elements")
        if stream is None:
            stream = torch.cuda.current_stream(sums.device).cuda_stream
        with torch.cuda.device(self.device):
            rc = self._lib.lib().ssdhot_allreduce_sums_peer(sums.data_ptr(), self._boxes_ptr, self.rank, self.world,
                                                            self.lag, self.flags.data_ptr(), stream)
        self._lib.check(rc, "ssdhot_allreduce_sums_peer")
        return sums

    def allreduce_partials(self, loss_work: torch.Tensor, batch: int, n_pos: Optional[torch.Tensor], sums: torch.Tensor,
                           stream: Optional[int] = None) -> torch.Tensor:
        """allreduce() fed by the per-image partial sums a loss forward left in `loss_work` (forward called with sums=None):
        the kernel folds them itself, so the loss branch of a step is loss kernel -> this kernel (no finalize launch)."""
        if sums.dtype != torch.float64 or sums.numel() != 3 or not sums.is_cuda:
            raise self._lib.SsdhotError("PeerSums.allreduce_partials needs a CUDA float64 tensor of 3 elements")
        if stream is None:
            stream = torch.cuda.current_stream(sums.device).cuda_stream
        with torch.cuda.device(self.device):
            rc = self._lib.lib().ssdhot_allreduce_partials_peer(loss_work.data_ptr(), int(batch), None if n_pos is None else n_pos.data_ptr(),
                                                                sums.data_ptr(), self._boxes_ptr, self.rank, self.world, self.lag,
                                                                self.flags.data_ptr(), stream)
        self._lib.check(rc, "ssdhot_allreduce_partials_peer")
        return sums

    def timed_out(self) -> bool:
        """True if some allreduce gave up waiting for a peer (host synchronisation)."""
        return bool(int(self.flags.item()) & 2)

    def close(self) -> None:
        L = self._lib.lib()
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            for m in self._opened:
                L.ssdhot_peer_close(m)
            self._opened = []
            if self._own:
                L.ssdhot_peer_free(self._own)
                self._own = None
            for m in getattr(self, "_virtual_boxes", []):
                L.ssdhot_peer_free(m)
            self._virtual_boxes = []
