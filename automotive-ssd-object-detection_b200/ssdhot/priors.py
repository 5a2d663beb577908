"""Default boxes of SSD300 (SSD_from_scratch.py:275-331) and their per-prior constants."""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch

from . import _lib

FMAPS = (38, 19, 10, 5, 3, 1)
RATIOS = ((2,), (2, 3), (2, 3), (2, 3), (2,), (2,))


def default_boxes(s_min: float = 0.2, s_max: float = 0.9, clip: bool = True) -> torch.Tensor:
    """[8732,4] (cx,cy,w,h) fp32 on the CPU.  Host-side one-off; shapes per cell are
    [s, sqrt(s s'), (s*sqrt(a), s/sqrt(a)), (s/sqrt(a), s*sqrt(a)) for a in ratios], cells
    row-major, levels 38/19/10/5/3/1, computed in float64 and rounded once to fp32."""
    levels = len(FMAPS)
    s = [s_min + (s_max - s_min) * (l / (levels - 1)) for l in range(levels)] + [1.0]
    per_level = []
    for l, n in enumerate(FMAPS):
        wh = [(s[l], s[l]), (math.sqrt(s[l] * s[l + 1]),) * 2]
        for a in RATIOS[l]:
            q = math.sqrt(a)
            wh += [(s[l] * q, s[l] / q), (s[l] / q, s[l] * q)]
        wh_t = torch.tensor(wh, dtype=torch.float64)                                   # [k,2]
        c = (torch.arange(n, dtype=torch.float64) + 0.5) / n
        cy, cx = torch.meshgrid(c, c, indexing="ij")                                    # [n,n]
        ctr = torch.stack((cx, cy), -1).reshape(n * n, 1, 2).expand(-1, wh_t.shape[0], -1)
        per_level.append(torch.cat((ctr, wh_t.expand(n * n, -1, -1)), -1).reshape(-1, 4))
    out = torch.cat(per_level, 0).to(torch.float32)
    if clip:
        out[:, :2].clamp_(0.0, 1.0)
        out[:, 2:].clamp_(1e-6, 1.0)
    return out.contiguous()


class PriorSet:
    """Device-resident prior tables: cxcywh, clamped xyxy, and the per-prior CIoU constants
    (area, centre, atan(w/h)) that `complete_box_iou` would recompute for every image."""

    def __init__(self, priors: torch.Tensor, priors_xyxy: Optional[torch.Tensor] = None,
                 variances: Tuple[float, float] = (0.1, 0.2), img_hw: Tuple[int, int] = (300, 300),
                 generic: bool = False):
        if not priors.is_cuda:
            raise _lib.SsdhotError("ssdhot needs CUDA tensors (no CPU fallback)")
        self.priors = priors.detach().to(torch.float32).contiguous()
        self.P = int(self.priors.shape[0])
        self.variances = (float(variances[0]), float(variances[1]))
        self.img_h, self.img_w = int(img_hw[0]), int(img_hw[1])
        # SSD300 grid structure (checked once on a host copy) enables the box-centric matching kernel
        host = self.priors.cpu().contiguous()
        self.layout = 0 if generic else int(_lib.lib().ssdhot_ssd300_layout_host(host.data_ptr(), self.P))
        stream = torch.cuda.current_stream(self.priors.device).cuda_stream
        self.aux = torch.empty((self.P, 4), dtype=torch.float32, device=self.priors.device)
        with torch.cuda.device(self.priors.device):
            if priors_xyxy is None:
                self.priors_xyxy = torch.empty_like(self.priors)
                _lib.check(_lib.lib().ssdhot_prior_tables(self.priors.data_ptr(), self.P, self.priors_xyxy.data_ptr(),
                                                          self.aux.data_ptr(), stream), "ssdhot_prior_tables")
            else:
                self.priors_xyxy = priors_xyxy.detach().to(torch.float32).contiguous()
                _lib.check(_lib.lib().ssdhot_prior_aux(self.priors_xyxy.data_ptr(), self.P, self.aux.data_ptr(), stream),
                           "ssdhot_prior_aux")
                if self.layout:
                    # the SSD300 fast path rebuilds the clamped corners from the grid position (train_path.cu: grid_prior);
                    # corners that are not box_convert(priors).clamp(0, 1) (SFS:34) keep the table-driven kernels
                    own = torch.empty_like(self.priors)
                    _lib.check(_lib.lib().ssdhot_prior_tables(self.priors.data_ptr(), self.P, own.data_ptr(), None, stream),
                               "ssdhot_prior_tables")
                    if not torch.equal(own, self.priors_xyxy):
                        self.layout = 0

    @property
    def device(self) -> torch.device:
        return self.priors.device

    @classmethod
    def default(cls, device="cuda", variances=(0.1, 0.2), generic: bool = False) -> "PriorSet":
        """The reference's 8732 default boxes.  generic=True withholds the SSD300 layout flag, which
        routes matching through the layout-agnostic kernels (used by the parity tests)."""
        return cls(default_boxes().to(device), None, variances, generic=generic)

    _cache = {}            # (buffer address, device, variances) -> PriorSet, least recently used first; at most CACHE entries
    CACHE = 8

    @classmethod
    def of(cls, model) -> "PriorSet":
        """PriorSet of a reference `mySSD` (uses its `priors` / `priors_xyxy` buffers and variances), cached per buffer --
        several models (a train and an eval copy, say) alternate without re-deriving their tables; a PriorSet is returned
        unchanged.  A cached entry keeps the model's prior buffer alive, so its address cannot be recycled under the key."""
        if isinstance(model, PriorSet):
            return model
        pri = model.priors
        key = (pri.data_ptr(), pri.device, getattr(model, "variance_center", 0.1), getattr(model, "variance_size", 0.2))
        hit = cls._cache.pop(key, None)
        if hit is None:
            hit = cls(pri, getattr(model, "priors_xyxy", None),
                      (getattr(model, "variance_center", 0.1), getattr(model, "variance_size", 0.2)),
                      (getattr(model, "img_h", 300), getattr(model, "img_w", 300)))
            hit.source = pri
            while len(cls._cache) >= cls.CACHE:
                cls._cache.pop(next(iter(cls._cache)))
        cls._cache[key] = hit
        return hit
