"""ssdhot -- B200-native (sm_100a) SSD300 multibox post-backbone hot path.

Drop-in replacements for the reference's `build_targets`, `CELoss_w_neg_mining`
(SSD_trainer.py) and `mySSD.encode_ssd`, `decode_ssd`, `iou_nms`, `predict`
(SSD_from_scratch.py), backed by hand-written CUDA kernels behind the C ABI declared in
include/ssdhot.h.  See DESIGN.md and INTEGRATION.md at the repository root.

Importing the package does not load the shared library (so `ssdhot.synth` and
`ssdhot.priors.default_boxes` work anywhere); the first compute call does, and raises if
libssdhot.so is missing -- there is no CPU fallback.
"""
__version__ = "0.1.0"

from ._lib import SsdhotError, launch_count, lib  # noqa: F401
from .priors import PriorSet, default_boxes  # noqa: F401
from .api import (CELoss_w_neg_mining, HeadSet, PackedTargets, build_targets, collate_detection, decode_ssd, eval_step, encode_ssd, forward_heads, iou_nms,  # noqa: F401
                  match_encode_batch, multibox_loss, multibox_loss_heads, nms_sets, pack_heads, pack_targets, patch, predict,
                  predict_heads, predict_heads_padded, predict_padded, smooth_l1_positive_loss)
from .trainer import SSD_test_step, SSD_train_step  # noqa: F401
