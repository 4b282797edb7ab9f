"""dg_tta_b200 — B200 (sm_100a) implementation of DG-TTA's input-transform hot path.

Public names mirror the reference package (multimodallearning/DG-TTA):
    dg_tta_b200.mind   <-> dg_tta/mind.py        MIND3D, mind_hook
    dg_tta_b200.gin    <-> dg_tta/gin.py         GINGroupConv, GradlessGCReplayNonlinBlock, gin_aug, gin_hook
    dg_tta_b200.utils  <-> dg_tta/utils.py       enable/disable/get_internal_augmentation_enabled
    dg_tta_b200.tta.augmentation_utils <-> dg_tta/tta/augmentation_utils.py  get_rand_affine, gin_mind_aug
                                        (+ affine_grid_sample / affine_label_argmax for the inline
                                         affine_grid + grid_sample pairs of tta.py and torch_utils.py)
    dg_tta_b200.tta.torch_utils <-> dg_tta/tta/torch_utils.py  get_batch, get_argmaxed_segs, soft_dice_loss
                                        (+ consistency_dice_loss: the loss assembly of tta.py:263-269 in one pass)
    dg_tta_b200.host_pipeline   —  pinned-host-in / pinned-host-out front end (H2D, transform, D2H overlapped)
Everything executes in libdgtta_sm100.so (hand-written CUDA, C ABI in include/dgtta.h).
"""
from .gin import GINGroupConv, GradlessGCReplayNonlinBlock, gin_aug, gin_hook  # noqa: F401
from .mind import MIND3D, mind_hook, mind_ssc  # noqa: F401
from .utils import (disable_internal_augmentation, enable_internal_augmentation,  # noqa: F401
                    get_internal_augmentation_enabled)

__version__ = "0.1.0"
