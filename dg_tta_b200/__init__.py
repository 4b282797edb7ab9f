"""dg_tta_b200 — B200 (sm_100a) implementation of DG-TTA's input-transform hot path.

Public names mirror the reference package (multimodallearning/DG-TTA):
    dg_tta_b200.mind   <-> dg_tta/mind.py        MIND3D, mind_hook
    dg_tta_b200.gin    <-> dg_tta/gin.py         GINGroupConv, GradlessGCReplayNonlinBlock, gin_aug, gin_hook
    dg_tta_b200.utils  <-> dg_tta/utils.py       enable/disable/get_internal_augmentation_enabled
    dg_tta_b200.tta.augmentation_utils <-> dg_tta/tta/augmentation_utils.py  get_rand_affine, gin_mind_aug
                                        (+ affine_grid_sample for the inline affine_grid/grid_sample pairs)
Everything executes in libdgtta_sm100.so (hand-written CUDA, C ABI in include/dgtta.h).
"""
from .gin import GINGroupConv, GradlessGCReplayNonlinBlock, gin_aug, gin_hook  # noqa: F401
from .mind import MIND3D, mind_hook, mind_ssc  # noqa: F401
from .utils import (disable_internal_augmentation, enable_internal_augmentation,  # noqa: F401
                    get_internal_augmentation_enabled)

__version__ = "0.1.0"
