"""Bench/test infrastructure (NOT part of the drop-in): a stand-in for DG-TTA's inner adaptation step.

`dg_tta.tta.tta` cannot be imported without nnunetv2, so BASELINE configs 3 and 5 are measured
on this restatement of `tta_main`'s inner loop (dg_tta/tta/tta.py:221-281) and `calc_branch` (:480-579) with
TEMPLATE_PLAN defaults (dg_tta/tta/config_log_utils.py:24-41: affine view augmentation in both branches, gradient
in branch_a, intensity augmentation off, GIN hook disabled during TTA, MIND hook on), driven by the drop-in ops:

    get_batch -> [per branch: get_rand_affine -> affine_grid_sample(border) -> model (mind_hook -> UNet)
                  -> channel selection (map_label, logits) -> affine_grid_sample(zeros, differentiable)]
              -> common-content mask, softmax, soft Dice (torch_utils.py:90-104) -> backward (branch_a only)

The backbone is a PlainConvUNet-shaped fixture built from plans.json:279-401 (5 stages, 32/64/128/256/320 features,
2 convs per stage, strides 1,2,2,2,2, InstanceNorm + LeakyReLU, transposed-conv upsampling, 12 input channels for
the MIND trainers, nnUNetTrainer_GIN_MIND.py:46) with random weights; it stays on PyTorch/cuDNN — out of scope for
the CUDA work, here only so that the transforms are timed in their real consumer.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


class _Block(nn.Sequential):
    def __init__(self, cin, cout, stride):
        super().__init__(nn.Conv3d(cin, cout, 3, stride, 1), nn.InstanceNorm3d(cout, eps=1e-5, affine=True), nn.LeakyReLU(0.01, True),
                         nn.Conv3d(cout, cout, 3, 1, 1), nn.InstanceNorm3d(cout, eps=1e-5, affine=True), nn.LeakyReLU(0.01, True))


class StandInUNet(nn.Module):
    def __init__(self, in_channels=12, num_classes=105, features=(32, 64, 128, 256, 320)):
        super().__init__()
        self.enc = nn.ModuleList()
        c = in_channels
        for i, f in enumerate(features):
            self.enc.append(_Block(c, f, 1 if i == 0 else 2))
            c = f
        self.up, self.dec = nn.ModuleList(), nn.ModuleList()
        for f in reversed(features[:-1]):
            self.up.append(nn.ConvTranspose3d(c, f, 2, 2))
            self.dec.append(_Block(2 * f, f, 1))
            c = f
        self.head = nn.Conv3d(c, num_classes, 1)

    def forward(self, x):
        skips = []
        for blk in self.enc:
            x = blk(x)
            skips.append(x)
        skips.pop()
        for up, dec in zip(self.up, self.dec):
            x = dec(torch.cat([up(x), skips.pop()], 1))
        return self.head(x)


def soft_dice_loss(smp_a, smp_b):
    """dg_tta/tta/torch_utils.py:90-104."""
    B, _, D, H, W = smp_a.shape
    nominator = (2.0 * smp_a * smp_b).reshape(B, -1, D * H * W).mean(2)
    denominator = 0.5 * ((smp_a + smp_b) ** 2).reshape(B, -1, D * H * W).mean(2)
    if denominator.sum() == 0.0:
        return nominator * 0.0 + 1.0
    return nominator / denominator


def build_model(transforms, in_channels=12, num_classes=105, features=(32, 64, 128, 256, 320), seed=0):
    """transforms: module-like namespace with gin_hook / mind_hook (the drop-in package or the torch-eager restatement)."""
    torch.manual_seed(seed)
    net = StandInUNet(in_channels, num_classes, features)
    net.register_forward_pre_hook(transforms.gin_hook)     # registration order of nnUNetTrainer_GIN_MIND.py:55-57
    net.register_forward_pre_hook(transforms.mind_hook)
    return net


def calc_branch(model, imgs, optimized_idx, with_grad, transforms):
    """tta.py:480-579 for spatial_aug_type='affine', do_spatial_aug_in='both', no intensity augmentation."""
    ctx = torch.enable_grad() if with_grad else torch.no_grad()
    with ctx:
        R, R_inverse = transforms.get_rand_affine(imgs.shape[0], flip=False)
        imgs_aug = transforms.warp(imgs, R, "border")                       # tta.py:549-551
        target = model(imgs_aug)                                            # pre-hooks: gin (off), mind
        target = target.transpose(0, 1)[optimized_idx].transpose(0, 1)      # map_label(..., "logits"), torch_utils.py:214-222
        return transforms.warp(target, R_inverse, "zeros")                  # tta.py:573-575


def tta_inner_step(model, volumes, patch_size, batch_size, optimized_idx, transforms, accum=16, rng=None):
    """One accumulation iteration of tta.py:221-275.  Returns the detached loss (device tensor; the reference syncs
    it to the host every step, tta.py:272 — the caller decides)."""
    import numpy as np
    idx = (rng or np.random).choice(range(len(volumes)), batch_size).tolist()
    with torch.no_grad():
        imgs, _ = transforms.get_batch(volumes, idx, patch_size, fixed_patch_idx=None, device=volumes[0].device)
    imgs = torch.cat(imgs, dim=0)
    target_a = calc_branch(model, imgs, optimized_idx, True, transforms)     # have_grad_in = branch_a
    target_b = calc_branch(model, imgs, optimized_idx, False, transforms)
    fused = getattr(transforms, "consistency_loss", None)
    if fused is not None:                                                    # drop-in: mask + softmaxes + Dice sums in one pass
        loss = fused(target_a, target_b, 1)
    else:                                                                    # the reference's chain, tta.py:263-269
        mask = (target_a.sum(1, keepdim=True) > 0.0).float() * (target_b.sum(1, keepdim=True) > 0.0).float()
        sm_a = target_a.softmax(1) * mask
        sm_b = target_b.softmax(1) * mask
        loss = 1 - soft_dice_loss(sm_a, sm_b)[:, 1:].mean()                  # START_CLASS = 1
    (loss / accum).backward()
    return loss.detach()


class DropInTransforms:
    """The B200 drop-in (dg_tta_b200)."""

    def __init__(self):
        from dg_tta_b200 import gin, mind, utils
        from dg_tta_b200.tta import augmentation_utils as au
        from dg_tta_b200.tta import torch_utils as tu
        utils.disable_internal_augmentation()           # tta.py:154
        self.gin_hook, self.mind_hook = gin.gin_hook, mind.mind_hook
        self.get_rand_affine, self.get_batch = au.get_rand_affine, tu.get_batch
        self._sample = au.affine_grid_sample
        self.consistency_loss = tu.consistency_dice_loss

    def warp(self, x, theta, padding):
        return self._sample(x, theta, padding_mode=padding)
