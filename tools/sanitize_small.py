#!/usr/bin/env python3
"""Small-shape pass through every kernel family for compute-sanitizer (memcheck / racecheck): MIND (TMA + LDG paths, noise,
fix-up pass), GIN stack (single / double segments, pointwise), Philox, sampler fwd / bwd / nearest / labels, get_batch,
consistency loss (plain and fused-warp), resize.  Results are compared loosely against the unfused / eager paths so that
the run also fails on wrong answers."""
import os
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests"))
from dg_tta_b200 import MIND3D, mind_ssc  # noqa: E402
from dg_tta_b200.gin import gin_aug  # noqa: E402
from dg_tta_b200.pretraining import resize_edge  # noqa: E402
from dg_tta_b200.tta.augmentation_utils import affine_grid_sample, get_rand_affine, gin_mind_aug  # noqa: E402
from dg_tta_b200.tta.torch_utils import consistency_dice_loss, consistency_dice_loss_warped, get_batch  # noqa: E402
from gpu_util import synth_volume  # noqa: E402


def main():
    torch.manual_seed(0)
    for shape in ((1, 1, 12, 20, 36), (2, 1, 9, 17, 23)):           # W % 4 == 0 -> TMA path; ragged -> LDG path
        x = synth_volume(shape, 3).cuda()
        a = MIND3D()(x)
        b = mind_ssc(x, noise=False)
        flat = torch.zeros_like(x) + 0.5
        flat[..., :4] += x[..., :4] * 50                              # huge variance contrast: the clamp fix-up pass runs
        c = mind_ssc(flat.contiguous(), noise=False)
        assert torch.isfinite(a).all() and torch.isfinite(b).all() and tuple(c.shape) == tuple(a.shape)
        for seed in range(8):                                         # many kernel-size patterns of the GIN stack
            torch.manual_seed(seed)
            g = gin_aug(x)
            assert torch.isfinite(g).all()
        d = gin_mind_aug(x)
        assert torch.isfinite(d).all()
    R, Ri = get_rand_affine(2, strength=0.1)
    lg = torch.randn(2, 5, 10, 12, 20, device="cuda").abs().requires_grad_(True)
    lb = torch.randn(2, 5, 10, 12, 20, device="cuda").abs()
    w = affine_grid_sample(lg, Ri)
    l1 = consistency_dice_loss(w, affine_grid_sample(lb, R))
    l2 = consistency_dice_loss_warped(lg, lb, Ri, R)
    (g1,) = torch.autograd.grad(l1, lg)
    (g2,) = torch.autograd.grad(l2, lg)
    assert abs(float(l1) - float(l2)) < 1e-5 and float((g1 - g2).abs().max()) <= 1e-4 * float(g1.abs().max())
    affine_grid_sample(lb, R, (7, 9, 11), mode="nearest", padding_mode="border")
    vol = synth_volume((1, 1, 14, 18, 22), 5)[0]
    lab = torch.zeros(3, 14, 18, 22)
    lab[1, 2:9, 3:12, 4:15] = 1
    b_img, b_lbl = get_batch([torch.cat([vol, lab], 0)], [0, 0], [8, 10, 12], device="cuda")
    assert b_lbl[0].dtype == torch.int64
    up = resize_edge(resize_edge(vol.cuda(), (7, 5, 6), 0), (14, 18, 22), 3)
    assert torch.isfinite(up).all()
    torch.cuda.synchronize()
    print("sanitize_small ok")


if __name__ == "__main__":
    main()
