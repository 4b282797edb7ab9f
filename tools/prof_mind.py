#!/usr/bin/env python3
"""Profiling driver: a few launches of one op so that ncu can capture it (never a bench number)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests"))
from gpu_util import synth_volume  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "mind"
shape = (2, 1, 192, 192, 192)
x = synth_volume(shape, 1).cuda()
if what == "mind":
    from dg_tta_b200 import mind_ssc
    for _ in range(3):
        mind_ssc(x, noise=False)
elif what == "mind_noise":
    from dg_tta_b200 import mind_ssc
    n = torch.randn(2, 12, 192, 192, 192, device="cuda")
    for _ in range(3):
        mind_ssc(x, noise=n)
elif what == "philox":
    from dg_tta_b200.mind import randn_like_reference
    for _ in range(3):
        randn_like_reference((2, 12, 192, 192, 192), "cuda")
elif what.startswith("gin"):
    from dg_tta_b200.gin import GINGroupConv, gin_forward
    want = [int(c) for c in what[3:]] if len(what) > 3 else [3, 3, 3, 3]
    net = GINGroupConv(dict(IN_CHANNELS=1, N_LAYER=4, INTERM_CHANNELS=2))
    seed = 0
    while True:
        torch.manual_seed(seed)
        alphas, kers, shifts = net.draw(x)
        if [k.shape[-1] for k in kers] == want:
            break
        seed += 1
    for _ in range(3):
        gin_forward(x, kers, shifts, alphas, 2)
elif what == "sampler":
    from dg_tta_b200.tta.augmentation_utils import affine_grid_sample, get_rand_affine
    torch.manual_seed(0)
    R, Ri = get_rand_affine(2)
    img = synth_volume((2, 1, 128, 128, 128), 3).cuda()
    affine_grid_sample(img, R, padding_mode="border")
    lg = torch.randn(2, 14, 128, 128, 128, device="cuda", requires_grad=True)
    out = affine_grid_sample(lg, Ri)
    out.backward(torch.randn_like(out))
torch.cuda.synchronize()
