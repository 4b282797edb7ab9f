#!/usr/bin/env python3
"""Developer probe: where the HOST time of one gin_mind_aug call goes (tiny volume, so the GPU is never the limit)."""
import cProfile
import pstats
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from dg_tta_b200.tta.augmentation_utils import gin_mind_aug  # noqa: E402

x = torch.randn(2, 1, 19, 19, 21, device="cuda")
for i in range(20):
    torch.manual_seed(i); gin_mind_aug(x)
torch.cuda.synchronize()
t = time.perf_counter()
for i in range(200):
    torch.manual_seed(i); gin_mind_aug(x)
torch.cuda.synchronize()
print("ms per call", (time.perf_counter() - t) / 200 * 1e3)
pr = cProfile.Profile()
pr.enable()
for i in range(200):
    torch.manual_seed(i); gin_mind_aug(x)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
