#!/usr/bin/env python3
"""Developer probe: one small MIND call with a noise tensor (TMA-staged path), synchronised."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from dg_tta_b200 import mind_ssc
shape = tuple(int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (24, 32, 64)
x = torch.randn((1, 1) + shape, device="cuda")
n = torch.randn((1, 12) + shape, device="cuda")
torch.cuda.synchronize()
out = mind_ssc(x, noise=n)
torch.cuda.synchronize()
print("ok", float(out.mean()))
