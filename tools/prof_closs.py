#!/usr/bin/env python3
"""Profiling driver: a few fused consistency-loss forward+backward calls for ncu (never a bench number)."""
import sys, torch
sys.path.insert(0, str(__import__('pathlib').Path(__file__).resolve().parents[1]))
from dg_tta_b200.tta.torch_utils import consistency_dice_loss
ta = torch.randn(2, 14, 128, 128, 128, device="cuda", requires_grad=True)
tb = torch.randn(2, 14, 128, 128, 128, device="cuda")
for _ in range(3):
    torch.autograd.grad(consistency_dice_loss(ta, tb), ta)
torch.cuda.synchronize()
