"""Helpers shared by the -m gpu parity tests."""
import numpy as np
import torch


def cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def synth_volume(shape, seed, modality="ct"):
    """SURVEY.md §8d synthetic CT/MR-like volume: low-frequency field + ellipsoid 'organs' + fine noise,
    normalised like nnU-Net's CTNormalization output (roughly [-1.9, 2.8])."""
    g = torch.Generator().manual_seed(seed)
    B, C, D, H, W = shape
    low = torch.randn(B, C, -(-D // 16) + 1, -(-H // 16) + 1, -(-W // 16) + 1, generator=g)
    x = torch.nn.functional.interpolate(low, size=(D, H, W), mode="trilinear", align_corners=True) * 0.4
    zz, yy, xx = torch.meshgrid(torch.linspace(-1, 1, D), torch.linspace(-1, 1, H), torch.linspace(-1, 1, W),
                                indexing="ij")
    for _ in range(12):
        c = torch.rand(3, generator=g) * 1.6 - 0.8
        r = torch.rand(3, generator=g) * 0.35 + 0.08
        val = float(torch.randn(1, generator=g)) * 0.9
        mask = ((zz - c[0]) / r[0]) ** 2 + ((yy - c[1]) / r[1]) ** 2 + ((xx - c[2]) / r[2]) ** 2 < 1
        x = x + mask.float() * val
    x = x + 0.05 * torch.randn(shape, generator=g)
    if modality == "mr":
        bias = 1 + 0.3 * torch.nn.functional.interpolate(
            torch.randn(B, C, 3, 3, 3, generator=g), size=(D, H, W), mode="trilinear", align_corners=True)
        x = x * bias
        x = (x - x.mean()) / x.std()
    return x.clamp(-1.85, 2.75).contiguous()
