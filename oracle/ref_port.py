"""TEST INFRASTRUCTURE ONLY — torch-CPU port of the reference's op sequence for the hot path.

Why a second oracle next to the plain-C one (oracle/dgtta_oracle.c)?  The reference's arithmetic
lives in PyTorch ops; this port issues the *same ATen ops in the same order*
(F.pad/conv3d for MIND, grouped conv3d for GIN, affine_grid + grid_sample for the view warp), so
timing it on the GPU box's host cores is the closest thing to "the reference's torch CPU path"
that can travel (the reference tree itself is absent there).  bench.py uses it for `cpu_baseline`
(kind "port") and for `--impl reference`; tests use it as a second checker.  It is pinned against
the golden fixtures (tests/test_oracle_golden.py), bit-for-bit where the op order is identical.

All random draws are arguments — the reference's draw order is restated in
dg_tta_b200/gin.py (product, host side) and traced in tests/golden/make_golden.py.
"""
import torch
import torch.nn.functional as F

# dg_tta/mind.py:104-136 — offsets (d,h,w) of the one-hot kernels, checked against
# tests/golden/mind_shift_table.npz in tests/test_oracle_golden.py
SHIFT1 = [(0, 0, -1), (0, -1, 0), (0, -1, 0), (0, 0, 1), (0, 0, 1), (1, 0, 0),
          (1, 0, 0), (1, 0, 0), (0, 1, 0), (0, 1, 0), (0, 1, 0), (0, 1, 0)]
SHIFT2 = [(-1, 0, 0), (-1, 0, 0), (0, 0, -1), (-1, 0, 0), (0, -1, 0), (0, 0, -1),
          (0, -1, 0), (0, 0, 1), (-1, 0, 0), (0, 0, -1), (0, 0, 1), (1, 0, 0)]


def _one_hot_kernels(table):
    k = torch.zeros(12, 1, 3, 3, 3)
    for c, (d, h, w) in enumerate(table):
        k[c, 0, d + 1, h + 1, w + 1] = 1.0
    return k


def gaussian_taps(sigma, device="cpu"):
    """dg_tta/mind.py:27-37 (same tensor ops, so an int sigma stays an int64 tensor as there)."""
    s = torch.tensor([sigma], device=device)
    n = int(torch.ceil(s * 3.0 / 2.0).long().item()) * 2 + 1
    w = torch.exp(-torch.pow(torch.linspace(-(n // 2), n // 2, n, device=device), 2) / (2 * torch.pow(s, 2)))
    return w / w.sum()


def _blur_axis(vol, taps, axis):
    """dg_tta/mind.py:5-24: replicate-pad one axis, 1-D cross-correlation along it."""
    n = taps.numel()
    pad = [0] * 6
    pad[4 - 2 * axis] = pad[5 - 2 * axis] = n // 2
    shape = [1, 1, 1, 1, 1]
    shape[axis + 2] = n
    b, c, d, h, w = vol.shape
    flat = vol.reshape(b * c, 1, d, h, w)
    return F.conv3d(F.pad(flat, pad, mode="replicate"), taps.view(shape)).view(b, c, d, h, w)


def mind_ssc(img, delta=1, sigma=1, randn_weighting=0.05, noise=None):
    """dg_tta/mind.py:142-164.  noise: the [B,12,D,H,W] tensor drawn at :150, or None to skip it."""
    k1, k2 = _one_hot_kernels(SHIFT1).to(img.device), _one_hot_kernels(SHIFT2).to(img.device)
    padded = F.pad(img, [delta] * 6, mode="replicate")
    edge = F.conv3d(padded, k1, dilation=delta) - F.conv3d(padded, k2, dilation=delta)
    if noise is not None:
        edge = edge + randn_weighting * noise
    taps = gaussian_taps(sigma, img.device)
    ssd = edge ** 2
    for axis in range(3):
        ssd = _blur_axis(ssd, taps, axis)
    mind = ssd - ssd.min(1, keepdim=True)[0]
    var = mind.mean(1, keepdim=True)
    var = torch.clamp(var, var.mean() * 0.001, var.mean() * 1000)
    return torch.exp(-(mind / var))


def gin(x, kers, shifts, alphas):
    """dg_tta/gin.py:94-113 per layer, :197-228 blend + Frobenius re-normalisation."""
    b, c = x.shape[:2]
    spatial = x.shape[2:]
    cur = x
    for layer, (ker, shift) in enumerate(zip(kers, shifts)):
        k = ker.shape[-1]
        cout = ker.shape[0] // b
        y = F.conv3d(cur.reshape(1, -1, *spatial), ker, stride=1, padding=k // 2, dilation=1, groups=b)
        y = y + shift.reshape(-1, 1, 1, 1)
        if layer != len(kers) - 1:
            y = F.leaky_relu(y)
        cur = y.reshape(b, cout, *spatial)
    a = alphas.reshape(b, 1, 1, 1, 1)
    mixed = a * cur + (1.0 - a) * x
    in_frob = torch.norm(x.reshape(b, c, -1), dim=(-1, -2), p="fro").reshape(b, 1, 1, 1, 1)
    self_frob = torch.norm(mixed.reshape(b, c, -1), dim=(-1, -2), p="fro").reshape(b, 1, 1, 1, 1)
    return mixed * (1.0 / (self_frob + 1e-5)) * in_frob


def affine_sample(src, theta, out_size, mode="bilinear", padding_mode="zeros"):
    """affine_grid + grid_sample pair of tta.py:524-551,571-575 / torch_utils.py:55-73."""
    grid = F.affine_grid(theta, list(out_size), align_corners=False)
    return F.grid_sample(src, grid, mode=mode, padding_mode=padding_mode, align_corners=False)


def tta_view_warp(src, theta, identity_grid, padding_mode):
    """The reference's exact grid arithmetic in calc_branch (tta.py:505,523-526,548-551):
    grid = 0*I + (affine_grid(R) - I) + I before sampling."""
    grid = 0.0 * identity_grid + (F.affine_grid(theta, list(src.shape[:1]) + [1] + list(src.shape[2:]),
                                                align_corners=False) - identity_grid)
    grid = grid + identity_grid
    return F.grid_sample(src, grid, padding_mode=padding_mode, align_corners=False)


def gin_mind(x, kers, shifts, alphas, noise=None, **mind_kw):
    """dg_tta/tta/augmentation_utils.py:173-174."""
    return mind_ssc(gin(x, kers, shifts, alphas), noise=noise, **mind_kw)


def argmaxed_segs(segs):
    """dg_tta/tta/torch_utils.py:79-82 (get_argmaxed_segs): background channel where no label is set, then argmax."""
    with_bg = torch.cat([(segs.sum(1, keepdim=True) < 1.0).float(), segs], dim=1)
    return with_bg.argmax(1, keepdim=True)


def consistency_loss(target_a, target_b, start_class=1):
    """dg_tta/tta/tta.py:263-269 with soft_dice_loss of dg_tta/tta/torch_utils.py:90-104 written out."""
    mask = (target_a.sum(1, keepdim=True) > 0.0).float() * (target_b.sum(1, keepdim=True) > 0.0).float()
    sm_a = target_a.softmax(1) * mask
    sm_b = target_b.softmax(1) * mask
    B, _, D, H, W = sm_a.shape
    nominator = (2.0 * sm_a * sm_b).reshape(B, -1, D * H * W).mean(2)
    denominator = 0.5 * ((sm_a + sm_b) ** 2).reshape(B, -1, D * H * W).mean(2)
    dice = (nominator * 0.0) + 1.0 if denominator.sum() == 0.0 else nominator / denominator
    return 1 - dice[:, start_class:].mean()
