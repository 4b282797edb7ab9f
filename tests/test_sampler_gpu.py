"""Affine sampler parity (forward modes, adjoint) vs golden fixtures (torch affine_grid+grid_sample run by
the reference's op sequence) and the oracle.  Tolerance 5e-5*max|x|: the reference itself adds and
subtracts the identity grid (tta.py:523-532,548), perturbing coordinates by ~1e-7*size voxels."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from gpu_util import cuda, synth_volume

pytestmark = pytest.mark.gpu


def test_tta_view_warp_forward_and_backward():
    from dg_tta_b200.tta.augmentation_utils import affine_grid_sample
    g = load_golden("affine_view")
    img = affine_grid_sample(cuda(g["imgs"]), torch.from_numpy(g["R"]), padding_mode="border")
    assert np.abs(img.cpu().numpy() - g["imgs_aug"]).max() <= 2e-5
    lg = cuda(g["logits"]).requires_grad_(True)
    warped = affine_grid_sample(lg, torch.from_numpy(g["R_inv"]))
    assert np.abs(warped.detach().cpu().numpy() - g["warped"]).max() <= 5e-5
    (warped * cuda(g["grad_out"])).sum().backward()
    assert np.abs(lg.grad.cpu().numpy() - g["grad_logits"]).max() <= 5e-5


def test_general_modes_and_sizes():
    from dg_tta_b200.tta.augmentation_utils import affine_grid_sample
    g = load_golden("affine_general")
    src, theta, size = cuda(g["src"]), torch.from_numpy(g["theta"]), g["out_size"].tolist()
    for mode in ("bilinear", "nearest"):
        for pad in ("zeros", "border"):
            out = affine_grid_sample(src, theta, size, mode=mode, padding_mode=pad).cpu().numpy()
            ref = g[f"{mode}_{pad}"]
            if mode == "nearest":
                assert np.array_equal(out, ref)          # index work: bit-exact
            else:
                assert np.abs(out - ref).max() <= 2e-6   # same coordinates as torch; only the corner summation order differs
    for pad in ("zeros", "border"):
        s = src.clone().requires_grad_(True)
        out = affine_grid_sample(s, theta, size, padding_mode=pad)
        (out * cuda(g["grad_out"])).sum().backward()
        assert np.abs(s.grad.cpu().numpy() - g[f"grad_src_{pad}"]).max() <= 5e-5


def test_identity_is_exact_and_adjoint_property():
    from dg_tta_b200.tta.augmentation_utils import affine_grid_sample
    x = synth_volume((2, 3, 17, 18, 19), 4).cuda()
    eye = torch.eye(3, 4)[None].repeat(2, 1, 1)
    y = affine_grid_sample(x, eye, padding_mode="border")
    assert (y - x).abs().max() <= 2e-6
    # <A x, g> == <x, A^T g>
    torch.manual_seed(0)
    theta = eye + 0.1 * torch.randn(2, 3, 4)
    g = torch.randn_like(x)
    xr = x.clone().requires_grad_(True)
    out = affine_grid_sample(xr, theta)
    (out * g).sum().backward()
    lhs = float((out.detach().double() * g.double()).sum())
    rhs = float((x.double() * xr.grad.double()).sum())
    assert abs(lhs - rhs) <= 1e-4 * max(1.0, abs(lhs))


def test_against_oracle_full_patch():
    """config-3 shapes (2x1x128^3 image, 2x14x128^3 logits is too big for the CPU oracle in seconds ->
    2x4 channels) through the C oracle"""
    from dg_tta_b200.tta.augmentation_utils import affine_grid_sample, get_rand_affine
    from oracle import cform
    torch.manual_seed(7)
    R, Ri = get_rand_affine(2)
    x = synth_volume((2, 1, 128, 128, 128), 31)
    out = affine_grid_sample(x.cuda(), R, padding_mode="border").cpu().numpy()
    ref = cform.affine_sample(x.numpy(), R.numpy(), x.shape, padding_mode="border")
    assert np.abs(out - ref).max() <= 2e-5 * np.abs(ref).max()
    lg = synth_volume((2, 4, 64, 72, 80), 32)
    out = affine_grid_sample(lg.cuda(), Ri).cpu().numpy()
    ref = cform.affine_sample(lg.numpy(), Ri.numpy(), lg.shape)
    assert np.abs(out - ref).max() <= 2e-5 * np.abs(ref).max()
    go = synth_volume((2, 4, 64, 72, 80), 33)
    s = lg.cuda().requires_grad_(True)
    (affine_grid_sample(s, Ri) * go.cuda()).sum().backward()
    gref = cform.affine_sample_bwd_input(go.numpy(), Ri.numpy(), lg.shape)
    assert np.abs(s.grad.cpu().numpy() - gref).max() <= 5e-5 * max(1.0, np.abs(gref).max())


@pytest.mark.parametrize("kind", ["tta", "shrink", "magnify", "flip", "many_channels"])
def test_deterministic_gather_backward(kind, monkeypatch):
    """DGTTA_SAMPLE_BWD_DETERMINISTIC=1: the adjoint as a gather over the source voxels (no atomics).  Same gradient as the
    scatter up to summation order and as the C oracle's adjoint, bit-identical from run to run; an affine that magnifies
    strongly makes the candidate boxes large and the call falls back to the scatter on the device (still correct)."""
    from dg_tta_b200.tta.augmentation_utils import affine_grid_sample, get_rand_affine
    from oracle import cform
    torch.manual_seed(11)
    _, Ri = get_rand_affine(2)
    if kind == "shrink":      # output covers a sub-region: source voxels outside receive nothing, inside ~(1/0.6)^3 outputs each
        Ri = Ri.clone(); Ri[:, :, :3] *= 0.6
    elif kind == "magnify":   # each source voxel is hit by ~(1/3)^-3 = 27x more outputs: candidate box > 6 -> scatter fallback
        Ri = Ri.clone(); Ri[:, :, :3] *= 3.0
    elif kind == "flip":
        Ri = Ri.clone(); Ri[:, 0] *= -1.0
    C = 19 if kind == "many_channels" else 5     # more than the 16 channel accumulators of one pass
    lg = synth_volume((2, C, 33, 40, 52), 41)
    go = synth_volume((2, C, 30, 44, 48), 42)

    def grad():
        s = lg.cuda().requires_grad_(True)
        (affine_grid_sample(s, Ri, out_size=go.shape[-3:]) * go.cuda()).sum().backward()
        return s.grad.clone()

    monkeypatch.delenv("DGTTA_SAMPLE_BWD_DETERMINISTIC", raising=False)
    g_scatter = grad()
    monkeypatch.setenv("DGTTA_SAMPLE_BWD_DETERMINISTIC", "1")
    g1, g2 = grad(), grad()
    assert torch.equal(g1, g2)
    scale = max(1.0, float(g_scatter.abs().max()))
    assert float((g1 - g_scatter).abs().max()) <= 2e-5 * scale
    gref = cform.affine_sample_bwd_input(go.numpy(), Ri.numpy(), lg.shape)
    assert np.abs(g1.cpu().numpy() - gref).max() <= 5e-5 * max(1.0, np.abs(gref).max())


def test_gin_mind_aug_fused_chain():
    from dg_tta_b200.gin import gin_forward
    from dg_tta_b200 import mind_ssc
    from conftest import gin_layers
    g = load_golden("gin_mind_aug")
    kers, shifts = gin_layers(g)
    mixed, scale = gin_forward(cuda(g["x"]), [torch.from_numpy(k) for k in kers], [torch.from_numpy(s) for s in shifts],
                               cuda(g["alphas"]), 2, defer_scale=True)
    out = mind_ssc(mixed, noise=cuda(g["noise"]), in_scale=scale).cpu().numpy()
    assert np.abs(out - g["out"]).max() <= 5e-5   # chained fp32 stages; see tests/test_oracle_golden.py
    from dg_tta_b200.tta.augmentation_utils import gin_mind_aug
    res = gin_mind_aug(cuda(g["x"]))
    assert tuple(res.shape) == tuple(g["out"].shape) and not torch.isnan(res).any()


def test_get_batch_matches_reference_crops():
    """torch_utils.get_batch (dg_tta/tta/torch_utils.py:13-76): random, centre and larger-than-volume crops."""
    from dg_tta_b200.tta.torch_utils import get_batch
    g = load_golden("get_batch")
    sample = torch.from_numpy(g["sample"])
    patch = g["patch"].tolist()
    torch.manual_seed(int(g["seed"]))
    b_img, b_lbl = get_batch([sample], [0, 0], patch, fixed_patch_idx=None, device="cuda")
    for got, ref in ((b_img[0], g["img0"]), (b_img[1], g["img1"])):
        assert np.abs(got.cpu().numpy() - ref).max() <= 2e-5
    for got, ref in ((b_lbl[0], g["lbl0"]), (b_lbl[1], g["lbl1"])):
        assert got.dtype == torch.int64 and np.array_equal(got.cpu().numpy(), ref)
    c_img, c_lbl = get_batch([sample.cuda()], [0], patch, fixed_patch_idx="center", device="cuda")
    assert np.abs(c_img[0].cpu().numpy() - g["img_c"]).max() <= 2e-5
    assert np.array_equal(c_lbl[0].cpu().numpy(), g["lbl_c"])
    torch.manual_seed(int(g["seed_large"]))
    l_img, l_lbl = get_batch([sample], [0], g["patch_large"].tolist(), device="cuda")
    assert np.abs(l_img[0].cpu().numpy() - g["img_l"]).max() <= 2e-5
    assert np.array_equal(l_lbl[0].cpu().numpy(), g["lbl_l"])
    img_only, none_lbl = get_batch([sample[:1]], [0], patch, fixed_patch_idx="center", device="cuda")
    assert none_lbl[0] is None and tuple(img_only[0].shape) == (1, 1, *patch)


def test_nearest_ties_are_bit_exact():
    """tests/golden/nearest_ties.npz (reference get_batch / grid_sample(mode="nearest") on tie-heavy geometry): every
    label voxel and every nearest-sampled value equals the reference's, through get_batch (int16 label-map path), the
    one-hot label kernel and the plain nearest sampler."""
    from dg_tta_b200.tta.augmentation_utils import affine_grid_sample, affine_label_argmax
    from dg_tta_b200.tta.torch_utils import get_batch, patch_affines
    g = load_golden("nearest_ties")
    sample = torch.from_numpy(g["sample"])
    for name in ("tie_center", "rand_a", "rand_b", "up", "same"):
        seed, patch = int(g[f"{name}_seed"]), g[f"{name}_patch"].tolist()
        fixed = "center" if seed < 0 else None
        if seed >= 0:
            torch.manual_seed(seed)
        b_img, b_lbl = get_batch([sample], [0, 0], patch, fixed_patch_idx=fixed, device="cuda")
        if seed >= 0:
            torch.manual_seed(seed)
        thetas = patch_affines(sample.shape[-3:], patch, 2, fixed)
        for i in range(2):
            assert np.array_equal(b_lbl[i].cpu().numpy(), g[f"{name}_lbl{i}"]), name
            assert np.abs(b_img[i].cpu().numpy() - g[f"{name}_img{i}"]).max() <= 2e-6, name
            onehot_path = affine_label_argmax(sample[1:][None].cuda(), thetas[i:i + 1], patch)
            assert np.array_equal(onehot_path.cpu().numpy(), g[f"{name}_lbl{i}"]), name
    src = torch.arange(int(np.prod(g["src_shape"])), dtype=torch.float32).view(*g["src_shape"].tolist()).cuda()
    for pad in ("zeros", "border"):
        out = affine_grid_sample(src, torch.from_numpy(g["theta"]), g["out_size"].tolist(), mode="nearest", padding_mode=pad)
        assert np.array_equal(out.cpu().numpy(), g[f"nearest_{pad}"])


def test_label_argmax_equals_the_unfused_chain():
    """dgtta_affine_label_argmax == get_argmaxed_segs(nearest grid_sample of the one-hot channels) (torch_utils.py:71-82),
    bit for bit: same coordinates, same tie rule; overlapping / empty / fractional label channels included."""
    from dg_tta_b200.tta.augmentation_utils import affine_grid_sample, affine_label_argmax, get_rand_affine
    from dg_tta_b200.tta.torch_utils import get_argmaxed_segs
    g = torch.Generator().manual_seed(3)
    lab = (torch.rand(2, 7, 20, 18, 22, generator=g) > 0.8).float()     # overlapping labels and unlabeled voxels
    lab[:, 5] *= 0.5                                                     # a fractional channel: sum < 1 with a set label
    lab = lab.cuda()
    torch.manual_seed(4)
    R, _ = get_rand_affine(2, strength=0.2)
    for size in (None, (9, 30, 11)):
        unfused = get_argmaxed_segs(affine_grid_sample(lab, R, size, mode="nearest", padding_mode="zeros"))
        fused = affine_label_argmax(lab, R, size)
        assert fused.dtype == torch.int64 and tuple(fused.shape) == tuple(unfused.shape)
        assert torch.equal(fused, unfused)
        # the same through torch's own ops (the reference's chain, torch_utils.py:71-82)
        smp = affine_grid_sample(lab, R, size, mode="nearest", padding_mode="zeros")
        ref = torch.cat([(smp.sum(1, keepdim=True) < 1.0).float(), smp], dim=1).argmax(1, keepdim=True)
        assert torch.equal(fused, ref)
