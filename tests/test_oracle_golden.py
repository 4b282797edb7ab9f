"""Pin both oracles (plain-C closed form, torch-CPU op port) against the golden fixtures that were
produced by running the unmodified reference (tests/golden/make_golden.py).  CPU only."""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, gin_layers, load_golden
from oracle import cform, ref_port

MIND_CASES = ["a", "b", "c", "d", "e", "f", "g", "h"]
GIN_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(str(GOLDEN / "gin_k*.npz"))) + \
    ["gin_odd", "gin_b3", "gin_thin"]
TOL = 1e-5  # north_star: max-abs-err <= 1e-5 (MIND output lies in (0,1])


def _sigma(g):
    return int(g["sigma"]) if bool(g["sigma_is_int"]) else float(g["sigma"])


def test_shift_table_matches_reference_kernels():
    g = load_golden("mind_shift_table")
    s1, s2 = cform.shift_table()
    assert np.array_equal(s1, g["shift1"]) and np.array_equal(s2, g["shift2"])
    assert np.array_equal(np.array(ref_port.SHIFT1), g["shift1"])
    assert np.array_equal(np.array(ref_port.SHIFT2), g["shift2"])


def test_gaussian_taps():
    t = cform.gaussian_taps(1)
    assert np.allclose(t, [0.05448869, 0.24420136, 0.40261996, 0.24420136, 0.05448869], atol=1e-8)
    for s in (1, 0.5, 2, 1.7):
        assert np.allclose(cform.gaussian_taps(s), ref_port.gaussian_taps(s).numpy(), atol=1e-7)
    assert len(cform.gaussian_taps(0.5)) == 3 and len(cform.gaussian_taps(2)) == 7


@pytest.mark.parametrize("tag", MIND_CASES)
def test_mind_c_oracle(tag):
    g = load_golden(f"mind_{tag}")
    kw = dict(delta=int(g["delta"]), sigma=_sigma(g))
    clean = cform.mind_ssc(g["x"], **kw, noise=None)
    assert np.abs(clean - g["out_clean"]).max() <= TOL
    noisy = cform.mind_ssc(g["x"], **kw, noise=g["noise"], randn_weighting=float(g["randn_weighting"]))
    assert np.abs(noisy - g["out_noisy"]).max() <= TOL
    # the fp64 truth brackets both: reference-vs-truth and oracle-vs-truth stay under the tolerance
    truth = cform.mind_ssc(g["x"], **kw, noise=None, precision="f64")
    assert np.abs(truth - g["out_clean"]).max() <= TOL
    assert np.abs(truth - clean).max() <= TOL


@pytest.mark.parametrize("tag", MIND_CASES)
def test_mind_torch_port(tag):
    g = load_golden(f"mind_{tag}")
    x = torch.from_numpy(g["x"])
    kw = dict(delta=int(g["delta"]), sigma=_sigma(g))
    clean = ref_port.mind_ssc(x, **kw, noise=None).numpy()
    noisy = ref_port.mind_ssc(x, **kw, noise=torch.from_numpy(g["noise"]),
                              randn_weighting=float(g["randn_weighting"])).numpy()
    # same ATen ops in the same order as the reference -> expected bit-identical; allow 1 ulp-ish
    assert np.abs(clean - g["out_clean"]).max() <= 1e-6
    assert np.abs(noisy - g["out_noisy"]).max() <= 1e-6


def test_mind_constant_image_is_nan_like_reference():
    g = load_golden("mind_const")
    assert np.isnan(g["out_clean"]).all()
    assert np.isnan(cform.mind_ssc(g["x"], noise=None)).all()
    assert torch.isnan(ref_port.mind_ssc(torch.from_numpy(g["x"]))).all()


def test_mind_clamp_active_case():
    g = load_golden("mind_clamp")
    out = cform.mind_ssc(g["x"], noise=None)
    assert np.abs(out - g["out_clean"]).max() <= TOL
    # the lower clamp really is active in this fixture: flat voxels map to exp(-0/lo) = 1
    assert (g["out_clean"] == 1.0).mean() > 0.5


def test_mind_hook_defaults():
    g = load_golden("mind_hook")
    out = cform.mind_ssc(g["x"], delta=1, sigma=1, randn_weighting=0.05, noise=g["noise"])
    assert np.abs(out - g["out"]).max() <= TOL


@pytest.mark.parametrize("name", GIN_CASES)
def test_gin_c_oracle(name):
    g = load_golden(name)
    kers, shifts = gin_layers(g)
    assert [k.shape[-1] for k in kers] == g["ksizes"].tolist()
    out = cform.gin(g["x"], kers, shifts, g["alphas"])
    scale = max(1.0, np.abs(g["out"]).max())
    assert np.abs(out - g["out"]).max() <= TOL * scale
    truth = cform.gin(g["x"], kers, shifts, g["alphas"], precision="f64")
    assert np.abs(truth - g["out"]).max() <= TOL * scale


@pytest.mark.parametrize("name", GIN_CASES)
def test_gin_torch_port(name):
    g = load_golden(name)
    kers, shifts = gin_layers(g)
    out = ref_port.gin(torch.from_numpy(g["x"]), [torch.from_numpy(k) for k in kers],
                       [torch.from_numpy(s) for s in shifts], torch.from_numpy(g["alphas"])).numpy()
    assert np.abs(out - g["out"]).max() <= 1e-6 * max(1.0, np.abs(g["out"]).max())


def test_gin_preserves_input_norm():
    g = load_golden("gin_k3333")
    kers, shifts = gin_layers(g)
    out = cform.gin(g["x"], kers, shifts, g["alphas"], precision="f64")
    for b in range(g["x"].shape[0]):
        assert abs(np.linalg.norm(out[b]) / np.linalg.norm(g["x"][b].astype(np.float64)) - 1) < 1e-4


def test_gin_mind_aug_composition():
    g = load_golden("gin_mind_aug")
    kers, shifts = gin_layers(g)
    mid = cform.gin(g["x"], kers, shifts, g["alphas"])
    out = cform.mind_ssc(mid, noise=g["noise"], randn_weighting=0.05)
    assert np.abs(out - g["out"]).max() <= 5e-5  # two chained fp32 stages; MIND amplifies input ulps
    t = ref_port.gin_mind(torch.from_numpy(g["x"]), [torch.from_numpy(k) for k in kers],
                          [torch.from_numpy(s) for s in shifts], torch.from_numpy(g["alphas"]),
                          noise=torch.from_numpy(g["noise"])).numpy()
    assert np.abs(t - g["out"]).max() <= 1e-6


def test_affine_view_warp():
    g = load_golden("affine_view")
    size = g["imgs"].shape
    # sampler tolerance: the reference adds and subtracts the identity grid (tta.py:523-532,548),
    # which perturbs coordinates by ~1e-7 * size voxels; values are O(1) with O(1) gradients
    img = cform.affine_sample(g["imgs"], g["R"], size, padding_mode="border")
    assert np.abs(img - g["imgs_aug"]).max() <= 2e-5
    wl = cform.affine_sample(g["logits"], g["R_inv"], g["logits"].shape, padding_mode="zeros")
    assert np.abs(wl - g["warped"]).max() <= 5e-5
    gi = cform.affine_sample_bwd_input(g["grad_out"], g["R_inv"], g["logits"].shape, padding_mode="zeros")
    assert np.abs(gi - g["grad_logits"]).max() <= 5e-5


def test_affine_general_modes():
    g = load_golden("affine_general")
    for mode in ("bilinear", "nearest"):
        for pad in ("zeros", "border"):
            out = cform.affine_sample(g["src"], g["theta"], g["out_size"], mode=mode, padding_mode=pad)
            ref = g[f"{mode}_{pad}"]
            # the oracle executes torch's coordinate arithmetic operation by operation (fma linspace, mul, div, fma-chained
            # bmm) and its corner order: index work is bit-exact, and so is the trilinear value on this fixture
            assert np.array_equal(out, ref)
            t = ref_port.affine_sample(torch.from_numpy(g["src"]), torch.from_numpy(g["theta"]),
                                       g["out_size"].tolist(), mode=mode, padding_mode=pad).numpy()
            assert np.array_equal(t, ref)
    for pad in ("zeros", "border"):
        gi = cform.affine_sample_bwd_input(g["grad_out"], g["theta"], g["src"].shape, padding_mode=pad)
        assert np.abs(gi - g[f"grad_src_{pad}"]).max() <= 5e-5


def test_consistency_loss_port_matches_reference_fixture():
    """oracle/ref_port.consistency_loss against the fixture produced with the reference's soft_dice_loss"""
    import torch
    from oracle import ref_port
    g = load_golden("consistency")
    ta = torch.from_numpy(g["target_a"]).requires_grad_(True)
    loss = ref_port.consistency_loss(ta, torch.from_numpy(g["target_b"]))
    assert abs(loss.item() - float(g["loss"])) <= 1e-6
    loss.backward()
    assert np.abs(ta.grad.numpy() - g["grad_a"]).max() <= 1e-7


def test_argmaxed_segs_port_matches_reference_fixture():
    import torch
    from oracle import ref_port
    g = load_golden("argmaxed_segs")
    assert np.array_equal(ref_port.argmaxed_segs(torch.from_numpy(g["segs"])).numpy(), g["out"])


def test_consistency_loss_c_oracle_matches_reference_fixture():
    g = load_golden("consistency")
    scale = np.abs(g["grad_a"]).max()
    for precision in ("f32", "f64"):                       # the fixture itself is fp32 torch
        loss, grad = cform.consistency_loss(g["target_a"], g["target_b"], precision=precision)
        assert abs(loss - float(g["loss"])) <= 2e-6
        assert np.abs(grad - g["grad_a"]).max() <= 1e-5 * scale


def test_label_argmax_c_oracle_matches_reference_fixture():
    g = load_golden("argmaxed_segs")
    theta = np.eye(3, 4, dtype=np.float32)[None]
    assert np.array_equal(cform.label_argmax(g["segs"], theta), g["out"])
    # and the reference crops of get_batch (nearest sampling + argmax through a real patch affine)
    gb = load_golden("get_batch")
    sample = gb["sample"]
    onehot = sample[1:][None]
    out = cform.label_argmax(onehot, theta, gb["lbl_c"].shape)   # centre crop = scale-only affine
    import torch
    t_patch = torch.as_tensor(gb["patch"].tolist(), dtype=torch.float32)
    t_in = torch.as_tensor(sample.shape[-3:], dtype=torch.float32)
    aff = torch.cat([(t_patch / t_in).flip(0), torch.tensor([1.0])]).diag()[:3][None].numpy()
    out = cform.label_argmax(onehot, aff, gb["lbl_c"].shape)
    assert np.array_equal(out, gb["lbl_c"])


def _reference_crops(g, name):
    """thetas of the reference's get_batch call behind fixture case `name` (same CPU draws), via the package's host code"""
    import torch
    from dg_tta_b200.tta.torch_utils import patch_affines
    seed = int(g[f"{name}_seed"])
    if seed >= 0:
        torch.manual_seed(seed)
    return patch_affines(g["sample"].shape[-3:], g[f"{name}_patch"].tolist(), 2, "center" if seed < 0 else None).numpy()


def test_nearest_ties_are_bit_exact_in_the_oracle():
    """Index work: label crops whose coordinates sit exactly on .5 ties (odd patch out of an even volume), random and
    up-sampling crops, and general nearest-mode affines — every voxel equal to the reference's output."""
    g = load_golden("nearest_ties")
    onehot = g["sample"][1:][None]
    for name in ("tie_center", "rand_a", "rand_b", "up", "same"):
        thetas = _reference_crops(g, name)
        for i in range(2):
            out = cform.label_argmax(onehot, thetas[i:i + 1], g[f"{name}_lbl{i}"].shape)
            assert np.array_equal(out, g[f"{name}_lbl{i}"]), name
            img = cform.affine_sample(g["sample"][:1][None] - g["sample"][0].min(), thetas[i:i + 1], g[f"{name}_img{i}"].shape)
            assert np.abs(img + g["sample"][0].min() - g[f"{name}_img{i}"]).max() <= 1e-6, name
    src = np.arange(np.prod(g["src_shape"]), dtype=np.float32).reshape(g["src_shape"])
    for pad in ("zeros", "border"):
        out = cform.affine_sample(src, g["theta"], g["out_size"], mode="nearest", padding_mode=pad)
        assert np.array_equal(out, g[f"nearest_{pad}"])
