#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ by RUNNING THE REFERENCE ITSELF.

The reference (multimodallearning/DG-TTA) has no tests and no golden vectors of its own
(SURVEY.md §4, §8c), so parity is pinned on outputs of the reference's own functions,
imported unmodified from /root/reference and executed on torch CPU fp32 in the build
container.  /root/reference does not exist on the GPU box, therefore the vectors are
committed as small .npz fixtures and this script is committed next to them.

Run (build container only):
    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Functions exercised (reference file:line):
    dg_tta/mind.py:98-164          MIND3D.__init__/forward  (noise off, noise injected)
    dg_tta/gin.py:59-122,168-230   GradlessGCReplayNonlinBlock / GINGroupConv.forward
    dg_tta/gin.py:233-247          gin_aug / gin_hook
    dg_tta/tta/augmentation_utils.py:156-174   get_rand_affine / gin_mind_aug
    dg_tta/tta/tta.py:505-551,571-575          affine_grid/grid_sample op sequence (restated
                                               inline here because tta.py needs nnunetv2)
    dg_tta/tta/torch_utils.py:13-76            get_batch
"""
import os
import sys
from pathlib import Path

REF = "/root/reference"
if not Path(REF).is_dir():
    sys.exit("reference tree not present; fixtures can only be regenerated in the build container")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
os.environ["DG_TTA_INTERNAL_AUGMENTATION"] = "true"

import numpy as np
import torch
import torch.nn.functional as F

from dg_tta.mind import MIND3D, mind_hook  # noqa: E402
from dg_tta import gin as ref_gin  # noqa: E402
from dg_tta.tta import augmentation_utils as ref_aug  # noqa: E402
from dg_tta.tta import torch_utils as ref_tu  # noqa: E402

OUT = Path(__file__).resolve().parent
torch.set_num_threads(4)


def volume(shape, seed, kind="smooth"):
    """Small deterministic test volume: low-frequency field + blobs + fine noise."""
    g = torch.Generator().manual_seed(seed)
    B, C, D, H, W = shape
    if kind == "randn":
        return torch.randn(shape, generator=g)
    low = torch.randn(B, C, max(D // 4, 1) + 1, max(H // 4, 1) + 1, max(W // 4, 1) + 1, generator=g)
    x = F.interpolate(low, size=(D, H, W), mode="trilinear", align_corners=True)
    x = x + (x > 0.3).float() * 0.8 + 0.05 * torch.randn(shape, generator=g)
    return x.contiguous()


def save(name, **arrays):
    arrays = {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrays.items()}
    np.savez_compressed(OUT / f"{name}.npz", **arrays)
    print(f"{name}: " + ", ".join(f"{k}{tuple(v.shape)}" for k, v in arrays.items()))


# ----------------------------------------------------------------------------- MIND-SSC
def gen_mind():
    cases = [
        # (tag, shape, delta, sigma, kind)
        ("a", (1, 1, 16, 18, 20), 1, 1, "smooth"),
        ("b", (2, 1, 9, 11, 13), 1, 1, "randn"),
        ("c", (1, 1, 12, 10, 21), 2, 1, "smooth"),
        ("d", (1, 1, 7, 9, 8), 3, 1, "randn"),
        ("e", (1, 1, 3, 4, 5), 1, 1, "randn"),      # smaller than the 2+delta halo
        ("f", (1, 1, 1, 1, 7), 2, 1, "randn"),      # degenerate axes
        ("g", (1, 1, 10, 12, 14), 1, 0.5, "smooth"),  # 3-tap Gaussian (sigma float)
        ("h", (1, 1, 10, 12, 14), 1, 2, "smooth"),    # 7-tap Gaussian
    ]
    for tag, shape, delta, sigma, kind in cases:
        x = volume(shape, 100 + ord(tag), kind)
        with torch.no_grad():
            clean = MIND3D(delta=delta, sigma=sigma, randn_weighting=0.0)(x)
        # injected noise: the reference draws torch.randn_like(edge_selection) (mind.py:150);
        # replace that single draw by a stored tensor so the CPU oracle and the CUDA path
        # can consume the same field.
        g = torch.Generator().manual_seed(7000 + ord(tag))
        noise = torch.randn((shape[0], 12) + tuple(shape[2:]), generator=g)
        orig = torch.randn_like
        torch.randn_like = lambda t, *a, **k: noise.clone()
        try:
            with torch.no_grad():
                noisy = MIND3D(delta=delta, sigma=sigma, randn_weighting=0.05)(x)
        finally:
            torch.randn_like = orig
        save(f"mind_{tag}", x=x, delta=delta, sigma=float(sigma), sigma_is_int=isinstance(sigma, int),
             out_clean=clean, noise=noise, randn_weighting=0.05, out_noisy=noisy)

    # constant image, no noise -> 0/0 -> NaN everywhere in the reference (mind.py:157-162)
    x = torch.full((1, 1, 4, 5, 6), 0.25)
    with torch.no_grad():
        out = MIND3D(randn_weighting=0.0)(x)
    save("mind_const", x=x, delta=1, sigma=1.0, sigma_is_int=True, out_clean=out)

    # clamp active: random-valued block inside a zero volume -> flat voxels have var==0 < 0.001*mean
    # (a lone symmetric spike would be ill-conditioned: equal ssd_c up to rounding -> 0/0-like noise)
    x = torch.zeros(1, 1, 14, 13, 12)
    g = torch.Generator().manual_seed(808)
    x[0, 0, 5:9, 4:9, 6:9] = torch.randn(4, 5, 3, generator=g) * 2.0
    with torch.no_grad():
        out = MIND3D(randn_weighting=0.0)(x)
    save("mind_clamp", x=x, delta=1, sigma=1.0, sigma_is_int=True, out_clean=out)

    # hook form (mind.py:167-168): defaults, noise via patched randn_like
    x = volume((2, 1, 8, 9, 10), 321)
    g = torch.Generator().manual_seed(99)
    noise = torch.randn(2, 12, 8, 9, 10, generator=g)
    orig = torch.randn_like
    torch.randn_like = lambda t, *a, **k: noise.clone()
    try:
        with torch.no_grad():
            out = mind_hook(None, (x,))
    finally:
        torch.randn_like = orig
    save("mind_hook", x=x, noise=noise, out=out)

    # shift table as the reference builds it (mind.py:104-136): offsets (d,h,w) in {-1,0,1}
    m = MIND3D()
    s1 = torch.nonzero(m.mshift1.view(12, 27))[:, 1]
    s2 = torch.nonzero(m.mshift2.view(12, 27))[:, 1]
    tab1 = torch.stack([s1 // 9 - 1, (s1 // 3) % 3 - 1, s1 % 3 - 1], 1)
    tab2 = torch.stack([s2 // 9 - 1, (s2 // 3) % 3 - 1, s2 % 3 - 1], 1)
    save("mind_shift_table", shift1=tab1, shift2=tab2)


# ----------------------------------------------------------------------------- GIN
def traced_gin(x, seed):
    """Run the reference gin_aug under torch.manual_seed(seed) while recording every random
    draw (gin.py:65,94,99,187) so the fixture is independent of torch's RNG streams."""
    rec = {"k": [], "randn": [], "rand": []}
    o_randint, o_randn, o_rand = torch.randint, torch.randn, torch.rand

    def randint(*a, **k):
        r = o_randint(*a, **k); rec["k"].append(int(r[0])); return r

    def randn(*a, **k):
        r = o_randn(*a, **k); rec["randn"].append(r.clone()); return r

    def rand(*a, **k):
        r = o_rand(*a, **k); rec["rand"].append(r.clone()); return r

    torch.manual_seed(seed)
    torch.randint, torch.randn, torch.rand = randint, randn, rand
    try:
        with torch.no_grad():
            out = ref_gin.gin_aug(x)
    finally:
        torch.randint, torch.randn, torch.rand = o_randint, o_randn, o_rand
    assert len(rec["k"]) == 4 and len(rec["randn"]) == 8 and len(rec["rand"]) == 1
    return out, rec


def gen_gin():
    seen = {}
    seed = 0
    x = volume((2, 1, 10, 11, 12), 555)
    # sweep seeds until all 16 kernel-size patterns of the 4 layers are covered
    while len(seen) < 16 and seed < 4000:
        torch.manual_seed(seed)
        torch.rand(2)
        ks = []
        for layer in range(4):
            k = [1, 3][int(torch.randint(high=2, size=(1,))[0])]
            cin = 1 if layer == 0 else 2
            cout = 1 if layer == 3 else 2
            torch.randn([cout * 2, cin, k, k, k]); torch.randn([cout * 2, 1, 1, 1])
            ks.append(k)
        seen.setdefault(tuple(ks), seed)
        seed += 1
    assert len(seen) == 16, seen
    for ks, s in sorted(seen.items()):
        out, rec = traced_gin(x, s)
        assert tuple([1, 3][i] for i in rec["k"]) == ks
        arrays = dict(x=x, seed=s, ksizes=np.array(ks), alphas=rec["rand"][0], out=out)
        for layer in range(4):
            arrays[f"ker{layer}"] = rec["randn"][2 * layer]
            arrays[f"shift{layer}"] = rec["randn"][2 * layer + 1]
        save("gin_k" + "".join(map(str, ks)), **arrays)

    # odd shapes / batch 1 / batch 3 / list input (gin.py:169-170)
    for tag, shape, s in [("odd", (1, 1, 5, 7, 9), 11), ("b3", (3, 1, 6, 6, 6), 12), ("thin", (1, 1, 1, 2, 17), 13)]:
        xx = volume(shape, 600 + s, "randn")
        out, rec = traced_gin(xx, s)
        arrays = dict(x=xx, seed=s, ksizes=np.array([[1, 3][i] for i in rec["k"]]), alphas=rec["rand"][0], out=out)
        for layer in range(4):
            arrays[f"ker{layer}"] = rec["randn"][2 * layer]
            arrays[f"shift{layer}"] = rec["randn"][2 * layer + 1]
        save(f"gin_{tag}", **arrays)

    # gin_hook on/off (gin.py:244-247)
    xx = volume((1, 1, 6, 7, 8), 77)
    os.environ["DG_TTA_INTERNAL_AUGMENTATION"] = "false"
    off = ref_gin.gin_hook(None, (xx,))
    assert isinstance(off, tuple) and off[0] is xx
    os.environ["DG_TTA_INTERNAL_AUGMENTATION"] = "true"
    torch.manual_seed(5)
    on = ref_gin.gin_hook(None, (xx,))
    save("gin_hook", x=xx, seed=5, out_on=on)

    # gin_mind_aug (augmentation_utils.py:173-174) with recorded GIN draws and injected MIND noise
    xx = volume((2, 1, 9, 10, 11), 88)
    g = torch.Generator().manual_seed(4242)
    noise = torch.randn(2, 12, 9, 10, 11, generator=g)
    o_rl = torch.randn_like
    torch.randn_like = lambda t, *a, **k: noise.clone()
    rec = {"k": [], "randn": [], "rand": []}
    o_randint, o_randn, o_rand = torch.randint, torch.randn, torch.rand

    def randint(*a, **k):
        r = o_randint(*a, **k); rec["k"].append(int(r[0])); return r

    def randn(*a, **k):
        r = o_randn(*a, **k); rec["randn"].append(r.clone()); return r

    def rand(*a, **k):
        r = o_rand(*a, **k); rec["rand"].append(r.clone()); return r

    torch.manual_seed(3)
    torch.randint, torch.randn, torch.rand = randint, randn, rand
    try:
        with torch.no_grad():
            out = ref_aug.gin_mind_aug(xx)
    finally:
        torch.randint, torch.randn, torch.rand = o_randint, o_randn, o_rand
        torch.randn_like = o_rl
    arrays = dict(x=xx, seed=3, ksizes=np.array([[1, 3][i] for i in rec["k"]]), alphas=rec["rand"][0],
                  noise=noise, out=out)
    for layer in range(4):
        arrays[f"ker{layer}"] = rec["randn"][2 * layer]
        arrays[f"shift{layer}"] = rec["randn"][2 * layer + 1]
    save("gin_mind_aug", **arrays)


# ----------------------------------------------------------------------------- affine sampling
def gen_affine():
    # get_rand_affine (augmentation_utils.py:156-170)
    torch.manual_seed(21)
    R, Ri = ref_aug.get_rand_affine(3, strength=0.05, flip=False)
    torch.manual_seed(22)
    Rf, Rfi = ref_aug.get_rand_affine(2, strength=0.1, flip=True)
    save("rand_affine", seed=21, R=R, R_inv=Ri, seed_flip=22, R_flip=Rf, R_flip_inv=Rfi)

    # view warp exactly as calc_branch composes it (tta.py:143-147, 505, 523-532, 548-551, 571-575)
    B, C, patch = 2, 3, [9, 10, 12]
    imgs = volume((B, 1, *patch), 900)
    logits = volume((B, C, *patch), 901, "randn")
    identity_grid = F.affine_grid(torch.eye(4).repeat(B, 1, 1)[:, :3], [B, 1] + patch, align_corners=False)
    torch.manual_seed(31)
    R, R_inverse = ref_aug.get_rand_affine(B, flip=False)
    zero_grid = 0.0 * identity_grid
    grid = zero_grid + (F.affine_grid(R, [B, 1] + patch, align_corners=False) - identity_grid)
    grid_inverse = zero_grid + (F.affine_grid(R_inverse, [B, 1] + patch, align_corners=False) - identity_grid)
    grid = grid + identity_grid
    imgs_aug = F.grid_sample(imgs, grid, padding_mode="border", align_corners=False)
    grid_inverse = grid_inverse + identity_grid
    lg = logits.clone().requires_grad_(True)
    warped = F.grid_sample(lg, grid_inverse, align_corners=False)
    gout = volume((B, C, *patch), 902, "randn")
    (warped * gout).sum().backward()
    save("affine_view", imgs=imgs, logits=logits, R=R, R_inv=R_inverse, imgs_aug=imgs_aug,
         warped=warped, grad_out=gout, grad_logits=lg.grad)

    # stronger affines incl. out-of-bounds, in != out size, nearest mode
    torch.manual_seed(41)
    theta = torch.eye(3, 4).unsqueeze(0).repeat(2, 1, 1) + 0.3 * torch.randn(2, 3, 4)
    src = volume((2, 2, 7, 8, 9), 903, "randn")
    out_size = [2, 2, 6, 11, 10]
    grid = F.affine_grid(theta, out_size, align_corners=False)
    res = {}
    for mode in ("bilinear", "nearest"):
        for pad in ("zeros", "border"):
            res[f"{mode}_{pad}"] = F.grid_sample(src, grid, mode=mode, padding_mode=pad, align_corners=False)
    s2 = src.clone().requires_grad_(True)
    o = F.grid_sample(s2, grid, mode="bilinear", padding_mode="zeros", align_corners=False)
    go = volume(tuple(out_size), 904, "randn")
    (o * go).sum().backward()
    s3 = src.clone().requires_grad_(True)
    o3 = F.grid_sample(s3, grid, mode="bilinear", padding_mode="border", align_corners=False)
    (o3 * go).sum().backward()
    save("affine_general", src=src, theta=theta, out_size=np.array(out_size), grad_out=go,
         grad_src_zeros=s2.grad, grad_src_border=s3.grad, **res)

    # get_batch (torch_utils.py:13-76): random crop + centre crop, image and one-hot labels
    vol = volume((1, 1, 14, 15, 16), 905)[0, 0]
    lab = torch.zeros(3, 14, 15, 16)
    lab[0, 2:8, 3:9, 4:10] = 1
    lab[1, 8:13, 1:6, 9:15] = 1
    lab[2, 5:10, 9:14, 2:7] = 1
    sample = torch.cat([vol[None], lab], 0)
    torch.manual_seed(51)
    b_img, b_lbl = ref_tu.get_batch([sample], [0, 0], [8, 10, 12], fixed_patch_idx=None, device="cpu")
    c_img, c_lbl = ref_tu.get_batch([sample], [0], [8, 10, 12], fixed_patch_idx="center", device="cpu")
    # patch larger than the volume along one axis (offset range clipped to 0, zeros outside)
    torch.manual_seed(52)
    l_img, l_lbl = ref_tu.get_batch([sample], [0], [16, 10, 20], fixed_patch_idx=None, device="cpu")
    save("get_batch", sample=sample, seed=51, patch=np.array([8, 10, 12]),
         img0=b_img[0], img1=b_img[1], lbl0=b_lbl[0], lbl1=b_lbl[1], img_c=c_img[0], lbl_c=c_lbl[0],
         seed_large=52, patch_large=np.array([16, 10, 20]), img_l=l_img[0], lbl_l=l_lbl[0])


def gen_consistency():
    # consistency loss of the TTA step: dg_tta/tta/tta.py:263-269 written out around the reference's own
    # soft_dice_loss (dg_tta/tta/torch_utils.py:90-104); loss and d loss / d target_a
    g = torch.Generator().manual_seed(77)
    ta = (torch.randn(2, 6, 10, 12, 14, generator=g) * 2 + 0.3)
    tb = (torch.randn(2, 6, 10, 12, 14, generator=g) * 2 + 0.3)
    ta[:, :, :2] = 0.0          # zeros warped in from outside the volume -> outside the common-content mask
    tb[:, :, :, -3:] = 0.0
    ta.requires_grad_(True)
    mask = (ta.sum(1, keepdim=True) > 0.0).float() * (tb.sum(1, keepdim=True) > 0.0).float()
    sm_a = ta.softmax(1) * mask
    sm_b = tb.softmax(1) * mask
    loss = 1 - ref_tu.soft_dice_loss(sm_a, sm_b)[:, 1:].mean()
    loss.backward()
    save("consistency", target_a=ta.detach(), target_b=tb, loss=loss.detach(), grad_a=ta.grad)
    # label crop: get_argmaxed_segs (torch_utils.py:79-82) on overlapping / empty / fractional channels
    seg = (torch.rand(1, 5, 6, 7, 8, generator=g) > 0.7).float()
    seg[:, 3] *= 0.5
    save("argmaxed_segs", segs=seg, out=ref_tu.get_argmaxed_segs(seg))


def gen_nearest_ties():
    """Index work must be bit-exact: label crops whose source coordinates sit exactly on .5 ties (a centre crop of an
    odd-sized patch out of an even-sized volume maps every voxel to k + 0.5 on that axis), random crops, an up-sampling
    crop, and general affines in nearest mode on a larger grid — all through the reference's get_batch
    (torch_utils.py:13-82) / F.affine_grid + F.grid_sample(mode="nearest")."""
    g = torch.Generator().manual_seed(1234)
    vol = volume((1, 1, 14, 18, 22), 906)[0, 0]
    ids = torch.randint(0, 6, (4, 5, 6), generator=g)                       # blocky label map, 0 = unlabeled
    ids = ids.repeat_interleave(4, 0)[:14].repeat_interleave(4, 1)[:, :18].repeat_interleave(4, 2)[:, :, :22]
    lab = torch.stack([(ids == l).float() for l in range(1, 6)])
    sample = torch.cat([vol[None], lab], 0)
    arrays = dict(sample=sample)
    cases = [("tie_center", [7, 9, 11], "center", None), ("rand_a", [7, 9, 11], None, 61), ("rand_b", [6, 10, 12], None, 62),
             ("up", [20, 18, 30], None, 63), ("same", [14, 18, 22], "center", None)]
    for name, patch, fixed, seed in cases:
        if seed is not None:
            torch.manual_seed(seed)
        b_img, b_lbl = ref_tu.get_batch([sample], [0, 0], patch, fixed_patch_idx=fixed, device="cpu")
        arrays[f"{name}_patch"] = np.array(patch)
        arrays[f"{name}_seed"] = np.array(-1 if seed is None else seed)
        for i in range(2):
            arrays[f"{name}_img{i}"] = b_img[i]
            arrays[f"{name}_lbl{i}"] = b_lbl[i]
    torch.manual_seed(64)
    theta = torch.eye(3, 4).unsqueeze(0).repeat(3, 1, 1) + 0.15 * torch.randn(3, 3, 4)
    theta[2] = torch.tensor([[0.5, 0, 0, 0.0], [0, 0.5, 0, 0.0], [0, 0, 0.5, 0.0]])   # exact ties on the odd axes below
    src = torch.arange(3 * 2 * 14 * 18 * 22, dtype=torch.float32).view(3, 2, 14, 18, 22)   # value == linear index
    out_size = [3, 2, 7, 9, 11]
    grid = F.affine_grid(theta, out_size, align_corners=False)
    arrays.update(theta=theta, src_shape=np.array(src.shape), out_size=np.array(out_size),
                  nearest_zeros=F.grid_sample(src, grid, mode="nearest", padding_mode="zeros", align_corners=False),
                  nearest_border=F.grid_sample(src, grid, mode="nearest", padding_mode="border", align_corners=False))
    save("nearest_ties", **arrays)


if __name__ == "__main__":
    groups = dict(mind=gen_mind, gin=gen_gin, affine=gen_affine, consistency=gen_consistency, nearest=gen_nearest_ties)
    for name in (sys.argv[1:] or list(groups)):
        groups[name]()
    total = sum(p.stat().st_size for p in OUT.glob("*.npz"))
    print(f"total fixture bytes: {total}")
