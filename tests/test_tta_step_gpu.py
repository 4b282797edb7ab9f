"""BASELINE.json configs[2]: GIN_MIND transforms feeding a PlainConvUNet-shaped TTA step with the affine grid_sample
consistency loss.  The stand-in loop (dg_tta_b200/tta/standin.py, a restatement of dg_tta/tta/tta.py:190-281, 480-579)
is run twice from the same seeds: once on the drop-in ops, once on the reference's torch-eager op sequence
(tests/eager_transforms.py).  Same generator consumption -> same crops, same affines, same MIND noise, so

* the consistency loss and the gradient it sends into the network must agree (one inner step);
* a whole adaptation (epochs x 16 accumulated patches, AdamW) must end in the same segmentation: per-class Dice within
  0.5 points (north_star's TTA tolerance);
* the CUDA-graph replay of the transform segment must be bit-identical to the eager call sequence.
"""
import copy
import json
from pathlib import Path

import numpy as np
import pytest
import torch

from gpu_util import synth_volume

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def _run(transforms, seed):
    from dg_tta_b200.tta import standin as ts
    _no_tf32()
    model = ts.build_model(transforms, num_classes=9, features=(8, 16, 32), seed=3).cuda()
    vol = [synth_volume((1, 1, 52, 60, 56), 77, "mr")[0].cuda()]
    torch.manual_seed(seed)
    rng = np.random.RandomState(seed)
    loss = ts.tta_inner_step(model, vol, [32, 40, 32], 2, list(range(1, 6)), transforms, rng=rng)
    grad = torch.cat([p.grad.reshape(-1) for p in model.parameters() if p.grad is not None])
    return float(loss), grad


def test_inner_step_matches_eager_reference_ops():
    from dg_tta_b200.tta import standin as ts
    from eager_transforms import EagerTorchTransforms
    loss_a, grad_a = _run(ts.DropInTransforms(), 11)
    loss_b, grad_b = _run(EagerTorchTransforms(), 11)
    assert abs(loss_a - loss_b) <= 2e-4 * max(1.0, abs(loss_b))
    denom = float(grad_b.abs().max())
    assert denom > 0
    assert float((grad_a - grad_b).abs().max()) <= 2e-2 * denom   # fp32 cuDNN backward + atomic scatter order


def test_view_graph_replay_is_bitwise_the_eager_sequence():
    """ViewGraph (one CUDA graph: params memcpy -> crops -> 2 x (view warp, Philox field, MIND)) against the eager call
    sequence get_batch -> get_rand_affine -> affine_grid_sample(border) -> MIND3D() from the same seeds, three steps in
    a row (the generator state carries over between replays)."""
    from dg_tta_b200 import MIND3D
    from dg_tta_b200.tta import standin as ts
    from dg_tta_b200.tta.augmentation_utils import affine_grid_sample, get_rand_affine
    from dg_tta_b200.tta.torch_utils import get_batch
    vol = synth_volume((1, 1, 44, 52, 48), 12, "mr")[0].cuda()
    patch, B = [24, 32, 28], 2
    views = ts.ViewGraph(vol, patch, B)
    torch.manual_seed(5)
    got = []
    for _ in range(3):
        da, db, ia, ib = views.step()
        got.append((da.clone(), db.clone(), ia.clone(), ib.clone()))
    after_graph = torch.cuda.default_generators[torch.cuda.current_device()].get_offset()
    torch.manual_seed(5)
    for da, db, ia, ib in got:
        imgs, _ = get_batch([vol], [0] * B, patch, device="cuda")
        imgs = torch.cat(imgs, 0)
        Ra, Ra_inv = get_rand_affine(B, flip=False)
        Rb, Rb_inv = get_rand_affine(B, flip=False)
        ea = MIND3D()(affine_grid_sample(imgs, Ra, padding_mode="border"))
        eb = MIND3D()(affine_grid_sample(imgs, Rb, padding_mode="border"))
        assert torch.equal(ea, da) and torch.equal(eb, db)
        assert torch.equal(Ra_inv, ia) and torch.equal(Rb_inv, ib)
    assert torch.cuda.default_generators[torch.cuda.current_device()].get_offset() == after_graph


def _labelled_pair(shape=(48, 56, 52)):
    """A CT-like source volume, an MR-like target volume of the same anatomy (different contrast + bias field) and the
    label map (0 = background, 1..4 = ellipsoid 'organs')."""
    g = torch.Generator().manual_seed(2024)
    D, H, W = shape
    zz, yy, xx = torch.meshgrid(torch.linspace(-1, 1, D), torch.linspace(-1, 1, H), torch.linspace(-1, 1, W), indexing="ij")
    lab = torch.zeros(shape, dtype=torch.int64)
    centres = [(-0.4, -0.3, -0.3), (0.35, 0.3, -0.2), (-0.2, 0.4, 0.45), (0.4, -0.4, 0.4)]
    radii = [(0.35, 0.4, 0.3), (0.3, 0.35, 0.4), (0.35, 0.3, 0.3), (0.3, 0.3, 0.35)]
    for i, (c, r) in enumerate(zip(centres, radii)):
        lab[((zz - c[0]) / r[0]) ** 2 + ((yy - c[1]) / r[1]) ** 2 + ((xx - c[2]) / r[2]) ** 2 < 1] = i + 1
    low = torch.nn.functional.interpolate(torch.randn(1, 1, 4, 4, 4, generator=g), size=shape, mode="trilinear", align_corners=True)[0, 0]
    ct_vals = torch.tensor([-0.8, 0.9, 0.2, 1.6, -0.1])
    mr_vals = torch.tensor([0.1, -0.7, 1.3, 0.4, 1.9])
    fine = torch.randn(shape, generator=g)
    src = ct_vals[lab] + 0.25 * low + 0.05 * fine
    tgt = (mr_vals[lab] + 0.05 * fine) * (1 + 0.3 * low)
    return src.contiguous(), ((tgt - tgt.mean()) / tgt.std()).contiguous(), lab


def test_tta_adaptation_ends_in_the_same_dice():
    """north_star: per-class Dice of a TTA run within 0.5 points.  A small UNet is pre-trained on the CT-like source
    (supervised, drop-in MIND features), then adapted to the MR-like target by the stand-in TTA loop — 3 epochs x 16
    accumulated patches, AdamW (tta.py:190-281; lr 1e-4, ten times the reference default, to make the two runs' updates
    count) — once on the drop-in ops and once on the reference's eager op sequence, from the same seeds; both adapted
    networks then segment the whole target volume."""
    from dg_tta_b200.tta import standin as ts
    from eager_transforms import EagerTorchTransforms
    _no_tf32()
    src, tgt, lab = _labelled_pair()
    ncls, patch = 5, [32, 40, 32]
    dropin = ts.DropInTransforms()
    base = ts.build_model(dropin, num_classes=ncls, features=(8, 16, 32), seed=1).cuda()
    # --- supervised pre-training on the source domain (random crops; plain torch indexing, not part of the parity claim)
    opt = torch.optim.Adam(base.parameters(), lr=3e-3)
    gen = torch.Generator().manual_seed(9)
    srcd, labd = src.cuda(), lab.cuda()
    torch.manual_seed(100)
    for _ in range(200):
        xs, ys = [], []
        for _b in range(2):
            o = [int(torch.randint(0, s - p + 1, (1,), generator=gen)) for s, p in zip(src.shape, patch)]
            sl = tuple(slice(a, a + p) for a, p in zip(o, patch))
            xs.append(srcd[sl][None, None])
            ys.append(labd[sl][None])
        loss = torch.nn.functional.cross_entropy(base(torch.cat(xs)), torch.cat(ys))
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
    with torch.no_grad():
        base.head.bias += 10.0      # softmax-invariant; makes sum_c logits > 0 inside the volume, i.e. the reference's
        #                             common-content mask (tta.py:263-264) non-empty for this freshly trained head
    state = copy.deepcopy(base.state_dict())

    def predict(model):
        torch.manual_seed(7)                      # same MIND noise for every prediction
        with torch.no_grad():
            return model(tgt.cuda()[None, None]).argmax(1)[0].cpu()

    dice_before = ts.dice_per_class(predict(base), lab, ncls)
    results = {}
    for name, tr in (("dropin", dropin), ("eager", EagerTorchTransforms())):
        model = ts.build_model(tr, num_classes=ncls, features=(8, 16, 32), seed=1).cuda()
        model.load_state_dict(state)
        torch.manual_seed(31)
        losses = ts.run_adaptation(model, [tgt.cuda()[None]], patch, 2, list(range(ncls)), tr, epochs=3, accum=16, lr=1e-4,
                                   rng=np.random.RandomState(31))
        results[name] = dict(dice=ts.dice_per_class(predict(model), lab, ncls), losses=losses)
    da, db = np.array(results["dropin"]["dice"]), np.array(results["eager"]["dice"])
    try:
        (ROOT / "gpurun_out").mkdir(exist_ok=True)
        (ROOT / "gpurun_out" / "tta_dice.json").write_text(json.dumps(
            dict(dice_before=dice_before, dice_dropin=da.tolist(), dice_eager=db.tolist(),
                 loss_first=[results[k]["losses"][0] for k in ("dropin", "eager")],
                 loss_last=[results[k]["losses"][-1] for k in ("dropin", "eager")]), indent=1))
    except OSError:
        pass
    assert np.all(np.abs(da - db) <= 0.5), (da, db)
    la, lb = np.array(results["dropin"]["losses"]), np.array(results["eager"]["losses"])
    assert np.abs(la - lb).max() <= 5e-3 * max(1.0, np.abs(lb).max())
    assert np.nanmean(db) > 20.0, "the pre-trained network should segment something (sanity of the test itself)"
