"""MultiRes low-resolution simulation on the GPU against the resampling oracle (which is pinned on
scipy.ndimage.zoom, tests/test_resize_oracle.py): nearest down-sampling bit-exact (index work), linear / cubic within
1e-6 of the float64 result after its rounding to float32; the transform's draw order and in-place semantics."""
import numpy as np
import pytest
import torch

from gpu_util import synth_volume

pytestmark = pytest.mark.gpu

CASES = [((24, 20, 28), (1 / 2, 1 / 4, 1 / 6)), ((31, 33, 29), (1 / 6, 1 / 2, 1 / 4)), ((19, 19, 21), (1 / 2, 1 / 2, 1 / 2)),
         ((56, 56, 64), (1 / 4, 1 / 6, 1 / 2)), ((128, 128, 128), (1 / 2, 1 / 4, 1 / 6))]


@pytest.mark.parametrize("shape,zooms", CASES)
def test_resize_edge_against_oracle(shape, zooms):
    from dg_tta_b200.pretraining import resize_edge
    from oracle import resize_oracle as ro
    x = synth_volume((1, 2) + shape, 60 + shape[0], "mr")[0]              # [2, D, H, W]
    target = np.round(np.array(shape) * np.array(zooms)).astype(int)
    down = resize_edge(x.cuda(), target, 0).cpu().numpy()
    for c in range(2):
        assert np.array_equal(down[c], ro.resize_edge(x[c].numpy(), target, 0).astype(np.float32))
    for order in (0, 1, 3):
        up = resize_edge(torch.from_numpy(down).cuda(), shape, order).cpu().numpy()
        for c in range(2):
            ref = ro.resize_edge(down[c], shape, order)
            scale = max(1.0, np.abs(ref).max())
            if order == 0:
                assert np.array_equal(up[c], ref.astype(np.float32))
            else:
                assert np.abs(up[c] - ref).max() <= 1e-6 * scale
    d1 = resize_edge(x.cuda(), target, 1).cpu().numpy()                   # the function's default order_downsample
    assert np.abs(d1[0] - ro.resize_edge(x[0].numpy(), target, 1)).max() <= 1e-6 * max(1.0, float(x.abs().max()))


def test_transform_matches_reference_semantics():
    """SimulateDiscreteLowResolutionTransform with the MultiRes trainer's arguments (nnUNetTrainer_GIN_MIND_MultiRes.py:60-66)
    against the oracle restatement of discrete_downsampling.py:8-75 from the same numpy seed."""
    from dg_tta_b200.pretraining import SimulateDiscreteLowResolutionTransform
    from oracle import resize_oracle as ro
    data = synth_volume((4, 1, 40, 48, 44), 81, "ct")
    tr = SimulateDiscreteLowResolutionTransform(zoom_range=(1 / 6, 1 / 4, 1 / 2), zoom_axes_invidually=True, per_channel=False,
                                                p_per_channel=1., order_downsample=0, order_upsample=3, p_per_sample=.5,
                                                ignore_axes=None)
    np.random.seed(5)
    got = tr(data=data.clone().cuda())["data"].cpu().numpy()
    np.random.seed(5)
    ref = data.clone().numpy()
    touched = 0
    for b in range(ref.shape[0]):
        if np.random.uniform() < 0.5:
            touched += 1
            ref[b] = ro.augment_discrete_linear_downsampling(ref[b], zoom_axes_invidually=True, p=1., order_downsample=0,
                                                             order_upsample=3)
    assert 0 < touched < 4                                                  # the seed exercises both branches
    assert np.abs(got - ref).max() <= 1e-6 * max(1.0, np.abs(ref).max())
    untouched = [b for b in range(4) if np.array_equal(ref[b], data[b].numpy())]
    for b in untouched:
        assert np.array_equal(got[b], data[b].numpy())
