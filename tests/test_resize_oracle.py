"""The resampling oracle of the MultiRes low-resolution simulation (oracle/resize_oracle.py) against scipy.ndimage.zoom —
the third-party routine skimage.transform.resize delegates to for discrete_downsampling.py:29-33 (skimage itself is not
installed: see the oracle's header for what is pinned and what is not)."""
import numpy as np
import pytest

from oracle import resize_oracle as ro

scipy = pytest.importorskip("scipy")

SHAPES = [((24, 20, 28), (1 / 2, 1 / 4, 1 / 6)), ((31, 33, 29), (1 / 6, 1 / 2, 1 / 4)), ((64, 48, 56), (1 / 4, 1 / 4, 1 / 2)),
          ((19, 19, 21), (1 / 2, 1 / 2, 1 / 2)), ((7, 40, 12), (1 / 6, 1 / 6, 1 / 6))]


@pytest.mark.parametrize("shape,zooms", SHAPES)
def test_resize_oracle_matches_scipy_zoom(shape, zooms):
    rng = np.random.default_rng(sum(shape))
    x = rng.standard_normal(shape).astype(np.float32)
    target = np.round(np.array(shape) * np.array(zooms)).astype(int)
    down = ro.resize_edge(x, target, 0)
    assert np.array_equal(down, ro.scipy_resize_edge(x, target, 0))          # index work: bit-exact
    down1 = ro.resize_edge(x, target, 1)
    assert np.abs(down1 - ro.scipy_resize_edge(x, target, 1)).max() <= 1e-14
    for order, tol in ((0, 0.0), (1, 1e-14), (3, 1e-12)):
        up = ro.resize_edge(down, shape, order)
        ref = ro.scipy_resize_edge(down, shape, order)
        assert up.shape == tuple(shape) and np.abs(up - ref).max() <= tol
        assert up.min() >= down.min() and up.max() <= down.max()             # skimage clip=True


def test_transform_draw_order_and_in_place_semantics():
    """discrete_downsampling.py:8-37: one np.random.choice(…, 3) for the zooms, one uniform per channel; channels that lose
    the draw stay untouched; the sample is modified in place and returned."""
    x = np.random.default_rng(3).standard_normal((2, 16, 18, 20)).astype(np.float32)
    np.random.seed(11)
    zooms = np.random.choice((1 / 6, 1 / 4, 1 / 2), 3, replace=True)
    u = [np.random.uniform() for _ in range(2)]
    np.random.seed(11)
    y = x.copy()
    out = ro.augment_discrete_linear_downsampling(y, zoom_axes_invidually=True, p=0.5, order_downsample=0, order_upsample=3)
    assert out is y
    target = np.round(np.array(x.shape[1:]) * zooms).astype(int)
    for c in range(2):
        if u[c] < 0.5:
            ref = ro.scipy_resize_edge(ro.scipy_resize_edge(x[c], target, 0), x.shape[1:], 3)
            assert np.abs(y[c] - ref.astype(np.float32)).max() <= 1e-6
        else:
            assert np.array_equal(y[c], x[c])
