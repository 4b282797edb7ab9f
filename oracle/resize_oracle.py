"""TEST INFRASTRUCTURE ONLY — CPU restatement (numpy, float64) of the resampling behind the MultiRes low-resolution
simulation, dg_tta/pretraining/discrete_downsampling.py:8-37 (augment_discrete_linear_downsampling_scipy; used by
nnUNetTrainer_GIN_MIND_MultiRes.py:57-69 with order_downsample=0, order_upsample=3).

Where the arithmetic lives: `skimage.transform.resize(..., mode='edge', anti_aliasing=False)` — third party, NOT under
/root/reference and NOT installed here (scikit-image >= 0.19, pulled in by nnunetv2 2.2.1 / batchgenerators).  Its published
algorithm for this call is: factors = in_shape / out_shape; out = scipy.ndimage.zoom(image.astype(float), 1 / factors,
order=order, mode='nearest', grid_mode=True); np.clip(out, image.min(), image.max()).  The wrapper is restated from the
skimage source from memory (PARITY UNPINNED for that thin layer); the numerical core is pinned: `resize_edge` below is
checked against scipy.ndimage.zoom itself (scipy 1.18 is installed) in tests/test_resize_oracle.py — orders 0 and 1
bit-for-bit, order 3 to 1e-13.  scipy sources followed: ndimage/src/ni_interpolation.c (NI_ZoomShift, grid_mode
coordinates), ni_splines.c (_apply_filter_gain, _init_causal_reflect, _init_anticausal_reflect, spline weights),
ndimage/_interpolation.py (_prepad_for_spline_filter: 12 edge samples for mode 'nearest').
"""
import numpy as np

NPAD = 12


def axis_coords(n_in, n_out):
    """NI_ZoomShift with grid_mode: cc = o; cc += 0.5; cc *= n_in / n_out; cc -= 0.5 (float64, this order)."""
    zoom = np.float64(n_in) / np.float64(n_out) if n_out > 0 else np.float64(1.0)
    cc = np.arange(n_out, dtype=np.float64)
    cc = cc + 0.5
    cc = cc * zoom
    return cc - 0.5


def _prefilter_axis(c, axis):
    """Cubic B-spline prefilter of ni_splines.c for the boundary scipy uses with mode 'nearest' (half-sample symmetric)."""
    z = np.sqrt(3.0) - 2.0
    c = np.moveaxis(c, axis, 0).copy()
    n = c.shape[0]
    c *= (1.0 - z) * (1.0 - 1.0 / z)
    if n > 1:
        z_n = z ** n
        c0 = c[0].copy()
        acc = c[0] + z_n * c[n - 1]
        z_i = z
        for i in range(1, n):
            acc = acc + z_i * (c[i] + z_n * c[n - 1 - i])
            z_i *= z
        c[0] = acc * (z / (1 - z_n * z_n)) + c0
        for i in range(1, n):
            c[i] += z * c[i - 1]
        c[n - 1] *= z / (z - 1)
        for i in range(n - 2, -1, -1):
            c[i] = z * (c[i + 1] - c[i])
    return np.moveaxis(c, 0, axis)


def resize_edge(x, out_shape, order):
    """skimage.transform.resize(x, out_shape, order=order, mode='edge', anti_aliasing=False, clip=True) for a 3-D array,
    order in {0, 1, 3}; float64 result."""
    x = np.asarray(x, np.float64)
    out_shape = tuple(int(v) for v in out_shape)
    if order == 0:
        idx = [np.floor(np.clip(axis_coords(a, b), 0.0, a - 1.0) + 0.5).astype(np.int64) for a, b in zip(x.shape, out_shape)]
        out = x[idx[0]][:, idx[1]][:, :, idx[2]]
        return np.clip(out, x.min(), x.max())
    coef, npad = x, 0
    if order == 3:
        npad = NPAD
        coef = np.pad(x, npad, mode="edge")
        for ax in range(3):
            coef = _prefilter_axis(coef, ax)
    starts, ws = [], []
    for a, b in zip(x.shape, out_shape):
        cc = axis_coords(a, b)
        if order == 3:
            f = np.floor(cc)
            y = cc - f
            z = 1.0 - y
            w1 = (y * y * (y - 2.0) * 3.0 + 4.0) / 6.0
            w2 = (z * z * (z - 2.0) * 3.0 + 4.0) / 6.0
            w0 = z * z * z / 6.0
            w3 = 1.0 - w0 - w1 - w2
            starts.append(f.astype(np.int64) - 1 + npad)
            ws.append(np.stack([w0, w1, w2, w3], -1))
        else:
            cc = np.clip(cc, 0.0, a - 1.0)
            f = np.floor(cc)
            starts.append(f.astype(np.int64))
            ws.append(np.stack([1.0 - (cc - f), cc - f], -1))
    out = np.zeros(out_shape, np.float64)
    for i in range(order + 1):
        di = np.clip(starts[0] + i, 0, coef.shape[0] - 1)
        for j in range(order + 1):
            hj = np.clip(starts[1] + j, 0, coef.shape[1] - 1)
            for k in range(order + 1):
                wk = np.clip(starts[2] + k, 0, coef.shape[2] - 1)
                out += coef[di][:, hj][:, :, wk] * (ws[0][:, i][:, None, None] * ws[1][:, j][None, :, None] * ws[2][:, k][None, None, :])
    return np.clip(out, x.min(), x.max())


def scipy_resize_edge(x, out_shape, order):
    """The same call through scipy.ndimage.zoom — the dependency skimage delegates to (used to pin resize_edge)."""
    from scipy import ndimage as ndi
    x = np.asarray(x, np.float64)
    factors = np.divide(x.shape, out_shape)
    out = ndi.zoom(x, [1 / f for f in factors], order=order, mode="nearest", cval=0, grid_mode=True)
    return np.clip(out, x.min(), x.max())


def augment_discrete_linear_downsampling(data_sample, zoom_range=(1 / 6, 1 / 4, 1 / 2), zoom_axes_invidually=False, p=.2,
                                         channels=None, order_downsample=1, order_upsample=0, ignore_axes=None, resize=resize_edge):
    """discrete_downsampling.py:8-37 with the same numpy.random draws in the same order (np.random.choice for the zooms,
    one np.random.uniform per channel); `resize` is the resampling oracle."""
    if not isinstance(zoom_range, (list, tuple, np.ndarray)):
        zoom_range = [zoom_range]
    shp = np.array(data_sample.shape[1:])
    zooms = np.random.choice(zoom_range, 3, replace=True) if zoom_axes_invidually else np.random.choice(zoom_range, 1)
    target_shape = np.round(shp * zooms).astype(int)
    if ignore_axes is not None:
        for i in ignore_axes:
            target_shape[i] = shp[i]
    if channels is None:
        channels = list(range(data_sample.shape[0]))
    for c in channels:
        if np.random.uniform() < p:
            down = resize(data_sample[c].astype(float), target_shape, order_downsample)
            data_sample[c] = resize(down, shp, order_upsample)
    return data_sample
