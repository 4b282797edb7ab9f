"""MultiRes low-resolution simulation on the GPU — drop-in for dg_tta/pretraining/discrete_downsampling.py
(augment_discrete_linear_downsampling_scipy :8-37, SimulateDiscreteLowResolutionTransform :41-75), which
nnUNetTrainer_GIN_MIND_MultiRes.py:57-69 installs with zoom_range=(1/6, 1/4, 1/2), zoom_axes_invidually=True,
order_downsample=0, order_upsample=3, p_per_sample=.5, p_per_channel=1.

Same names, arguments and numpy.random draw order as the reference (np.random.uniform per sample, np.random.choice for
the zooms, np.random.uniform per channel), so np.random.seed(s) selects the same samples, zooms and channels.  The two
`skimage.transform.resize(..., mode='edge', anti_aliasing=False)` calls per channel run in the CUDA library
(csrc/resize.cu behind dgtta_resize_edge; orders 0, 1 and 3) on CUDA tensors instead of float64 numpy arrays in the
data-loader workers; the data stays on the device for the GIN / MIND hooks that follow.  No CPU fallback.
"""
import numpy as np
import torch

from .. import _lib


def resize_edge(volumes, out_shape, order):
    """skimage.transform.resize(v, out_shape, order=order, mode='edge', anti_aliasing=False) for every volume of a CUDA
    float32 tensor [N, D, H, W] (all share the geometry) -> [N, *out_shape] float32; order in {0, 1, 3}."""
    _lib.require_cuda_f32(volumes, "volumes")
    if volumes.dim() != 4:
        raise ValueError("resize_edge expects [N, D, H, W]")
    if order not in (0, 1, 3):
        raise NotImplementedError(f"interpolation order {order} is not built (0, 1 and 3 are)")
    L = _lib.lib()
    x = volumes.contiguous()
    N, Di, Hi, Wi = x.shape
    Do, Ho, Wo = (int(v) for v in out_shape)
    with torch.cuda.device(x.device):
        out = torch.empty((N, Do, Ho, Wo), device=x.device, dtype=torch.float32)
        nbytes = L.dgtta_resize_edge_workspace_bytes(N, Di, Hi, Wi, Do, Ho, Wo, order)
        ws = torch.empty(nbytes + 256, device=x.device, dtype=torch.uint8)
        base = (ws.data_ptr() + 255) // 256 * 256
        _lib.check(L.dgtta_resize_edge(x.data_ptr(), out.data_ptr(), N, Di, Hi, Wi, Do, Ho, Wo, order, base, nbytes,
                                       _lib.stream_ptr()), "dgtta_resize_edge")
    return out


def augment_discrete_linear_downsampling_scipy(data_sample, zoom_range=(1 / 6, 1 / 4, 1 / 2), zoom_axes_invidually=False, p=.2,
                                               channels=None, order_downsample=1, order_upsample=0, ignore_axes=None):
    """discrete_downsampling.py:8-37 on a CUDA tensor [C, D, H, W]: modified in place and returned, like the reference."""
    _lib.require_cuda_f32(data_sample, "data_sample")
    if not isinstance(zoom_range, (list, tuple, np.ndarray)):
        zoom_range = [zoom_range]
    shp = np.array(data_sample.shape[1:])
    if zoom_axes_invidually:
        zooms = np.random.choice(zoom_range, 3, replace=True)
    else:
        zooms = np.random.choice(zoom_range, 1)
    target_shape = np.round(shp * zooms).astype(int)
    if ignore_axes is not None:
        for i in ignore_axes:
            target_shape[i] = shp[i]
    if channels is None:
        channels = list(range(data_sample.shape[0]))
    picked = [c for c in channels if np.random.uniform() < p]      # one draw per channel, in channel order
    if picked:
        # the picked channels share the geometry: one batched call per resampling step
        src = data_sample[picked] if len(picked) > 1 else data_sample[picked[0]][None]
        down = resize_edge(src, target_shape, order_downsample)
        up = resize_edge(down, shp, order_upsample)
        for i, c in enumerate(picked):
            data_sample[c] = up[i]
    return data_sample


class SimulateDiscreteLowResolutionTransform:
    """discrete_downsampling.py:41-75 (a batchgenerators AbstractTransform there; batchgenerators is a data-loader
    dependency outside this path, so this is a plain callable with the same constructor and __call__(**data_dict))."""

    def __init__(self, zoom_range=(1 / 6, 1 / 4, 1 / 2), zoom_axes_invidually=False, per_channel=False, p_per_channel=1,
                 channels=None, order_downsample=1, order_upsample=0, data_key="data", p_per_sample=1, ignore_axes=None):
        self.order_upsample = order_upsample
        self.order_downsample = order_downsample
        self.channels = channels
        self.per_channel = per_channel
        self.p_per_channel = p_per_channel
        self.p_per_sample = p_per_sample
        self.data_key = data_key
        self.zoom_range = zoom_range
        self.zoom_axes_invidually = zoom_axes_invidually
        self.ignore_axes = ignore_axes

    def __call__(self, **data_dict):
        data = data_dict[self.data_key]
        for b in range(len(data)):
            if np.random.uniform() < self.p_per_sample:
                data[b] = augment_discrete_linear_downsampling_scipy(
                    data[b], zoom_range=self.zoom_range, zoom_axes_invidually=self.zoom_axes_invidually,
                    p=self.p_per_channel, channels=self.channels, order_downsample=self.order_downsample,
                    order_upsample=self.order_upsample, ignore_axes=self.ignore_axes)
        return data_dict
