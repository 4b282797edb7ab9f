"""MIND-SSC descriptor — drop-in for dg_tta/mind.py (MIND3D :98-164, mind_hook :167-168).

Same call signatures as the reference; the arithmetic runs in one fused sm_100a stencil kernel
(csrc/mind_ssc.cu) reached through the C ABI (include/dgtta.h: dgtta_mind_ssc_fwd).
"""
import ctypes
import math

import torch

from . import _lib

_TAPS_CACHE = {}


def gaussian_taps(sigma):
    """Taps of smooth() (dg_tta/mind.py:27-37), evaluated with the same float32 tensor ops on the
    host: N = ceil(1.5 sigma)*2+1, w = exp(-x^2 / (2 sigma^2)), w /= sum(w).  Cached per sigma."""
    key = (type(sigma).__name__, float(sigma))
    taps = _TAPS_CACHE.get(key)
    if taps is None:
        s = torch.tensor([sigma])
        n = int(torch.ceil(s * 3.0 / 2.0).long().item()) * 2 + 1
        w = torch.exp(-torch.pow(torch.linspace(-(n // 2), n // 2, n), 2) / (2 * torch.pow(s, 2)))
        w = (w / w.sum()).to(torch.float32).contiguous()
        if not 3 <= n <= 9:
            raise ValueError(f"sigma={sigma} gives a {n}-tap Gaussian; the kernel supports 3..9 taps")
        taps = (ctypes.c_float * n)(*w.tolist())
        _TAPS_CACHE[key] = taps
    return taps


def _advance_generator(device, B, D, H, W):
    """Consume exactly the generator state torch.randn_like(edge_selection) (mind.py:150) would."""
    gen = torch.cuda.default_generators[device.index]
    props = torch.cuda.get_device_properties(device)
    inc = _lib.lib().dgtta_mind_philox_offset_increment(B, D, H, W, props.multi_processor_count,
                                                        props.max_threads_per_multi_processor)
    gen.set_offset(gen.get_offset() + inc)


def mind_ssc(img, delta=1, sigma=1, randn_weighting=0.05, noise=None, in_scale=None):
    """Functional form.  img [B,1,D,H,W] CUDA float32 -> [B,12,D,H,W].

    noise: None  -> draw the field the reference draws (torch.randn of the edge tensor's shape on
                    img's device generator; skipped but accounted for when randn_weighting == 0);
           False -> noise-free (E = I(p+s1) - I(p+s2)), generator untouched;
           Tensor [B,12,D,H,W] -> use it (what tests inject to compare with the CPU oracle).
    in_scale: optional [B,2] CUDA float32; the kernel reads I = (img*a_b)*c_b (deferred GIN rescale)."""
    _lib.require_cuda_f32(img, "img")
    if img.dim() != 5 or img.shape[1] != 1:
        raise ValueError(f"MIND3D expects [B,1,D,H,W], got {tuple(img.shape)}")
    L = _lib.lib()
    img = img.contiguous()
    B, _, D, H, W = img.shape
    delta = int(delta)
    if delta < 1:
        raise ValueError("delta must be >= 1")
    taps = gaussian_taps(sigma)
    rw = float(randn_weighting)
    with torch.cuda.device(img.device):
        if noise is None:
            if rw != 0.0:
                noise = torch.randn((B, 12, D, H, W), device=img.device, dtype=img.dtype)
            else:
                _advance_generator(img.device, B, D, H, W)
                noise = False
        if noise is False:
            mode, nptr = 0, None
        else:
            _lib.require_cuda_f32(noise, "noise")
            if tuple(noise.shape) != (B, 12, D, H, W):
                raise ValueError("noise must have shape [B,12,D,H,W]")
            noise = noise.contiguous()
            mode, nptr = 1, noise.data_ptr()
        sptr = None
        if in_scale is not None:
            _lib.require_cuda_f32(in_scale, "in_scale")
            if tuple(in_scale.shape) != (B, 2):
                raise ValueError("in_scale must have shape [B,2]")
            in_scale = in_scale.contiguous()
            sptr = in_scale.data_ptr()
        out = torch.empty((B, 12, D, H, W), device=img.device, dtype=torch.float32)
        nbytes = L.dgtta_mind_workspace_bytes(B, D, H, W)
        ws = torch.empty(max(nbytes, 16), device=img.device, dtype=torch.uint8)
        rc = L.dgtta_mind_ssc_fwd(img.data_ptr(), out.data_ptr(), sptr, B, D, H, W, delta, taps, len(taps),
                                  rw, mode, nptr, 0, 0, ws.data_ptr(), nbytes, _lib.stream_ptr())
        _lib.check(rc, "dgtta_mind_ssc_fwd")
    return out


class MIND3D(torch.nn.Module):
    """Same constructor, attributes and call behaviour as dg_tta/mind.py:97-164: a parameter-free
    module (survives deepcopy) mapping [B,1,D,H,W] to the 12-channel self-similarity descriptor,
    with Gaussian noise of weight `randn_weighting` added to the edge differences on every call."""

    def __init__(self, delta=1, sigma=1, randn_weighting=0.05) -> None:
        super().__init__()
        self.delta = delta
        self.sigma = sigma
        self.out_channels = 12
        self.randn_weighting = randn_weighting

    def forward(self, img, noise=None):
        return mind_ssc(img, self.delta, self.sigma, self.randn_weighting, noise=noise)

    def extra_repr(self):
        return f"delta={self.delta}, sigma={self.sigma}, randn_weighting={self.randn_weighting}"


def mind_hook(module, input):
    """forward-pre-hook (dg_tta/mind.py:167-168): replaces the module input by its descriptor."""
    return MIND3D().forward(*input)
