"""GIN augmentation — drop-in for dg_tta/gin.py (GradlessGCReplayNonlinBlock :35-122,
GINGroupConv :125-230, gin_aug :233-241, gin_hook :244-247), 5-D (volumetric) inputs.

The random draws stay on the host in the reference's exact order and on the reference's devices
(SURVEY.md §8a6): alphas = torch.rand(B, device=x.device); then per layer
randint(2,(1,)) / randn(ker) / randn(shift) on the CPU generator — so torch.manual_seed(s)
reproduces the reference's weights.  The convolution stack, blend and Frobenius re-normalisation
run in the CUDA library (csrc/gin.cu, csrc/gin_fused.cu) behind dgtta_gin_fwd.
"""
import ctypes

import torch
import torch.nn as nn

from . import _lib
from .utils import get_internal_augmentation_enabled


def _as_5d(x_in):
    if x_in.dim() == 5:
        return x_in
    if x_in.dim() == 4:
        raise NotImplementedError("2-D GIN (4-D input, gin.py:75-90) is outside the B200 hot path; pass [B,C,D,H,W]")
    raise ValueError()


class GradlessGCReplayNonlinBlock(nn.Module):
    """One random conv + shift + leaky-ReLU layer (gin.py:35-122); same constructor arguments."""

    def __init__(self, out_channel=32, in_channel=3, scale_pool=[1, 3], layer_id=0, use_act=True,
                 requires_grad=False):
        super().__init__()
        self.in_channel = in_channel
        self.out_channel = out_channel
        self.scale_pool = scale_pool
        self.layer_id = layer_id
        self.use_act = use_act
        self.requires_grad = requires_grad
        assert requires_grad == False  # noqa: E712  (gin.py:57)

    def draw(self, nb):
        """The three host draws of gin.py:65-66,94-103 in reference order."""
        k = self.scale_pool[int(torch.randint(high=len(self.scale_pool), size=(1,)))]
        ker = torch.randn([self.out_channel * nb, self.in_channel, k, k, k])
        shift = torch.randn([self.out_channel * nb, 1, 1, 1])   # (the reference's "* 1.0" changes no value)
        return k, ker, shift

    def forward(self, x_in, requires_grad=False):
        x_in = _as_5d(x_in)
        _lib.require_cuda_f32(x_in, "x_in")
        nb, nc, nx, ny, nz = x_in.shape
        if nc != self.in_channel:
            raise ValueError(f"expected {self.in_channel} input channels, got {nc}")
        k, ker, shift = self.draw(nb)
        if k not in (1, 3):
            raise NotImplementedError("kernel sizes other than 1 and 3 are not built")
        L = _lib.lib()
        x = x_in.contiguous()
        out = torch.empty((nb, self.out_channel, nx, ny, nz), device=x.device, dtype=torch.float32)
        ker = ker.contiguous()
        shift = shift.contiguous()
        nbytes = (ker.numel() + shift.numel()) * 4
        with torch.cuda.device(x.device):
            ws = torch.empty(nbytes, device=x.device, dtype=torch.uint8)
            rc = L.dgtta_gin_layer_fwd(x.data_ptr(), out.data_ptr(), ker.data_ptr(), shift.data_ptr(), nb,
                                       self.in_channel, self.out_channel, k, nx, ny, nz, int(self.use_act),
                                       ws.data_ptr(), nbytes, _lib.stream_ptr())
            _lib.check(rc, "dgtta_gin_layer_fwd")
        return out


def gin_forward(x, kers, shifts, alphas, interm_channels, defer_scale=False):
    """Run the whole stack with explicit draws.  x [B,C,D,H,W] CUDA f32; kers/shifts: per-layer CPU
    tensors in the reference's shapes; alphas [B] on x's device.  Returns out, or (mixed, scale[B,2])
    when defer_scale (the consumer applies out = (mixed*scale[:,0])*scale[:,1], gin.py:228)."""
    _lib.require_cuda_f32(x, "x_in")
    _lib.require_cuda_f32(alphas, "alphas")
    _lib.require_no_grad(x, "x_in", "GINGroupConv")
    L = _lib.lib()
    x = x.contiguous()
    B, C, D, H, W = x.shape
    n_layer = len(kers)
    ksizes = [int(k.shape[-1]) for k in kers]
    if any(t.is_cuda for t in kers) or any(t.is_cuda for t in shifts):
        raise TypeError("GIN weights are host draws (gin.py:94-103); pass CPU tensors")
    with torch.no_grad():
        params = torch.cat([t.reshape(-1) for pair in zip(kers, shifts) for t in pair]).to(torch.float32)
    ks = (ctypes.c_int * n_layer)(*ksizes)
    alphas = alphas.reshape(-1).contiguous()
    if alphas.numel() != B:
        raise ValueError("alphas must have one entry per sample")
    with torch.cuda.device(x.device):
        out = torch.empty_like(x)
        scale = torch.empty((B, 2), device=x.device, dtype=torch.float32) if defer_scale else None
        nbytes = L.dgtta_gin_workspace_bytes(B, D, H, W, C, n_layer, interm_channels)
        ws = torch.empty(nbytes + 256, device=x.device, dtype=torch.uint8)
        base = (ws.data_ptr() + 255) // 256 * 256
        rc = L.dgtta_gin_fwd(x.data_ptr(), out.data_ptr(), params.data_ptr(), ks, alphas.data_ptr(), B, D, H, W, C,
                             n_layer, interm_channels, scale.data_ptr() if defer_scale else None, base, nbytes,
                             _lib.stream_ptr())
        _lib.check(rc, "dgtta_gin_fwd")
    return (out, scale) if defer_scale else out


class GINGroupConv(nn.Module):
    """Random shallow conv net + alpha blend + Frobenius re-normalisation (gin.py:125-230)."""

    def __init__(self, cfg):
        super().__init__()
        self.scale_pool = [1, 3]
        self.n_layer = cfg["N_LAYER"]
        self.out_norm = "frob"
        self.out_channel = cfg["IN_CHANNELS"]
        in_channel = cfg["IN_CHANNELS"]
        interm_channel = cfg["INTERM_CHANNELS"]
        self.interm_channel = interm_channel
        layers = [GradlessGCReplayNonlinBlock(out_channel=interm_channel, in_channel=in_channel,
                                              scale_pool=self.scale_pool, layer_id=0)]
        for ii in range(self.n_layer - 2):
            layers.append(GradlessGCReplayNonlinBlock(out_channel=interm_channel, in_channel=interm_channel,
                                                      scale_pool=self.scale_pool, layer_id=ii + 1))
        layers.append(GradlessGCReplayNonlinBlock(out_channel=self.out_channel, in_channel=interm_channel,
                                                  scale_pool=self.scale_pool, layer_id=self.n_layer - 1,
                                                  use_act=False))
        self.layers = nn.ModuleList(layers)

    def draw(self, x_in):
        """All random draws of one forward, in the reference's order (gin.py:187-190, then per layer)."""
        nb = x_in.shape[0]
        alphas = torch.rand(nb, device=x_in.device)
        kers, shifts = [], []
        for blk in self.layers:
            _, ker, shift = blk.draw(nb)
            kers.append(ker)
            shifts.append(shift)
        return alphas, kers, shifts

    def forward(self, x_in, defer_scale=False):
        if isinstance(x_in, list):
            x_in = torch.cat(x_in, dim=0)
        x_in = _as_5d(x_in)
        _lib.require_cuda_f32(x_in, "x_in")
        if x_in.shape[1] != self.out_channel:
            raise ValueError(f"expected {self.out_channel} channels, got {x_in.shape[1]}")
        alphas, kers, shifts = self.draw(x_in)
        return gin_forward(x_in, kers, shifts, alphas, self.interm_channel, defer_scale=defer_scale)


_GIN_CFG = dict(IN_CHANNELS=1, N_LAYER=4, INTERM_CHANNELS=2)


_DEFAULT_NET = None


def default_gin():
    """The GINGroupConv of gin_aug's fixed configuration.  The reference builds a new one on every call
    (gin.py:234-240); the module holds no parameters or state — every forward draws fresh weights — so one shared
    instance is equivalent and saves ~0.1 ms of nn.Module construction per call (the whole call is ~0.4 ms of host
    time, which is what bounds the transform on patch-sized inputs)."""
    global _DEFAULT_NET
    if _DEFAULT_NET is None:
        _DEFAULT_NET = GINGroupConv(dict(_GIN_CFG))
    return _DEFAULT_NET


def gin_aug(input):
    """gin.py:233-241: the 4-layer, 2-intermediate-channel GIN with fresh random weights per call."""
    return default_gin()(input)


def gin_hook(module, input):
    """forward-pre-hook (gin.py:244-247): augments iff DG_TTA_INTERNAL_AUGMENTATION == "true",
    otherwise hands the input tuple back unchanged."""
    if get_internal_augmentation_enabled():
        return gin_aug(*input)
    return input
