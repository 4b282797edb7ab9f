"""View augmentation of the TTA consistency loop — drop-in for the live part of
dg_tta/tta/augmentation_utils.py (get_rand_affine :156-170, gin_mind_aug :173-174) plus the
affine_grid + grid_sample op pair the reference writes inline (tta.py:523-551,571-575;
torch_utils.py:55-73), fused here into `affine_grid_sample`.

The deformable branch of the reference (augmentation_utils.py:8-153) is dead code there
(get_disp_field raises TypeError) and is not rebuilt.
"""
import torch

from .. import _lib
from ..gin import default_gin
from ..mind import mind_ssc, randn_like_reference

_INTERP = {"bilinear": 0, "nearest": 1}
_PAD = {"zeros": 0, "border": 1}


def get_rand_affine(batch_size, strength=0.05, flip=False):
    """augmentation_utils.py:156-170 — host draws in the same order; returns (R[:, :3], R^-1[:, :3])."""
    affine = torch.cat(
        (torch.randn(batch_size, 3, 4) * strength + torch.eye(3, 4).unsqueeze(0),
         torch.tensor([0, 0, 0, 1]).view(1, 1, 4).repeat(batch_size, 1, 1)), 1)
    if flip:
        signs = 2 * (torch.rand(3) > 0.5).float() - 1
        affine = affine @ torch.diag(torch.cat([signs, torch.tensor([1.0])]))
    return affine[:, :3], affine.inverse()[:, :3]


class _PinnedRing:
    """Small host-side staging ring for the per-call affines (48 bytes per sample).  Allocating a pinned tensor per call
    (`.pin_memory()`) costs tens of microseconds — more than the 1-channel warp kernel itself — so the slots are
    allocated once; a slot is reused only after the copy that read it has completed (one event per slot)."""
    SLOTS, FLOATS = 64, 12 * 64

    def __init__(self):
        self.buf = torch.empty((self.SLOTS, self.FLOATS), dtype=torch.float32).pin_memory()
        self.events = [None] * self.SLOTS
        self.next = 0

    def stage(self, theta, device):
        n = theta.numel()
        i = self.next
        self.next = (i + 1) % self.SLOTS
        if self.events[i] is not None:
            self.events[i].synchronize()
        slot = self.buf[i, :n].view(theta.shape)
        slot.copy_(theta)
        out = slot.to(device=device, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(device))
        self.events[i] = ev
        return out


_RING = None


def _theta_on(device, theta, B):
    global _RING
    if tuple(theta.shape) != (B, 3, 4):
        raise ValueError(f"theta must have shape [{B},3,4], got {tuple(theta.shape)}")
    theta = theta.to(dtype=torch.float32)
    if theta.is_cuda:
        return theta.to(device=device).contiguous()
    if theta.numel() > _PinnedRing.FLOATS or torch.cuda.is_current_stream_capturing():
        return theta.contiguous().pin_memory().to(device=device, non_blocking=True)
    if _RING is None:
        _RING = _PinnedRing()
    # 48 bytes per sample: staged through pinned memory so the upload is a truly asynchronous copy
    return _RING.stage(theta.contiguous(), device)


class _AffineSample(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input, theta, out_size, interp, padding):
        L = _lib.lib()
        B, C, Di, Hi, Wi = input.shape
        Do, Ho, Wo = out_size
        x = input.contiguous()
        with torch.cuda.device(x.device):
            out = torch.empty((B, C, Do, Ho, Wo), device=x.device, dtype=torch.float32)
            rc = L.dgtta_affine_sample_fwd(x.data_ptr(), theta.data_ptr(), out.data_ptr(), B, C, Di, Hi, Wi,
                                           Do, Ho, Wo, interp, padding, _lib.stream_ptr())
            _lib.check(rc, "dgtta_affine_sample_fwd")
        ctx.save_for_backward(theta)
        ctx.geom = (B, C, Di, Hi, Wi, Do, Ho, Wo, interp, padding)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (theta,) = ctx.saved_tensors
        B, C, Di, Hi, Wi, Do, Ho, Wo, interp, padding = ctx.geom
        if interp != 0:
            return torch.zeros((B, C, Di, Hi, Wi), device=grad_out.device), None, None, None, None
        L = _lib.lib()
        g = grad_out.contiguous().to(torch.float32)
        with torch.cuda.device(g.device):
            grad_in = torch.empty((B, C, Di, Hi, Wi), device=g.device, dtype=torch.float32)
            rc = L.dgtta_affine_sample_bwd_input(g.data_ptr(), theta.data_ptr(), grad_in.data_ptr(), B, C, Di, Hi,
                                                 Wi, Do, Ho, Wo, padding, _lib.stream_ptr())
            _lib.check(rc, "dgtta_affine_sample_bwd_input")
        return grad_in, None, None, None, None


def affine_grid_sample(input, theta, out_size=None, mode="bilinear", padding_mode="zeros", align_corners=False):
    """F.grid_sample(input, F.affine_grid(theta, out_size, align_corners=False), mode, padding_mode,
    align_corners=False) in one kernel, differentiable w.r.t. `input` (not theta — the reference never
    needs that gradient).  theta [B,3,4] may live on the host (as get_rand_affine returns it)."""
    if align_corners:
        raise NotImplementedError("the reference only uses align_corners=False")
    _lib.require_cuda_f32(input, "input")
    if input.dim() != 5:
        raise ValueError("affine_grid_sample expects [B,C,D,H,W]")
    if mode not in _INTERP or padding_mode not in _PAD:
        raise ValueError(f"unsupported mode/padding_mode: {mode}/{padding_mode}")
    B = input.shape[0]
    size = tuple(int(v) for v in (out_size[-3:] if out_size is not None else input.shape[-3:]))
    theta = _theta_on(input.device, theta, B)
    return _AffineSample.apply(input, theta, size, _INTERP[mode], _PAD[padding_mode])


def affine_label_argmax(onehot, theta, out_size=None):
    """get_argmaxed_segs(F.grid_sample(onehot, F.affine_grid(theta, ...), mode="nearest", padding_mode="zeros"))
    (dg_tta/tta/torch_utils.py:71-82) in one kernel: [B,L,D,H,W] float32 one-hot labels -> [B,1,*out_size] int64
    label map with 0 = background."""
    _lib.require_cuda_f32(onehot, "onehot")
    if onehot.dim() != 5:
        raise ValueError("affine_label_argmax expects [B,L,D,H,W]")
    B, L, Di, Hi, Wi = onehot.shape
    size = tuple(int(v) for v in (out_size[-3:] if out_size is not None else onehot.shape[-3:]))
    theta = _theta_on(onehot.device, theta, B)
    x = onehot.contiguous()
    with torch.cuda.device(x.device):
        out = torch.empty((B, 1) + size, device=x.device, dtype=torch.int64)
        rc = _lib.lib().dgtta_affine_label_argmax(x.data_ptr(), theta.data_ptr(), out.data_ptr(), B, L, Di, Hi, Wi,
                                                  size[0], size[1], size[2], _lib.stream_ptr())
        _lib.check(rc, "dgtta_affine_label_argmax")
    return out


def affine_label_gather(label_map, theta, out_size=None):
    """Nearest-mode crop of an integer label map (see label_map_from_onehot): [B,D,H,W] int16 -> [B,1,*out_size] int64,
    0 outside the volume.  Equal, voxel for voxel, to affine_label_argmax on the one-hot volume the map was built from."""
    if not (isinstance(label_map, torch.Tensor) and label_map.is_cuda and label_map.dtype == torch.int16 and label_map.dim() == 4):
        raise TypeError("label_map must be a CUDA int16 tensor [B,D,H,W] (label_map_from_onehot)")
    B, Di, Hi, Wi = label_map.shape
    size = tuple(int(v) for v in (out_size[-3:] if out_size is not None else label_map.shape[-3:]))
    theta = _theta_on(label_map.device, theta, B)
    m = label_map.contiguous()
    with torch.cuda.device(m.device):
        out = torch.empty((B, 1) + size, device=m.device, dtype=torch.int64)
        rc = _lib.lib().dgtta_affine_label_gather(m.data_ptr(), theta.data_ptr(), out.data_ptr(), B, Di, Hi, Wi,
                                                  size[0], size[1], size[2], _lib.stream_ptr())
        _lib.check(rc, "dgtta_affine_label_gather")
    return out


def label_map_from_onehot(onehot):
    """get_argmaxed_segs (dg_tta/tta/torch_utils.py:79-82) applied to a whole one-hot volume once: [B,L,D,H,W] float32
    -> [B,D,H,W] int16 with 0 = background (the L channels sum to < 1), else 1 + argmax."""
    _lib.require_cuda_f32(onehot, "onehot")
    if onehot.dim() != 5:
        raise ValueError("label_map_from_onehot expects [B,L,D,H,W]")
    B, L = onehot.shape[:2]
    x = onehot.contiguous()
    with torch.cuda.device(x.device):
        out = torch.empty((B,) + tuple(x.shape[2:]), device=x.device, dtype=torch.int16)
        rc = _lib.lib().dgtta_label_map_from_onehot(x.data_ptr(), out.data_ptr(), B, L, x[0, 0].numel(), _lib.stream_ptr())
        _lib.check(rc, "dgtta_label_map_from_onehot")
    return out


def volume_min(x):
    """x.min() as a 1-element CUDA tensor (two small kernels, no host sync); NaN-propagating like torch.min."""
    _lib.require_cuda_f32(x, "x")
    L = _lib.lib()
    x = x.contiguous()
    with torch.cuda.device(x.device):
        out = torch.empty(1, device=x.device, dtype=torch.float32)
        nbytes = L.dgtta_volume_min_workspace_bytes()
        ws = torch.empty(nbytes, device=x.device, dtype=torch.uint8)
        _lib.check(L.dgtta_volume_min(x.data_ptr(), x.numel(), out.data_ptr(), ws.data_ptr(), nbytes, _lib.stream_ptr()),
                   "dgtta_volume_min")
    return out


def affine_crop_shifted(input, theta, shift, out_size=None):
    """grid_sample(input - shift, affine_grid(theta), zeros) + shift in one gather (dg_tta/tta/torch_utils.py:58-62 with
    shift = input.min()); shift: CUDA float32 tensor with one entry per sample.  No autograd (get_batch runs under
    no_grad in the reference)."""
    _lib.require_cuda_f32(input, "input")
    _lib.require_cuda_f32(shift, "shift")
    if input.dim() != 5:
        raise ValueError("affine_crop_shifted expects [B,C,D,H,W]")
    B, C, Di, Hi, Wi = input.shape
    if shift.numel() != B:
        raise ValueError("shift must have one entry per sample")
    size = tuple(int(v) for v in (out_size[-3:] if out_size is not None else input.shape[-3:]))
    theta = _theta_on(input.device, theta, B)
    x = input.contiguous()
    with torch.cuda.device(x.device):
        out = torch.empty((B, C) + size, device=x.device, dtype=torch.float32)
        rc = _lib.lib().dgtta_affine_crop_shifted_fwd(x.data_ptr(), theta.data_ptr(), shift.contiguous().data_ptr(), out.data_ptr(),
                                                     B, C, Di, Hi, Wi, size[0], size[1], size[2], _lib.stream_ptr())
        _lib.check(rc, "dgtta_affine_crop_shifted_fwd")
    return out


_SIDE_STREAMS = {}


def gin_mind_aug(input):
    """augmentation_utils.py:173-174: MIND3D()(gin_aug(input)).

    * GIN's final rescale is deferred into MIND's loads (one pass over the volume less); the values MIND sees are
      bit-identical to the unfused chain because the same two multiplications are applied in the same order.
    * The N(0,1) field MIND adds to the edges (mind.py:150) does not depend on GIN, so it is drawn on a side stream
      while the GIN kernels run.  The generator is consumed in the reference's host order (GIN alphas first, then the
      MIND noise), so the values are the ones `gin_aug` followed by `MIND3D()` would draw."""
    _lib.require_cuda_f32(input, "input")
    if input.dim() != 5 or input.shape[1] != 1:
        raise ValueError(f"gin_mind_aug expects [B,1,D,H,W], got {tuple(input.shape)}")
    net = default_gin()
    alphas, kers, shifts = net.draw(input)                    # host draws + device rand(B), reference order
    B, _, D, H, W = input.shape
    dev = input.device
    main = torch.cuda.current_stream(dev)
    side = _SIDE_STREAMS.get(dev.index)
    if side is None:
        side = _SIDE_STREAMS[dev.index] = torch.cuda.Stream(dev)
    # persistent 12-channel noise buffer of this (device, stream): the fill below waits for everything already queued
    # on `main` — in particular the previous call's MIND, its last reader — so reuse is ordered
    noise = _lib.scratch(dev, "gin_mind_aug_noise", B * 12 * D * H * W).view(B, 12, D, H, W)
    side.wait_stream(main)
    with torch.cuda.stream(side):
        randn_like_reference((B, 12, D, H, W), dev, out=noise)   # == torch.randn(...), generated by our Philox kernel
    from ..gin import gin_forward
    mixed, scale = gin_forward(input, kers, shifts, alphas, net.interm_channel, defer_scale=True)
    main.wait_stream(side)
    return mind_ssc(mixed, noise=noise, in_scale=scale)
