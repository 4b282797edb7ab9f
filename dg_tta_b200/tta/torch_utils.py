"""Patch sampling of the TTA loop — drop-in for `get_batch` / `get_argmaxed_segs` of
dg_tta/tta/torch_utils.py:13-82.  The affine_grid + grid_sample pairs (:55-62 image, zeros padding;
:71-73 labels, nearest, fused with get_argmaxed_segs :79-82 into one label-map kernel) run in the sampler kernels
(csrc/affine_sample.cu); the host draws
(`2*rand(3)-1` per batch element, CPU generator, :39) keep the reference's order.

Volumes may already live on the device: the reference re-uploads the whole volume on every call
(:58-60); pass CUDA tensors (e.g. `[t.cuda() for t in tensor_list]` once per sample) to avoid that.
"""
import torch

from .. import _lib
from .augmentation_utils import (affine_crop_shifted, affine_grid_sample, affine_label_argmax, affine_label_gather,
                                 label_map_from_onehot, volume_min)


def get_argmaxed_segs(segs):
    """Same result as torch_utils.py:79-82 ([B,L,D,H,W] one-hot -> [B,1,D,H,W] int64: 0 where the L channels sum to
    < 1, else 1 + argmax, lowest index on ties), evaluated by the label-map kernel in one pass instead of
    sum / cat / argmax over an (L+1)-channel copy."""
    return label_map_from_onehot(segs)[:, None].long()


class _ResidentVolume:
    """Per-volume state get_batch needs on every call but which only depends on the volume: the device copy of the
    image channel, its minimum (torch_utils.py:58) and the int16 label map that replaces the one-hot channels
    (torch_utils.py:71-82, see augmentation_utils.label_map_from_onehot).  The reference recomputes / re-uploads all of
    it per call (SURVEY.md §7 hard part 8)."""
    __slots__ = ("version", "device", "img", "img_min", "label_map", "ref")

    def __init__(self, data, device):
        self.version = data._version
        self.device = device
        d = data.to(device=device, dtype=torch.float32)
        self.img = d[0][None, None].contiguous()
        self.img_min = volume_min(self.img)
        self.label_map = label_map_from_onehot(d[1:][None]) if d[1:].numel() else None


_RESIDENT = {}
_RESIDENT_MAX = 16


def resident_volume(data, device):
    """Cached _ResidentVolume of a sample tensor ([1+L, D, H, W], CPU or CUDA).  Keyed on the tensor object (weak
    reference: the entry dies with the tensor) and its version counter (in-place edits rebuild it)."""
    import weakref
    key = (id(data), str(device))
    ent = _RESIDENT.get(key)
    if ent is not None and ent.ref() is data and ent.version == data._version:
        return ent
    ent = _ResidentVolume(data, device)
    ent.ref = weakref.ref(data, lambda _r, k=key: _RESIDENT.pop(k, None))
    if len(_RESIDENT) >= _RESIDENT_MAX:
        _RESIDENT.pop(next(iter(_RESIDENT)))
    _RESIDENT[key] = ent
    return ent


def release_resident_volumes():
    _RESIDENT.clear()


def patch_affines(input_shape, patch_size, n, fixed_patch_idx=None):
    """The host arithmetic of get_batch (torch_utils.py:23-45): n patch affines [n,3,4] (CPU float32) for crops of
    `patch_size` out of a volume of `input_shape`, drawing `2*rand(3)-1` per patch from the CPU generator in the
    reference's order (no draw for the centre crop)."""
    t_patch_size = torch.as_tensor(patch_size)
    t_input_shape = torch.as_tensor(tuple(input_shape))
    scales = t_patch_size / t_input_shape
    scales = torch.cat([scales.flip(0), torch.tensor([1.0])], dim=0)
    patch_affine = scales.diag()
    thetas = []
    for _ in range(n):
        if fixed_patch_idx == "center":
            pass
        else:
            rand_offset = 2.0 * torch.rand(3) - 1.0
            offset_range = ((t_input_shape - t_patch_size) / t_input_shape).clip(min=0.0)
            ranged_offset = torch.cat([(rand_offset * offset_range).flip(0), torch.tensor([1.0])], dim=0)
            patch_affine[:, -1] = ranged_offset
        thetas.append(patch_affine[:3].clone())
    return torch.stack(thetas).to(torch.float32)


def get_batch(tensor_list, batch_idxs, patch_size, fixed_patch_idx=None, device="cuda"):
    """Same arguments and return value as the reference: (list of [1,1,*patch] image patches, list of
    [1,1,*patch] int64 label patches or None).  `device` must be a CUDA device (no CPU fallback).

    Per call and batch element this is two small gathers: the image crop with the `- min ... + min` shift folded in
    (one kernel, the minimum cached per volume) and the label crop from the cached int16 label map."""
    assert fixed_patch_idx in range(8) or fixed_patch_idx is None or fixed_patch_idx == "center"
    device = torch.device(device)
    if device.type != "cuda":
        raise TypeError("dg_tta_b200.get_batch samples on a CUDA device; there is no CPU path")
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    B = len(batch_idxs)
    b_img, b_label = [], []
    out_size = (int(patch_size[0]), int(patch_size[1]), int(patch_size[2]))
    thetas = patch_affines(tensor_list[0].shape[-3:], patch_size, B, fixed_patch_idx)
    with torch.no_grad():
        for b in range(B):
            vol = resident_volume(tensor_list[batch_idxs[b]], device)
            theta = thetas[b][None]
            b_img.append(affine_crop_shifted(vol.img, theta, vol.img_min, out_size))
            if vol.label_map is None:
                b_label.append(None)   # no GT label available for this sample
            else:
                b_label.append(affine_label_gather(vol.label_map, theta, out_size))
    return b_img, b_label


def soft_dice_loss(smp_a, smp_b):
    """Per-(sample, class) soft Dice of two probability maps, the quantity torch_utils.py:90-104 returns:
    dice[b,c] = mean_v(2 a b) / (0.5 mean_v((a+b)^2)), all ones when every denominator is 0, no epsilon.  API mirror
    for callers that already hold the two masked softmax maps; the TTA step itself uses consistency_dice_loss, which
    never materialises them."""
    n_vox = smp_a[0, 0].numel()
    a, b = smp_a.flatten(2), smp_b.flatten(2)
    num = (a * b).mul(2.0).sum(2) / n_vox
    den = (a + b).square().sum(2).mul(0.5) / n_vox
    if bool(den.sum() == 0.0):
        return num * 0.0 + 1.0          # stays attached to the graph (zero gradient), like the reference's special case
    return num / den


class _ConsistencySums(torch.autograd.Function):
    """sums[b,c] = (sum_v 2 sm_a sm_b, sum_v (sm_a + sm_b)^2) with the common-content mask and both channel softmaxes
    computed on the fly (csrc/consistency_loss.cu); differentiable w.r.t. either logit tensor."""

    @staticmethod
    def forward(ctx, target_a, target_b):
        L = _lib.lib()
        a, b = target_a.contiguous(), target_b.contiguous()
        B, C = a.shape[:2]
        V = a[0, 0].numel()
        with torch.cuda.device(a.device):
            sums = torch.empty((B, C, 2), device=a.device, dtype=torch.float64)
            _lib.check(L.dgtta_consistency_sums_fwd(a.data_ptr(), b.data_ptr(), sums.data_ptr(), B, C, V,
                                                    _lib.stream_ptr()), "dgtta_consistency_sums_fwd")
        ctx.save_for_backward(a, b)
        return sums.to(torch.float32)

    @staticmethod
    def backward(ctx, grad_sums):
        a, b = ctx.saved_tensors
        L = _lib.lib()
        B, C = a.shape[:2]
        V = a[0, 0].numel()
        g = grad_sums.contiguous().to(torch.float32)
        grads = [None, None]
        with torch.cuda.device(a.device):
            for i, (x, y) in enumerate(((a, b), (b, a))):       # the sums are symmetric in the two branches
                if ctx.needs_input_grad[i]:
                    gx = torch.empty_like(x)
                    _lib.check(L.dgtta_consistency_sums_bwd(x.data_ptr(), y.data_ptr(), g.data_ptr(), gx.data_ptr(), B, C,
                                                            V, _lib.stream_ptr()), "dgtta_consistency_sums_bwd")
                    grads[i] = gx
        return tuple(grads)


def consistency_dice_loss(target_a, target_b, start_class=1):
    """The consistency loss of the TTA step (dg_tta/tta/tta.py:263-269):
        mask = (target_a.sum(1) > 0) * (target_b.sum(1) > 0); sm_x = target_x.softmax(1) * mask
        loss = 1 - soft_dice_loss(sm_a, sm_b)[:, start_class:].mean()
    with mask, both softmaxes and the per-(sample, class) sums in one pass over the two [B,C,D,H,W] logit tensors
    (and one pass for the gradient), instead of ~10 elementwise passes.  Same special cases as soft_dice_loss."""
    _lib.require_cuda_f32(target_a, "target_a")
    _lib.require_cuda_f32(target_b, "target_b")
    if target_a.shape != target_b.shape or target_a.dim() != 5:
        raise ValueError("consistency_dice_loss expects two [B,C,D,H,W] tensors of the same shape")
    V = target_a[0, 0].numel()
    sums = _ConsistencySums.apply(target_a, target_b)
    nominator = sums[..., 0] / V
    denominator = 0.5 * sums[..., 1] / V
    if denominator.sum() == 0.0:
        dice = (nominator * 0.0) + 1.0
    else:
        dice = nominator / denominator
    return 1 - dice[:, start_class:].mean()


class _ConsistencyWarpSums(torch.autograd.Function):
    """_ConsistencySums with the two inverse warps fused in (csrc/consistency_loss.cu, *_warp kernels): the inputs are the
    un-warped logits and the inverse affines; differentiable w.r.t. either logit tensor."""

    @staticmethod
    def forward(ctx, logits_a, logits_b, theta_a, theta_b):
        L = _lib.lib()
        a, b = logits_a.contiguous(), logits_b.contiguous()
        B, C, D, H, W = a.shape
        with torch.cuda.device(a.device):
            sums = torch.empty((B, C, 2), device=a.device, dtype=torch.float64)
            _lib.check(L.dgtta_consistency_warp_sums_fwd(a.data_ptr(), b.data_ptr(), theta_a.data_ptr(), theta_b.data_ptr(),
                                                         sums.data_ptr(), B, C, D, H, W, _lib.stream_ptr()),
                       "dgtta_consistency_warp_sums_fwd")
        ctx.save_for_backward(a, b, theta_a, theta_b)
        return sums.to(torch.float32)

    @staticmethod
    def backward(ctx, grad_sums):
        a, b, theta_a, theta_b = ctx.saved_tensors
        L = _lib.lib()
        B, C, D, H, W = a.shape
        g = grad_sums.contiguous().to(torch.float32)
        grads = [None, None, None, None]
        with torch.cuda.device(a.device):
            for i, (x, y, tx, ty) in enumerate(((a, b, theta_a, theta_b), (b, a, theta_b, theta_a))):   # symmetric in the branches
                if ctx.needs_input_grad[i]:
                    gx = torch.empty_like(x)
                    _lib.check(L.dgtta_consistency_warp_sums_bwd(x.data_ptr(), y.data_ptr(), tx.data_ptr(), ty.data_ptr(),
                                                                 g.data_ptr(), gx.data_ptr(), B, C, D, H, W, _lib.stream_ptr()),
                               "dgtta_consistency_warp_sums_bwd")
                    grads[i] = gx
        return tuple(grads)


def consistency_dice_loss_warped(logits_a, logits_b, theta_a, theta_b, start_class=1):
    """consistency_dice_loss(affine_grid_sample(logits_a, theta_a), affine_grid_sample(logits_b, theta_b), start_class)
    — the inverse warps of dg_tta/tta/tta.py:571-575 and the loss of :263-269 — without materialising the warped logits.
    theta_*: [B,3,4] inverse affines (host or device).  More than 16 channels fall back to warp + consistency_dice_loss
    (still this library's kernels)."""
    from .augmentation_utils import _theta_on
    _lib.require_cuda_f32(logits_a, "logits_a")
    _lib.require_cuda_f32(logits_b, "logits_b")
    if logits_a.shape != logits_b.shape or logits_a.dim() != 5:
        raise ValueError("consistency_dice_loss_warped expects two [B,C,D,H,W] tensors of the same shape")
    B, C = logits_a.shape[:2]
    if C > 16:
        return consistency_dice_loss(affine_grid_sample(logits_a, theta_a), affine_grid_sample(logits_b, theta_b), start_class)
    theta_a = _theta_on(logits_a.device, theta_a, B)
    theta_b = _theta_on(logits_a.device, theta_b, B)
    V = logits_a[0, 0].numel()
    sums = _ConsistencyWarpSums.apply(logits_a, logits_b, theta_a, theta_b)
    nominator = sums[..., 0] / V
    denominator = 0.5 * sums[..., 1] / V
    if denominator.sum() == 0.0:
        dice = (nominator * 0.0) + 1.0
    else:
        dice = nominator / denominator
    return 1 - dice[:, start_class:].mean()
