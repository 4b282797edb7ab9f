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
from .augmentation_utils import affine_grid_sample, affine_label_argmax


def get_argmaxed_segs(segs):
    """torch_utils.py:79-82: prepend a background channel (no label set) and take the argmax."""
    segs_oh_w_bg = torch.cat([(segs.sum(1, keepdim=True) < 1.0).float(), segs], dim=1)
    return segs_oh_w_bg.argmax(1, keepdim=True)


def get_batch(tensor_list, batch_idxs, patch_size, fixed_patch_idx=None, device="cuda"):
    """Same arguments and return value as the reference: (list of [1,1,*patch] image patches, list of
    [1,1,*patch] int64 label patches or None).  `device` must be a CUDA device (no CPU fallback)."""
    assert fixed_patch_idx in range(8) or fixed_patch_idx is None or fixed_patch_idx == "center"
    device = torch.device(device)
    if device.type != "cuda":
        raise TypeError("dg_tta_b200.get_batch samples on a CUDA device; there is no CPU path")
    B = len(batch_idxs)
    b_img, b_label = [], []
    t_patch_size = torch.as_tensor(patch_size)
    t_input_shape = torch.as_tensor(tensor_list[0].shape[-3:])
    scales = t_patch_size / t_input_shape
    scales = torch.cat([scales.flip(0), torch.tensor([1.0])], dim=0)
    patch_affine = scales.diag()
    out_size = (int(patch_size[0]), int(patch_size[1]), int(patch_size[2]))
    with torch.no_grad():
        for b in range(B):
            data = tensor_list[batch_idxs[b]]
            if fixed_patch_idx == "center":
                pass
            else:
                rand_offset = 2.0 * torch.rand(3) - 1.0
                offset_range = ((t_input_shape - t_patch_size) / t_input_shape).clip(min=0.0)
                ranged_offset = torch.cat([(rand_offset * offset_range).flip(0), torch.tensor([1.0])], dim=0)
                patch_affine[:, -1] = ranged_offset
            theta = patch_affine[:3][None]
            data = data.to(device=device, dtype=torch.float32)
            img = data[0][None, None]
            img_min = img.min()
            img_patch = affine_grid_sample(img - img_min, theta, out_size, padding_mode="zeros") + img_min
            b_img.append(img_patch)
            if data[1:].numel() == 0:
                b_label.append(None)   # no GT label available for this sample
            else:
                # nearest-mode crop of the one-hot channels + background + argmax (:71-82), fused: no L-channel patch
                b_label.append(affine_label_argmax(data[1:][None], theta, out_size))
    return b_img, b_label


def soft_dice_loss(smp_a, smp_b):
    """torch_utils.py:90-104, unchanged (torch ops): per-(sample, class) soft Dice of two masked softmax maps."""
    B, _, D, H, W = smp_a.shape
    nominator = (2.0 * smp_a * smp_b).reshape(B, -1, D * H * W).mean(2)
    denominator = 0.5 * ((smp_a + smp_b) ** 2).reshape(B, -1, D * H * W).mean(2)
    if denominator.sum() == 0.0:
        return (nominator * 0.0) + 1.0
    return nominator / denominator   # "Do not add an eps here, it disturbs the consistency"


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
