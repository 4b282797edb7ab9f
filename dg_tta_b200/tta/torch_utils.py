"""Patch sampling of the TTA loop — drop-in for `get_batch` / `get_argmaxed_segs` of
dg_tta/tta/torch_utils.py:13-82.  The affine_grid + grid_sample pairs (:55-62 image, zeros padding;
:71-73 labels, nearest, fused with get_argmaxed_segs :79-82 into one label-map kernel) run in the sampler kernels
(csrc/affine_sample.cu); the host draws
(`2*rand(3)-1` per batch element, CPU generator, :39) keep the reference's order.

Volumes may already live on the device: the reference re-uploads the whole volume on every call
(:58-60); pass CUDA tensors (e.g. `[t.cuda() for t in tensor_list]` once per sample) to avoid that.
"""
import torch

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
