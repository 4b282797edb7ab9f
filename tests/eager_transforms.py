"""Test infrastructure: the reference's op sequence (torch eager) behind the same small interface as
tools/tta_standin.DropInTransforms, for kernel-vs-eager comparisons on the same GPU (tests/perf_eager_gpu.py).
Imports the torch port under oracle/ — allowed here because this file lives in tests/."""
import torch
import torch.nn.functional as F


class EagerTorchTransforms:
    """The reference's op sequence on the same GPU (torch eager: affine_grid + grid arithmetic + grid_sample,
    MIND as pad/conv3d chains) — the kernel-vs-kernel bar of SURVEY.md §2b.  Uses the torch port under oracle/,
    which is test infrastructure: only tests/ and bench legs may construct this class."""

    def __init__(self):
        import os
        from oracle import ref_port
        os.environ["DG_TTA_INTERNAL_AUGMENTATION"] = "false"
        self._rp = ref_port
        self._identity = {}
        from dg_tta_b200.tta import augmentation_utils as au
        self.get_rand_affine = au.get_rand_affine          # pure host code, identical to the reference's

    def gin_hook(self, module, input):
        return input

    def mind_hook(self, module, input):
        (x,) = input
        return self._rp.mind_ssc(x, noise=torch.randn((x.shape[0], 12) + tuple(x.shape[2:]), device=x.device))

    def warp(self, x, theta, padding):
        key = (tuple(x.shape[2:]), x.shape[0], x.device)
        if key not in self._identity:
            eye = torch.eye(4, device=x.device).repeat(x.shape[0], 1, 1)[:, :3]
            self._identity[key] = F.affine_grid(eye, [x.shape[0], 1] + list(x.shape[2:]), align_corners=False)
        return self._rp.tta_view_warp(x, theta.to(x.device), self._identity[key], padding)

    def get_batch(self, tensor_list, batch_idxs, patch_size, fixed_patch_idx=None, device="cuda"):
        b_img = []
        t_patch = torch.as_tensor(patch_size)
        t_in = torch.as_tensor(tensor_list[0].shape[-3:])
        scales = torch.cat([(t_patch / t_in).flip(0), torch.tensor([1.0])], dim=0)
        aff = scales.diag()
        for b in batch_idxs:
            data = tensor_list[b]
            off = (2.0 * torch.rand(3) - 1.0) * ((t_in - t_patch) / t_in).clip(min=0.0)
            aff[:, -1] = torch.cat([off.flip(0), torch.tensor([1.0])], dim=0)
            grid = F.affine_grid(aff[:3][None].to(device), (1, 1, *patch_size), align_corners=False)
            mn = data[0].min()
            b_img.append(F.grid_sample(data[0][None, None].to(device) - mn, grid, align_corners=False) + mn)
        return b_img, [None] * len(batch_idxs)
