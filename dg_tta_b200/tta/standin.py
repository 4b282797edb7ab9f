"""Stand-in for DG-TTA's inner adaptation loop, driven by the drop-in ops.

`dg_tta.tta.tta` cannot be imported without nnunetv2, so BASELINE configs 3 and 5 are measured on
this restatement of `tta_main`'s loop (dg_tta/tta/tta.py:190-281) and `calc_branch` (:480-579) with TEMPLATE_PLAN
defaults (dg_tta/tta/config_log_utils.py:24-41: affine view augmentation in both branches, gradient in branch_a,
intensity augmentation off, GIN hook disabled during TTA, MIND hook on, AdamW lr 1e-5, 16 accumulated patches):

    get_batch -> [per branch: get_rand_affine -> affine_grid_sample(border) -> model (mind_hook -> UNet)
                  -> channel selection (map_label, logits) -> affine_grid_sample(zeros, differentiable)]
              -> common-content mask, softmax, soft Dice (torch_utils.py:90-104) -> backward (branch_a only)

The backbone is a PlainConvUNet-shaped fixture built from plans.json:279-401 (5 stages, 32/64/128/256/320 features,
2 convs per stage, strides 1,2,2,2,2, InstanceNorm + LeakyReLU, transposed-conv upsampling, 12 input channels for the
MIND trainers, nnUNetTrainer_GIN_MIND.py:46) with random weights; it stays on PyTorch/cuDNN — out of scope for the CUDA
work, here only so that the transforms are timed (and their parity is judged) in their real consumer.

`ViewGraph` captures the whole pre-network transform segment of a step (patch crops -> two view warps -> two Philox
noise fields -> two MIND descriptors) in ONE CUDA graph: the per-step random state (crop offsets, the two affines,
the generator's seed/offset) is written into a pinned parameter block on the host and uploaded by a memcpy node
inside the graph, so a step costs one graph launch instead of ~25 kernel launches plus their Python glue, and there is
no host synchronisation anywhere in the segment.  Host draws keep the reference's order and generators.
"""
import numpy as np
import torch
import torch.nn as nn

from .. import _lib


class _Block(nn.Sequential):
    def __init__(self, cin, cout, stride):
        super().__init__(nn.Conv3d(cin, cout, 3, stride, 1), nn.InstanceNorm3d(cout, eps=1e-5, affine=True), nn.LeakyReLU(0.01, True),
                         nn.Conv3d(cout, cout, 3, 1, 1), nn.InstanceNorm3d(cout, eps=1e-5, affine=True), nn.LeakyReLU(0.01, True))


class StandInUNet(nn.Module):
    def __init__(self, in_channels=12, num_classes=105, features=(32, 64, 128, 256, 320)):
        super().__init__()
        self.enc = nn.ModuleList()
        c = in_channels
        for i, f in enumerate(features):
            self.enc.append(_Block(c, f, 1 if i == 0 else 2))
            c = f
        self.up, self.dec = nn.ModuleList(), nn.ModuleList()
        for f in reversed(features[:-1]):
            self.up.append(nn.ConvTranspose3d(c, f, 2, 2))
            self.dec.append(_Block(2 * f, f, 1))
            c = f
        self.head = nn.Conv3d(c, num_classes, 1)

    def forward(self, x):
        skips = []
        for blk in self.enc:
            x = blk(x)
            skips.append(x)
        skips.pop()
        for up, dec in zip(self.up, self.dec):
            x = dec(torch.cat([up(x), skips.pop()], 1))
        return self.head(x)


def build_model(transforms, in_channels=12, num_classes=105, features=(32, 64, 128, 256, 320), seed=0):
    """transforms: namespace with gin_hook / mind_hook (this package, or a torch-eager restatement of the reference)."""
    torch.manual_seed(seed)
    net = StandInUNet(in_channels, num_classes, features)
    net.register_forward_pre_hook(transforms.gin_hook)     # registration order of nnUNetTrainer_GIN_MIND.py:55-57
    net.register_forward_pre_hook(transforms.mind_hook)
    return net


def _reference_soft_dice(smp_a, smp_b):
    from .torch_utils import soft_dice_loss
    return soft_dice_loss(smp_a, smp_b)


def calc_branch(model, imgs, optimized_idx, with_grad, transforms, warp_back=True):
    """tta.py:480-579 for spatial_aug_type='affine', do_spatial_aug_in='both', no intensity augmentation.
    warp_back=False returns (selected logits, R_inverse) for a consumer that fuses the inverse warp."""
    ctx = torch.enable_grad() if with_grad else torch.no_grad()
    with ctx:
        R, R_inverse = transforms.get_rand_affine(imgs.shape[0], flip=False)
        imgs_aug = transforms.warp(imgs, R, "border")                       # tta.py:549-551
        target = model(imgs_aug)                                            # pre-hooks: gin (off), mind
        target = target.transpose(0, 1)[optimized_idx].transpose(0, 1)      # map_label(..., "logits"), torch_utils.py:214-222
        if not warp_back:
            return target, R_inverse
        return transforms.warp(target, R_inverse, "zeros")                  # tta.py:573-575


def consistency(target_a, target_b, transforms):
    fused = getattr(transforms, "consistency_loss", None)
    if fused is not None:                                                    # drop-in: mask + softmaxes + Dice sums in one pass
        return fused(target_a, target_b, 1)
    mask = (target_a.sum(1, keepdim=True) > 0.0).float() * (target_b.sum(1, keepdim=True) > 0.0).float()   # tta.py:263-269
    sm_a = target_a.softmax(1) * mask
    sm_b = target_b.softmax(1) * mask
    return 1 - _reference_soft_dice(sm_a, sm_b)[:, 1:].mean()                # START_CLASS = 1


def tta_inner_step(model, volumes, patch_size, batch_size, optimized_idx, transforms, accum=16, rng=None):
    """One accumulation iteration of tta.py:221-275.  Returns the detached loss (device tensor; the reference syncs
    it to the host every step, tta.py:272 — the caller decides)."""
    idx = (rng or np.random).choice(range(len(volumes)), batch_size).tolist()
    with torch.no_grad():
        imgs, _ = transforms.get_batch(volumes, idx, patch_size, fixed_patch_idx=None, device=volumes[0].device)
    imgs = torch.cat(imgs, dim=0)
    fused = getattr(transforms, "consistency_loss_warped", None)
    if fused is not None and len(optimized_idx) <= 16:
        # drop-in: inverse warps + mask + softmaxes + Dice sums in one pass (and one for the gradient)
        la, Ra_inv = calc_branch(model, imgs, optimized_idx, True, transforms, warp_back=False)    # have_grad_in = branch_a
        lb, Rb_inv = calc_branch(model, imgs, optimized_idx, False, transforms, warp_back=False)
        loss = fused(la, lb, Ra_inv, Rb_inv, 1)
    else:
        target_a = calc_branch(model, imgs, optimized_idx, True, transforms)
        target_b = calc_branch(model, imgs, optimized_idx, False, transforms)
        loss = consistency(target_a, target_b, transforms)
    (loss / accum).backward()
    return loss.detach()


class DropInTransforms:
    """The B200 drop-in (dg_tta_b200) behind the small interface the loop needs."""

    def __init__(self, fused_warp=False):
        # fused_warp: route the loss through consistency_dice_loss_warped (inverse warps inside the reductions, nothing
        # materialised).  Measured on B200 (2x14x128^3, fwd+bwd): 2.81 ms vs 1.80 ms for warp -> consistency_dice_loss —
        # holding all channels of both branches per voxel caps the latency-bound gather at 16 warps/SM — so the default
        # is the unfused (faster) sequence; see profiles/experiments/README.md.
        from .. import gin, mind, utils
        from . import augmentation_utils as au
        from . import torch_utils as tu
        utils.disable_internal_augmentation()           # tta.py:154
        self.gin_hook, self.mind_hook = gin.gin_hook, mind.mind_hook
        self.get_rand_affine, self.get_batch = au.get_rand_affine, tu.get_batch
        self._sample = au.affine_grid_sample
        self.consistency_loss = tu.consistency_dice_loss
        if fused_warp:
            self.consistency_loss_warped = tu.consistency_dice_loss_warped

    def warp(self, x, theta, padding):
        return self._sample(x, theta, padding_mode=padding)


class ViewGraph:
    """The pre-network transform segment of one TTA step as a single CUDA graph.

    Captured once per (volume, patch, batch): memcpy(params) -> B image crops (min shift folded) -> view warp a (border)
    -> Philox field a -> MIND a -> view warp b -> Philox field b -> MIND b.  `step()` performs the host draws of the
    reference step in the reference's order (CPU generator: crop offsets, R_a, R_b; device generator: the two MIND
    noise fields, consumed as (seed, offset) pairs), writes them into the pinned parameter block and replays the graph.
    Returns (desc_a, desc_b, R_a_inverse, R_b_inverse): the two 12-channel network inputs (static tensors, valid until
    the next step) and the inverse affines for the prediction warps.  Values are bit-identical to the eager call
    sequence get_batch -> affine_grid_sample(border) -> MIND3D() from the same seeds."""

    def __init__(self, volume, patch_size, batch_size, device=None):
        from . import augmentation_utils as au
        from .torch_utils import resident_volume
        from ..mind import mind_ssc
        L = _lib.lib()
        self.device = torch.device(device) if device is not None else volume.device
        if self.device.type != "cuda":
            raise TypeError("ViewGraph needs a CUDA device")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.B = B = int(batch_size)
        self.patch = tuple(int(v) for v in patch_size)
        self.in_shape = tuple(volume.shape[-3:])
        self.vol = resident_volume(volume, self.device)
        self._keep = volume
        props = torch.cuda.get_device_properties(self.device)
        self._sms, self._mt = props.multi_processor_count, props.max_threads_per_multi_processor
        n_noise = B * 12 * self.patch[0] * self.patch[1] * self.patch[2]
        self._inc = L.dgtta_philox_normal_offset_increment(n_noise, self._sms, self._mt)
        # parameter block: 3*B thetas (crop, R_a, R_b; 12 floats each) then 2 x {seed, offset} (uint64)
        nfloat = 3 * B * 12
        nfloat += nfloat % 2                                  # 8-byte alignment of the uint64 tail
        self._nfloat = nfloat
        self.h_params = torch.zeros(nfloat * 4 + 32, dtype=torch.uint8).pin_memory()
        self.h_theta = self.h_params[:3 * B * 48].view(torch.float32).view(3, B, 3, 4)
        self.h_state = self.h_params[nfloat * 4:].numpy().view(np.uint64)    # shares the pinned memory
        with torch.cuda.device(self.device):
            self.d_params = torch.zeros(nfloat * 4 + 32, dtype=torch.uint8, device=self.device)
            d_theta = self.d_params[:3 * B * 48].view(torch.float32).view(3, B, 3, 4)
            d_state = self.d_params[nfloat * 4:].view(torch.int64)
            noise = self.noise = [torch.empty((B, 12) + self.patch, device=self.device) for _ in range(2)]   # graph inputs: keep alive
            vol = self.vol

            def body():
                self.d_params.copy_(self.h_params, non_blocking=True)
                crops = [au.affine_crop_shifted(vol.img, d_theta[0, b:b + 1], vol.img_min, self.patch) for b in range(B)]
                imgs = torch.cat(crops, dim=0)
                descs = []
                for k in range(2):
                    view = au.affine_grid_sample(imgs, d_theta[1 + k], padding_mode="border")
                    _lib.check(L.dgtta_philox_normal_fill_graphsafe(noise[k].data_ptr(), n_noise, d_state[2 * k:].data_ptr(),
                                                                    self._sms, self._mt, _lib.stream_ptr()),
                               "dgtta_philox_normal_fill_graphsafe")
                    descs.append(mind_ssc(view, noise=noise[k]))
                return imgs, descs

            self._write_params(torch.eye(3, 4).repeat(3, B, 1, 1), 0, 0)
            side = torch.cuda.Stream(self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):                      # warm-up outside capture (lazy module loads, allocator)
                body()
            torch.cuda.current_stream(self.device).wait_stream(side)
            torch.cuda.synchronize(self.device)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.imgs, (self.desc_a, self.desc_b) = body()
        self.replays = 0

    def _write_params(self, thetas, seed, offset):
        self.h_theta.copy_(thetas)
        self.h_state[:] = (seed % 2 ** 64, offset, seed % 2 ** 64, offset + self._inc)

    def step(self):
        from .augmentation_utils import get_rand_affine
        from .torch_utils import patch_affines
        from ..mind import _device_generator
        crop = patch_affines(self.in_shape, self.patch, self.B)          # CPU draws, get_batch order
        R_a, R_a_inv = get_rand_affine(self.B, flip=False)                # calc_branch(a)
        R_b, R_b_inv = get_rand_affine(self.B, flip=False)                # calc_branch(b)
        gen = _device_generator(self.device)
        seed, offset = gen.initial_seed(), gen.get_offset()
        if offset % 4:
            raise _lib.DgttaError("generator offset is not a multiple of 4: the Philox stream cannot be continued")
        gen.set_offset(offset + 2 * self._inc)                            # the two randn_like draws of mind.py:150
        self._write_params(torch.stack([crop, R_a, R_b]), seed, offset)
        self.graph.replay()
        self.replays += 1
        return self.desc_a, self.desc_b, R_a_inv, R_b_inv


def tta_inner_step_graphed(model_no_hooks, views, optimized_idx, transforms, accum=16):
    """The same accumulation iteration with the transform segment replayed from `views` (a ViewGraph).  The network is
    called on the descriptors directly (no mind_hook: MIND already ran inside the graph)."""
    desc_a, desc_b, R_a_inv, R_b_inv = views.step()
    sel = lambda t: t.transpose(0, 1)[optimized_idx].transpose(0, 1)
    fused = getattr(transforms, "consistency_loss_warped", None)
    la = sel(model_no_hooks(desc_a))
    with torch.no_grad():
        lb = sel(model_no_hooks(desc_b))
    if fused is not None and len(optimized_idx) <= 16:
        loss = fused(la, lb, R_a_inv, R_b_inv, 1)
    else:
        with torch.no_grad():
            target_b = transforms.warp(lb, R_b_inv, "zeros")
        loss = consistency(transforms.warp(la, R_a_inv, "zeros"), target_b, transforms)
    (loss / accum).backward()
    return loss.detach()


def run_adaptation(model, volumes, patch_size, batch_size, optimized_idx, transforms, epochs=2, accum=16, lr=1e-5, rng=None):
    """tta.py:190-281: `epochs` x `accum` accumulated patches, one AdamW step per epoch.  Returns the per-step losses
    (one host read per epoch, not per step)."""
    opt = torch.optim.AdamW(model.parameters(), lr=lr)
    losses = []
    model.train()
    for _ in range(epochs):
        opt.zero_grad(set_to_none=True)
        step_losses = [tta_inner_step(model, volumes, patch_size, batch_size, optimized_idx, transforms, accum, rng)
                       for _ in range(accum)]
        opt.step()
        losses.extend(float(v) for v in torch.stack(step_losses).cpu())
    return losses


def dice_per_class(pred, target, num_classes):
    """Hard Dice in percent points per class 1..num_classes-1 (NaN where the class is absent from both)."""
    out = []
    for c in range(1, num_classes):
        p, t = pred == c, target == c
        den = int(p.sum()) + int(t.sum())
        out.append(float("nan") if den == 0 else 200.0 * int((p & t).sum()) / den)
    return out
