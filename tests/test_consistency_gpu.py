"""Fused consistency loss (dg_tta_b200.tta.torch_utils.consistency_dice_loss) against the reference's op chain
(dg_tta/tta/tta.py:263-269 + torch_utils.py:90-104) written out in torch: loss value and gradient w.r.t. target_a."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from gpu_util import cuda

pytestmark = pytest.mark.gpu


def test_golden_from_the_reference_soft_dice_loss():
    """fixture generated with the reference's own soft_dice_loss (tests/golden/make_golden.py::gen_consistency)"""
    from dg_tta_b200.tta.torch_utils import consistency_dice_loss
    g = load_golden("consistency")
    ta = cuda(g["target_a"]).requires_grad_(True)
    loss = consistency_dice_loss(ta, cuda(g["target_b"]))
    assert abs(loss.item() - float(g["loss"])) <= 1e-5
    loss.backward()
    assert np.abs(ta.grad.cpu().numpy() - g["grad_a"]).max() <= 2e-5 * np.abs(g["grad_a"]).max()


def test_label_argmax_golden_identity_crop():
    """get_argmaxed_segs fixture from the reference; an identity crop must reproduce it exactly"""
    from dg_tta_b200.tta.augmentation_utils import affine_label_argmax
    g = load_golden("argmaxed_segs")
    theta = torch.eye(3, 4)[None]
    out = affine_label_argmax(cuda(g["segs"]), theta)
    assert np.array_equal(out.cpu().numpy(), g["out"])


def reference_loss(target_a, target_b, start_class=1):
    mask = (target_a.sum(1, keepdim=True) > 0.0).float() * (target_b.sum(1, keepdim=True) > 0.0).float()
    sm_a = target_a.softmax(1) * mask
    sm_b = target_b.softmax(1) * mask
    B, _, D, H, W = sm_a.shape
    nominator = (2.0 * sm_a * sm_b).reshape(B, -1, D * H * W).mean(2)
    denominator = 0.5 * ((sm_a + sm_b) ** 2).reshape(B, -1, D * H * W).mean(2)
    dice = (nominator * 0.0) + 1.0 if denominator.sum() == 0.0 else nominator / denominator
    return 1 - dice[:, start_class:].mean()


@pytest.mark.parametrize("shape", [(2, 5, 12, 14, 16), (1, 14, 20, 24, 28), (2, 17, 9, 10, 11), (1, 40, 6, 7, 8),
                                   (2, 14, 64, 64, 64)])
def test_loss_and_gradient_match_the_reference_chain(shape):
    from dg_tta_b200.tta.torch_utils import consistency_dice_loss
    g = torch.Generator(device="cuda").manual_seed(shape[1])
    a = torch.randn(shape, device="cuda", generator=g) * 2 + 0.3
    b = torch.randn(shape, device="cuda", generator=g) * 2 + 0.3
    a[:, :, :3] = 0.0                      # warped-in zeros: outside the common content (sum == 0 -> masked)
    b[:, :, :, -2:] = 0.0
    a1 = a.clone().requires_grad_(True)
    a2 = a.clone().requires_grad_(True)
    ref = reference_loss(a1, b)
    got = consistency_dice_loss(a2, b)
    assert abs(got.item() - ref.item()) <= 1e-5
    ref.backward()
    got.backward()
    scale = float(a1.grad.abs().max())
    assert scale > 0
    assert float((a2.grad - a1.grad).abs().max()) <= 2e-5 * scale + 1e-10
    assert float(a2.grad[:, :, :3].abs().max()) == 0.0          # masked voxels get no gradient


def test_gradient_for_both_branches_and_empty_overlap():
    from dg_tta_b200.tta.torch_utils import consistency_dice_loss
    a = torch.randn(1, 6, 8, 8, 8, device="cuda").requires_grad_(True)
    b = torch.randn(1, 6, 8, 8, 8, device="cuda").requires_grad_(True)
    a_ref, b_ref = a.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)
    consistency_dice_loss(a, b).backward()
    reference_loss(a_ref, b_ref).backward()
    for x, y in ((a, a_ref), (b, b_ref)):
        assert float((x.grad - y.grad).abs().max()) <= 2e-5 * float(y.grad.abs().max()) + 1e-10
    # no common content at all: denominator.sum() == 0 -> dice = 1 -> loss = 0 (torch_utils.py:97-98)
    z = torch.zeros(1, 6, 4, 4, 4, device="cuda")
    assert float(consistency_dice_loss(z, z)) == 0.0


def test_against_the_c_oracle_fp64_truth():
    """CUDA path vs the plain-C restatement evaluated in double precision (oracle/dgtta_oracle_impl.h)"""
    from dg_tta_b200.tta.torch_utils import consistency_dice_loss
    from oracle import cform
    g = torch.Generator().manual_seed(12)
    a = torch.randn(2, 9, 14, 18, 20, generator=g) * 3 + 0.2
    b = torch.randn(2, 9, 14, 18, 20, generator=g) * 3 + 0.2
    a[:, :, -2:] = 0.0
    truth_loss, truth_grad = cform.consistency_loss(a.numpy(), b.numpy(), precision="f64")
    ad = a.cuda().requires_grad_(True)
    loss = consistency_dice_loss(ad, b.cuda())
    loss.backward()
    assert abs(loss.item() - truth_loss) <= 1e-5
    assert np.abs(ad.grad.cpu().numpy() - truth_grad).max() <= 2e-5 * np.abs(truth_grad).max()


def test_label_argmax_against_the_c_oracle():
    from dg_tta_b200.tta.augmentation_utils import affine_label_argmax, get_rand_affine
    from oracle import cform
    g = torch.Generator().manual_seed(5)
    lab = (torch.rand(2, 6, 16, 18, 20, generator=g) > 0.75).float()
    torch.manual_seed(8)
    R, _ = get_rand_affine(2, strength=0.15)
    got = affine_label_argmax(lab.cuda(), R, (12, 20, 9)).cpu().numpy()
    ref = cform.label_argmax(lab.numpy(), R.numpy(), (12, 20, 9))
    assert np.array_equal(got, ref)              # index work is bit-exact: kernel and oracle execute torch's coordinate arithmetic


def test_fused_inverse_warp_loss_matches_unfused_and_oracle():
    """consistency_dice_loss_warped (inverse warps fused into the reductions, tta.py:571-575 + :263-269) against
    (a) the unfused product path warp -> consistency_dice_loss and (b) the C oracle chain in double precision:
    loss value and d loss / d logits_a (and, with the roles swapped, d loss / d logits_b)."""
    from dg_tta_b200.tta.augmentation_utils import affine_grid_sample, get_rand_affine
    from dg_tta_b200.tta.torch_utils import consistency_dice_loss, consistency_dice_loss_warped
    from oracle import cform
    g = torch.Generator().manual_seed(21)
    B, C, shape = 2, 6, (18, 22, 40)
    # |N(0,1)| logits: the channel sum is far from 0 inside the volume, so the common-content mask (sum > 0) is decided
    # by the zero padding alone and cannot flip between fp32 and fp64 interpolation
    la = torch.randn((B, C) + shape, generator=g).abs() * 2 + 0.4
    lb = torch.randn((B, C) + shape, generator=g).abs() * 2 + 0.4
    torch.manual_seed(3)
    _, Ra = get_rand_affine(B, strength=0.08)
    _, Rb = get_rand_affine(B, strength=0.08)
    a1 = la.cuda().requires_grad_(True)
    b1 = lb.cuda().requires_grad_(True)
    fused = consistency_dice_loss_warped(a1, b1, Ra, Rb)
    fused.backward()
    a2 = la.cuda().requires_grad_(True)
    b2 = lb.cuda().requires_grad_(True)
    unfused = consistency_dice_loss(affine_grid_sample(a2, Ra), affine_grid_sample(b2, Rb))
    unfused.backward()
    assert abs(fused.item() - unfused.item()) <= 2e-6
    for got, ref in ((a1.grad, a2.grad), (b1.grad, b2.grad)):
        scale = float(ref.abs().max())
        assert scale > 0 and float((got - ref).abs().max()) <= 2e-5 * scale
    # oracle chain: warp (f64) -> loss + gradient w.r.t. the warped logits (f64) -> adjoint warp (f64)
    wa = cform.affine_sample(la.numpy(), Ra.numpy(), la.shape, precision="f64")
    wb = cform.affine_sample(lb.numpy(), Rb.numpy(), lb.shape, precision="f64")
    loss64, gwa = cform.consistency_loss(wa.astype(np.float32), wb.astype(np.float32), precision="f64")
    ga = cform.affine_sample_bwd_input(gwa.astype(np.float32), Ra.numpy(), la.shape, precision="f64")
    assert abs(fused.item() - loss64) <= 1e-5
    assert np.abs(a1.grad.cpu().numpy() - ga).max() <= 1e-4 * np.abs(ga).max()


def test_fused_inverse_warp_loss_shapes_and_fallback():
    """ragged sizes (W not a multiple of 32, H not of 8), C = 14 (the bench's class subset) and the > 16-channel fallback"""
    from dg_tta_b200.tta.augmentation_utils import affine_grid_sample, get_rand_affine
    from dg_tta_b200.tta.torch_utils import consistency_dice_loss, consistency_dice_loss_warped
    g = torch.Generator().manual_seed(5)
    for C, shape in ((14, (9, 13, 37)), (3, (7, 8, 33)), (18, (6, 9, 20))):
        la = (torch.randn((1, C) + shape, generator=g).abs() + 0.3).cuda().requires_grad_(True)
        lb = (torch.randn((1, C) + shape, generator=g).abs() + 0.3).cuda()
        torch.manual_seed(C)
        _, Ra = get_rand_affine(1, strength=0.1)
        _, Rb = get_rand_affine(1, strength=0.1)
        f = consistency_dice_loss_warped(la, lb, Ra, Rb)
        (ga_f,) = torch.autograd.grad(f, la)
        u = consistency_dice_loss(affine_grid_sample(la, Ra), affine_grid_sample(lb, Rb))
        (ga_u,) = torch.autograd.grad(u, la)
        assert abs(f.item() - u.item()) <= 2e-6
        assert float((ga_f - ga_u).abs().max()) <= 2e-5 * float(ga_u.abs().max())
