"""BASELINE.json configs[3]: GIN_MIND_MultiRes transforms on volumes resampled to 1.5/3.0/6.0/9.0 mm and on the patch
sizes of those resolutions (median shape [231,228,242] x {1, 1/2, 1/4, 1/6}; patches [56,56,64],
[28,28,32], [19,19,21]).  W % 4 != 0 shapes run the LDG noise path, the others the TMA-staged one; both against the
plain-C oracle at the north_star tolerances (GIN 1e-5 * max|ref|; MIND 1e-5 on identical input; whole chain 2e-4)."""
import numpy as np
import pytest
import torch

from gpu_util import synth_volume

pytestmark = pytest.mark.gpu

SHAPES = [(231, 228, 242), (116, 114, 121), (58, 57, 60), (38, 38, 40), (56, 56, 64), (28, 28, 32), (19, 19, 21)]


@pytest.mark.parametrize("dhw", SHAPES)
def test_gin_then_mind_against_oracle(dhw):
    from dg_tta_b200 import mind_ssc
    from dg_tta_b200.gin import GINGroupConv, gin_forward
    from oracle import cform
    B = 1 if dhw[0] > 200 else 2
    x = synth_volume((B, 1) + dhw, 40 + dhw[0], "mr")
    net = GINGroupConv(dict(IN_CHANNELS=1, N_LAYER=4, INTERM_CHANNELS=2))
    torch.manual_seed(dhw[2])
    alphas, kers, shifts = net.draw(x.cuda())
    noise = torch.randn((B, 12) + dhw, generator=torch.Generator().manual_seed(1))
    xd = x.cuda()
    gin_out = gin_forward(xd, kers, shifts, alphas, 2)
    gin_ref = cform.gin(x.numpy(), [k.numpy() for k in kers], [s.numpy() for s in shifts], alphas.cpu().numpy())
    assert np.abs(gin_out.cpu().numpy() - gin_ref).max() <= 1e-5 * np.abs(gin_ref).max()
    # MIND on the CUDA path's own GIN output (deferred rescale applied on MIND's loads) against the oracle's MIND of
    # that same tensor: the per-operator bar of north_star (1e-5) ...
    mixed, scale = gin_forward(xd, kers, shifts, alphas, 2, defer_scale=True)
    got = mind_ssc(mixed, noise=noise.cuda(), in_scale=scale).cpu().numpy()
    ref_same_input = cform.mind_ssc(gin_out.cpu().numpy(), noise=noise.numpy(), randn_weighting=0.05)
    assert np.abs(got - ref_same_input).max() <= 1e-5
    # ... and the whole chain against the oracle's chain: MIND divides differences of smoothed squares by their mean, so
    # the (within-tolerance) accumulation-order differences of GIN are amplified by ssd / var on flat MR background
    ref = cform.mind_ssc(gin_ref, noise=noise.numpy(), randn_weighting=0.05)
    assert np.abs(got - ref).max() <= 2e-4
    assert (got.max(1) == 1.0).all() and (got > 0).all()


@pytest.mark.parametrize("dhw", [(231, 228, 242), (58, 57, 60)])
def test_hook_chain_on_multires_volume(dhw):
    """hooks in the trainers' registration order (nnUNetTrainer_GIN_MIND.py:55-57): GIN first, then MIND -> 12 channels"""
    from dg_tta_b200 import gin_hook, mind_hook
    from dg_tta_b200.utils import disable_internal_augmentation, enable_internal_augmentation
    seen = {}

    class Probe(torch.nn.Module):
        def forward(self, x):
            seen["shape"] = tuple(x.shape)
            return x

    net = Probe()
    net.register_forward_pre_hook(gin_hook)
    net.register_forward_pre_hook(mind_hook)
    x = synth_volume((1, 1) + dhw, 5, "mr").cuda()
    try:
        enable_internal_augmentation()
        torch.manual_seed(0)
        a = net(x)
        assert seen["shape"] == (1, 12) + dhw
        torch.manual_seed(0)
        assert torch.equal(a, net(x))            # seeds reproduce the whole chain (GIN draws + MIND noise)
        disable_internal_augmentation()
        torch.manual_seed(0)
        b = net(x)                               # MIND only
        assert not torch.equal(a, b) and tuple(b.shape) == (1, 12) + dhw
    finally:
        disable_internal_augmentation()
