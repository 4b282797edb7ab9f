"""Boundary contract of the drop-in: argument errors are raised loudly, non-contiguous and list
inputs behave like the reference's, nothing silently falls back to another device or dtype."""
import pytest
import torch

from gpu_util import synth_volume

pytestmark = pytest.mark.gpu


def test_mind_argument_errors_and_noncontiguous_input():
    from dg_tta_b200 import MIND3D, mind_ssc
    m = MIND3D()
    with pytest.raises(ValueError):
        m(torch.zeros(1, 8, 8, 8, device="cuda"))                      # 4-D
    with pytest.raises(ValueError):
        m(torch.zeros(1, 2, 8, 8, 8, device="cuda"))                   # the shift kernels are [12,1,3,3,3]: C must be 1
    with pytest.raises(TypeError):
        m(torch.zeros(1, 1, 8, 8, 8, device="cuda", dtype=torch.float64))   # the reference is fp32-only as well
    with pytest.raises(TypeError):
        m(torch.zeros(1, 1, 8, 8, 8))                                  # CPU tensor: no fallback
    with pytest.raises(ValueError):
        mind_ssc(torch.zeros(1, 1, 8, 8, 8, device="cuda"), noise=torch.zeros(1, 12, 8, 8, 4, device="cuda"))
    x = synth_volume((1, 1, 20, 24, 28), 5).cuda()
    xt = x.permute(0, 1, 4, 3, 2)                                       # a non-contiguous view of another volume
    assert not xt.is_contiguous()
    assert torch.equal(mind_ssc(xt, noise=False), mind_ssc(xt.contiguous(), noise=False))
    assert MIND3D().out_channels == 12 and isinstance(MIND3D(), torch.nn.Module)
    import copy
    assert isinstance(copy.deepcopy(MIND3D(delta=2)), MIND3D)           # get_model_from_network deep-copies hooks' owners


def test_gin_list_input_and_channel_check():
    from dg_tta_b200.gin import GINGroupConv, gin_aug
    x = synth_volume((2, 1, 12, 16, 20), 9).cuda()
    net = GINGroupConv(dict(IN_CHANNELS=1, N_LAYER=4, INTERM_CHANNELS=2))
    torch.manual_seed(2)
    a = net([x[:1], x[1:]])                                            # gin.py:169-170: a list is concatenated on dim 0
    torch.manual_seed(2)
    b = net(x)
    assert torch.equal(a, b)
    with pytest.raises(ValueError):
        net(torch.zeros(1, 2, 8, 8, 8, device="cuda"))
    with pytest.raises(NotImplementedError):
        gin_aug(torch.zeros(1, 1, 8, 8, device="cuda"))                # 2-D GIN is outside the hot path
    out = gin_aug(x)
    # the augmentation keeps each sample's Frobenius norm (gin.py:200-228)
    assert torch.allclose(out.flatten(1).norm(dim=1), x.flatten(1).norm(dim=1), rtol=1e-4)


def test_sampler_argument_errors():
    from dg_tta_b200.tta.augmentation_utils import affine_grid_sample, affine_label_argmax
    x = torch.zeros(2, 3, 8, 8, 8, device="cuda")
    eye = torch.eye(3, 4)[None].repeat(2, 1, 1)
    with pytest.raises(ValueError):
        affine_grid_sample(x, eye[:1])                                  # theta batch mismatch
    with pytest.raises(NotImplementedError):
        affine_grid_sample(x, eye, align_corners=True)
    with pytest.raises(ValueError):
        affine_grid_sample(x, eye, mode="bicubic")
    with pytest.raises(ValueError):
        affine_label_argmax(x[0], eye)
    assert torch.equal(affine_grid_sample(x + 1, eye, mode="nearest"), x + 1)   # identity crop, nearest: exact
