"""dgtta_philox_normal_fill regenerates torch's CUDA randn stream: bitwise equality with torch.randn for the same
generator state, same generator advance, across sizes that exercise torch's grid clamp and the ragged tail."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(1, 12, 5, 7, 9), (1, 12, 24, 20, 36), (2, 12, 33, 17, 65), (1, 12, 64, 64, 64),
                                   (1, 12, 128, 128, 128), (3, 1, 1, 1, 1), (1, 12, 100, 101, 103)])
@pytest.mark.parametrize("pre_draws", [0, 3])
def test_fill_is_bitwise_torch_randn(shape, pre_draws):
    from dg_tta_b200.mind import randn_like_reference
    torch.cuda.init()
    gen = torch.cuda.default_generators[torch.cuda.current_device()]
    torch.manual_seed(1234)
    for _ in range(pre_draws):
        torch.rand(1000, device="cuda")          # move the offset away from 0
    start = gen.get_offset()
    want = torch.randn(shape, device="cuda")
    after = gen.get_offset()
    torch.manual_seed(1234)
    for _ in range(pre_draws):
        torch.rand(1000, device="cuda")
    assert gen.get_offset() == start
    got = randn_like_reference(shape, "cuda")
    assert gen.get_offset() == after
    assert torch.equal(got.view(torch.int32), want.view(torch.int32))


def test_default_mind_call_uses_the_regenerated_stream():
    """MIND3D()(x) with noise=None == MIND3D()(x, noise=torch.randn(...)) for the same seed, bit for bit."""
    from dg_tta_b200 import MIND3D
    x = torch.randn(2, 1, 20, 24, 40, device="cuda")
    torch.manual_seed(7)
    n = torch.randn(2, 12, 20, 24, 40, device="cuda")
    a = MIND3D()(x, noise=n)
    torch.manual_seed(7)
    b = MIND3D()(x)
    assert torch.equal(a, b)


def test_full_size_stream(tmp_path):
    """2x12x192^3 (BASELINE configs[1]): 170 M normals, compared on the device"""
    from dg_tta_b200.mind import randn_like_reference
    shape = (2, 12, 192, 192, 192)
    torch.manual_seed(99)
    want = torch.randn(shape, device="cuda")
    torch.manual_seed(99)
    got = randn_like_reference(shape, "cuda")
    assert torch.equal(got, want)
