"""Host-buffer front end (dg_tta_b200/host_pipeline.py): same values as direct device calls, same seeds."""
import pytest
import torch

from gpu_util import synth_volume

pytestmark = pytest.mark.gpu


def test_pipeline_equals_direct_calls():
    from dg_tta_b200.host_pipeline import HostPipeline
    from dg_tta_b200.tta.augmentation_utils import gin_mind_aug
    shape = (2, 1, 24, 32, 40)
    xs = [synth_volume(shape, 300 + i) for i in range(3)]
    direct = []
    for i, x in enumerate(xs):
        torch.manual_seed(50 + i)
        direct.append(gin_mind_aug(x.cuda()).cpu())
    pipe = HostPipeline()
    h_in = [x.pin_memory() for x in xs]
    h_out = [torch.empty((2, 12, 24, 32, 40)).pin_memory() for _ in xs]
    events = []
    for i in range(3):
        torch.manual_seed(50 + i)
        events.append(pipe.submit(h_in[i], h_out[i]))
    for e in events:
        e.synchronize()
    for a, b in zip(direct, h_out):
        assert torch.equal(a, b)
    assert pipe.h2d_bytes == 3 * xs[0].numel() * 4 and pipe.d2h_bytes == 3 * h_out[0].numel() * 4


def test_pipeline_rejects_pageable_and_device_tensors():
    from dg_tta_b200.host_pipeline import HostPipeline
    pipe = HostPipeline()
    x = torch.zeros(1, 1, 8, 8, 8)
    y = torch.zeros(1, 12, 8, 8, 8)
    with pytest.raises(ValueError):
        pipe.submit(x, y.pin_memory())
    with pytest.raises(TypeError):
        pipe.submit(x.cuda(), y.pin_memory())
