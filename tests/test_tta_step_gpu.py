"""BASELINE.json configs[2]: GIN_MIND transforms feeding a PlainConvUNet-shaped TTA step with the affine grid_sample
consistency loss.  The stand-in step (tools/tta_standin.py, a restatement of dg_tta/tta/tta.py:221-281, 480-579) is
run twice from the same seeds: once on the drop-in ops, once on the reference's torch-eager op sequence
(tests/eager_transforms.py).  Same generator consumption -> same crops, same affines, same MIND noise -> the
consistency loss and the gradient it sends into the network must agree."""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

from gpu_util import synth_volume

pytestmark = pytest.mark.gpu
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tools"))


def _run(transforms, seed):
    import tta_standin as ts
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    model = ts.build_model(transforms, num_classes=9, features=(8, 16, 32), seed=3).cuda()
    vol = [synth_volume((1, 1, 52, 60, 56), 77, "mr")[0].cuda()]
    torch.manual_seed(seed)
    rng = np.random.RandomState(seed)
    loss = ts.tta_inner_step(model, vol, [32, 40, 32], 2, list(range(1, 6)), transforms, rng=rng)
    grad = torch.cat([p.grad.reshape(-1) for p in model.parameters() if p.grad is not None])
    return float(loss), grad


def test_inner_step_matches_eager_reference_ops():
    import tta_standin as ts
    from eager_transforms import EagerTorchTransforms
    loss_a, grad_a = _run(ts.DropInTransforms(), 11)
    loss_b, grad_b = _run(EagerTorchTransforms(), 11)
    assert abs(loss_a - loss_b) <= 2e-4 * max(1.0, abs(loss_b))
    denom = float(grad_b.abs().max())
    assert denom > 0
    assert float((grad_a - grad_b).abs().max()) <= 2e-2 * denom   # fp32 cuDNN backward + atomic scatter order
