"""Host-side logic of the drop-in (no GPU): RNG draw order, env flag, shapes/contracts, isolation."""
import ast
import os
from pathlib import Path

import numpy as np
import pytest
import torch

from conftest import gin_layers, load_golden

ROOT = Path(__file__).resolve().parents[1]


def test_get_rand_affine_matches_reference_draws():
    from dg_tta_b200.tta.augmentation_utils import get_rand_affine
    g = load_golden("rand_affine")
    torch.manual_seed(int(g["seed"]))
    R, Ri = get_rand_affine(3, strength=0.05, flip=False)
    assert np.array_equal(R.numpy(), g["R"]) and np.array_equal(Ri.numpy(), g["R_inv"])
    torch.manual_seed(int(g["seed_flip"]))
    R, Ri = get_rand_affine(2, strength=0.1, flip=True)
    assert np.array_equal(R.numpy(), g["R_flip"]) and np.array_equal(Ri.numpy(), g["R_flip_inv"])


@pytest.mark.parametrize("name", ["gin_k3131", "gin_k1333", "gin_b3", "gin_odd"])
def test_gin_draw_order_reproduces_reference_weights(name):
    """torch.manual_seed(s) followed by GINGroupConv.draw yields the reference's alphas (CPU input ->
    CPU generator), kernel sizes, kernels and shifts (gin.py:187, 65-66, 94-103)."""
    from dg_tta_b200.gin import GINGroupConv
    g = load_golden(name)
    x = torch.from_numpy(g["x"])
    torch.manual_seed(int(g["seed"]))
    alphas, kers, shifts = GINGroupConv(dict(IN_CHANNELS=1, N_LAYER=4, INTERM_CHANNELS=2)).draw(x)
    gk, gs = gin_layers(g)
    assert np.array_equal(alphas.numpy(), g["alphas"])
    assert [k.shape[-1] for k in kers] == g["ksizes"].tolist()
    for a, b in zip(kers, gk):
        assert np.array_equal(a.numpy(), b)
    for a, b in zip(shifts, gs):
        assert np.array_equal(a.numpy(), b)


def test_gaussian_taps_host():
    from dg_tta_b200.mind import gaussian_taps
    from oracle import cform
    for s in (1, 0.5, 2):
        assert np.allclose(list(gaussian_taps(s)), cform.gaussian_taps(s), atol=1e-7)
    with pytest.raises(ValueError):
        gaussian_taps(3.5)  # 13 taps


def test_internal_augmentation_flag_semantics(monkeypatch):
    from dg_tta_b200 import gin, utils
    monkeypatch.delenv("DG_TTA_INTERNAL_AUGMENTATION", raising=False)
    with pytest.raises(AttributeError):  # reference: None.lower() (utils.py:17-18)
        utils.get_internal_augmentation_enabled()
    utils.disable_internal_augmentation()
    assert os.environ["DG_TTA_INTERNAL_AUGMENTATION"] == "false"
    x = torch.zeros(1, 1, 2, 2, 2)
    out = gin.gin_hook(None, (x,))
    assert isinstance(out, tuple) and out[0] is x      # gin.py:247 returns the input tuple untouched
    utils.check_internal_augmentation_disabled()
    utils.enable_internal_augmentation()
    assert utils.get_internal_augmentation_enabled()
    with pytest.raises(TypeError):                      # enabled + CPU tensor: no CPU fallback
        gin.gin_hook(None, (x,))


def test_module_surface_matches_reference():
    import dg_tta_b200 as pkg
    m = pkg.MIND3D()
    assert (m.delta, m.sigma, m.out_channels, m.randn_weighting) == (1, 1, 12, 0.05)
    assert len(list(m.parameters())) == 0
    import copy
    assert copy.deepcopy(m).delta == 1
    g = pkg.GINGroupConv(dict(IN_CHANNELS=1, N_LAYER=4, INTERM_CHANNELS=2))
    assert len(g.layers) == 4 and g.layers[-1].use_act is False and g.layers[0].use_act is True
    assert [(b.in_channel, b.out_channel) for b in g.layers] == [(1, 2), (2, 2), (2, 2), (2, 1)]
    with pytest.raises(AssertionError):
        pkg.GradlessGCReplayNonlinBlock(requires_grad=True)


def test_cpu_tensors_are_rejected_loudly():
    import dg_tta_b200 as pkg
    from dg_tta_b200.tta.augmentation_utils import affine_grid_sample
    with pytest.raises(TypeError):
        pkg.MIND3D()(torch.zeros(1, 1, 4, 4, 4))
    with pytest.raises(TypeError):
        pkg.gin_aug(torch.zeros(1, 1, 4, 4, 4))
    with pytest.raises(TypeError):
        affine_grid_sample(torch.zeros(1, 1, 4, 4, 4), torch.eye(3, 4)[None])


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under dg_tta_b200/ may import or open it."""
    for py in (ROOT / "dg_tta_b200").rglob("*.py"):
        tree = ast.parse(py.read_text())
        for node in ast.walk(tree):
            names = []
            if isinstance(node, ast.Import):
                names = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                names = [node.module or ""]
            assert not any(n.split(".")[0] in ("oracle", "tests") for n in names), py
        assert "oracle/" not in py.read_text().replace("oracle/ is test", "").replace("under oracle/", "")
    for cu in (ROOT / "dg_tta_b200" / "csrc").glob("*"):
        assert "oracle" not in cu.read_text()
