#!/usr/bin/env python3
"""Not a pytest file: kernel-vs-eager timings on one GPU ("the bar to beat is the reference's own
torch-eager op sequence on the same B200").  Runs the torch port of the reference ops (oracle/ref_port.py) on CUDA
tensors with TF32 off (parity mode) and on (the reference's default), next to the drop-in kernels.
    python tests/perf_eager_gpu.py [--tta]
"""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from gpu_util import synth_volume  # noqa: E402
from oracle import ref_port  # noqa: E402


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


def main():
    from dg_tta_b200 import mind_ssc
    from dg_tta_b200.gin import GINGroupConv, gin_forward
    from dg_tta_b200.tta.augmentation_utils import affine_grid_sample, get_rand_affine
    res = {}
    x = synth_volume((2, 1, 192, 192, 192), 1).cuda()
    net = GINGroupConv(dict(IN_CHANNELS=1, N_LAYER=4, INTERM_CHANNELS=2))
    torch.manual_seed(8)   # a mixed kernel-size pattern
    alphas, kers, shifts = net.draw(x)
    kd, sd = [k.cuda() for k in kers], [s.cuda() for s in shifts]
    res["ksizes"] = [k.shape[-1] for k in kers]
    for tf32 in (False, True):
        torch.backends.cudnn.allow_tf32 = tf32
        tag = "tf32" if tf32 else "fp32"
        with torch.no_grad():
            res[f"eager_mind_2x192_{tag}_ms"] = timeit(lambda: ref_port.mind_ssc(x, noise=torch.randn(2, 12, 192, 192, 192, device="cuda")))
            res[f"eager_gin_2x192_{tag}_ms"] = timeit(lambda: ref_port.gin(x, kd, sd, alphas))
    torch.backends.cudnn.allow_tf32 = False
    res["ours_mind_2x192_ms"] = timeit(lambda: mind_ssc(x))
    res["ours_gin_2x192_ms"] = timeit(lambda: gin_forward(x, kers, shifts, alphas, 2))
    torch.manual_seed(0)
    R, Ri = get_rand_affine(2)
    img = synth_volume((2, 1, 128, 128, 128), 3).cuda()
    ident = torch.nn.functional.affine_grid(torch.eye(4, device="cuda").repeat(2, 1, 1)[:, :3], [2, 1, 128, 128, 128], align_corners=False)
    Rd = R.cuda()
    res["eager_warp_img_2x128_ms"] = timeit(lambda: ref_port.tta_view_warp(img, Rd, ident, "border"))
    res["ours_warp_img_2x128_ms"] = timeit(lambda: affine_grid_sample(img, R, padding_mode="border"))
    lg = torch.randn(2, 14, 128, 128, 128, device="cuda")
    Rid = Ri.cuda()
    res["eager_warp_logits_2x14x128_ms"] = timeit(lambda: ref_port.tta_view_warp(lg, Rid, ident, "zeros"))
    res["ours_warp_logits_2x14x128_ms"] = timeit(lambda: affine_grid_sample(lg, Ri))
    if "--tta" in sys.argv:
        import numpy as np
        sys.path.insert(0, str(ROOT / "tools"))
        import tta_standin as ts
        from eager_transforms import EagerTorchTransforms
        vol = [synth_volume((1, 1, 160, 160, 176), 5, "mr")[0].cuda()]
        for name, tr in (("ours", ts.DropInTransforms()), ("eager", EagerTorchTransforms())):
            model = ts.build_model(tr, num_classes=105).cuda()
            opt = torch.optim.AdamW(model.parameters(), lr=1e-5)
            idx = list(range(1, 15))
            rng = np.random.RandomState(0)

            def step():
                ts.tta_inner_step(model, vol, [112, 112, 128], 1, idx, tr, rng=rng)
            res[f"tta_step_{name}_ms"] = timeit(step, iters=4, warm=2)
            opt.zero_grad(set_to_none=True)
            del model, opt
            torch.cuda.empty_cache()
    print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in res.items()}))


if __name__ == "__main__":
    main()
