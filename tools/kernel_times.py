#!/usr/bin/env python3
"""Developer probe: CUDA-event timings of the individual kernels (not the bench contract; see bench.py)."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests"))
from dg_tta_b200 import mind_ssc  # noqa: E402
from dg_tta_b200.gin import GINGroupConv, gin_forward  # noqa: E402
from dg_tta_b200.tta.augmentation_utils import affine_grid_sample, get_rand_affine  # noqa: E402
from gpu_util import synth_volume  # noqa: E402


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    res = {}
    only = sys.argv[1] if len(sys.argv) > 1 else ""   # "mind": MIND timings only
    for shape in ([] if only in ("gin", "sampler") else [(1, 1, 128, 128, 128), (2, 1, 192, 192, 192)]):
        x = synth_volume(shape, 1).cuda()
        vox = x.numel()
        for delta in (1, 2):
            med, best = timeit(lambda: mind_ssc(x, delta=delta, noise=False))
            res[f"mind_clean_d{delta}_{shape[0]}x{shape[2]}"] = dict(ms=med, best=best, gvox_s=vox / med / 1e6, gbs=vox * 52 / med / 1e6)
        noise = torch.randn((shape[0], 12) + shape[2:], device="cuda")
        med, best = timeit(lambda: mind_ssc(x, noise=noise))
        res[f"mind_noise_tensor_{shape[0]}x{shape[2]}"] = dict(ms=med, best=best, gvox_s=vox / med / 1e6, gbs=vox * 100 / med / 1e6)
        med, best = timeit(lambda: mind_ssc(x))
        res[f"mind_default_randn_{shape[0]}x{shape[2]}"] = dict(ms=med, best=best, gvox_s=vox / med / 1e6)
    from dg_tta_b200.mind import randn_like_reference
    for shape in ([] if only in ("gin", "sampler") else [(2, 12, 192, 192, 192)]):
        med, best = timeit(lambda: torch.randn(shape, device="cuda"))
        res["torch_randn_2x12x192"] = dict(ms=med, best=best, gbs=2 * 12 * 192 ** 3 * 4 / med / 1e6)
        med, best = timeit(lambda: randn_like_reference(shape, "cuda"))
        res["ours_philox_normal_2x12x192"] = dict(ms=med, best=best, gbs=2 * 12 * 192 ** 3 * 4 / med / 1e6)
    if only == "mind":
        for k, v in res.items():
            print(k, json.dumps({a: round(b, 4) for a, b in v.items()}))
        return
    x = synth_volume((2, 1, 192, 192, 192), 2).cuda()
    net = GINGroupConv(dict(IN_CHANNELS=1, N_LAYER=4, INTERM_CHANNELS=2))
    for want in ([] if only == "sampler" else [[1, 1, 1, 1], [3, 3, 3, 3], [3, 1, 3, 1]]):
        seed = 0
        while True:
            torch.manual_seed(seed)
            alphas, kers, shifts = net.draw(x)
            if [k.shape[-1] for k in kers] == want:
                break
            seed += 1
        med, best = timeit(lambda: gin_forward(x, kers, shifts, alphas, 2))
        res["gin_k" + "".join(map(str, want))] = dict(ms=med, best=best, gvox_s=x.numel() / med / 1e6, gbs=x.numel() * 8 / med / 1e6)
    # gin_aug over seeds 0..15 (all 16 draws of the four kernel sizes occur with equal probability)
    from dg_tta_b200.gin import gin_aug
    per_seed = []
    for seed in range(0 if only == "sampler" else 16):
        def run(seed=seed):
            torch.manual_seed(seed)
            return gin_aug(x)
        med, _ = timeit(run, iters=5, warm=2)
        per_seed.append(med)
    if per_seed:
      res["gin_aug_seeds0_15_2x192"] = dict(ms=sum(per_seed) / 16, best=min(per_seed), worst=max(per_seed),
                                          gvox_s=x.numel() / (sum(per_seed) / 16) / 1e6)
    # GIN -> MIND on the MultiRes volume / patch shapes
    from dg_tta_b200.tta.augmentation_utils import gin_mind_aug
    for dhw in ([] if only == "sampler" else [(231, 228, 242), (116, 114, 121), (58, 57, 60), (38, 38, 40), (56, 56, 64), (28, 28, 32), (19, 19, 21)]):
        xm = synth_volume((2, 1) + dhw, 4).cuda()
        def run(xm=xm):
            torch.manual_seed(1)
            return gin_mind_aug(xm)
        med, best = timeit(run, iters=5, warm=2)
        res["gin_mind_aug_2x%dx%dx%d" % dhw] = dict(ms=med, best=best, gvox_s=xm.numel() / med / 1e6)
    torch.manual_seed(0)
    R, Ri = get_rand_affine(2)
    img = synth_volume((2, 1, 128, 128, 128), 3).cuda()
    med, best = timeit(lambda: affine_grid_sample(img, R, padding_mode="border"))
    res["sample_img_border_2x128"] = dict(ms=med, best=best, gbs=img.numel() * 8 / med / 1e6)
    lg = torch.randn(2, 14, 128, 128, 128, device="cuda", requires_grad=True)
    med, best = timeit(lambda: affine_grid_sample(lg, Ri))
    res["sample_logits_fwd_2x14x128"] = dict(ms=med, best=best, gbs=lg.numel() * 8 / med / 1e6)
    out = affine_grid_sample(lg, Ri)
    go = torch.randn_like(out)
    med, best = timeit(lambda: torch.autograd.grad(out, lg, go, retain_graph=True))
    res["sample_logits_bwd_2x14x128"] = dict(ms=med, best=best, gbs=lg.numel() * 12 / med / 1e6)
    # consistency loss: fused sums (+ gradient) vs the reference's elementwise chain, on the warped logits' shape
    from dg_tta_b200.tta.torch_utils import consistency_dice_loss
    ta = torch.randn(2, 14, 128, 128, 128, device="cuda", requires_grad=True)
    tb = torch.randn(2, 14, 128, 128, 128, device="cuda")

    def chain():
        m = (ta.sum(1, keepdim=True) > 0.0).float() * (tb.sum(1, keepdim=True) > 0.0).float()
        sa, sb = ta.softmax(1) * m, tb.softmax(1) * m
        n = (2.0 * sa * sb).reshape(2, -1, 128 ** 3).mean(2)
        d = 0.5 * ((sa + sb) ** 2).reshape(2, -1, 128 ** 3).mean(2)
        return 1 - (n / d)[:, 1:].mean()

    for name, fn in (("closs_fused_fwd_bwd_2x14x128", lambda: torch.autograd.grad(consistency_dice_loss(ta, tb), ta)),
                     ("closs_torch_fwd_bwd_2x14x128", lambda: torch.autograd.grad(chain(), ta))):
        med, best = timeit(fn)
        res[name] = dict(ms=med, best=best, gbs=ta.numel() * 4 * 5 / med / 1e6)   # read a,b twice + write grad: 5 passes
    # the whole TTA epilogue (tta.py:571-575 + 263-269, forward + backward w.r.t. branch a): unfused product path
    # (2 warps + fused sums + their backwards) vs the fused-warp kernels
    from dg_tta_b200.tta.torch_utils import consistency_dice_loss_warped
    la = torch.randn(2, 14, 128, 128, 128, device="cuda", requires_grad=True)
    lb = torch.randn(2, 14, 128, 128, 128, device="cuda")
    _, Rb_inv = get_rand_affine(2)
    for name, fn in (("tta_epilogue_unfused_fwd_bwd_2x14x128", lambda: torch.autograd.grad(consistency_dice_loss(affine_grid_sample(la, Ri), affine_grid_sample(lb, Rb_inv)), la)),
                     ("tta_epilogue_fused_fwd_bwd_2x14x128", lambda: torch.autograd.grad(consistency_dice_loss_warped(la, lb, Ri, Rb_inv), la))):
        med, best = timeit(fn)
        res[name] = dict(ms=med, best=best)
    from dg_tta_b200.tta.torch_utils import get_batch
    vol = synth_volume((1, 1, 231, 228, 242), 9)[0].cuda()
    lab = torch.zeros(104, 231, 228, 242, device="cuda")
    lab[3, 50:150, 40:160, 60:180] = 1
    sample = torch.cat([vol, lab], 0)
    get_batch([sample], [0, 0], [128, 128, 128], device="cuda")
    med, best = timeit(lambda: get_batch([sample], [0, 0], [128, 128, 128], device="cuda"))
    res["get_batch_2_crops_104_labels_231x228x242_to_128"] = dict(ms=med, best=best)
    del sample, lab
    from dg_tta_b200.pretraining import resize_edge
    xr = synth_volume((1, 2, 128, 128, 128), 10)[0].cuda()
    med, best = timeit(lambda: resize_edge(resize_edge(xr, (64, 32, 21), 0), (128, 128, 128), 3))
    res["lowres_sim_2ch_128_down_64x32x21_up_cubic"] = dict(ms=med, best=best)
    # torch eager comparison for the sampler
    import torch.nn.functional as F
    Rd = R.cuda()
    med, best = timeit(lambda: F.grid_sample(img, F.affine_grid(Rd, list(img.shape), align_corners=False), padding_mode="border", align_corners=False))
    res["torch_sample_img_border_2x128"] = dict(ms=med, best=best)
    for k, v in res.items():
        print(k, json.dumps({a: round(b, 4) for a, b in v.items()}))


if __name__ == "__main__":
    main()
