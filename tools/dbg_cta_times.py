#!/usr/bin/env python3
"""Developer probe (needs a -DDGTTA_CTA_TIMES build (tools/build_variant.sh times -DDGTTA_CTA_TIMES) selected with DGTTA_LIB_PATH): per-CTA start / end times of the
MIND pass-1 kernel (mind_fast_kernel)."""
import ctypes
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests"))
from dg_tta_b200 import mind_ssc  # noqa: E402
from dg_tta_b200 import _lib  # noqa: E402
from gpu_util import synth_volume  # noqa: E402

shape = tuple(int(v) for v in sys.argv[1].split("x")) if len(sys.argv) > 1 else (2, 1, 192, 192, 192)
what = sys.argv[2] if len(sys.argv) > 2 else "mind"      # "mind" | "gin" (the LAST gin_stack_kernel launch of an all-3x3x3 stack)
x = synth_volume(shape, 1).cuda()
lib = ctypes.CDLL(str(_lib.LIB_PATH))
buf = np.zeros((2048, 4), dtype=np.uint64)
if what == "mind":
    n = torch.randn((shape[0], 12) + shape[2:], device="cuda")
    for _ in range(3):
        mind_ssc(x, noise=n)
    torch.cuda.synchronize()
    lib.dgtta_debug_cta_times.argtypes = [ctypes.c_void_p, ctypes.c_int]
    rc = lib.dgtta_debug_cta_times(buf.ctypes.data, 1024)
else:
    from dg_tta_b200.gin import GINGroupConv, gin_forward
    net = GINGroupConv(dict(IN_CHANNELS=1, N_LAYER=4, INTERM_CHANNELS=2))
    seed = 0
    while True:
        torch.manual_seed(seed)
        alphas, kers, shifts = net.draw(x)
        if [k.shape[-1] for k in kers] == [3, 3, 3, 3]:
            break
        seed += 1
    for _ in range(3):
        gin_forward(x, kers, shifts, alphas, 2)
    torch.cuda.synchronize()
    lib.dgtta_debug_gin_cta_times.argtypes = [ctypes.c_void_p, ctypes.c_int]
    rc = lib.dgtta_debug_gin_cta_times(buf.ctypes.data, 2048)
ncta = int((buf[:, 1] > 0).sum())
b = buf[:ncta].astype(np.int64)
t0 = b[:, 1].min()
start, end = b[:, 1] - t0, b[:, 2] - t0
print(f"rc {rc} ctas {ncta} distinct SMs {len(set(b[:, 0]))}")
print("start  min/med/max us", start.min() / 1e3, np.median(start) / 1e3, start.max() / 1e3)
print("end    min/med/max us", end.min() / 1e3, np.median(end) / 1e3, end.max() / 1e3)
dur = (end - start) / 1e3
print("dur    min/med/max us", dur.min(), np.median(dur), dur.max())
print("cycles min/med/max", b[:, 3].min(), int(np.median(b[:, 3])), b[:, 3].max())
order = np.argsort(dur)
print("fastest CTAs (cta, smid, dur):", [(int(i), int(b[i, 0]), round(float(dur[i]), 1)) for i in order[:8]])
print("slowest CTAs (cta, smid, dur):", [(int(i), int(b[i, 0]), round(float(dur[i]), 1)) for i in order[-8:]])
hist, edges = np.histogram(dur, bins=10)
print("hist", list(zip(np.round(edges[:-1], 0).tolist(), hist.tolist())))
if what == "gin":
    # 1-D grid, chunk index slowest: lin = cd * (npatch * nb) + sample * npatch + patch; an all-3x3x3 stack on 192^3 has 4 chunks
    per = {}
    nchunk = 4
    per_chunk = ncta // nchunk
    npatch = per_chunk // shape[0]
    for i in range(ncta):
        cd, smp = i // per_chunk, (i % per_chunk) // npatch
        per.setdefault((smp, cd), []).append(float(dur[i]))
    for k in sorted(per):
        v = np.array(per[k])
        print("sample %d chunk %d: n %d  min %.1f med %.1f max %.1f" % (k[0], k[1], len(v), v.min(), np.median(v), v.max()))
    sm_of = {}
    for i in range(ncta):
        sm_of.setdefault(int(b[i, 0]), []).append(i)
    pairs = [(round(float(dur[v[0]]), 1), round(float(dur[v[1]]), 1), v) for v in sm_of.values() if len(v) == 2][:12]
    print("co-resident pairs (dur, dur, ctas):", pairs)
