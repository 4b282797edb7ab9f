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
x = synth_volume(shape, 1).cuda()
n = torch.randn((shape[0], 12) + shape[2:], device="cuda")
for _ in range(3):
    mind_ssc(x, noise=n)
torch.cuda.synchronize()
lib = ctypes.CDLL(str(_lib.LIB_PATH))
buf = np.zeros((1024, 4), dtype=np.uint64)
lib.dgtta_debug_cta_times.argtypes = [ctypes.c_void_p, ctypes.c_int]
rc = lib.dgtta_debug_cta_times(buf.ctypes.data, 1024)
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
