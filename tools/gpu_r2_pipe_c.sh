#!/bin/bash
mkdir -p gpurun_out
export DGTTA_LIB_PATH=$PWD/gpurun_variants/lib_dbg.so
for shp in 2x1x192x192x192 2x1x191x192x192 2x1x193x192x192 2x1x192x192x208; do
  echo "=== $shp"; python tools/dbg_pipe_times.py $shp 2>&1 | tail -9
done
unset DGTTA_LIB_PATH
python - <<'PY'
import sys, torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from dg_tta_b200 import mind_ssc
from gpu_util import synth_volume
import os
def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]
for mode in ("pipe", "batch"):
    if mode == "batch": os.environ["DGTTA_MIND_NO_PIPE"] = "1"
    for shp in [(2,1,192,192,192), (2,1,191,192,192), (2,1,193,192,192), (2,1,192,192,208), (2,1,200,192,192)]:
        x = synth_volume(shp, 1).cuda(); n = torch.randn((shp[0], 12) + shp[2:], device="cuda")
        t = timeit(lambda: mind_ssc(x, noise=n)); tc = timeit(lambda: mind_ssc(x, noise=False))
        print(mode, shp, "noise %.4f ms (%.2f ns/kvox)  clean %.4f ms" % (t, t * 1e6 / (x.numel() / 1e3), tc))
PY
