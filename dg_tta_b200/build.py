"""Build recipe for libdgtta_sm100.so (hand-written sm_100a CUDA behind the C ABI of include/dgtta.h).

The library is built IN-TREE (dg_tta_b200/lib/) with nvcc directly — no JIT cache, no torch
extension machinery — so the .so travels with the repository snapshot to the GPU box.
    python -m dg_tta_b200.build [--force] [--verbose]
"""
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB_DIR = PKG / "lib"
LIB_PATH = LIB_DIR / "libdgtta_sm100.so"
OBJ_DIR = PKG / "build"
SOURCES = ["api.cu", "mind_ssc.cu", "mind_fast.cu", "mind_general.cu", "gin.cu", "gin_stack.cu", "affine_sample.cu", "philox_normal.cu", "consistency_loss.cu", "resize.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-O3",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: libdgtta_sm100.so cannot be built (there is no CPU fallback)")


def _deps_common():
    return list(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "dgtta.h", Path(__file__)]


def _compile_one(src, obj, verbose, env):
    cmd = [_nvcc(), *NVCC_FLAGS, "-c", "-o", str(obj), str(src)]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
        print(" ".join(cmd), flush=True)
    res = subprocess.run(cmd, capture_output=True, text=True, env=env)
    return src, res


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ for sm_100a (one object per source, in parallel; only stale objects are rebuilt)
    and link them into one shared library; returns its path."""
    from concurrent.futures import ThreadPoolExecutor
    LIB_DIR.mkdir(exist_ok=True)
    OBJ_DIR.mkdir(exist_ok=True)
    env = dict(os.environ)
    env.pop("CC", None)   # the image exports a gcc wrapper nvcc should not be forced onto
    env.pop("CXX", None)
    common = max(d.stat().st_mtime for d in _deps_common())
    jobs, objs = [], []
    for s in SOURCES:
        src = CSRC / s
        if not src.exists():
            continue
        obj = OBJ_DIR / (src.stem + ".o")
        objs.append(obj)
        if force or not obj.exists() or obj.stat().st_mtime < max(common, src.stat().st_mtime):
            jobs.append((src, obj))
    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as ex:
            for src, res in ex.map(lambda j: _compile_one(j[0], j[1], verbose, env), jobs):
                if verbose or res.returncode:
                    sys.stderr.write(res.stdout + res.stderr)
                if res.returncode:
                    raise RuntimeError(f"nvcc failed on {src.name}")
    if jobs or not LIB_PATH.exists() or any(o.stat().st_mtime > LIB_PATH.stat().st_mtime for o in objs):
        cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-Xcompiler", "-fPIC", "-o", str(LIB_PATH),
               *map(str, objs)]
        res = subprocess.run(cmd, capture_output=True, text=True, env=env)
        if res.returncode:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("nvcc failed linking libdgtta_sm100.so")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
