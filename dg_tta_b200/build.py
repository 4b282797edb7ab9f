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
SOURCES = ["api.cu", "mind_ssc.cu", "mind_fast.cu", "mind_general.cu", "gin.cu", "gin_fused.cu", "affine_sample.cu", "philox_normal.cu", "consistency_loss.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "--shared",
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-O3",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: libdgtta_sm100.so cannot be built (there is no CPU fallback)")


def _stale():
    if not LIB_PATH.exists():
        return True
    t = LIB_PATH.stat().st_mtime
    deps = [CSRC / s for s in SOURCES if (CSRC / s).exists()] + list(CSRC.glob("*.cuh")) + \
        [PKG.parent / "include" / "dgtta.h", Path(__file__)]
    return any(d.stat().st_mtime > t for d in deps)


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ for sm_100a into one shared library; returns its path."""
    if not force and not _stale():
        return LIB_PATH
    LIB_DIR.mkdir(exist_ok=True)
    srcs = [str(CSRC / s) for s in SOURCES if (CSRC / s).exists()]
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", str(LIB_PATH), *srcs]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
        print(" ".join(cmd), flush=True)
    env = dict(os.environ)
    env.pop("CC", None)   # the image exports a gcc wrapper nvcc should not be forced onto
    env.pop("CXX", None)
    res = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if verbose or res.returncode:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode:
        raise RuntimeError("nvcc failed building libdgtta_sm100.so")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
