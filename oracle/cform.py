"""TEST INFRASTRUCTURE ONLY — ctypes front-end of the plain-C oracle (oracle/dgtta_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package dg_tta_b200 never does (tests/test_product_isolation.py
enforces that).  Each wrapper takes/returns numpy arrays in the reference's NCDHW float32 layout.

Reference lines restated by the C code: dg_tta/mind.py:98-164, dg_tta/gin.py:59-122,168-230,
dg_tta/tta/tta.py:505-551,571-575, dg_tta/tta/torch_utils.py:55-73 (details in
dgtta_oracle_impl.h).
"""
import ctypes
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "_build" / "libdgtta_oracle.so"
_lib = None

_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int)


def build(force=False):
    """Compile the C oracle with the committed Makefile (gcc only, no reference sources involved)."""
    src_mtime = max((_HERE / n).stat().st_mtime for n in ("dgtta_oracle.c", "dgtta_oracle_impl.h", "Makefile"))
    if force or not _SO.exists() or _SO.stat().st_mtime < src_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-s", "-B"], check=True, capture_output=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(str(_SO))
    return _lib


def _ptr(a, ct=ctypes.c_float):
    return a.ctypes.data_as(ctypes.POINTER(ct))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def gaussian_taps(sigma):
    """Taps of dg_tta/mind.py:27-37 evaluated the way the reference does (float32 tensor math):
    N = ceil(1.5*sigma)*2+1, w = exp(-x^2/(2 sigma^2)), w /= sum(w)."""
    s = np.float32(sigma)
    n = int(np.ceil(s * np.float32(3.0) / np.float32(2.0))) * 2 + 1
    xs = np.linspace(-(n // 2), n // 2, n, dtype=np.float32)
    w = np.exp(-(xs ** 2) / (np.float32(2) * s * s)).astype(np.float32)
    return (w / w.sum(dtype=np.float32)).astype(np.float32)


def shift_table():
    s1 = np.zeros((12, 3), np.int32)
    s2 = np.zeros((12, 3), np.int32)
    lib().oracle_mind_shift_table(_ptr(s1, ctypes.c_int), _ptr(s2, ctypes.c_int))
    return s1, s2


def mind_ssc(img, delta=1, sigma=1, randn_weighting=0.05, noise=None, precision="f32", taps=None):
    """MIND3D.forward (dg_tta/mind.py:142-164).  noise=None means randn_weighting is ignored
    (noise-free); otherwise `noise` is the tensor the reference would have drawn at mind.py:150."""
    # precision "f64x": double arithmetic AND a double input image (chain truth: GIN's f64 output is not rounded)
    img = np.ascontiguousarray(img, dtype=np.float64) if precision == "f64x" else _f32(img)
    B, C, D, H, W = img.shape
    assert C == 1
    taps = gaussian_taps(sigma) if taps is None else _f32(taps)
    dt = np.float32 if precision == "f32" else np.float64
    out = np.empty((B, 12, D, H, W), dt)
    nz = None
    if noise is not None:
        nz = _f32(noise)
        assert nz.shape == out.shape
    fn = getattr(lib(), f"oracle_mind_ssc_{precision}")
    fn.restype = ctypes.c_int
    rc = fn(_ptr(img, ctypes.c_double if precision == "f64x" else ctypes.c_float), _ptr(nz) if nz is not None else None,
            _ptr(out, ctypes.c_float if precision == "f32" else ctypes.c_double),
            B, D, H, W, int(delta), _ptr(taps), len(taps), ctypes.c_float(randn_weighting))
    if rc:
        raise RuntimeError(f"oracle_mind_ssc failed: {rc}")
    return out


def pack_gin_params(kers, shifts):
    """Layer-major packing shared by the oracle and the C-ABI: ker_L then shift_L, L = 0..n-1."""
    parts = []
    for k, s in zip(kers, shifts):
        parts.append(_f32(k).ravel())
        parts.append(_f32(s).ravel())
    return np.concatenate(parts)


def gin(x, kers, shifts, alphas, precision="f32", interm_channels=2):
    """GINGroupConv.forward (dg_tta/gin.py:168-230) with the random draws supplied by the caller."""
    x = _f32(x)
    B, C, D, H, W = x.shape
    ks = np.array([k.shape[-1] for k in kers], np.int32)
    params = pack_gin_params(kers, shifts)
    al = _f32(alphas).ravel()
    dt = np.float32 if precision == "f32" else np.float64
    out = np.empty(x.shape, dt)
    fn = getattr(lib(), f"oracle_gin_{precision}")
    fn.restype = ctypes.c_int
    rc = fn(_ptr(x), _ptr(out, ctypes.c_float if precision == "f32" else ctypes.c_double), _ptr(params),
            _ptr(ks, ctypes.c_int), _ptr(al), B, D, H, W, C, len(kers), interm_channels)
    if rc:
        raise RuntimeError(f"oracle_gin failed: {rc}")
    return out


_MODES = {"bilinear": 0, "nearest": 1}
_PADS = {"zeros": 0, "border": 1}


def affine_sample(src, theta, out_size, mode="bilinear", padding_mode="zeros", precision="f32"):
    """F.grid_sample(src, F.affine_grid(theta, out_size, align_corners=False), mode, padding_mode,
    align_corners=False) — the op pair at tta.py:524-551,571-575 and torch_utils.py:55-73."""
    src = _f32(src)
    theta = _f32(theta)
    B, C, Di, Hi, Wi = src.shape
    Do, Ho, Wo = [int(v) for v in out_size[-3:]]
    dt = np.float32 if precision == "f32" else np.float64
    out = np.empty((B, C, Do, Ho, Wo), dt)
    fn = getattr(lib(), f"oracle_affine_sample_{precision}")
    fn.restype = ctypes.c_int
    fn(_ptr(src), _ptr(theta), _ptr(out, ctypes.c_float if precision == "f32" else ctypes.c_double),
       B, C, Di, Hi, Wi, Do, Ho, Wo, _MODES[mode], _PADS[padding_mode])
    return out


def affine_sample_bwd_input(grad_out, theta, in_size, padding_mode="zeros", precision="f32"):
    """Gradient of the trilinear sampler w.r.t. its input (autograd through tta.py:573-575)."""
    grad_out = _f32(grad_out)
    theta = _f32(theta)
    B, C, Do, Ho, Wo = grad_out.shape
    Di, Hi, Wi = [int(v) for v in in_size[-3:]]
    dt = np.float32 if precision == "f32" else np.float64
    gin_ = np.empty((B, C, Di, Hi, Wi), dt)
    fn = getattr(lib(), f"oracle_affine_sample_bwd_input_{precision}")
    fn.restype = ctypes.c_int
    fn(_ptr(grad_out), _ptr(theta), _ptr(gin_, ctypes.c_float if precision == "f32" else ctypes.c_double),
       B, C, Di, Hi, Wi, Do, Ho, Wo, _PADS[padding_mode])
    return gin_


def consistency_loss(target_a, target_b, start_class=1, precision="f32", with_grad=True):
    """Loss of tta.py:263-269 (+ soft_dice_loss, torch_utils.py:90-104) and its gradient w.r.t. target_a."""
    ta, tb = _f32(target_a), _f32(target_b)
    B, C = ta.shape[:2]
    V = int(np.prod(ta.shape[2:]))
    dt, ct = (np.float32, ctypes.c_float) if precision == "f32" else (np.float64, ctypes.c_double)
    grad = np.empty(ta.shape, dt) if with_grad else None
    fn = getattr(lib(), f"oracle_consistency_loss_{precision}")
    fn.restype = ct
    loss = fn(_ptr(ta), _ptr(tb), _ptr(grad, ct) if with_grad else None, B, C, ctypes.c_long(V), int(start_class))
    return (float(loss), grad) if with_grad else float(loss)


def label_argmax(onehot, theta, out_size=None, precision="f32"):
    """get_argmaxed_segs(grid_sample(onehot, affine_grid(theta), mode="nearest", zeros)) — torch_utils.py:71-82."""
    oh, theta = _f32(onehot), _f32(theta)
    B, L, Di, Hi, Wi = oh.shape
    Do, Ho, Wo = [int(v) for v in (out_size[-3:] if out_size is not None else oh.shape[-3:])]
    out = np.empty((B, 1, Do, Ho, Wo), np.int64)
    fn = getattr(lib(), f"oracle_label_argmax_{precision}")
    fn.restype = ctypes.c_int
    fn(_ptr(oh), _ptr(theta), out.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)), B, L, Di, Hi, Wi, Do, Ho, Wo)
    return out
