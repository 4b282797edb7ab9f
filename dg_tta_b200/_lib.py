"""ctypes binding of libdgtta_sm100.so (C ABI declared in include/dgtta.h).

There is deliberately no fallback: if the CUDA library is missing or fails to load, every
operator of this package raises.  The oracle under oracle/ is test infrastructure and is never
imported from here.
"""
import ctypes
from ctypes import c_char_p, c_float, c_int, c_size_t, c_uint64, c_void_p
from pathlib import Path

import os

# DGTTA_LIB_PATH lets a developer A/B-test an experimental build of the same CUDA library; there is still no non-CUDA path.
LIB_PATH = Path(os.environ.get("DGTTA_LIB_PATH") or Path(__file__).resolve().parent / "lib" / "libdgtta_sm100.so")
_lib = None

# name -> (restype, argtypes); mirrors include/dgtta.h one to one (checked by tests/test_abi_symbols.py)
SIGNATURES = {
    "dgtta_abi_version": (c_int, []),
    "dgtta_last_error": (c_char_p, []),
    "dgtta_launch_count": (c_uint64, []),
    "dgtta_preload_kernels": (c_int, []),
    "dgtta_mind_workspace_bytes": (c_size_t, [c_int] * 4),
    "dgtta_mind_ssc_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int,
                                   c_float, c_int, c_void_p, c_uint64, c_uint64, c_void_p, c_size_t, c_void_p]),
    "dgtta_mind_philox_offset_increment": (c_uint64, [c_int] * 6),
    "dgtta_philox_normal_fill": (c_int, [c_void_p, c_uint64, c_uint64, c_uint64, c_int, c_int, c_void_p]),
    "dgtta_philox_normal_offset_increment": (c_uint64, [c_uint64, c_int, c_int]),
    "dgtta_philox_normal_fill_graphsafe": (c_int, [c_void_p, c_uint64, c_void_p, c_int, c_int, c_void_p]),
    "dgtta_gin_workspace_bytes": (c_size_t, [c_int] * 7),
    "dgtta_gin_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                              c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "dgtta_gin_layer_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                    c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "dgtta_affine_sample_fwd": (c_int, [c_void_p, c_void_p, c_void_p] + [c_int] * 10 + [c_void_p]),
    "dgtta_affine_sample_bwd_input": (c_int, [c_void_p, c_void_p, c_void_p] + [c_int] * 9 + [c_void_p]),
    "dgtta_consistency_sums_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, ctypes.c_longlong, c_void_p]),
    "dgtta_consistency_sums_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, ctypes.c_longlong, c_void_p]),
    "dgtta_consistency_warp_sums_fwd": (c_int, [c_void_p] * 5 + [c_int] * 5 + [c_void_p]),
    "dgtta_consistency_warp_sums_bwd": (c_int, [c_void_p] * 6 + [c_int] * 5 + [c_void_p]),
    "dgtta_affine_label_argmax": (c_int, [c_void_p, c_void_p, c_void_p] + [c_int] * 8 + [c_void_p]),
    "dgtta_label_map_from_onehot": (c_int, [c_void_p, c_void_p, c_int, c_int, ctypes.c_longlong, c_void_p]),
    "dgtta_affine_label_gather": (c_int, [c_void_p, c_void_p, c_void_p] + [c_int] * 7 + [c_void_p]),
    "dgtta_affine_crop_shifted_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p] + [c_int] * 8 + [c_void_p]),
    "dgtta_volume_min_workspace_bytes": (c_size_t, []),
    "dgtta_resize_edge_workspace_bytes": (c_size_t, [c_int] * 8),
    "dgtta_resize_edge": (c_int, [c_void_p, c_void_p] + [c_int] * 8 + [c_void_p, c_size_t, c_void_p]),
    "dgtta_volume_min": (c_int, [c_void_p, ctypes.c_longlong, c_void_p, c_void_p, c_size_t, c_void_p]),
}


class DgttaError(RuntimeError):
    pass


def lib():
    """Load the shared library once.  Raises if it has not been built (python -m dg_tta_b200.build)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise DgttaError(
                f"{LIB_PATH} is missing: build it with `python -m dg_tta_b200.build` "
                "(nvcc, sm_100a).  dg_tta_b200 has no CPU or PyTorch fallback.")
        handle = ctypes.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name, None)
            if fn is None:
                continue  # optional symbols are checked where they are used
            fn.restype = res
            fn.argtypes = args
        if handle.dgtta_abi_version() != 1:
            raise DgttaError("libdgtta_sm100.so ABI version mismatch")
        _lib = handle
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().dgtta_last_error()
        raise DgttaError(f"{what} failed (status {rc}): {msg.decode() if msg else ''}")


_preloaded = set()


def stream_ptr():
    """Current torch stream of the current device as a cudaStream_t.  Every operator passes through here, so this is
    also where the library's kernels are loaded into a device's context the first time that device is used."""
    import torch
    dev = torch.cuda.current_device()
    if dev not in _preloaded:
        _preloaded.add(dev)
        check(lib().dgtta_preload_kernels(), "dgtta_preload_kernels")
    return c_void_p(torch.cuda.current_stream().cuda_stream)


_scratch = {}            # (device, stream, tag) -> tensor; insertion order = age
_SCRATCH_MAX = 4         # buffers kept at most (each can be hundreds of MB: 679 MB for MIND's noise field at 2x192^3)


def scratch(device, tag, numel):
    """A float32 work buffer of >= numel elements that persists between calls (keyed by device, current stream and
    tag), for intermediates that never leave an operator: the 12-channel noise field of MIND is 679 MB at 2x192^3, and
    re-allocating it per call makes the caching allocator's pool depth — and with it cudaMalloc stalls on the host —
    depend on how far the host runs ahead of the GPU.  Reuse is ordered by the stream the buffer is keyed on.

    The cache is bounded: at most _SCRATCH_MAX buffers live (least recently used first out — a dead stream's buffer
    cannot pile up), and a buffer more than twice as large as the request is dropped and re-allocated at the requested
    size.  release_scratch() frees everything (e.g. before the nnU-Net backbone needs the memory)."""
    import torch
    key = (device.index if device.index is not None else torch.cuda.current_device(),
           torch.cuda.current_stream(device).cuda_stream, tag)
    buf = _scratch.pop(key, None)
    if buf is not None and (buf.numel() < numel or buf.numel() > 2 * max(numel, 1)):
        buf = None
    if buf is None:
        while len(_scratch) >= _SCRATCH_MAX:
            _scratch.pop(next(iter(_scratch)))
        buf = torch.empty(numel, device=device, dtype=torch.float32)
    _scratch[key] = buf          # re-inserted: most recently used
    return buf[:numel]


def release_scratch():
    _scratch.clear()


def require_no_grad(t, name, what):
    """MIND / GIN are forward-only here (the reference never back-propagates through them: inputs carry no gradient and
    gin.py:57 asserts it for its weights).  Refuse loudly instead of silently detaching."""
    import torch
    if torch.is_grad_enabled() and t.requires_grad:
        raise RuntimeError(f"{what}: {name} requires grad, but this operator has no backward (the DG-TTA loops never need "
                           "one: MIND/GIN run on inputs without gradient); call it under torch.no_grad() or detach the input")


def require_cuda_f32(t, name):
    """Boundary contract: CUDA, float32; made contiguous by the caller."""
    import torch
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise TypeError(f"{name} must live on a CUDA device: dg_tta_b200 has no CPU path")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
