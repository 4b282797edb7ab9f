"""Host-buffer front end of the hot path: pinned host tensors in, pinned host tensors out.

The reference moves every batch host -> device itself (nnU-Net's train_step / tta.py:510 call `.to(device)` on the
loader's CPU tensors) and keeps the transformed batch on the device; a caller that wants the descriptor back on the
host (a loader process that pre-computes MIND features, a test harness, bench.py's `e2e` leg) pays a 12x larger
device -> host copy.  This module overlaps the three legs of consecutive calls on three streams:

    copy-in stream :  H2D(i+1)
    caller's stream:            transform(i)           transform(i+1)
    copy-out stream:                         D2H(i)                     D2H(i+1)

so the steady-state cost per call is max(H2D, transform, D2H) instead of their sum.  Nothing here computes: the
transform is the package's ordinary CUDA path (default `gin_mind_aug`), and there is no CPU fallback.
"""
import torch

from . import _lib
from .tta.augmentation_utils import gin_mind_aug


class HostPipeline:
    """submit(h_in, h_out) enqueues H2D -> fn -> D2H and returns a CUDA event that fires when h_out is complete.

    h_in / h_out must be pinned (torch.Tensor.pin_memory()) for the copies to be asynchronous; the caller must not
    touch h_out (or overwrite h_in) before the returned event has fired.  Random draws inside `fn` happen in
    submission order on the caller's thread and stream, so seeds reproduce exactly like direct calls."""

    def __init__(self, device=None, fn=gin_mind_aug):
        _lib.lib()  # fail loudly if the CUDA library is missing
        if not torch.cuda.is_available():
            raise _lib.DgttaError("HostPipeline needs a CUDA device: dg_tta_b200 has no CPU path")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.fn = fn
        self.copy_in = torch.cuda.Stream(self.device)
        self.copy_out = torch.cuda.Stream(self.device)
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def submit(self, h_in, h_out):
        if h_in.is_cuda or h_out.is_cuda:
            raise TypeError("HostPipeline.submit takes host tensors (use the operators directly for device tensors)")
        if not (h_in.is_pinned() and h_out.is_pinned()):
            raise ValueError("h_in and h_out must be pinned host tensors (tensor.pin_memory())")
        main = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self.copy_in):
            x = h_in.to(self.device, non_blocking=True)
            arrived = torch.cuda.Event()
            arrived.record(self.copy_in)
        main.wait_event(arrived)
        x.record_stream(main)
        y = self.fn(x)
        if tuple(y.shape) != tuple(h_out.shape) or y.dtype != h_out.dtype:
            raise ValueError(f"h_out must be {tuple(y.shape)} {y.dtype}, got {tuple(h_out.shape)} {h_out.dtype}")
        computed = torch.cuda.Event()
        computed.record(main)
        self.copy_out.wait_event(computed)
        with torch.cuda.stream(self.copy_out):
            h_out.copy_(y, non_blocking=True)
            done = torch.cuda.Event()
            done.record(self.copy_out)
        y.record_stream(self.copy_out)
        self.h2d_bytes += h_in.numel() * h_in.element_size()
        self.d2h_bytes += h_out.numel() * h_out.element_size()
        return done

    def drain(self):
        self.copy_in.synchronize()
        torch.cuda.current_stream(self.device).synchronize()
        self.copy_out.synchronize()
