"""Drop-in for `matcha.utils.monotonic_align.maximum_path` (Matcha-TTS/matcha/utils/monotonic_align/__init__.py:7-22),
the training-side alignment search MatchaTTS.forward calls at models/matcha_tts.py:198.  The dynamic programme runs in
libemojivoice_b200.so (csrc/mas.cu) on the tensor's own device -- no host round trip, no CPU fallback."""
from __future__ import annotations

import torch

from . import _lib

_ctx = {}


def _context(device):
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    if key not in _ctx:
        _ctx[key] = _lib.Context(torch.device("cuda", key))
    return _ctx[key]


@torch.no_grad()
def maximum_path(value: torch.Tensor, mask: torch.Tensor, max_neg_val: float = -1e9) -> torch.Tensor:
    """value, mask: [b, t_x, t_y] -> 0/1 path of value's dtype on value's device (same contract as the reference)."""
    if value.dim() != 3 or mask.shape != value.shape:
        raise ValueError("value and mask must both be [b, t_x, t_y]")
    if not value.is_cuda:
        raise RuntimeError("emojivoice_b200.maximum_path runs on CUDA tensors only (there is no CPU fallback)")
    ctx, L = _context(value.device), _lib.lib()
    dtype = value.dtype
    with torch.cuda.device(value.device):
        v = (value * mask).to(torch.float32).contiguous()                  # __init__.py:13,16
        t_x = mask.sum(1)[:, 0].to(torch.int32).contiguous()               # __init__.py:20
        t_y = mask.sum(2)[:, 0].to(torch.int32).contiguous()               # __init__.py:21
        b, tx, ty = v.shape
        path = torch.empty(b, tx, ty, dtype=torch.int32, device=v.device)
        ws = ctx.workspace(L.ev_maximum_path_workspace_bytes(ctx.handle, b, tx, ty))
        ctx.check(L.ev_maximum_path(ctx.handle, _lib.ptr(v), _lib.ptr(t_x), _lib.ptr(t_y), b, tx, ty, float(max_neg_val),
                                    _lib.ptr(path), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()), "ev_maximum_path")
    return path.to(dtype)
