"""TEST INFRASTRUCTURE ONLY -- ctypes loader of oracle/mas_oracle.c (CPU restatement of
Matcha-TTS/matcha/utils/monotonic_align/core.pyx:11-47) and of the python wrapper
Matcha-TTS/matcha/utils/monotonic_align/__init__.py:7-22.  Never imported by the product path."""
import ctypes as C
import os

import numpy as np

from . import build_oracle

_lib = None


def _load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build_oracle.build_c())
        _lib.mas_oracle_maximum_path.restype = None
        _lib.mas_oracle_maximum_path.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float]
    return _lib


def maximum_path_c(paths, values, t_xs, t_ys, max_neg_val=-1e9):
    """core.pyx:42-47 -- same argument meaning: int32 paths (zero-filled) and float32 values are modified in place."""
    assert paths.dtype == np.int32 and values.dtype == np.float32 and paths.flags.c_contiguous and values.flags.c_contiguous
    t_xs = np.ascontiguousarray(t_xs, dtype=np.int32)
    t_ys = np.ascontiguousarray(t_ys, dtype=np.int32)
    b, tx, ty = values.shape
    _load().mas_oracle_maximum_path(paths.ctypes.data, values.ctypes.data, t_xs.ctypes.data, t_ys.ctypes.data, b, tx, ty,
                                    C.c_float(max_neg_val))


def maximum_path(value, mask):
    """monotonic_align/__init__.py:7-22 on torch tensors (value, mask: [b, t_x, t_y])."""
    import torch

    value = value * mask
    device, dtype = value.device, value.dtype
    v = np.ascontiguousarray(value.detach().cpu().numpy().astype(np.float32))
    path = np.zeros_like(v).astype(np.int32)
    m = mask.detach().cpu().numpy()
    t_x_max = m.sum(1)[:, 0].astype(np.int32)
    t_y_max = m.sum(2)[:, 0].astype(np.int32)
    maximum_path_c(path, v, t_x_max, t_y_max)
    return torch.from_numpy(path).to(device=device, dtype=dtype)
