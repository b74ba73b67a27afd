#!/usr/bin/env python
"""Times the fused feed-forward kernel (ff_tc.cu) alone through ev_test_ff_block at the decoder's two shapes
(B = 32, T = 668 and 334); EV_FF_DEBUG isolates its phases (timing only)."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from emojivoice_b200 import _lib  # noqa: E402

ctx = _lib.Context()
D, inner = 256, 1024
g = torch.Generator().manual_seed(1)
mk = lambda *s: torch.randn(*s, generator=g).cuda()
ln_g, ln_b, w1, b1, w2, b2 = mk(D), mk(D), mk(inner, D) / 16, mk(inner), mk(D, inner) / 32, mk(D)
sa, sb = torch.exp(0.3 * mk(inner)), torch.exp(0.3 * mk(inner))
for B, T in ((32, 668), (32, 334), (32, 128), (8, 128)):
    x = mk(B, T, D)
    frac = float(os.environ.get("EV_FF_LEN_FRAC", "1.0"))     # < 1: ragged lengths U[frac*T, T] as in the bench batch
    lens = (torch.rand(B, generator=g) * (1 - frac) * T + frac * T).long().clamp(1, T).cuda()
    lens[0] = T
    out = torch.empty(B, T, D, device="cuda")
    us = C.c_float(0.0)
    for _ in range(2):
        ctx.check(_lib.lib().ev_test_ff_block(ctx.handle, *[_lib.ptr(t) for t in (x, ln_g, ln_b, w1, b1, sa, sb, w2, b2)], _lib.ptr(lens),
                                              B, T, inner, 0, _lib.ptr(out), 50, C.byref(us), _lib.stream_ptr()), "ev_test_ff_block")
    tiles = B * ((T + 127) // 128)
    fl = 4.0 * B * T * D * inner
    print(f"EV_FF_DEBUG={os.environ.get('EV_FF_DEBUG', '0')} B={B} T={T} tiles={tiles}: {us.value:8.1f} us  {fl / us.value / 1e6:7.1f} TFLOP/s")

if int(os.environ.get("EV_FF_DEBUG", "0")) & 16:
    import numpy as np
    buf = np.zeros(192, dtype=np.uint64)
    ctx.check(_lib.lib().ev_test_ff_trace(ctx.handle, buf.ctypes.data_as(C.c_void_p), 192), "ev_test_ff_trace")
    t0 = int(buf[64])
    rel = lambda i: int(buf[i]) - t0 if buf[i] else None
    print("worker: pdl_wait 0, LN batch-1 loads issued", rel(65), "batch-2 loads issued", rel(67), "a_ready arrive", rel(66), "| issuer: loop", rel(0), "a_ready seen", rel(1), "G1(0) issued", rel(2), "G1(1) issued", rel(3))
    for c in range(8):
        print(f"chunk {c}: worker acc1_full {rel(72 + 6 * c)} ld {rel(73 + 6 * c)} math {rel(74 + 6 * c)} p_free {rel(75 + 6 * c)} arrive {rel(76 + 6 * c)}"
              f" | issuer p_ready {rel(8 + 4 * c)} G2 issued {rel(9 + 4 * c)} G1(c+2) issued {rel(10 + 4 * c)}")
    print("acc2_full", rel(130), "output done", rel(131))
    for h in range(2):
        print(f"output block {h}: loads issued {rel(132 + 4 * h)} tmem_ld {rel(133 + 4 * h)} staged {rel(134 + 4 * h)} rows stored {rel(135 + 4 * h)}")
