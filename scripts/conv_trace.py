#!/usr/bin/env python
"""Timeline of ONE tensor-core conv launch (EV_TC_TRACE=1): clock64 stamps written by each CTA at the pipeline's milestones
(include/emojivoice_b200.h: ev_test_conv_trace).  Prints medians over CTAs, in clocks relative to CTA entry.

    EV_TC_TRACE=1 python scripts/conv_trace.py [B Cin T Cout K [precision]]      (default: a decoder conv, 32 256 334 256 3;
    precision 1 = bf16, 2 = the text encoder's 3xTF32 path)
"""
import ctypes as C
import os
import sys

os.environ.setdefault("EV_TC_TRACE", "1")
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from emojivoice_b200 import _lib  # noqa: E402

NAMES = ["entry", "prologue done", "after pdl wait", "first TMA issued", "first operands landed", "first tile MMAs committed",
         "last commit", "first accumulator seen (epilogue)", "first tile stored", "last tile stored", "exit", "(globaltimer)",
         "producer: tile loop entered", "producer: first tile decoded", "producer: first slot free", "epilogue warp 0: first block done",
         "MMA warp: clk waiting for weight tiles", "MMA warp: clk waiting for activation tiles", "producer: clk waiting for a free weight slot",
         "producer: clk waiting for a free activation slot", "MMA warp: clk waiting for a drained accumulator"]


def main():
    a = [int(v) for v in sys.argv[1:6]] if len(sys.argv) >= 6 else [32, 256, 334, 256, 3]
    B, Cin, T, Cout, K = a
    prec = int(sys.argv[6]) if len(sys.argv) >= 7 else 1
    ctx = _lib.Context()
    x = torch.randn(B, Cin, T, device="cuda")
    w = torch.randn(Cout, Cin, K, device="cuda") / (Cin * K) ** 0.5
    b = torch.randn(Cout, device="cuda")
    y = torch.empty(B, Cout, T, device="cuda")
    L = _lib.lib()
    for rep in range(3):
        ctx.check(L.ev_test_conv1d(ctx.handle, _lib.ptr(x), _lib.ptr(w), _lib.ptr(b), B, Cin, T, Cout, K, 1, K // 2, 1, 0, prec,
                                   _lib.ptr(y), _lib.stream_ptr()), "ev_test_conv1d")
    buf = np.zeros(512 * 24, dtype=np.uint64)
    ctx.check(L.ev_test_conv_trace(ctx.handle, buf.ctypes.data_as(C.c_void_p), buf.size), "ev_test_conv_trace")
    t = buf.reshape(512, 24).astype(np.int64)
    live = t[:, 0] != 0
    t = t[live]
    print(f"conv B={B} Cin={Cin} T={T} Cout={Cout} K={K} precision={prec}: {t.shape[0]} CTAs traced")
    rel = t[:, :21] - t[:, :1]
    rel[:, 16:21] = t[:, 16:21]          # accumulated waits, not time stamps
    for i, n in enumerate(NAMES):
        if i == 11:
            continue
        col = rel[:, i][t[:, i] != 0] if i < 16 else rel[:, i]
        if col.size:
            print(f"  {i:2d} {n:<36} median {int(np.median(col)):>7} clk   min {int(col.min()):>7}   max {int(col.max()):>7}")
    gt = t[:, 11]
    print(f"  CTA start skew (globaltimer): {int(gt.max() - gt.min())} ns")


if __name__ == "__main__":
    main()
