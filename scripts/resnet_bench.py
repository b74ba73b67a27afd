#!/usr/bin/env python
"""Timing + in-kernel clock trace of the fused ResNet-block kernel (resnet_tc.cu) at the decoder's shapes (one B200).
    EV_RN_TRACE=1 python scripts/resnet_bench.py"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from emojivoice_b200 import _lib  # noqa: E402

NAMES = {0: "entry", 15: "first input tile landed", 13: "constants folded", 16: "apply 1: first tmem_ld done", 17: "math done", 18: "stored", 14: "apply 1 half 0 done", 1: "conv1 issued", 2: "acc1 seen", 3: "barrier 1 passed", 4: "apply 1 done", 7: "conv2 issued", 8: "acc2 seen",
         5: "barrier 2 passed", 9: "apply 2 -> TMEM done", 10: "res issued", 11: "res half 0 seen", 6: "xr stored", 12: "exit of epilogue"}


def main():
    ctx = _lib.Context()
    L = _lib.lib()
    D = 256
    for (B, T, Cin, shift, full) in [(32, 334, 256, 1, 1), (32, 334, 512, 1, 1), (32, 668, 224, 0, 1), (32, 668, 512, 0, 1), (32, 668, 256, 0, 0),
                                     (8, 334, 256, 1, 1), (1, 300, 256, 0, 1)]:
        g = torch.Generator().manual_seed(1)
        r = lambda *s, k=1.0: (torch.randn(*s, generator=g) * k)
        w = {"conv1.weight": r(D, Cin, 3, k=(2 / (3 * Cin)) ** 0.5), "conv1.bias": r(D, k=0.1), "gn1.weight": 1 + r(D, k=0.1), "gn1.bias": r(D, k=0.1),
             "temb": r(D, k=0.5), "conv2.weight": r(D, D, 3, k=(2 / (3 * D)) ** 0.5), "conv2.bias": r(D, k=0.1), "gn2.weight": 1 + r(D, k=0.1),
             "gn2.bias": r(D, k=0.1), "res.weight": r(D, Cin, 1, k=Cin ** -0.5), "res.bias": r(D, k=0.1), "ln.weight": 1 + r(D, k=0.1), "ln.bias": r(D, k=0.1)}
        arr, keep = _lib.tensor_list(w, ctx.device)
        x = r(B, Cin, T).cuda()
        lens = torch.randint((T << shift) // 2, (T << shift) + 1, (B,), generator=g).cuda()
        outs = [torch.empty(B, T, D, device="cuda") for _ in range(3)]
        us = C.c_float(0)
        ctx.check(L.ev_test_resnet_block(ctx.handle, arr, len(keep), _lib.ptr(x), _lib.ptr(lens), B, T, Cin, shift, full, *[_lib.ptr(o) for o in outs],
                                         50, C.byref(us), _lib.stream_ptr()), "resnet")
        flops = 2.0 * B * T * D * (3 * Cin + (3 * D + Cin if full else 0))
        print(f"B={B} T={T} C_in={Cin} full={full}: {us.value:7.2f} us per launch (incl. a 2 KB memset)  {flops / us.value / 1e6:7.1f} TFLOP/s")
        if os.environ.get("EV_RN_TRACE"):
            buf = (C.c_uint64 * 32)()
            ctx.check(L.ev_test_resnet_trace(ctx.handle, buf, 32), "trace")
            t0 = buf[0]
            print("   trace (clk from entry, CTA 0): " + ", ".join(f"{NAMES[i]} {buf[i] - t0}" for i in (15, 1, 2, 3, 13, 16, 17, 18, 14, 4, 7, 8, 5, 9, 10, 11, 6, 12) if buf[i] >= t0 and buf[i] != 0))


if __name__ == "__main__":
    main()
