#!/usr/bin/env python
"""Text-encoder attention alone at the bench shape (B = 32, Tx = 177, 2 heads of 128): fp32 CUDA-core kernel vs the tcgen05
3xFP16 kernel -- average launch time (CUDA events) and each one's error against a float64 evaluation of
text_encoder.py:223-246.  EV_ENC_ATTN_TRACE=1 adds the tcgen05 kernel's clock stamps."""
import ctypes as C
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from emojivoice_b200 import _lib  # noqa: E402


def rope64(x, d=64, base=10000.0):
    """rotate-half RoPE on the first d features of (B, H, T, 128), float64"""
    t = x.shape[2]
    theta = 1.0 / (base ** (torch.arange(0, d, 2, device=x.device).float() / d))
    ang = torch.einsum("n,d->nd", torch.arange(t, device=x.device).float(), theta)
    ang = torch.cat([ang, ang], dim=1).double()
    xr, xp = x[..., :d], x[..., d:]
    neg_half = torch.cat([-xr[..., d // 2:], xr[..., : d // 2]], dim=-1)
    return torch.cat([xr * ang.cos() + neg_half * ang.sin(), xp], dim=-1)


ctx = _lib.Context("cuda:0")
shapes = ((32, 177), (32, 120), (32, 384), (4, 177)) if len(sys.argv) < 2 else ((32, int(sys.argv[1])),)
for B, T in shapes:
    g = torch.Generator().manual_seed(1)
    qkv = (torch.randn(B, T, 768, generator=g) * 1.7).cuda()
    lens = torch.randint(T // 2, T + 1, (B,), generator=g).cuda()
    q, k, v = (z.reshape(B, T, 2, 128).transpose(1, 2).double() for z in qkv.chunk(3, dim=2))
    q, k = rope64(q), rope64(k)
    m = (torch.arange(T, device="cuda")[None, :] < lens[:, None]).double()
    sc = (q @ k.transpose(-2, -1)) / math.sqrt(128)
    sc = sc.masked_fill((m[:, None, :, None] * m[:, None, None, :]) == 0, -1e4)
    ref = (torch.softmax(sc, dim=-1) @ v).transpose(1, 2).reshape(B, T, 256)
    for impl in (0, 1):
        out = torch.empty(B, T, 256, device="cuda")
        us = C.c_float(0.0)
        ctx.check(_lib.lib().ev_test_encoder_attention(ctx.handle, _lib.ptr(qkv), _lib.ptr(lens), B, T, 2, impl, _lib.ptr(out), 50,
                                                       C.byref(us), _lib.stream_ptr()), "ev_test_encoder_attention")
        err = float((out.double() - ref).norm() / ref.norm())
        print(f"B={B} T={T} impl={impl}: {us.value:8.2f} us per launch, rel-L2 vs float64 {err:.3e}")
