#!/usr/bin/env python
"""Stage-by-stage numerical diagnostics on the GPU box (prints, never asserts)."""
import ctypes as C
import os
import sys
import time
import traceback

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import emojivoice_b200 as ev  # noqa: E402
from emojivoice_b200 import _lib, synthetic  # noqa: E402
from emojivoice_b200.config import HIFIGAN_V1, VCTK  # noqa: E402
from oracle import hifigan_oracle as ho  # noqa: E402
from oracle import matcha_oracle as mo  # noqa: E402


def rel(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float(torch.linalg.norm(a - b) / torch.linalg.norm(b).clamp_min(1e-30))


def conv_case(ctx, B, Cin, T, Cout, K, stride, pad, dil, transposed, prec, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, Cin, T, generator=g)
    w = torch.randn((Cin, Cout, K) if transposed else (Cout, Cin, K), generator=g) / (Cin * K) ** 0.5
    b = torch.randn(Cout, generator=g)
    if prec == 1:
        xr, wr = x.bfloat16().float(), w.bfloat16().float()
    else:
        xr, wr = x, w
    if transposed:
        ref = F.conv_transpose1d(xr.double(), wr.double(), b.double(), stride=stride, padding=pad)
    else:
        ref = F.conv1d(xr.double(), wr.double(), b.double(), stride=stride, padding=pad, dilation=dil)
    y = torch.empty(ref.shape, device="cuda")
    xd, wd, bd = x.cuda(), w.cuda(), b.cuda()
    rc = _lib.lib().ev_test_conv1d(ctx.handle, _lib.ptr(xd), _lib.ptr(wd), _lib.ptr(bd), B, Cin, T, Cout, K, stride, pad,
                                   dil, int(transposed), prec, _lib.ptr(y), _lib.stream_ptr())
    torch.cuda.synchronize()
    if rc != 0:
        return f"rc={rc} {_lib.lib().ev_last_error(ctx.handle).decode()}"
    return f"rel={rel(y, ref):.3e}"


def main():
    print(torch.cuda.get_device_name(0), torch.__version__)
    ctx = _lib.Context()
    cases = [
        (2, 64, 200, 64, 1, 1, 0, 1, False), (2, 256, 300, 256, 3, 1, 1, 1, False), (1, 224, 130, 256, 3, 1, 1, 1, False),
        (2, 80, 77, 512, 7, 1, 3, 1, False), (2, 32, 1000, 32, 11, 1, 25, 5, False), (2, 128, 500, 128, 7, 1, 9, 3, False),
        (2, 256, 128, 256, 3, 2, 1, 1, False), (2, 256, 64, 256, 4, 2, 1, 1, True), (2, 512, 50, 256, 16, 8, 4, 1, True),
        (1, 64, 300, 32, 4, 2, 1, 1, True), (2, 256, 100, 80, 1, 1, 0, 1, False), (1, 1024, 200, 256, 1, 1, 0, 1, False),
        (2, 256, 140, 384, 1, 1, 0, 1, False),
    ]
    for prec in (0, 1):
        for c in cases:
            try:
                print("conv", "fp32" if prec == 0 else "bf16", c, conv_case(ctx, *c, prec), flush=True)
            except Exception as e:  # noqa
                print("conv EXC", c, e, flush=True)
                traceback.print_exc()
                return
    # ---- aten sum order
    rng = np.random.default_rng(0)
    bad = 0
    for n in (3, 7, 9, 17, 151, 333, 513, 1100):
        for ls in (0.8, 0.9, 1.1, 1.2):
            w = torch.from_numpy((np.ceil(np.exp(rng.normal(0.9, 0.6, size=(16, n)))) * ls).astype(np.float32))
            out = torch.empty(16, device="cuda")
            _lib.lib().ev_test_row_sum(ctx.handle, _lib.ptr(w.cuda()), 16, n, _lib.ptr(out), _lib.stream_ptr())
            bad += int((out.cpu() != torch.sum(w.view(16, 1, n), [1, 2])).sum())
    print("aten row-sum mismatches:", bad, flush=True)

    # ---- matcha
    sd = synthetic.matcha_state_dict(VCTK, seed=1234)
    model = ev.MatchaTTS(**VCTK.constructor_kwargs())
    t0 = time.time(); model.load_state_dict(sd); print("load matcha", time.time() - t0, flush=True)
    x, xl, spk = synthetic.phoneme_batch(3, 4, 30, seed=2)
    for ls in (1.0, 0.8):
        probe = mo.synthesise(sd, VCTK, x, xl, 1, 0.667, spk, ls)
        z = synthetic.prior_noise(3, 80, probe["t_pad"], seed=6)
        ref = mo.synthesise(sd, VCTK, x, xl, 4, 0.667, spk, ls, z=z, return_steps=True)
        for dtype in ("fp32", "bf16"):
            try:
                out = model.synthesise(x, xl, 4, 0.667, spk, ls, z=z, dtype=dtype)
            except Exception as e:  # noqa
                print("synthesise EXC", dtype, e, flush=True)
                continue
            torch.cuda.synchronize()
            print(f"ls={ls} {dtype}: lens gpu {out['mel_lengths'].tolist()} ref {ref['mel_lengths'].tolist()}",
                  "logw", f"{rel(out['logw'], ref['logw']):.2e}", "mu_x", f"{rel(out['mu_x'], ref['mu_x']):.2e}",
                  "w_ceil eq", bool(torch.equal(out['w_ceil'].cpu(), ref['w_ceil'])), flush=True)
            if out["t_pad"] == ref["t_pad"]:
                print("   attn eq", bool(torch.equal(out["attn"].cpu(), ref["attn"])), "enc_out",
                      f"{rel(out['encoder_outputs'], ref['encoder_outputs']):.2e}", "dec",
                      f"{rel(out['decoder_outputs'], ref['decoder_outputs']):.2e}", "mel", f"{rel(out['mel'], ref['mel']):.2e}",
                      "dec_full", f"{rel(out['decoder_outputs_full'], ref['decoder_outputs_full']):.2e}", flush=True)
        # one Euler step only, to separate estimator error from accumulation
        ref1 = mo.synthesise(sd, VCTK, x, xl, 1, 0.667, spk, ls, z=z)
        for dtype in ("fp32", "bf16"):
            out1 = model.synthesise(x, xl, 1, 0.667, spk, ls, z=z, dtype=dtype)
            print(f"   1-step {dtype} dec_full", f"{rel(out1['decoder_outputs_full'], ref1['decoder_outputs_full']):.2e}", flush=True)
    # ---- hifigan
    for name, kw in (("stock", dict(seed=4321)), ("gain1", dict(seed=4321, gain=1.0))):
        hs = synthetic.hifigan_state_dict(HIFIGAN_V1, **kw)
        gen = ev.Generator(HIFIGAN_V1)
        gen.load_state_dict(hs); gen.remove_weight_norm()
        mel = synthetic.synthetic_mel(2, 37, seed=3)
        ref = ho.generator(hs, HIFIGAN_V1, mel)
        for dtype in ("fp32", "bf16"):
            try:
                wav = gen(mel, dtype=dtype)
                torch.cuda.synchronize()
                print("hifigan", name, dtype, "rel", f"{rel(wav, ref):.3e}", "absmax", float(ref.abs().max()), flush=True)
            except Exception as e:  # noqa
                print("hifigan EXC", name, dtype, e, flush=True)
    print("launches matcha", model.launch_count(), flush=True)


if __name__ == "__main__":
    main()
