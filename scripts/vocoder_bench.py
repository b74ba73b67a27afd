#!/usr/bin/env python
"""Vocoder only on a config-2-shaped mel batch (B = 32, T_pad = 668, lengths 368..548): total time (CUDA events, eager),
per-layer table (EV_PROF_DETAIL) and a hash of the valid samples -- run under different EV_RB_* switches to compare
schedules of the fused ResBlock kernel (the hash must not change where the arithmetic is the same).

    EV_RB_CLUSTER=1 python scripts/vocoder_bench.py ; EV_RB_CLUSTER=2 python scripts/vocoder_bench.py [--dense] [--table]
"""
import argparse
import hashlib
import os
import sys

os.environ.setdefault("EV_PROF_DETAIL", "1")
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import emojivoice_b200 as ev  # noqa: E402
from emojivoice_b200 import synthetic  # noqa: E402
from emojivoice_b200.config import HIFIGAN_V1  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--frames", type=int, default=668)
    ap.add_argument("--dense", action="store_true")
    ap.add_argument("--table", action="store_true")
    ap.add_argument("--reps", type=int, default=10)
    a = ap.parse_args()
    voc = ev.Generator(HIFIGAN_V1, precision="bf16", cuda_graphs=False)
    voc.load_state_dict(synthetic.hifigan_state_dict(HIFIGAN_V1, seed=4321))
    voc.remove_weight_norm()
    g = torch.Generator().manual_seed(77)
    mel = (torch.randn(a.batch, 80, a.frames, generator=g) * 1.5 - 5.0).cuda()
    lens = torch.randint(int(a.frames * 0.55), int(a.frames * 0.82), (a.batch,), generator=g)
    lens[0] = a.frames
    lens = lens.cuda()
    kw = {} if a.dense else {"lengths": lens}
    for _ in range(3):
        wav = voc(mel, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        wav = voc(mel, **kw)
    e1.record()
    torch.cuda.synchronize()
    h = hashlib.sha256()
    w = wav.float().cpu()
    for b in range(a.batch):
        h.update(w[b, 0, : int(lens[b]) * 256].numpy().tobytes())
    sw = {k: v for k, v in os.environ.items() if k.startswith("EV_RB")}
    print(f"vocoder {'dense' if a.dense else 'ragged'} {sw}: {e0.elapsed_time(e1) / a.reps:.3f} ms per call, sha {h.hexdigest()[:16]}, "
          f"finite {bool(torch.isfinite(w).all())}, rms {float(w.pow(2).mean().sqrt()):.4f}")
    if a.table:
        voc._ctx.profile_begin()
        voc(mel, **kw)
        stats = voc._ctx.profile_end()
        tot = sum(s["total_ms"] for s in stats)
        print(f"total {tot:.2f} ms over {sum(s['launches'] for s in stats)} launches")
        for s in sorted(stats, key=lambda s: -s["total_ms"]):
            sec = s["total_ms"] / 1e3
            print(f"{s['name']:<46} n={s['launches']:<4} {s['total_ms']:8.3f} ms {100 * s['total_ms'] / tot:5.1f}%  "
                  f"{s['total_ms'] * 1e3 / s['launches']:8.1f} us/launch  {s['flops'] / sec / 1e12:7.1f} TF/s  {s['bytes'] / sec / 1e9:7.0f} GB/s")


if __name__ == "__main__":
    main()
