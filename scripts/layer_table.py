#!/usr/bin/env python
"""Per-layer-shape timing table of one bench step (CUDA events around every launch, EV_PROF_DETAIL=1).

    python scripts/layer_table.py [--batch 32] [--steps 10] > gpurun_out/layers.txt
"""
import argparse
import os
import sys

os.environ["EV_PROF_DETAIL"] = "1"
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import emojivoice_b200 as ev  # noqa: E402
from emojivoice_b200 import synthetic  # noqa: E402
from emojivoice_b200.config import HIFIGAN_V1, VCTK  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--steps", type=int, default=10)
    a = ap.parse_args()
    model = ev.MatchaTTS(**VCTK.constructor_kwargs(), precision="bf16", cuda_graphs=False)
    model.load_state_dict(synthetic.matcha_state_dict(VCTK, seed=1234))
    voc = ev.Generator(HIFIGAN_V1, precision="bf16", cuda_graphs=False)
    voc.load_state_dict(synthetic.hifigan_state_dict(HIFIGAN_V1, seed=4321))
    voc.remove_weight_norm()
    x, xl, spk = synthetic.phoneme_batch(a.batch, 60, 90, seed=2000)
    x, xl, spk = x.cuda(), xl.cuda(), spk.cuda()

    def step():
        out = model.synthesise(x, xl, a.steps, 0.667, spk, 0.8)
        return out, voc(out["mel"]).clamp(-1, 1)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    model._ctx.profile_begin(); voc._ctx.profile_begin()
    step()
    stats = model._ctx.profile_end() + voc._ctx.profile_end()
    tot = sum(s["total_ms"] for s in stats)
    print(f"total {tot:.2f} ms over {sum(s['launches'] for s in stats)} launches")
    for s in sorted(stats, key=lambda s: -s["total_ms"]):
        sec = s["total_ms"] / 1e3
        print(f"{s['name']:<46} n={s['launches']:<4} {s['total_ms']:8.3f} ms {100 * s['total_ms'] / tot:5.1f}%  "
              f"{s['total_ms'] * 1e3 / s['launches']:8.1f} us/launch  {s['flops'] / sec / 1e12:7.1f} TF/s  {s['bytes'] / sec / 1e9:7.0f} GB/s")


if __name__ == "__main__":
    main()
