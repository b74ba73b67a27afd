#!/usr/bin/env python
"""One bench step bracketed by cudaProfilerStart/Stop, for `ncu --profile-from-start off` (launch list / full capture).

    python scripts/profile_step.py [--batch 32] [--precision bf16] [--part all|matcha|vocoder]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import emojivoice_b200 as ev  # noqa: E402
from emojivoice_b200 import synthetic  # noqa: E402
from emojivoice_b200.config import HIFIGAN_V1, VCTK  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--part", default="all")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--dense-vocoder", action="store_true", help="vocode the padded frames too (bench.py's `padded_vocoder`)")
    a = ap.parse_args()
    model = ev.MatchaTTS(**VCTK.constructor_kwargs(), precision=a.precision)
    model.load_state_dict(synthetic.matcha_state_dict(VCTK, seed=1234))
    voc = ev.Generator(HIFIGAN_V1, precision=a.precision)
    voc.load_state_dict(synthetic.hifigan_state_dict(HIFIGAN_V1, seed=4321))
    voc.remove_weight_norm()
    x, xl, spk = synthetic.phoneme_batch(a.batch, 60, 90, seed=2000)
    x, xl, spk = x.cuda(), xl.cuda(), spk.cuda()

    def step():
        out = model.synthesise(x, xl, a.steps, 0.667, spk, 0.8)
        if a.part == "matcha":
            return out, None
        return out, voc(out["mel"], lengths=None if a.dense_vocoder else out["mel_lengths"]).clamp(-1, 1)

    for _ in range(2):
        out, wav = step()
    torch.cuda.synchronize()
    mel, mel_len = out["mel"].clone(), out["mel_lengths"].clone()
    torch.cuda.profiler.start()
    if a.part == "vocoder":
        voc(mel, lengths=None if a.dense_vocoder else mel_len)
    else:
        step()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("frames", int(out["mel_lengths"].sum()), "t_pad", out["t_pad"], "launches", model.launch_count(), voc.launch_count())


if __name__ == "__main__":
    main()
