#!/usr/bin/env python
"""Throughput of the other BASELINE.json configurations on one B200 (not bench lines; recorded under profiles/).

  config 1: 1 utterance (Tx = 151, spk 107), n_timesteps 10            -> latency / RTF of the interactive apps
  config 3: 1024 mixed-length utterances through synthesise_corpus     -> micro-batched corpus throughput (1 GPU's view)
  config 4: config-2 batch x n_timesteps {2,4,10,50} x length_scale {0.8,1.0,1.2}
  config 5: vocoder only, synthetic mel, 10/20/30/60 s segments
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import emojivoice_b200 as ev  # noqa: E402
from emojivoice_b200 import synthetic  # noqa: E402
from emojivoice_b200.config import HIFIGAN_V1, VCTK  # noqa: E402

SR, HOP = 22050, 256


def timed(fn, reps=5, warm=3):
    for _ in range(warm):
        out = fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 1e3)
    return best, out


def main():
    model = ev.MatchaTTS(**VCTK.constructor_kwargs(), precision="bf16")
    model.load_state_dict(synthetic.matcha_state_dict(VCTK, seed=1234))
    voc = ev.Generator(HIFIGAN_V1, precision="bf16")
    voc.load_state_dict(synthetic.hifigan_state_dict(HIFIGAN_V1, seed=4321))
    voc.remove_weight_norm()

    def step(x, xl, spk, n=10, ls=0.8):
        out = model.synthesise(x, xl, n, 0.667, spk, ls)
        return out, voc(out["mel"], lengths=out["mel_lengths"]).clamp(-1, 1)

    print("# config 1: one utterance, Tx=151, n_timesteps=10, length_scale 0.8 (feel_me.py operating point)")
    x, xl, _ = synthetic.phoneme_batch(1, 75, 75, seed=1)
    spk = torch.tensor([107])
    x, xl, spk = x.cuda(), xl.cuda(), spk.cuda()
    t, (out, wav) = timed(lambda: step(x, xl, spk))
    secs = float(out["mel_lengths"].sum()) * HOP / SR
    print(f"B=1: {secs:.2f} audio-s in {t * 1e3:.2f} ms  -> RTF {t / secs:.5f}  ({secs / t:.0f}x real-time)")

    print("# config 4: config-2 batch (B=32), ODE steps x length_scale")
    x, xl, spk = synthetic.phoneme_batch(32, 60, 90, seed=2000)
    x, xl, spk = x.cuda(), xl.cuda(), spk.cuda()
    for n in (2, 4, 10, 50):
        for ls in (0.8, 1.0, 1.2):
            t, (out, wav) = timed(lambda: step(x, xl, spk, n, ls), reps=3, warm=3)
            secs = float(out["mel_lengths"].sum()) * HOP / SR
            print(f"n_timesteps={n:<3} length_scale={ls}: {secs:7.1f} audio-s in {t * 1e3:7.2f} ms -> {secs / t:8.1f} audio-s/s  (T_pad {out['t_pad']})")

    print("# config 5: vocoder only, synthetic mel (B chosen so that B*T ~ 21k frames)")
    for sec_len in (10, 20, 30, 60):
        T = int(round(sec_len * SR / HOP))
        B = max(1, 21376 // T)
        mel = synthetic.synthetic_mel(B, T, seed=5).cuda()
        t, wav = timed(lambda: voc(mel), reps=3, warm=3)
        secs = B * T * HOP / SR
        print(f"{sec_len:>2}-s segments: B={B:<3} T={T:<5}: {secs:7.1f} audio-s in {t * 1e3:7.2f} ms -> {secs / t:8.1f} audio-s/s")

    print("# config 3: 1024 mixed-length utterances (P~U[20,150]), 11 emoji speakers, micro-batches of 32 sorted by length")
    g = torch.Generator().manual_seed(1237)
    voices = list(ev.EMOJI_MAPPING_FEMALE.values())
    utts = []
    for i in range(1024):
        p = int(torch.randint(20, 151, (1,), generator=g))
        ids = ev.intersperse(torch.randint(1, 178, (p,), generator=g).tolist())
        utts.append((ids, voices[i % len(voices)]))
    for rep in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res, stats = ev.synthesise_corpus(model, voc, utts, batch_size=32, n_timesteps=10, temperature=0.667, length_scale=0.8)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        print(f"pass {rep}: {stats.utterances} utterances, {stats.audio_seconds:.1f} audio-s: device {stats.seconds:.3f} s -> "
              f"{stats.audio_seconds / stats.seconds:.1f} audio-s/s; wall incl. D2H + crop {wall:.3f} s -> {stats.audio_seconds / wall:.1f} audio-s/s")


if __name__ == "__main__":
    main()
