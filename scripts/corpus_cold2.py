#!/usr/bin/env python
"""bench.py's order of events: config-2 steps on three lanes (graphs captured, workspaces sized for Tx = 181), THEN the first pass of the
1024-utterance corpus.  Prints the host wall time at which each micro-batch was enqueued.  Diagnostic."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import emojivoice_b200 as ev  # noqa: E402
from emojivoice_b200 import batch, synthetic  # noqa: E402
from emojivoice_b200.config import HIFIGAN_V1, VCTK  # noqa: E402

model = ev.MatchaTTS(**VCTK.constructor_kwargs(), precision="bf16")
model.load_state_dict(synthetic.matcha_state_dict(VCTK, seed=1234))
voc = ev.Generator(HIFIGAN_V1, precision="bf16")
voc.load_state_dict(synthetic.hifigan_state_dict(HIFIGAN_V1, seed=4321))
voc.remove_weight_norm()
lanes = ev.lanes_for(model, voc, 3)
x, xl, spk = synthetic.phoneme_batch(32, 60, 90, seed=2000)
x, xl, spk = x.cuda(), xl.cuda(), spk.cuda()
for i in range(12):
    m, v, st = lanes.models[i % 3], lanes.vocoders[i % 3], lanes.streams[i % 3]
    with torch.cuda.stream(st):
        out = m.synthesise(x, xl, 10, 0.667, spk, 0.8)
        v(out["mel"], lengths=out["mel_lengths"]).clamp(-1, 1)
torch.cuda.synchronize()
utts = synthetic.mixed_length_corpus(1024)
marks = []
orig = batch.collate


def collate(u, items):
    marks.append(time.perf_counter())
    return orig(u, items)


batch.collate = collate
for rep in range(2):
    marks.clear()
    a0 = torch.cuda.memory_stats()["num_device_alloc"]
    t0 = time.perf_counter()
    res, st = ev.synthesise_corpus(model, voc, utts, batch_size=32, n_timesteps=10, temperature=0.667, length_scale=0.8, lanes=lanes)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"pass {rep}: {dt:.3f} s, {st.audio_seconds / dt:.0f} audio-s/s, {torch.cuda.memory_stats()['num_device_alloc'] - a0} cudaMallocs, reserved {torch.cuda.memory_reserved() / 2**30:.1f} GiB")
    print("   enqueue times of the micro-batches (ms):", " ".join(f"{1e3 * (m - t0):.0f}" for m in marks))
