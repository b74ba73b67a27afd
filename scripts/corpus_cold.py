#!/usr/bin/env python
"""First pass of synthesise_corpus over 1024 never-seen shapes in a fresh process, by lanes in flight: wall time, cudaMalloc
calls of the caching allocator, time in the first micro-batches.  Diagnostic (where does a cold corpus lose its time?)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import emojivoice_b200 as ev  # noqa: E402
from emojivoice_b200 import synthetic  # noqa: E402
from emojivoice_b200.config import HIFIGAN_V1, VCTK  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
model = ev.MatchaTTS(**VCTK.constructor_kwargs(), precision="bf16")
model.load_state_dict(synthetic.matcha_state_dict(VCTK, seed=1234))
voc = ev.Generator(HIFIGAN_V1, precision="bf16")
voc.load_state_dict(synthetic.hifigan_state_dict(HIFIGAN_V1, seed=4321))
voc.remove_weight_norm()
t0 = time.perf_counter()
lanes = ev.lanes_for(model, voc, n)
torch.cuda.synchronize()
print(f"in_flight {n}: lanes built in {time.perf_counter() - t0:.3f} s")
utts = synthetic.mixed_length_corpus(1024)
for rep in range(3):
    a0 = torch.cuda.memory_stats()["num_device_alloc"]
    t0 = time.perf_counter()
    res, st = ev.synthesise_corpus(model, voc, utts, batch_size=32, n_timesteps=10, temperature=0.667, length_scale=0.8, lanes=lanes)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"  pass {rep}: {dt:.3f} s wall, {st.audio_seconds / dt:8.1f} audio-s/s, {torch.cuda.memory_stats()['num_device_alloc'] - a0} cudaMallocs, "
          f"reserved {torch.cuda.memory_reserved() / 2**30:.1f} GiB")
