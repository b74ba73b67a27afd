#!/usr/bin/env python
"""Synthesises one fixed small batch (bf16 mode, decoder + vocoder) and prints a SHA-256 of mel and waveform: tests run it
in subprocesses under different scheduling knobs (EV_DEC_LANES, EV_RB_WAVE, EV_RB_OCC2, EV_PDL) -- schedules must not change
a single bit."""
import hashlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import emojivoice_b200 as ev  # noqa: E402
from emojivoice_b200 import synthetic  # noqa: E402
from emojivoice_b200.config import HIFIGAN_V1, VCTK  # noqa: E402

model = ev.MatchaTTS(**VCTK.constructor_kwargs(), precision="bf16", cuda_graphs=False)
model.load_state_dict(synthetic.matcha_state_dict(VCTK, seed=1234))
voc = ev.Generator(HIFIGAN_V1, precision="bf16", cuda_graphs=False)
voc.load_state_dict(synthetic.hifigan_state_dict(HIFIGAN_V1, seed=4321, gain=1.0))
voc.remove_weight_norm()
x, xl, spk = synthetic.phoneme_batch(9, 20, 60, seed=5)
probe = model.synthesise(x, xl, 1, 0.667, spk, 0.8)
z = synthetic.prior_noise(9, 80, probe["t_pad"], seed=6)
out = model.synthesise(x, xl, 4, 0.667, spk, 0.8, z=z)
wav = voc(out["mel"])
torch.cuda.synchronize()
h = hashlib.sha256()
h.update(out["mel"].cpu().numpy().tobytes())
h.update(wav.cpu().numpy().tobytes())
print("HASH", h.hexdigest())
