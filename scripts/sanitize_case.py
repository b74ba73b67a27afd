#!/usr/bin/env python
"""A small end-to-end case for `compute-sanitizer --tool memcheck|racecheck|synccheck python scripts/sanitize_case.py`: bf16
synthesise (fused ResNet / tail kernels, ragged tile lists), dense + ragged vocoder (paired ResBlock windows), the training-side
forward pass and the denoiser on a 3-utterance batch.  Prints a checksum so that a plain run and a sanitized run can be compared."""
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
x, xl, spk = synthetic.phoneme_batch(3, 10, 40, seed=5)
out = model.synthesise(x, xl, 2, 0.667, spk, 0.9)
wav_d = voc(out["mel"])
wav_r = voc(out["mel"], lengths=out["mel_lengths"])
tx, txl, tspk, y, yl = synthetic.training_batch(3, 6, 12, 71, VCTK.n_feats)
t, z = synthetic.training_draws(3, VCTK.n_feats, y.shape[-1], 72)
losses = [float(v) for v in model.forward(tx, txl, y, yl, spks=tspk, t=t, z=z, dtype="bf16")[:3]]
den = ev.Denoiser(voc)(wav_d.squeeze(1), strength=0.00025)
torch.cuda.synchronize()
n = int(out["mel_lengths"][0]) * 256
print("OK mel %.6f wav %.6f ragged==dense %s losses %s den %.6f" % (float(out["mel"].double().abs().mean()), float(wav_d.double().abs().mean()),
      bool(torch.equal(wav_d[0, 0, :n], wav_r[0, 0, :n])), ["%.5f" % v for v in losses], float(den.double().abs().mean())))
