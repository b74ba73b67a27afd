#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) in the build container.

The reference has no tests or golden vectors for the synthesis path, so the pins for the
oracle are outputs of the reference's own python files (imported through oracle/reference_shim.py) on small
seeded inputs.  Weights and inputs are not stored: they are re-drawn from the same numpy Philox streams
(emojivoice_b200/synthetic.py) and guarded by a checksum stored in the fixture.

    python scripts/make_golden.py            # writes tests/golden/{matcha_*,hifigan_*}.npz
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from emojivoice_b200 import synthetic  # noqa: E402
from emojivoice_b200.config import HIFIGAN_V1, VCTK  # noqa: E402
from oracle import reference_shim as shim  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

# name -> (batch, p_lo, p_hi, input seed, n_timesteps, temperature, length_scale, z seed)
MATCHA_CASES = {
    "matcha_b2_ls10": (2, 8, 12, 1, 4, 0.667, 1.0, 5),
    "matcha_b2_ls08": (2, 8, 12, 1, 4, 0.667, 0.8, 5),
    "matcha_b3_ragged": (3, 3, 30, 2, 2, 1.0, 1.1, 6),
    "matcha_cfg1": (1, 75, 75, 1234, 10, 0.667, 0.8, 1235),   # BASELINE.json configs[0]
}
# name -> (batch, frames, mel seed, weight kwargs)
HIFIGAN_CASES = {
    "hifigan_stock": (2, 40, 3, dict(seed=4321)),
    "hifigan_gain1": (1, 33, 4, dict(seed=4321, gain=1.0)),
}


def matcha_case(name, sd, ref):
    b, plo, phi, seed, n, temp, ls, zseed = MATCHA_CASES[name]
    x, xl, spk = synthetic.phoneme_batch(b, plo, phi, seed=seed)
    if name == "matcha_cfg1":
        spk = torch.tensor([107])
    # first pass only to learn T_pad (the reference draws z with that shape)
    probe = ref.synthesise(x, xl, n_timesteps=1, temperature=temp, spks=spk, length_scale=ls)
    y_max = int(probe["mel_lengths"].max())
    t_pad = -(-y_max // 4) * 4
    z = synthetic.prior_noise(b, VCTK.n_feats, t_pad, seed=zseed)
    with shim.injected_noise(z):
        out = ref.synthesise(x, xl, n_timesteps=n, temperature=temp, spks=spk, length_scale=ls)
    attn = out["attn"][:, 0]                                      # (B, Tx, T) 0/1
    np.savez_compressed(
        os.path.join(GOLD, name + ".npz"),
        meta=np.array([b, plo, phi, seed, n, zseed, t_pad], dtype=np.int64),
        temperature=np.float64(temp), length_scale=np.float64(ls),
        x=x.numpy(), x_lengths=xl.numpy(), spks=spk.numpy(),
        mel_lengths=out["mel_lengths"].numpy(),
        attn_bits=np.packbits(attn.numpy().astype(np.uint8), axis=-1), attn_shape=np.array(attn.shape),
        encoder_outputs=out["encoder_outputs"].numpy(), decoder_outputs=out["decoder_outputs"].numpy(),
        mel=out["mel"].numpy(), weights_checksum=np.float64(synthetic.checksum(sd)))
    print(f"{name}: T_pad={t_pad} mel_lengths={out['mel_lengths'].tolist()}")


def hifigan_case(name):
    b, frames, seed, wk = HIFIGAN_CASES[name]
    sd = synthetic.hifigan_state_dict(HIFIGAN_V1, **wk)
    gen = shim.build_hifigan(HIFIGAN_V1, sd)
    mel = synthetic.synthetic_mel(b, frames, seed=seed)
    with torch.inference_mode():
        wav = gen(mel)
        den = shim.build_denoiser(gen)
        clean = den(wav.clamp(-1, 1).squeeze(1), strength=0.00025)
    np.savez_compressed(
        os.path.join(GOLD, name + ".npz"), meta=np.array([b, frames, seed], dtype=np.int64),
        wav=wav.numpy(), bias_spec=den.bias_spec.numpy(), denoised=clean.numpy(),
        weights_checksum=np.float64(synthetic.checksum(sd)))
    print(f"{name}: wav {tuple(wav.shape)} absmax {wav.abs().max():.4f}")


def main():
    if not shim.available():
        sys.exit("reference tree not present; golden fixtures can only be regenerated in the build container")
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    sd = synthetic.matcha_state_dict(VCTK, seed=1234)
    ref = shim.build_matcha(VCTK, sd)
    n_params = sum(p.numel() for p in ref.parameters())
    assert n_params == 20857569, n_params
    for name in MATCHA_CASES:
        matcha_case(name, sd, ref)
    for name in HIFIGAN_CASES:
        hifigan_case(name)


if __name__ == "__main__":
    main()
