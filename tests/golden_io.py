"""Helpers to read tests/golden/*.npz (written by scripts/make_golden.py from the reference itself)."""
import os

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MATCHA = ["matcha_b2_ls10", "matcha_b2_ls08", "matcha_b3_ragged", "matcha_cfg1"]
HIFIGAN = {"hifigan_stock": dict(seed=4321), "hifigan_gain1": dict(seed=4321, gain=1.0)}


def load(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    d = {k: g[k] for k in g.files}
    if "attn_bits" in d:
        shp = tuple(int(v) for v in d["attn_shape"])
        d["attn"] = torch.from_numpy(np.unpackbits(d["attn_bits"], axis=-1)[..., : shp[-1]].astype(np.float32))
    return d


def matcha_inputs(g):
    from emojivoice_b200 import synthetic

    b, _plo, _phi, _seed, n, zseed, t_pad = (int(v) for v in g["meta"])
    z = synthetic.prior_noise(b, 80, t_pad, seed=zseed)
    return dict(x=torch.from_numpy(g["x"]), x_lengths=torch.from_numpy(g["x_lengths"]),
                spks=torch.from_numpy(g["spks"]), n_timesteps=n, temperature=float(g["temperature"]),
                length_scale=float(g["length_scale"]), z=z)
