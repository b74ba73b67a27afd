#!/usr/bin/env python
"""Generate tests/golden/train_forward_*.npz: loss values and alignments of the UNMODIFIED reference's MatchaTTS.forward
(matcha_tts.py:154-245; its own monotonic_align wrapper on its own Cython kernel compiled under oracle/_ref) on small seeded
batches, with the random draws of CFM.compute_loss injected through torch.rand / torch.randn_like.  Build container only.

    python -m oracle.build_oracle && python scripts/make_golden_forward.py
"""
import os
import sys
from unittest import mock

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from emojivoice_b200 import synthetic  # noqa: E402
from emojivoice_b200.config import VCTK  # noqa: E402
from oracle import reference_shim as shim  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
# name -> (batch, p_lo, p_hi, input seed, draw seed, out_size, cut offsets)
CASES = {
    "train_forward_b3": (3, 4, 14, 41, 42, None, None),
    "train_forward_b4_cut": (4, 10, 24, 43, 44, 32, [3, 0, 7, 11]),
}


def main():
    sd = synthetic.matcha_state_dict(VCTK, seed=1234)
    ref = shim.build_matcha(VCTK, sd).eval()
    for name, (b, plo, phi, seed, dseed, out_size, offs) in CASES.items():
        x, xl, spk, y, yl = synthetic.training_batch(b, plo, phi, seed, VCTK.n_feats)
        frames = out_size or y.shape[-1]
        t, z = synthetic.training_draws(b, VCTK.n_feats, frames, dseed)
        patches = [mock.patch("torch.rand", lambda *a, **k: t.reshape(b, 1, 1).clone()),
                   mock.patch("torch.randn_like", lambda ref_t, *a, **k: z.clone())]
        if out_size is not None:     # the reference draws the offsets with random.choice(range(start, end)), matcha_tts.py:213-216
            it = iter(offs)
            patches.append(mock.patch("random.choice", lambda rng: next(it)))
        with torch.no_grad():
            for p in patches:
                p.start()
            try:
                dur, prior, diff, attn = ref(x, xl, y, yl, spks=spk, out_size=out_size)
            finally:
                for p in patches:
                    p.stop()
        np.savez_compressed(os.path.join(GOLD, name + ".npz"),
                            meta=np.array([b, plo, phi, seed, dseed, out_size or 0], dtype=np.int64),
                            offsets=np.array(offs if offs else [], dtype=np.int64),
                            x_lengths=xl.numpy(), y_lengths=yl.numpy(), y_checksum=np.float64(y.double().sum()),
                            losses=np.array([float(dur), float(prior), float(diff)], dtype=np.float64),
                            attn=np.packbits(attn.numpy().astype(np.uint8), axis=-1), attn_shape=np.array(attn.shape, dtype=np.int64))
        print(name, "dur %.6f prior %.6f diff %.6f" % (float(dur), float(prior), float(diff)), tuple(attn.shape))


if __name__ == "__main__":
    main()
