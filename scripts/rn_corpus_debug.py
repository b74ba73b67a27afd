#!/usr/bin/env python
"""Debug aid: run MatchaTTS.synthesise over the micro-batches of the config-3 corpus one at a time (blocking launches) and
report the first shape that fails.   CUDA_LAUNCH_BLOCKING=1 python scripts/rn_corpus_debug.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import emojivoice_b200 as ev  # noqa: E402
from emojivoice_b200 import sharding, synthetic  # noqa: E402
from emojivoice_b200.batch import collate  # noqa: E402
from emojivoice_b200.config import VCTK  # noqa: E402


def main():
    model = ev.MatchaTTS(**VCTK.constructor_kwargs()).eval()
    model.load_state_dict(synthetic.matcha_state_dict(VCTK, seed=1234))
    model.cuda_graphs = False
    utts = synthetic.mixed_length_corpus(1024)
    plan = sharding.shard([len(u[0]) for u in utts], 32, 0, 1, n_timesteps=2, sort=True)
    plan = sorted(plan, key=lambda m: -m.cost)
    for k, mb in enumerate(plan):
        x, xl, spks = collate(utts, mb.items)
        try:
            out = model.synthesise(x, xl, 2, 0.667, spks, 0.8)
            torch.cuda.synchronize()
            print(k, "ok  B", x.shape[0], "Tx", x.shape[1], "T_pad", out["mel"].shape[-1], flush=True)
        except Exception as e:  # noqa: BLE001
            print(k, "FAILED  B", x.shape[0], "Tx", x.shape[1], str(e)[:200], flush=True)
            break


if __name__ == "__main__":
    main()
