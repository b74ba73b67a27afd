#!/usr/bin/env python
"""Host-side cost of the launch path (one B200): how long the host needs to ENQUEUE one config-2 step eagerly (no graphs),
how long the GPU needs to run it, what a graph capture costs, and what a replay costs.  Decides whether never-seen shapes
(config 3's first pass) are host-bound.   python scripts/host_cost.py"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import emojivoice_b200 as ev  # noqa: E402
from emojivoice_b200 import synthetic  # noqa: E402
from emojivoice_b200.config import HIFIGAN_V1, VCTK  # noqa: E402


def main():
    model = ev.MatchaTTS(**VCTK.constructor_kwargs(), precision="bf16", cuda_graphs=False)
    model.load_state_dict(synthetic.matcha_state_dict(VCTK, seed=1234))
    voc = ev.Generator(HIFIGAN_V1, precision="bf16", cuda_graphs=False)
    voc.load_state_dict(synthetic.hifigan_state_dict(HIFIGAN_V1, seed=4321))
    voc.remove_weight_norm()
    x, xl, spk = synthetic.phoneme_batch(32, 60, 90, seed=2000)
    x, xl, spk = x.cuda(), xl.cuda(), spk.cuda()

    def step():
        out = model.synthesise(x, xl, 10, 0.667, spk, 0.8)
        return out, voc(out["mel"], lengths=out["mel_lengths"])

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    for graphs in (False, True):
        model.cuda_graphs = voc.cuda_graphs = graphs
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        host, total = [], []
        for _ in range(8):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            step()
            t1 = time.perf_counter()
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            host.append((t1 - t0) * 1e3)
            total.append((t2 - t0) * 1e3)
        print(f"graphs={graphs}: host enqueue (incl. the y_max read-back wait) {min(host):.2f} ms, step wall {min(total):.2f} ms, "
              f"launches/step {model.launch_count(reset=True) + voc.launch_count(reset=True)} over 14 steps")
    # per stage, eager: host time of the C calls alone
    model.cuda_graphs = voc.cuda_graphs = False
    out = model.synthesise(x, xl, 10, 0.667, spk, 0.8)
    torch.cuda.synchronize()
    for name, fn in (("synthesise", lambda: model.synthesise(x, xl, 10, 0.667, spk, 0.8)),
                     ("vocoder ragged", lambda: voc(out["mel"], lengths=out["mel_lengths"])),
                     ("vocoder dense", lambda: voc(out["mel"]))):
        ts = []
        for _ in range(5):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn()
            t1 = time.perf_counter()
            torch.cuda.synchronize()
            ts.append(((t1 - t0) * 1e3, (time.perf_counter() - t0) * 1e3))
        print(f"eager {name}: host {min(t[0] for t in ts):.2f} ms, wall {min(t[1] for t in ts):.2f} ms")
    # capture cost of a fresh shape
    model.cuda_graphs = voc.cuda_graphs = True
    x2, xl2, spk2 = synthetic.phoneme_batch(32, 40, 70, seed=2001)
    x2, xl2, spk2 = x2.cuda(), xl2.cuda(), spk2.cuda()
    for i in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        o = model.synthesise(x2, xl2, 10, 0.667, spk2, 0.8)
        voc(o["mel"], lengths=o["mel_lengths"])
        torch.cuda.synchronize()
        print(f"fresh shape, call {i}: {(time.perf_counter() - t0) * 1e3:.1f} ms wall")


if __name__ == "__main__":
    main()
