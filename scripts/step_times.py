#!/usr/bin/env python
"""Per-step GPU/host timing of the bench workload (diagnostic): prints event-timed and wall-clock ms for each step,
first without and then with a host sync + waveform read-back per step."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import emojivoice_b200 as ev  # noqa: E402
from emojivoice_b200 import synthetic  # noqa: E402
from emojivoice_b200.config import HIFIGAN_V1, VCTK  # noqa: E402

model = ev.MatchaTTS(**VCTK.constructor_kwargs(), precision="bf16")
model.load_state_dict(synthetic.matcha_state_dict(VCTK, seed=1234))
voc = ev.Generator(HIFIGAN_V1, precision="bf16")
voc.load_state_dict(synthetic.hifigan_state_dict(HIFIGAN_V1, seed=4321))
voc.remove_weight_norm()
x, xl, spk = synthetic.phoneme_batch(int(os.environ.get("EV_BATCH", "32")), 60, 90, seed=2000)
x, xl, spk = x.cuda(), xl.cuda(), spk.cuda()
host = None
for mode in ("async", "sync+d2h"):
    for i in range(16):
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        t0 = time.perf_counter()
        e0.record()
        out = model.synthesise(x, xl, 10, 0.667, spk, 0.8)
        t1 = time.perf_counter()
        e1.record()
        wav = voc(out["mel"]).clamp(-1, 1)
        t2 = time.perf_counter()
        if mode != "async":
            if host is None:
                host = torch.empty(wav.shape).pin_memory()
            host.copy_(wav, non_blocking=True)
        e2.record()
        if mode != "async":
            torch.cuda.current_stream().synchronize()
        t3 = time.perf_counter()
        torch.cuda.synchronize()
        print(f"{mode} step {i}: gpu matcha {e0.elapsed_time(e1):7.2f} voc(+d2h) {e1.elapsed_time(e2):7.2f} ms | host synth {1e3 * (t1 - t0):7.2f} "
              f"voc {1e3 * (t2 - t1):7.2f} tail {1e3 * (t3 - t2):7.2f} ms")
