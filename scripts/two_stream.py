#!/usr/bin/env python
"""Throughput of the bench workload with more than one batch in flight (one host thread).  Diagnostic.
  serial      one stream, one batch at a time (the round-1 / early round-2 bench loop)
  pairs N     N (model, vocoder) instances on N streams, batch i on stream i % N
  split       ONE model + vocoder: Matcha (encoder, alignment, decoder) on a high-priority stream, the vocoder of the previous
              batch on a second stream -- the latency-bound decoder runs in the shadow of the tensor-bound vocoder"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import emojivoice_b200 as ev  # noqa: E402
from emojivoice_b200 import synthetic  # noqa: E402
from emojivoice_b200.config import HIFIGAN_V1, VCTK  # noqa: E402

msd = synthetic.matcha_state_dict(VCTK, seed=1234)
hsd = synthetic.hifigan_state_dict(HIFIGAN_V1, seed=4321)
pairs = []
for _ in range(3):
    m = ev.MatchaTTS(**VCTK.constructor_kwargs(), precision="bf16")
    m.load_state_dict(msd)
    v = ev.Generator(HIFIGAN_V1, precision="bf16")
    v.load_state_dict(hsd)
    v.remove_weight_norm()
    pairs.append((m, v))
x, xl, spk = synthetic.phoneme_batch(32, 60, 90, seed=2000)
x, xl, spk = x.cuda(), xl.cuda(), spk.cuda()
streams = [torch.cuda.Stream() for _ in range(3)]
K = int(os.environ.get("EV_STEPS", "40"))


def run_pairs(n, steps):
    out = None
    for i in range(steps):
        m, v = pairs[i % n]
        with torch.cuda.stream(streams[i % n]):
            out = m.synthesise(x, xl, 10, 0.667, spk, 0.8)
            v(out["mel"], lengths=out["mel_lengths"]).clamp(-1, 1)
    return out


def run_split(steps, prio, max_ahead=2):
    m, v = pairs[0]
    sm = torch.cuda.Stream(priority=prio)
    sv = torch.cuda.Stream()
    done = []
    out = None
    for i in range(steps):
        with torch.cuda.stream(sm):
            out = m.synthesise(x, xl, 10, 0.667, spk, 0.8)
            ready = torch.cuda.Event()
            ready.record()
        with torch.cuda.stream(sv):
            sv.wait_event(ready)
            out["mel"].record_stream(sv)
            out["mel_lengths"].record_stream(sv)
            v(out["mel"], lengths=out["mel_lengths"]).clamp(-1, 1)
            e = torch.cuda.Event()
            e.record()
        done.append(e)
        if len(done) > max_ahead:
            done.pop(0).synchronize()
    return out


modes = [("serial", lambda s: run_pairs(1, s)), ("pairs 2", lambda s: run_pairs(2, s)), ("pairs 3", lambda s: run_pairs(3, s)),
         ("split prio 0", lambda s: run_split(s, 0)), ("split prio -1", lambda s: run_split(s, -1)),
         ("split prio -1 ahead 3", lambda s: run_split(s, -1, 3)), ("serial", lambda s: run_pairs(1, s))]
for name, fn in modes:
    out = fn(8)
    torch.cuda.synchronize()
    audio = float(out["mel_lengths"].sum()) * 256 / 22050
    t0 = time.perf_counter()
    fn(K)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{name:24s}: {1e3 * dt / K:7.3f} ms per step (wall), {audio * K / dt:8.1f} audio-s/s")
