#!/usr/bin/env python
"""Headline benchmark: audio-seconds synthesised per second (22.05 kHz, mel + vocoder) on N x B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one pass of the hot path over one batch: MatchaTTS.synthesise(x, x_lengths, n_timesteps=10,
temperature=0.667, spks=emoji ids, length_scale=0.8) followed by vocoder(mel).clamp(-1, 1), on BASELINE.json
configs[1] (batch 32 emoji-tagged utterances of ~5 s, bf16, random-init VCTK Matcha-TTS + HiFi-GAN v1).
For N > 1 the driver launches one rank per GPU with torch.distributed.run; utterances shard by batch, every
rank synthesises its own 32 utterances (weak scaling), no collective on the data path.

One JSON line is printed by rank 0 (see the task contract): `value` is measured with the inputs resident in HBM,
`e2e` through the same public API with pinned HOST inputs and the waveform read back, `roofline` for the dominant
kernel from a CUDA-event-instrumented step, `cpu_baseline` = the CPU oracle (a port of the reference's PyTorch
path) timed on this box's host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SR, HOP = 22050, 256
N_TIMESTEPS, TEMPERATURE, LENGTH_SCALE = 10, 0.667, 0.8      # feel_me.py:71-77 (the operating point of every app)
BATCH, P_LO, P_HI = 32, 60, 90                                # Tx = 2P+1 in [121, 181]
CPU_SAMPLE = 8                                                # utterances of the batch the CPU baseline synthesises (~3 s of CPU work per pass)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm_gbs=float(p["hbm_gbs"]), tflops=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                    source="MEASURED_PEAKS.json (sustained bf16, copy bandwidth)")
    return dict(hbm_gbs=6650.0, tflops=1400.0, source="fallback of B200_PROFILING.md")


class ClockSampler:
    """SM clock + throttle reasons sampled every 100 ms while the timed region runs, through NVML in-process
    (spawning `nvidia-smi -lms` was measurably perturbing the launch path on the GPU box)."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index: int):
        self.gpu, self.sm, self.bits, self.stop_flag, self.thread, self.h, self.mx = gpu_index, [], 0, False, None, None, None
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.h = None

    def _loop(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    self.bits |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    self.bits |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:
                pass
            time.sleep(0.1)

    def start(self):
        if self.h is not None:
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()

    def stop(self):
        if self.h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        self.stop_flag = True
        self.thread.join(timeout=1.0)
        reasons = sorted(n for b, n in self.REASONS.items() if self.bits & b)
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.mx, "reasons": reasons,
                "samples": len(self.sm)}


def measured_traffic(kernel_class: str):
    """DRAM bytes per launch of the dominant kernel from the committed ncu launch list of this same command
    (profiles/r01_launch_summary.json, written by scripts/ncu_launch_summary.py); None when no capture is committed."""
    import re

    path = os.path.join(ROOT, "profiles", "r01_launch_summary.json")
    if not os.path.exists(path):
        return None, None
    k = json.load(open(path))["kernels"]
    m = re.match(r"conv_tc_(bn|tf32x)(\d+)", kernel_class)
    fn = "conv_tc_kernel<%s>" % (m.group(2) if m.group(1) == "bn" else "128") if m else None
    if kernel_class.startswith("resblock_tc"):
        hits = [v for n, v in k.items() if n.startswith("resblock_tc_kernel")]
        if hits:
            n = sum(h["launches"] for h in hits)
            return sum(h["dram_bytes_per_launch"] * h["launches"] for h in hits) / n, "resblock_tc_kernel<*>"
    if fn in k:
        return k[fn]["dram_bytes_per_launch"], fn
    return None, None


def audio_seconds(mel_lengths) -> float:
    return float(mel_lengths.sum()) * HOP / SR


def cpu_oracle_rate(x, xl, spks, n_utts, repeats=2):
    """The reference's CPU PyTorch path (restated in oracle/) on the first n_utts utterances, all host threads."""
    from emojivoice_b200 import synthetic
    from emojivoice_b200.config import HIFIGAN_V1, VCTK
    from oracle import hifigan_oracle as ho
    from oracle import matcha_oracle as mo

    torch.set_num_threads(os.cpu_count() or 1)
    sd = synthetic.matcha_state_dict(VCTK, seed=1234)
    hsd = synthetic.hifigan_state_dict(HIFIGAN_V1, seed=4321)
    n = min(n_utts, x.shape[0])
    xs, ls, ss = x[:n, : int(xl[:n].max())].contiguous(), xl[:n], spks[:n]

    def once():
        t0 = time.perf_counter()
        out = mo.synthesise(sd, VCTK, xs, ls, N_TIMESTEPS, TEMPERATURE, ss, LENGTH_SCALE)
        wav = ho.generator(hsd, HIFIGAN_V1, out["mel"]).clamp(-1, 1)
        dt = time.perf_counter() - t0
        return audio_seconds(out["mel_lengths"]), dt, wav

    return once, n


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path (oracle port; the python reference cannot
    travel to the GPU box), all host threads, each step a bounded sample of the same workload."""
    if rank != 0:
        return
    from emojivoice_b200 import synthetic

    x, xl, spks = synthetic.phoneme_batch(BATCH, P_LO, P_HI, seed=2000)
    once, n = cpu_oracle_rate(x, xl, spks, CPU_SAMPLE)
    for _ in range(args.warmup):
        once()
    t0 = time.perf_counter()
    secs = 0.0
    for _ in range(args.steps):
        a, _, _ = once()
        secs += a
    dt = time.perf_counter() - t0
    val = secs / dt
    sample = f"first {n} utterances of the batch per step ({secs / args.steps:.1f} audio-s), fp32, torch CPU"
    print(json.dumps({
        "impl": "reference", "metric": "audio_seconds_per_second", "value": val, "unit": "audio-s/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(),
        "cpu_baseline": {"value": val, "unit": "audio-s/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "rtf": 1.0 / val}))


def workload_config():
    return {"workload": "BASELINE.json configs[1]: batch 32 emoji-tagged utterances (~5 s each), n_timesteps=10, "
                        "Matcha-TTS VCTK arch + HiFi-GAN v1, random init, synthetic blank-interspersed phoneme ids",
            "batch_per_gpu": BATCH, "n_timesteps": N_TIMESTEPS, "temperature": TEMPERATURE, "length_scale": LENGTH_SCALE,
            "tokens_per_utt": f"2P+1, P~U[{P_LO},{P_HI}]", "speakers": "11 emoji voices (feel_me.py:84-96)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dense-vocoder", action="store_true",
                    help="vocode the padded frames of every utterance too (the default skips the time tiles past each utterance's "
                         "own length: identical waveform on [: length*256], zero beyond -- what cli.py:307-311 crops away)")
    ap.add_argument("--profile-one-step", action="store_true",
                    help="after warm-up, bracket ONE step with cudaProfilerStart/Stop and exit (for `ncu --profile-from-start off`)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.warmup < 3:
        args.warmup = 3

    import torch.distributed as dist

    import emojivoice_b200 as ev
    from emojivoice_b200 import _lib as _lib_mod
    from emojivoice_b200 import synthetic
    from emojivoice_b200.config import HIFIGAN_V1, VCTK

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the CUDA path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    model = ev.MatchaTTS(**VCTK.constructor_kwargs(), device=dev, precision=args.precision).eval()
    model.load_state_dict(synthetic.matcha_state_dict(VCTK, seed=1234))
    voc = ev.Generator(HIFIGAN_V1, device=dev, precision=args.precision)
    voc.load_state_dict(synthetic.hifigan_state_dict(HIFIGAN_V1, seed=4321))
    voc.eval()
    voc.remove_weight_norm()

    # every rank synthesises its own batch (weak scaling); same distribution, different seed
    x, xl, spks = synthetic.phoneme_batch(BATCH, P_LO, P_HI, seed=2000 + rank)
    x_pin, xl_pin, spk_pin = x.pin_memory(), xl.pin_memory(), spks.pin_memory()
    x_dev, xl_dev, spk_dev = x.to(dev), xl.to(dev), spks.to(dev)

    ragged = [not args.dense_vocoder]

    def vocode(out):
        # to_waveform (feel_me.py:183).  Ragged: item b is vocoded up to mel_lengths[b] (+ receptive field) only
        return voc(out["mel"], lengths=out["mel_lengths"] if ragged[0] else None).clamp(-1, 1)

    def step_resident():
        out = model.synthesise(x_dev, xl_dev, N_TIMESTEPS, TEMPERATURE, spk_dev, LENGTH_SCALE)
        return out, vocode(out)

    wav_host = [None]
    len_host = torch.empty(BATCH, dtype=torch.int64).pin_memory()

    def step_e2e():
        xd = x_pin.to(dev, non_blocking=True)
        ld = xl_pin.to(dev, non_blocking=True)
        sd_ = spk_pin.to(dev, non_blocking=True)
        out = model.synthesise(xd, ld, N_TIMESTEPS, TEMPERATURE, sd_, LENGTH_SCALE)
        wav = vocode(out)
        if wav_host[0] is None or wav_host[0].shape != wav.shape:
            wav_host[0] = torch.empty(wav.shape, dtype=wav.dtype).pin_memory()
        wav_host[0].copy_(wav, non_blocking=True)               # .cpu() of to_waveform (feel_me.py:187)
        len_host.copy_(out["mel_lengths"], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return out, wav

    # warm-up: W steps at least (graphs are captured on the second sight of a shape), then keep stepping until the
    # per-step wall time has settled (a freshly booted box stalls the host for 100s of ms now and then) or 15 s passed
    t_w0, recent = time.perf_counter(), []
    n_warm = 0
    while True:
        t_s = time.perf_counter()
        out, wav = step_resident()
        torch.cuda.synchronize()
        recent.append(time.perf_counter() - t_s)
        n_warm += 1
        if n_warm >= max(args.warmup, 4):
            last = recent[-6:]
            if (len(last) >= 6 and max(last) < 1.15 * min(last)) or time.perf_counter() - t_w0 > 15.0:
                break
    secs_per_step = audio_seconds(out["mel_lengths"].cpu())
    frames = int(out["mel_lengths"].sum())
    t_pad = int(out["t_pad"])
    ws_bytes = (_lib_mod.lib().ev_vocode_workspace_bytes(voc._ctx.handle, BATCH, int(out["mel"].shape[2])) +
                _lib_mod.lib().ev_decode_workspace_bytes(model._ctx.handle, BATCH, t_pad, N_TIMESTEPS))

    if args.profile_one_step:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        step_resident()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        print(json.dumps({"profiled_step": True, "t_pad": t_pad, "frames": frames}))
        return

    # ---------------- timed region: K steps, inputs resident in HBM
    sampler = ClockSampler(local_rank)
    sampler.start()
    model.launch_count(reset=True); voc.launch_count(reset=True)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_resident()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = model.launch_count() + voc.launch_count()
    clocks = sampler.stop()

    # ---------------- the same K steps with the padded frames vocoded too (reference-identical waveform everywhere)
    ms_dense = None
    if ragged[0]:
        ragged[0] = False
        for _ in range(3):
            step_resident()
        barrier()
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d0.record()
        for _ in range(args.steps):
            step_resident()
        d1.record()
        barrier()
        ms_dense = d0.elapsed_time(d1)
        ragged[0] = True

    # ---------------- e2e: same steps through the public API with pinned host inputs and the waveform read back
    recent = []
    while True:                                                  # settle like the warm-up above
        t_s = time.perf_counter()
        step_e2e()
        recent.append(time.perf_counter() - t_s)
        last = recent[-4:]
        if (len(recent) >= 4 and max(last) < 1.15 * min(last)) or len(recent) >= 40:
            break
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        step_e2e()
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)

    tt = torch.tensor([ms, ms_e2e, -secs_per_step, ms_dense or 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)                # slowest rank bounds the job
        tot = torch.tensor([secs_per_step, float(launches)], dtype=torch.float64, device=dev)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        total_secs, launches = float(tot[0]), int(tot[1])
    else:
        total_secs = secs_per_step
    ms, ms_e2e = float(tt[0]), float(tt[1])
    ms_dense = float(tt[3]) if ms_dense is not None else None
    value = total_secs * args.steps / (ms / 1e3)
    value_e2e = total_secs * args.steps / (ms_e2e / 1e3)
    h2d = x.numel() * 8 + xl.numel() * 8 + spks.numel() * 8
    d2h = int(wav.numel()) * 4 + BATCH * 8

    # ---------------- roofline of the dominant kernel: one more step with CUDA events around every launch
    peaks = load_peaks()
    model.cuda_graphs = voc.cuda_graphs = False                  # per-launch events need eager launches, not a graph replay
    model._ctx.profile_begin(); voc._ctx.profile_begin()
    step_resident()
    stats = model._ctx.profile_end() + voc._ctx.profile_end()
    model.cuda_graphs = voc.cuda_graphs = True
    agg = {}
    for s in stats:
        a = agg.setdefault(s["name"], dict(name=s["name"], launches=0, total_ms=0.0, flops=0.0, bytes=0.0))
        for k in ("launches", "total_ms", "flops", "bytes"):
            a[k] += s[k]
    klist = sorted(agg.values(), key=lambda a: -a["total_ms"])
    ksum = sum(a["total_ms"] for a in klist) or 1.0
    table = []
    for a in klist[:10]:
        sec = a["total_ms"] / 1e3
        table.append({"name": a["name"], "launches": a["launches"], "ms": round(a["total_ms"], 3),
                      "share": round(a["total_ms"] / ksum, 4), "tflops": round(a["flops"] / sec / 1e12, 2) if sec else 0,
                      "gbs": round(a["bytes"] / sec / 1e9, 1) if sec else 0})
    # dominant kernel = the __global__ function with the largest share of the step (classes "fn/stage" share a function)
    byfn = {}
    for a in klist:
        f = byfn.setdefault(a["name"].split("/")[0], dict(name=a["name"].split("/")[0], launches=0, total_ms=0.0, flops=0.0, bytes=0.0))
        for k in ("launches", "total_ms", "flops", "bytes"):
            f[k] += a[k]
    top = max(byfn.values(), key=lambda a: a["total_ms"])
    sec = top["total_ms"] / 1e3
    ridge = peaks["tflops"] * 1e12 / (peaks["hbm_gbs"] * 1e9)
    ai = top["flops"] / max(top["bytes"], 1.0)
    if top["flops"] > 0 and ai >= ridge:
        ach, peak, unit, bound = top["flops"] / sec / 1e12, peaks["tflops"], "TFLOP/s", "tensor"
    else:
        ach, peak, unit, bound = top["bytes"] / sec / 1e9, peaks["hbm_gbs"], "GB/s", "hbm"
    roofline = {"kernel": top["name"], "bound": bound, "achieved": round(ach, 2), "peak": peak, "unit": unit,
                "frac": round(ach / peak, 4), "traffic": None, "launches": top["launches"],
                "algorithmic_bytes_per_launch": round(top["bytes"] / max(top["launches"], 1)),
                "algorithmic_flops_per_launch": round(top["flops"] / max(top["launches"], 1)),
                "avg_launch_us": round(top["total_ms"] * 1e3 / max(top["launches"], 1), 2),
                "share_of_step": round(top["total_ms"] / ksum, 4), "arith_intensity": round(ai, 1),
                "peak_source": peaks["source"], "how": "CUDA events around every launch on the launching stream, one "
                "instrumented step right after the timed region; algorithmic FLOPs/bytes (valid un-padded work)"}
    tr, tr_fn = measured_traffic(top["name"])
    if tr is not None:
        roofline["traffic"] = round(tr)
        roofline["traffic_source"] = ("dram__bytes_read.sum + dram__bytes_write.sum per launch of %s, averaged over the launches of one "
                                      "bench step (profiles/r01_launch_summary.json, ncu --clock-control none)" % tr_fn)
    # whole-path algorithmic FLOPs as a fraction of the tensor roofline
    tx = xl.double()
    flops_step = float((tx * (19309056 + 6144 * tx)).sum()) + 0.0
    ml = out["mel_lengths"].double().cpu()
    flops_step += float((N_TIMESTEPS * ml * (11116544 + 1536 * ml)).sum()) + float((ml * 614105088).sum())
    path_tflops = flops_step * args.steps / (ms / 1e3) / 1e12 * (world if world > 1 else 1)

    result = {
        "metric": "audio_seconds_per_second", "value": round(value, 2), "unit": "audio-s/s", "n_gpus": world,
        "steps": args.steps, "warmup": n_warm, "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.precision if args.precision != "fp32" else "f32",
        "data": "synthetic",
        "config": dict(workload_config(), audio_seconds_per_step_per_gpu=round(secs_per_step, 2), mel_frames_per_step_per_gpu=frames,
                       t_pad=t_pad, l2="no flush: each step streams a %.1f GB activation workspace (>> 126 MB L2); weights "
                       "(35 MB bf16) stay L2-resident as they would in service" % (ws_bytes / 1e9),
                       launch="encoder, alignment + decoder and vocoder replayed as CUDA graphs (captured during warm-up)",
                       vocoder=("ragged: time tiles past each utterance's own length (+ the receptive field behind each layer: 13 frames at conv_pre ... 2 at the last stage) are "
                                "not computed; waveform bit-identical to the dense generator on [: mel_length*256] and zero beyond "
                                "-- the part the reference's batched caller crops away (cli.py:307-311); `padded_vocoder` is the "
                                "same run with the padded frames vocoded too") if not args.dense_vocoder else "dense (padded frames vocoded too)"),
        "rtf": round(1.0 / value, 8), "x_realtime": round(value, 1),
        "clocks": clocks,
        "e2e": {"value": round(value_e2e, 2), "unit": "audio-s/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": round(ms_e2e / args.steps, 3)},
        "gpu_launches": int(launches),
        "padded_vocoder": ({"value": round(total_secs * args.steps / (ms_dense / 1e3), 2), "unit": "audio-s/s",
                            "ms_per_step": round(ms_dense / args.steps, 3)} if ms_dense else None),
        "roofline": roofline,
        "path_tflops": round(path_tflops, 2), "path_frac_of_tensor_peak": round(path_tflops / (peaks["tflops"] * world), 4),
        "kernels": table,
    }
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        once, n = cpu_oracle_rate(x, xl, spks, CPU_SAMPLE)
        once()
        best = None
        for _ in range(3):
            a, dt, _ = once()
            best = (a, dt) if best is None or dt < best[1] else best
        result["cpu_baseline"] = {"value": round(best[0] / best[1], 3), "unit": "audio-s/s", "cores": torch.get_num_threads(),
                                  "kind": "port", "sample": f"first {n} utterances of the batch ({best[0]:.1f} audio-s), fp32 "
                                  f"torch CPU oracle, warm-up 1, best of 3, {os.cpu_count()} host cpus"}
    elif rank == 0:
        result["cpu_baseline"] = None
    if rank == 0:
        print(json.dumps(result))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
