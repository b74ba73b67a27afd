"""Drop-in for the reference vocoder objects: `Generator(h)` (hifigan/models.py:148-206) and `Denoiser`
(hifigan/denoiser.py), plus `to_waveform` (feel_me.py:181-187 / cli.py:121-126)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .config import HIFIGAN_V1


class Generator:
    def __init__(self, h=None, device=None, precision="bf16", cuda_graphs=True):
        self.h = dict(h) if h is not None else dict(HIFIGAN_V1)
        if str(self.h.get("resblock", "1")) != "1":
            raise ValueError("only ResBlock1 (config v1) is implemented")
        self.num_kernels = len(self.h["resblock_kernel_sizes"])
        self.num_upsamples = len(self.h["upsample_rates"])
        self.precision = precision
        self._device = torch.device(device) if device is not None else None
        self._sd = None
        self._ctx = None
        self.cuda_graphs = cuda_graphs            # replay the generator as a CUDA graph once a (B, T) shape repeats
        self._graphs = _lib.GraphCache()
        self._replayed_launches = 0
        self.hop = 1
        for u in self.h["upsample_rates"]:
            self.hop *= int(u)

    def to(self, device):
        self._device = torch.device(device)
        return self

    def cuda(self, device=None):
        return self.to(torch.device("cuda", torch.cuda.current_device() if device is None else device))

    def eval(self):
        return self

    def parameters(self):
        return iter((self._sd or {}).values())

    @property
    def device(self):
        return self._ctx.device if self._ctx is not None else self._device

    def load_state_dict(self, state_dict, strict=True):
        """Accepts plain `weight` keys or the checkpoint's weight-norm form (`weight_g`/`weight_v`,
        feel_me.py:161-167); the library folds the latter as `remove_weight_norm()` would."""
        self._sd = {k: v for k, v in state_dict.items()}
        return self

    def remove_weight_norm(self):
        self._materialise()

    def replica(self):
        """Another generator on the same device with its own context (weights, workspace, graph cache): see MatchaTTS.replica."""
        if self._sd is None:
            raise RuntimeError("load_state_dict() first")
        g = Generator(self.h, device=self.device, precision=self.precision, cuda_graphs=self.cuda_graphs)
        g.load_state_dict(self._sd)
        g._materialise()
        return g

    def _materialise(self):
        if self._ctx is not None:
            return
        if self._sd is None:
            raise RuntimeError("load_state_dict() first")
        self._ctx = _lib.Context(self._device)
        h = self.h
        cfg = _lib.EvHifiganCfg()
        cfg.num_mels = int(h.get("num_mels", 80))
        cfg.upsample_initial_channel = int(h["upsample_initial_channel"])
        cfg.n_ups, cfg.n_kernels = self.num_upsamples, self.num_kernels
        for i, (u, k) in enumerate(zip(h["upsample_rates"], h["upsample_kernel_sizes"])):
            cfg.upsample_rates[i], cfg.upsample_kernel_sizes[i] = int(u), int(k)
        for j, (k, dil) in enumerate(zip(h["resblock_kernel_sizes"], h["resblock_dilation_sizes"])):
            cfg.resblock_kernel_sizes[j] = int(k)
            if len(dil) != 3:
                raise ValueError("ResBlock1 needs three dilations per kernel size")
            for l in range(3):
                cfg.resblock_dilation_sizes[j][l] = int(dil[l])
        arr, keep = _lib.tensor_list(self._sd, self._ctx.device)
        with torch.cuda.device(self._ctx.device):
            rc = _lib.lib().ev_load_hifigan(self._ctx.handle, arr, len(keep), C.byref(cfg), _lib.stream_ptr())
        self._ctx.check(rc, "ev_load_hifigan")

    @torch.inference_mode()
    def __call__(self, mel, dtype=None, lengths=None):
        """mel (B, num_mels, T) -> wav (B, 1, T*hop); tanh output already clamped to [-1, 1].

        lengths (B,) int, optional: valid mel frames per item (the `mel_lengths` of synthesise) of a padded batch.  The
        waveform of item b is then bit-identical on [: lengths[b]*hop] -- all the reference's batched caller keeps
        (cli.py:307-311) -- and zero beyond; time tiles further than the generator's receptive field past an utterance's
        end are never computed (ev_vocode_ragged)."""
        self._materialise()
        ctx, L = self._ctx, _lib.lib()
        dev = ctx.device
        if mel.dim() == 2:
            mel = mel.unsqueeze(0)
        with torch.cuda.device(dev):
            mel = mel.to(device=dev, dtype=torch.float32).contiguous()
            B, _, T = mel.shape
            prec = _lib.PREC[dtype if dtype is not None else self.precision]
            nb = L.ev_vocode_workspace_bytes(ctx.handle, B, T)
            if lengths is not None:
                lengths = torch.as_tensor(lengths).reshape(-1).to(device=dev, dtype=torch.int64).contiguous()
                if lengths.numel() != B:
                    raise ValueError(f"lengths has {lengths.numel()} entries for a batch of {B}")

            def call(mel_, wav_, ws_, len_=None):
                ctx.check(L.ev_vocode_ragged(ctx.handle, _lib.ptr(mel_), _lib.ptr(len_) if len_ is not None else None, B, T, prec,
                                             _lib.ptr(wav_), _lib.ptr(ws_), ws_.numel(), _lib.stream_ptr()), "ev_vocode")

            key = (B, T, prec, lengths is not None)
            ws_shared = ctx.workspace(nb)                   # (may grow the workspace: do it before looking graphs up)
            ent = self._graphs.get(key, ctx.ws_version) if self.cuda_graphs else None
            if ent is None and self.cuda_graphs and self._graphs.should_capture(key):
                ent = dict(mel=mel.clone(), wav=torch.empty(B, 1, T * self.hop, device=dev), ws=ws_shared, ws_version=ctx.ws_version,
                           len=lengths.clone() if lengths is not None else None)
                ent["graph"], ent["launches"] = _lib.capture(ctx, lambda: call(ent["mel"], ent["wav"], ent["ws"], ent["len"]))
                self._graphs.put(key, ent)
            if ent is None:
                wav = torch.empty(B, 1, T * self.hop, device=dev)
                call(mel, wav, ws_shared, lengths)
            else:
                ent["mel"].copy_(mel)
                if lengths is not None:
                    ent["len"].copy_(lengths)
                ent["graph"].replay()
                self._replayed_launches += ent["launches"]
                wav = ent["wav"].clone()
        return wav

    forward = __call__

    def launch_count(self, reset=False):
        """Kernels launched by this generator's context, graph replays included."""
        n = self._ctx.launch_count(reset) + self._replayed_launches
        if reset:
            self._replayed_launches = 0
        return n


class Denoiser:
    """hifigan/denoiser.py:7-64 (mode="zeros").  bias_spec is computed on the GPU with the given generator."""

    def __init__(self, vocoder: Generator, filter_length=1024, n_overlap=4, win_length=1024, mode="zeros"):
        if mode != "zeros":
            raise Exception(f"Mode {mode} if not supported")
        if (filter_length, n_overlap, win_length) != (1024, 4, 1024):
            raise ValueError("the denoiser kernel is built for filter_length=1024, n_overlap=4, win_length=1024")
        vocoder._materialise()
        self.vocoder = vocoder
        self.device = vocoder.device
        ctx, L = vocoder._ctx, _lib.lib()
        self.bias_spec = torch.empty(1, 513, 1, device=self.device)
        with torch.cuda.device(self.device):
            ws = ctx.workspace(L.ev_denoise_workspace_bytes(ctx.handle, 1, 88 * vocoder.hop) +
                               L.ev_vocode_workspace_bytes(ctx.handle, 1, 88) + (1 << 20))
            ctx.check(L.ev_denoiser_init(ctx.handle, _lib.ptr(self.bias_spec), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()),
                      "ev_denoiser_init")

    @torch.inference_mode()
    def __call__(self, audio, strength=0.0005):
        ctx, L = self.vocoder._ctx, _lib.lib()
        squeeze = audio.dim() == 1
        a = audio.reshape(-1, audio.shape[-1]).to(device=self.device, dtype=torch.float32).contiguous()
        B, n = a.shape
        out = torch.empty(B, (n // 256) * 256, device=self.device)
        with torch.cuda.device(self.device):
            ws = ctx.workspace(L.ev_denoise_workspace_bytes(ctx.handle, B, n))
            ctx.check(L.ev_denoise(ctx.handle, _lib.ptr(a), B, n, float(strength), _lib.ptr(out), _lib.ptr(ws), ws.numel(),
                                   _lib.stream_ptr()), "ev_denoise")
        return out[0] if squeeze else out

    forward = __call__


@torch.inference_mode()
def to_waveform(mel, vocoder, denoiser=None, strength=0.00025, lengths=None):
    """feel_me.py:181-187: vocoder(mel).clamp(-1, 1) -> optional denoiser -> .cpu().squeeze().

    lengths: optional valid frames per item of a padded batch (see Generator.__call__)."""
    audio = (vocoder(mel, lengths=lengths) if lengths is not None else vocoder(mel)).clamp(-1, 1)
    if denoiser is not None:
        audio = denoiser(audio.squeeze(1), strength=strength)
    return audio.cpu().squeeze()
