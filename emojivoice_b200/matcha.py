"""Drop-in for the reference's `MatchaTTS` inference surface (Matcha-TTS/matcha/models/matcha_tts.py:26-152).

Same constructor kwargs, `load_state_dict` with the reference key names, `eval()`, `synthesise(x, x_lengths,
n_timesteps, temperature, spks, length_scale)` with the same return dict.  All arithmetic runs in
libemojivoice_b200.so; this file only allocates tensors and sequences four C-ABI calls.
"""
from __future__ import annotations

import ctypes as C
import datetime as dt

import torch

from . import _lib
from .config import MatchaConfig


class MatchaTTS:
    def __init__(self, n_vocab, n_spks, spk_emb_dim, n_feats, encoder, decoder, cfm, data_statistics, out_size=None,
                 optimizer=None, scheduler=None, prior_loss=True, use_precomputed_durations=False, device=None,
                 precision="bf16", cuda_graphs=True):
        self.cfg = MatchaConfig.from_constructor_kwargs(n_vocab, n_spks, spk_emb_dim, n_feats, encoder, decoder, cfm,
                                                        data_statistics)
        self.n_vocab, self.n_spks, self.spk_emb_dim, self.n_feats = n_vocab, n_spks, spk_emb_dim, n_feats
        self.out_size, self.prior_loss, self.use_precomputed_durations = out_size, prior_loss, use_precomputed_durations
        self.hparams = dict(n_vocab=n_vocab, n_spks=n_spks, spk_emb_dim=spk_emb_dim, n_feats=n_feats, encoder=encoder,
                            decoder=decoder, cfm=cfm, data_statistics=data_statistics, out_size=out_size)
        self.mel_mean = torch.tensor(self.cfg.mel_mean)
        self.mel_std = torch.tensor(self.cfg.mel_std)
        self.precision = precision
        self._device = torch.device(device) if device is not None else None
        self._ctx = None
        self._loaded = False
        self.cuda_graphs = cuda_graphs            # replay the decoder as a CUDA graph once a (B, T_pad, n, ...) key repeats
        self._graphs = _lib.GraphCache()

    # -- nn.Module-ish surface the callers touch (feel_me.py:156-159, cli.py:110-118)
    def eval(self):
        return self

    def to(self, device):
        if self._loaded and torch.device(device) != self.device:
            raise RuntimeError("weights already live on " + str(self.device))
        self._device = torch.device(device)
        return self

    def cuda(self, device=None):
        return self.to(torch.device("cuda", torch.cuda.current_device() if device is None else device))

    @property
    def device(self):
        return self._ctx.device if self._ctx is not None else self._device

    @classmethod
    def load_from_checkpoint(cls, checkpoint_path, map_location=None, **kw):
        """Lightning checkpoint shim: needs `hyper_parameters` + `state_dict` (feel_me.py:156-159)."""
        ckpt = torch.load(checkpoint_path, map_location="cpu", weights_only=False)
        hp = dict(ckpt["hyper_parameters"])
        model = cls(**{k: hp.get(k) for k in ("n_vocab", "n_spks", "spk_emb_dim", "n_feats", "encoder", "decoder", "cfm",
                                               "data_statistics", "out_size")}, device=map_location, **kw)
        model.load_state_dict(ckpt["state_dict"])
        return model

    def load_state_dict(self, state_dict, strict=True):
        if self._loaded:
            raise RuntimeError("weights are packed once per model instance; create a new MatchaTTS to reload")
        self._ctx = _lib.Context(self._device)
        c = self.cfg
        if len(c.dec_channels) != 2 or c.dec_channels[0] != c.dec_channels[1] or c.dec_n_blocks != 1:
            raise ValueError("only the reference decoder layout channels=(c,c), n_blocks=1 is implemented")
        if c.dec_act != "snakebeta":
            raise ValueError("only act_fn='snakebeta' is implemented (configs/model/decoder/default.yaml:7)")
        cfg = _lib.EvMatchaCfg(c.n_vocab, c.n_spks, c.spk_emb_dim, c.n_feats, c.enc_channels, c.enc_filter_channels,
                               c.enc_filter_channels_dp, c.enc_heads, c.enc_layers, c.enc_kernel, int(c.enc_prenet),
                               c.dec_channels[0], c.dec_heads, c.dec_head_dim, c.dec_mid_blocks,
                               float(state_dict.get("mel_mean", c.mel_mean)), float(state_dict.get("mel_std", c.mel_std)))
        self.mel_mean = torch.tensor(cfg.mel_mean, device=self._ctx.device)
        self.mel_std = torch.tensor(cfg.mel_std, device=self._ctx.device)
        arr, keep = _lib.tensor_list(state_dict, self._ctx.device)
        with torch.cuda.device(self._ctx.device):
            rc = _lib.lib().ev_load_matcha(self._ctx.handle, arr, len(keep), C.byref(cfg), _lib.stream_ptr())
        self._ctx.check(rc, "ev_load_matcha")
        del keep
        self._loaded = True
        return self

    @staticmethod
    def fix_len_compatibility(length: int, num_downsamplings_in_unet: int = 2) -> int:
        """utils/model.py:14-20"""
        f = 2 ** num_downsamplings_in_unet
        return -(-int(length) // f) * f

    @torch.inference_mode()
    def synthesise(self, x, x_lengths, n_timesteps, temperature=1.0, spks=None, length_scale=1.0, z=None, dtype=None):
        """Same contract as matcha_tts.py:77-152.  Extras: `z` injects the prior noise (B, n_feats, T_pad) *before*
        temperature scaling (the reference draws it at flow_matching.py:51); `dtype` picks "fp32" | "bf16"."""
        if not self._loaded:
            raise RuntimeError("load_state_dict() / load_from_checkpoint() first")
        t0 = dt.datetime.now()
        ctx, L = self._ctx, _lib.lib()
        dev = ctx.device
        prec = _lib.PREC[dtype if dtype is not None else self.precision]
        with torch.cuda.device(dev):
            x = x.to(device=dev, dtype=torch.int64).contiguous()
            x_lengths = x_lengths.to(device=dev, dtype=torch.int64).contiguous()
            B, Tx = x.shape
            if self.n_spks > 1:
                if spks is None:
                    raise ValueError("multi-speaker model: `spks` (speaker / emoji ids) is required")
                spks = spks.to(device=dev).long().contiguous()          # matcha_tts.py:118 spks.long()
                if spks.numel() != B:
                    raise ValueError("spks must hold one id per utterance")
                spk_emb = torch.empty(B, self.spk_emb_dim, device=dev)
            else:
                spk_emb = None
            F = self.n_feats
            mu_x = torch.empty(B, F, Tx, device=dev)
            logw = torch.empty(B, 1, Tx, device=dev)
            w_ceil = torch.empty(B, 1, Tx, device=dev)
            y_lengths = torch.empty(B, dtype=torch.int64, device=dev)
            st = _lib.stream_ptr()
            nb = L.ev_encode_workspace_bytes(ctx.handle, B, Tx)
            ws = ctx.workspace(nb)
            ctx.check(L.ev_encode(ctx.handle, _lib.ptr(x), _lib.ptr(x_lengths), _lib.ptr(spks), B, Tx, float(length_scale),
                                  _lib.ptr(spk_emb), _lib.ptr(mu_x), _lib.ptr(logw), _lib.ptr(w_ceil), _lib.ptr(y_lengths),
                                  _lib.ptr(ws), ws.numel(), st), "ev_encode")
            y_max_length = int(y_lengths.max().item())                    # the reference's one host sync (utils/model.py:18)
            T_pad = self.fix_len_compatibility(y_max_length)
            attn = torch.empty(B, Tx, T_pad, device=dev)
            mu_y = torch.empty(B, F, T_pad, device=dev)
            y_mask = torch.empty(B, 1, T_pad, device=dev)
            ws = ctx.workspace(L.ev_align_workspace_bytes(ctx.handle, B, Tx, T_pad))
            ctx.check(L.ev_align(ctx.handle, _lib.ptr(w_ceil), _lib.ptr(x_lengths), _lib.ptr(y_lengths), _lib.ptr(mu_x), B, Tx,
                                 T_pad, _lib.ptr(attn), _lib.ptr(mu_y), _lib.ptr(y_mask), _lib.ptr(ws), ws.numel(), st), "ev_align")
            if z is None:
                z = torch.randn_like(mu_y)                               # flow_matching.py:51
            else:
                z = z.to(device=dev, dtype=torch.float32).contiguous()
                if tuple(z.shape) != (B, F, T_pad):
                    raise ValueError(f"z must have shape {(B, F, T_pad)}, got {tuple(z.shape)}")
            dec, mel = self._decode(mu_y, y_lengths, z, spk_emb, B, T_pad, int(n_timesteps), float(temperature), prec)
        t = (dt.datetime.now() - t0).total_seconds()
        rtf = t * 22050 / (max(y_max_length, 1) * 256)                    # matcha_tts.py:142-143 (host clock, no sync)
        return {
            "encoder_outputs": mu_y[:, :, :y_max_length],
            "decoder_outputs": dec[:, :, :y_max_length],
            # NB the reference slices the 4-D (B,1,Tx,T_pad) tensor with [:, :, :y_max_length], i.e. along the TOKEN
            # axis (matcha_tts.py:148) -- a no-op whenever y_max_length >= Tx.  Kept verbatim for drop-in parity.
            "attn": attn.unsqueeze(1)[:, :, :y_max_length],
            "mel": mel[:, :, :y_max_length],
            "mel_lengths": y_lengths,
            "rtf": rtf,
            # extras (not in the reference dict): intermediates the parity tests compare
            "logw": logw, "w_ceil": w_ceil, "mu_x": mu_x, "y_mask": y_mask, "t_pad": T_pad, "z": z,
            "decoder_outputs_full": dec, "mel_full": mel,
        }

    def _decode(self, mu_y, y_lengths, z, spk_emb, B, T_pad, n_timesteps, temperature, prec):
        """ev_decode, eagerly or -- once the same shape key has been seen before -- as a CUDA-graph replay over static
        buffers (inputs are copied in, outputs copied out, so callers still own fresh tensors as with the reference)."""
        ctx, L, dev, F = self._ctx, _lib.lib(), self._ctx.device, self.n_feats

        def call(mu_y, y_lengths, z, spk_emb, dec, mel, ws):
            ctx.check(L.ev_decode(ctx.handle, _lib.ptr(mu_y), _lib.ptr(y_lengths), _lib.ptr(z), _lib.ptr(spk_emb), B, T_pad,
                                  n_timesteps, temperature, prec, _lib.ptr(dec), _lib.ptr(mel), _lib.ptr(ws), ws.numel(),
                                  _lib.stream_ptr()), "ev_decode")

        nb = L.ev_decode_workspace_bytes(ctx.handle, B, T_pad, n_timesteps)
        key = (B, T_pad, n_timesteps, temperature, prec)
        ws_shared = ctx.workspace(nb)                       # (may grow the workspace: do it before looking graphs up)
        ent = self._graphs.get(key, ctx.ws_version) if self.cuda_graphs else None
        if ent is None and self.cuda_graphs and self._graphs.should_capture(key):
            st = dict(mu_y=torch.empty_like(mu_y), y_lengths=torch.empty_like(y_lengths), z=torch.empty_like(z),
                      spk_emb=None if spk_emb is None else torch.empty_like(spk_emb), dec=torch.empty(B, F, T_pad, device=dev),
                      mel=torch.empty(B, F, T_pad, device=dev), ws=ws_shared, ws_version=ctx.ws_version)
            st["mu_y"].copy_(mu_y); st["y_lengths"].copy_(y_lengths); st["z"].copy_(z)
            if spk_emb is not None:
                st["spk_emb"].copy_(spk_emb)
            st["graph"], st["launches"] = _lib.capture(
                ctx, lambda: call(st["mu_y"], st["y_lengths"], st["z"], st["spk_emb"], st["dec"], st["mel"], st["ws"]))
            self._graphs.put(key, st)
            ent = st
        if ent is None:
            dec = torch.empty(B, F, T_pad, device=dev)
            mel = torch.empty(B, F, T_pad, device=dev)
            call(mu_y, y_lengths, z, spk_emb, dec, mel, ws_shared)
            return dec, mel
        ent["mu_y"].copy_(mu_y); ent["y_lengths"].copy_(y_lengths); ent["z"].copy_(z)
        if spk_emb is not None:
            ent["spk_emb"].copy_(spk_emb)
        ent["graph"].replay()
        self._replayed_launches += ent["launches"]
        return ent["dec"].clone(), ent["mel"].clone()

    _replayed_launches = 0

    def launch_count(self, reset=False):
        """Kernels launched by this model's context, graph replays included (a replay launches every captured kernel)."""
        n = self._ctx.launch_count(reset) + self._replayed_launches
        if reset:
            self._replayed_launches = 0
        return n
