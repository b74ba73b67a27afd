"""Seeded synthetic weights and inputs (checkpoints are unavailable offline).

State-dict key names and shapes are the reference's (SURVEY.md §8a "weights contract"); the values come from
numpy Philox streams so that the container, the GPU box and every rank draw identical tensors regardless of
torch version or thread count.  Initial distributions follow the reference constructors
(text_encoder.py:57-58,216-221,344; decoder.py:345-361; hifigan/xutils.py:25-28) and, because those leave
several paths numerically dead (zero prenet.proj, zero biases, unit norm affines, SnakeBeta alpha=beta=0),
every bias / norm affine / alpha,beta / prenet.proj tensor gets an extra N(0, 0.1^2) perturbation
(SURVEY.md §8d "coverage caveat").
"""
from __future__ import annotations

import math
import zlib

import numpy as np
import torch

from .config import HIFIGAN_V1, VCTK, MatchaConfig, EMOJI_MAPPING_FEMALE


class _Stream:
    def __init__(self, seed: int):
        self.seed = int(seed)

    def _rng(self, name: str):
        # one independent Philox stream per tensor name: insertion order cannot change values
        return np.random.Generator(np.random.Philox(key=[self.seed, zlib.crc32(name.encode())]))

    def normal(self, name, shape, std=1.0, mean=0.0):
        a = self._rng(name).standard_normal(size=tuple(shape), dtype=np.float32)
        return torch.from_numpy(a * np.float32(std) + np.float32(mean))

    def uniform(self, name, shape, bound):
        a = self._rng(name).random(size=tuple(shape), dtype=np.float32)
        return torch.from_numpy((a * 2.0 - 1.0).astype(np.float32) * np.float32(bound))


def matcha_state_dict(cfg: MatchaConfig = VCTK, seed: int = 1234, dur_scale: float = 2.4,
                      perturb: float = 0.1) -> dict:
    """Random-init Matcha-TTS weights under the reference `state_dict()` names."""
    s = _Stream(seed)
    sd: dict[str, torch.Tensor] = {}
    pt = perturb

    def conv(name, cout, cin, k, bias=True, kind="default", fan_in=None):
        fi = fan_in if fan_in is not None else cin * k
        if kind == "default":      # nn.Conv1d default: kaiming_uniform(a=sqrt(5)) == U(+-1/sqrt(fan_in))
            sd[name + ".weight"] = s.uniform(name + ".weight", (cout, cin, k), 1.0 / math.sqrt(fi))
        elif kind == "xavier":     # text_encoder.py:216-221
            bound = math.sqrt(6.0 / (cin * k + cout * k))
            sd[name + ".weight"] = s.uniform(name + ".weight", (cout, cin, k), bound)
        elif kind == "kaiming":    # decoder.py:345-361
            sd[name + ".weight"] = s.normal(name + ".weight", (cout, cin, k), math.sqrt(2.0 / fi))
        if bias:
            if kind == "kaiming":
                sd[name + ".bias"] = s.normal(name + ".bias", (cout,), pt)
            else:
                sd[name + ".bias"] = s.uniform(name + ".bias", (cout,), 1.0 / math.sqrt(fi))

    def linear(name, cout, cin, bias=True):
        sd[name + ".weight"] = s.normal(name + ".weight", (cout, cin), math.sqrt(2.0 / cin))
        if bias:
            sd[name + ".bias"] = s.normal(name + ".bias", (cout,), pt)

    def affine(name, c, wkey="weight", bkey="bias"):
        sd[f"{name}.{wkey}"] = s.normal(f"{name}.{wkey}", (c,), pt, 1.0)
        sd[f"{name}.{bkey}"] = s.normal(f"{name}.{bkey}", (c,), pt)

    C, H = cfg.enc_channels, cfg.enc_hidden
    if cfg.n_spks > 1:
        sd["spk_emb.weight"] = s.normal("spk_emb.weight", (cfg.n_spks, cfg.spk_emb_dim))
    sd["encoder.emb.weight"] = s.normal("encoder.emb.weight", (cfg.n_vocab, C), C ** -0.5)
    if cfg.enc_prenet:
        for i in range(3):
            conv(f"encoder.prenet.conv_layers.{i}", C, C, 5)
            affine(f"encoder.prenet.norm_layers.{i}", C, "gamma", "beta")
        sd["encoder.prenet.proj.weight"] = s.normal("encoder.prenet.proj.weight", (C, C, 1), pt)
        sd["encoder.prenet.proj.bias"] = s.normal("encoder.prenet.proj.bias", (C,), pt)
    for i in range(cfg.enc_layers):
        p = f"encoder.encoder.attn_layers.{i}"
        conv(p + ".conv_q", H, H, 1, kind="xavier")
        conv(p + ".conv_k", H, H, 1, kind="xavier")
        conv(p + ".conv_v", H, H, 1, kind="xavier")
        conv(p + ".conv_o", H, H, 1)
        for q in ("conv_q", "conv_k", "conv_v"):  # xavier touches the weight only; bias keeps the default init
            sd[f"{p}.{q}.bias"] = s.uniform(f"{p}.{q}.bias", (H,), 1.0 / math.sqrt(H))
        affine(f"encoder.encoder.norm_layers_1.{i}", H, "gamma", "beta")
        conv(f"encoder.encoder.ffn_layers.{i}.conv_1", cfg.enc_filter_channels, H, cfg.enc_kernel)
        conv(f"encoder.encoder.ffn_layers.{i}.conv_2", H, cfg.enc_filter_channels, cfg.enc_kernel)
        affine(f"encoder.encoder.norm_layers_2.{i}", H, "gamma", "beta")
    conv("encoder.proj_m", cfg.n_feats, H, 1)
    F = cfg.enc_filter_channels_dp
    conv("encoder.proj_w.conv_1", F, H, 3)
    affine("encoder.proj_w.norm_1", F, "gamma", "beta")
    conv("encoder.proj_w.conv_2", F, F, 3)
    affine("encoder.proj_w.norm_2", F, "gamma", "beta")
    conv("encoder.proj_w.proj", 1, F, 1)
    # speech-like durations from random weights: exp(bias) ~ frames per token
    sd["encoder.proj_w.proj.bias"] = torch.full((1,), math.log(dur_scale), dtype=torch.float32)

    E = "decoder.estimator."
    D, TD = cfg.dec_channels[0], cfg.time_dim
    linear(E + "time_mlp.linear_1", TD, cfg.dec_in)
    linear(E + "time_mlp.linear_2", TD, TD)

    def resnet(p, cin, cout):
        linear(p + ".mlp.1", cout, TD)
        conv(p + ".block1.block.0", cout, cin, 3, kind="kaiming")
        affine(p + ".block1.block.1", cout)
        conv(p + ".block2.block.0", cout, cout, 3, kind="kaiming")
        affine(p + ".block2.block.1", cout)
        conv(p + ".res_conv", cout, cin, 1, kind="kaiming")

    def transformer(p, dim):
        inner = cfg.dec_heads * cfg.dec_head_dim
        affine(p + ".norm1", dim)
        for q in ("to_q", "to_k", "to_v"):
            linear(f"{p}.attn1.{q}", inner, dim, bias=False)
        linear(p + ".attn1.to_out.0", dim, inner)
        affine(p + ".norm3", dim)
        linear(p + ".ff.net.0.proj", dim * 4, dim)
        sd[p + ".ff.net.0.alpha"] = s.normal(p + ".ff.net.0.alpha", (dim * 4,), 2 * pt)
        sd[p + ".ff.net.0.beta"] = s.normal(p + ".ff.net.0.beta", (dim * 4,), 2 * pt)
        linear(p + ".ff.net.2", dim, dim * 4)

    chans = cfg.dec_channels
    out_c = cfg.dec_in
    for i, c in enumerate(chans):
        in_c, out_c = out_c, c
        resnet(f"{E}down_blocks.{i}.0", in_c, out_c)
        for j in range(cfg.dec_n_blocks):
            transformer(f"{E}down_blocks.{i}.1.{j}", out_c)
        if i < len(chans) - 1:
            conv(f"{E}down_blocks.{i}.2.conv", out_c, out_c, 3, kind="kaiming")
        else:
            conv(f"{E}down_blocks.{i}.2", out_c, out_c, 3, kind="kaiming")
    for i in range(cfg.dec_mid_blocks):
        resnet(f"{E}mid_blocks.{i}.0", chans[-1], chans[-1])
        for j in range(cfg.dec_n_blocks):
            transformer(f"{E}mid_blocks.{i}.1.{j}", chans[-1])
    up = tuple(chans[::-1]) + (chans[0],)
    for i in range(len(up) - 1):
        resnet(f"{E}up_blocks.{i}.0", 2 * up[i], up[i + 1])
        for j in range(cfg.dec_n_blocks):
            transformer(f"{E}up_blocks.{i}.1.{j}", up[i + 1])
        if i < len(up) - 2:
            # nn.ConvTranspose1d keeps its default init (decoder.py:345 only matches nn.Conv1d); weight (in,out,k)
            n = f"{E}up_blocks.{i}.2.conv"
            bound = 1.0 / math.sqrt(up[i + 1] * 4)
            sd[n + ".weight"] = s.uniform(n + ".weight", (up[i + 1], up[i + 1], 4), bound)
            sd[n + ".bias"] = s.uniform(n + ".bias", (up[i + 1],), bound)
        else:
            conv(f"{E}up_blocks.{i}.2", up[i + 1], up[i + 1], 3, kind="kaiming")
    conv(E + "final_block.block.0", up[-1], up[-1], 3, kind="kaiming")
    affine(E + "final_block.block.1", up[-1])
    conv(E + "final_proj", cfg.n_feats, up[-1], 1, kind="kaiming")
    sd["mel_mean"] = torch.tensor(cfg.mel_mean, dtype=torch.float32)
    sd["mel_std"] = torch.tensor(cfg.mel_std, dtype=torch.float32)
    return sd


def hifigan_state_dict(h=HIFIGAN_V1, seed: int = 4321, std: float = 0.01, gain: float | None = None,
                       weight_norm: bool = False) -> dict:
    """HiFi-GAN v1 generator weights under the reference names (hifigan/models.py:148-179).

    `std` is the reference's init_weights N(0, 0.01) (hifigan/xutils.py:25-28).  With `gain` set, every
    resblock / upsampler conv instead gets std = gain/sqrt(fan_in) so the convolutions matter as much as the
    residual stream (a harder numerical test than the stock near-identity init).  `weight_norm=True` emits the
    checkpoint form (`weight_g`, `weight_v`) that feel_me.py:161-167 loads before `remove_weight_norm()`.
    """
    s = _Stream(seed)
    sd: dict[str, torch.Tensor] = {}

    def put(name, shape, fan_in, bias_n, default=False):
        if default:
            w = s.uniform(name + ".weight", shape, 1.0 / math.sqrt(fan_in))
        else:
            w = s.normal(name + ".weight", shape, (gain / math.sqrt(fan_in)) if gain else std)
        b = s.uniform(name + ".bias", (bias_n,), 1.0 / math.sqrt(fan_in))
        if weight_norm:
            g = w.flatten(1).norm(dim=1).reshape(-1, *([1] * (w.dim() - 1)))
            sd[name + ".weight_g"] = g * s.normal(name + ".g", g.shape, 0.05, 1.0)
            sd[name + ".weight_v"] = w
        else:
            sd[name + ".weight"] = w
        sd[name + ".bias"] = b

    c0 = h["upsample_initial_channel"]
    put("conv_pre", (c0, h["num_mels"], 7), h["num_mels"] * 7, c0, default=True)
    ch = c0
    for i, (u, k) in enumerate(zip(h["upsample_rates"], h["upsample_kernel_sizes"])):
        cin, ch = c0 // (2 ** i), c0 // (2 ** (i + 1))
        # ConvTranspose1d weight is (in, out, k); torch computes its fan_in from dim 1
        put(f"ups.{i}", (cin, ch, k), ch * k, ch)
        for j, (rk, dil) in enumerate(zip(h["resblock_kernel_sizes"], h["resblock_dilation_sizes"])):
            for l in range(len(dil)):
                put(f"resblocks.{i * 3 + j}.convs1.{l}", (ch, ch, rk), ch * rk, ch)
                put(f"resblocks.{i * 3 + j}.convs2.{l}", (ch, ch, rk), ch * rk, ch)
    put("conv_post", (1, ch, 7), ch * 7, 1)
    return sd


def phoneme_batch(batch: int, p_lo: int, p_hi: int, seed: int, n_vocab: int = 178,
                  speakers=None) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Blank-interspersed synthetic phoneme ids (utils/utils.py:131-135): Tx = 2P+1, odd slots in [1, n_vocab).

    Returns x (B, Tx_max) int64 zero-padded, x_lengths (B,), spks (B,) drawn from the 11 female emoji voices
    (feel_me.py:84-96).
    """
    rng = np.random.Generator(np.random.Philox(key=[int(seed), 77]))
    speakers = list(speakers or EMOJI_MAPPING_FEMALE.values())
    P = rng.integers(p_lo, p_hi + 1, size=batch)
    lens = 2 * P + 1
    x = np.zeros((batch, int(lens.max())), dtype=np.int64)
    for b in range(batch):
        x[b, 1:lens[b]:2] = rng.integers(1, n_vocab, size=P[b])
    spk = np.asarray(speakers, dtype=np.int64)[rng.integers(0, len(speakers), size=batch)]
    return torch.from_numpy(x), torch.from_numpy(lens.astype(np.int64)), torch.from_numpy(spk)


def mixed_length_corpus(n: int = 1024, p_lo: int = 20, p_hi: int = 150, seed: int = 1237, n_vocab: int = 178, speakers=None):
    """BASELINE.json configs[2]: `n` mixed-length utterances (P ~ U[p_lo, p_hi] phonemes, Tx = 2P + 1 blank-interspersed ids,
    about 1.5-12 s of speech each at length_scale 0.8) cycling through the 11 emoji voices.  -> list of (ids, speaker id),
    the input format of `synthesise_corpus` (cli.py:277-317 feeds the same pairs from a text file)."""
    rng = np.random.Generator(np.random.Philox(key=[int(seed), 33]))
    speakers = list(speakers or EMOJI_MAPPING_FEMALE.values())
    utts = []
    for i in range(n):
        p = int(rng.integers(p_lo, p_hi + 1))
        ids = [0] * (2 * p + 1)
        ids[1::2] = rng.integers(1, n_vocab, size=p).tolist()
        utts.append((ids, int(speakers[i % len(speakers)])))
    return utts


def prior_noise(batch: int, n_feats: int, t_pad: int, seed: int) -> torch.Tensor:
    """The injected prior-noise tensor z (B, n_feats, T_pad), before temperature scaling (flow_matching.py:51)."""
    rng = np.random.Generator(np.random.Philox(key=[int(seed), 99]))
    return torch.from_numpy(rng.standard_normal(size=(batch, n_feats, t_pad), dtype=np.float32))


def synthetic_mel(batch: int, frames: int, seed: int, cfg: MatchaConfig = VCTK) -> torch.Tensor:
    """Vocoder-only input: clip(N(mel_mean, mel_std^2), -11.51, 2.0), (B, 80, T)."""
    rng = np.random.Generator(np.random.Philox(key=[int(seed), 55]))
    m = rng.standard_normal(size=(batch, cfg.n_feats, frames), dtype=np.float32) * cfg.mel_std + cfg.mel_mean
    return torch.from_numpy(np.clip(m, -11.51, 2.0).astype(np.float32))


def checksum(sd: dict) -> float:
    """Order-independent float64 digest used to confirm both sides drew the same tensors."""
    tot = 0.0
    for k in sorted(sd):
        v = sd[k].double()
        tot += float(v.sum()) + float((v * v).sum()) * 1e-3
    return tot


def training_batch(batch: int, p_lo: int, p_hi: int, seed: int, n_feats: int = 80):
    """Text + target mel for the training-side forward pass (MatchaTTS.forward): mel lengths exceed the text lengths (the
    alignment search needs t_y >= t_x) and the padded length is a multiple of 4 (utils/model.py:14-20, as the reference's
    collate pads).  -> x, x_lengths, spks, y (B, n_feats, Ty) zero beyond its length, y_lengths."""
    x, xl, spk = phoneme_batch(batch, p_lo, p_hi, seed=seed)
    g = torch.Generator().manual_seed(seed + 1)
    yl = (xl * 2 + torch.randint(0, 9, (batch,), generator=g)).long()
    ty = int(-(-int(yl.max()) // 4) * 4)
    y = torch.randn(batch, n_feats, ty, generator=g) * 0.8
    y = y * (torch.arange(ty)[None, None, :] < yl[:, None, None])
    return x, xl, spk, y, yl


def training_draws(batch: int, n_feats: int, frames: int, seed: int):
    """The random draws of CFM.compute_loss (flow_matching.py:106-108) from a seeded generator: t (B,), z (B, n_feats, frames)."""
    g = torch.Generator().manual_seed(seed)
    return torch.rand(batch, generator=g), torch.randn(batch, n_feats, frames, generator=g)
