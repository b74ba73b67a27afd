"""CPU ORACLE (test infrastructure, not product code) for `MatchaTTS.synthesise`.

A functional, fp32, single-device restatement of the reference inference graph that works straight off a
reference-named `state_dict` and accepts an injected prior-noise tensor.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module; the product path
(emojivoice_b200/) never does.

Pinning: the reference ships no tests or golden vectors for this path, so the oracle is pinned
against the reference's own python files executed in the build container (oracle/reference_shim.py +
scripts/make_golden.py -> tests/golden/*.npz; tests/test_oracle_vs_reference.py re-derives them when
/root/reference is present).  The one third-party piece that is not under /root/reference --
diffusers==0.25.0 `Attention`/`AttnProcessor2_0` (Matcha-TTS/requirements.txt:40) -- is restated from that
release's published behaviour (see `_diffusers_attention`), structurally pinned by the parameter count
18,204,193 printed in synthesis.ipynb:127.

Every function cites the reference lines it follows (paths relative to /root/reference/Matcha-TTS/matcha/).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------- utils/model.py
def sequence_mask(length: torch.Tensor, max_length: int) -> torch.Tensor:
    """utils/model.py:7-11"""
    pos = torch.arange(max_length, dtype=length.dtype, device=length.device)
    return pos.unsqueeze(0) < length.unsqueeze(1)


def fix_len_compatibility(length: torch.Tensor, num_downsamplings: int = 2) -> int:
    """utils/model.py:14-20 (float division, ceil, times 2**n)."""
    factor = torch.scalar_tensor(2).pow(num_downsamplings)
    return int(((length / factor).ceil() * factor).int().item())


def generate_path(duration: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """utils/model.py:29-41: float cumsum, `arange < cum`, first difference along tokens, times mask."""
    b, t_x, t_y = mask.shape
    cum = torch.cumsum(duration, 1)
    path = sequence_mask(cum.view(b * t_x), t_y).to(mask.dtype).view(b, t_x, t_y)
    path = path - F.pad(path, [0, 0, 1, 0, 0, 0])[:, :-1]
    return path * mask


# ----------------------------------------------------------------------------- text_encoder.py
def _channel_layer_norm(x, gamma, beta, eps=1e-4):
    """models/components/text_encoder.py:15-33 (normalises over dim 1, biased variance, eps inside rsqrt)."""
    mean = torch.mean(x, 1, keepdim=True)
    var = torch.mean((x - mean) ** 2, 1, keepdim=True)
    x = (x - mean) * torch.rsqrt(var + eps)
    return x * gamma.view(1, -1, 1) + beta.view(1, -1, 1)


def _rope(x: torch.Tensor, d: int, base: float = 10000.0) -> torch.Tensor:
    """text_encoder.py:97-172. x: (B, H, T, Dh); rotates the first `d` features with rotate-half pairing."""
    t = x.shape[2]
    theta = 1.0 / (base ** (torch.arange(0, d, 2).float() / d))
    idx = torch.einsum("n,d->nd", torch.arange(t).float(), theta)
    idx2 = torch.cat([idx, idx], dim=1)
    cos, sin = idx2.cos()[None, None], idx2.sin()[None, None]      # (1,1,T,d)
    xr, xp = x[..., :d], x[..., d:]
    neg_half = torch.cat([-xr[..., d // 2:], xr[..., : d // 2]], dim=-1)
    return torch.cat([xr * cos + neg_half * sin, xp], dim=-1)


def _enc_attention(sd, p, x, attn_mask, n_heads):
    """text_encoder.py:223-252: 1x1 conv q/k/v/o, RoPE on half of each head, -1e4 masked fill, softmax."""
    q = F.conv1d(x, sd[p + ".conv_q.weight"], sd[p + ".conv_q.bias"])
    k = F.conv1d(x, sd[p + ".conv_k.weight"], sd[p + ".conv_k.bias"])
    v = F.conv1d(x, sd[p + ".conv_v.weight"], sd[p + ".conv_v.bias"])
    b, d, t = k.shape
    kc = d // n_heads
    q, k, v = (z.view(b, n_heads, kc, t).transpose(2, 3) for z in (q, k, v))   # b h t c
    q, k = _rope(q, kc // 2), _rope(k, kc // 2)
    scores = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(kc)
    scores = scores.masked_fill(attn_mask == 0, -1e4)
    out = torch.matmul(F.softmax(scores, dim=-1), v)
    out = out.transpose(2, 3).contiguous().view(b, d, t)
    return F.conv1d(out, sd[p + ".conv_o.weight"], sd[p + ".conv_o.bias"])


def text_encoder(sd, cfg, x, x_lengths, spk_emb):
    """text_encoder.py:378-410 (+ ConvReluNorm :60-67, Encoder :314-325, FFN :267-273, DurationPredictor :84-94)."""
    P = "encoder."
    h = F.embedding(x, sd[P + "emb.weight"]) * math.sqrt(cfg.enc_channels)
    h = h.transpose(1, -1)
    x_mask = sequence_mask(x_lengths, h.size(2)).unsqueeze(1).to(h.dtype)
    if cfg.enc_prenet:
        org = h
        for i in range(3):
            h = F.conv1d(h * x_mask, sd[f"{P}prenet.conv_layers.{i}.weight"], sd[f"{P}prenet.conv_layers.{i}.bias"],
                         padding=2)
            h = _channel_layer_norm(h, sd[f"{P}prenet.norm_layers.{i}.gamma"], sd[f"{P}prenet.norm_layers.{i}.beta"])
            h = torch.relu(h)
        h = org + F.conv1d(h, sd[P + "prenet.proj.weight"], sd[P + "prenet.proj.bias"])
        h = h * x_mask
    if cfg.n_spks > 1:
        h = torch.cat([h, spk_emb.unsqueeze(-1).repeat(1, 1, h.shape[-1])], dim=1)
    attn_mask = x_mask.unsqueeze(2) * x_mask.unsqueeze(-1)
    pad = cfg.enc_kernel // 2
    for i in range(cfg.enc_layers):
        h = h * x_mask
        y = _enc_attention(sd, f"{P}encoder.attn_layers.{i}", h, attn_mask, cfg.enc_heads)
        h = _channel_layer_norm(h + y, sd[f"{P}encoder.norm_layers_1.{i}.gamma"], sd[f"{P}encoder.norm_layers_1.{i}.beta"])
        f = f"{P}encoder.ffn_layers.{i}"
        y = F.conv1d(h * x_mask, sd[f + ".conv_1.weight"], sd[f + ".conv_1.bias"], padding=pad)
        y = torch.relu(y)
        y = F.conv1d(y * x_mask, sd[f + ".conv_2.weight"], sd[f + ".conv_2.bias"], padding=pad) * x_mask
        h = _channel_layer_norm(h + y, sd[f"{P}encoder.norm_layers_2.{i}.gamma"], sd[f"{P}encoder.norm_layers_2.{i}.beta"])
    h = h * x_mask
    mu = F.conv1d(h, sd[P + "proj_m.weight"], sd[P + "proj_m.bias"]) * x_mask
    W = P + "proj_w."
    d = F.conv1d(h * x_mask, sd[W + "conv_1.weight"], sd[W + "conv_1.bias"], padding=1)
    d = _channel_layer_norm(torch.relu(d), sd[W + "norm_1.gamma"], sd[W + "norm_1.beta"])
    d = F.conv1d(d * x_mask, sd[W + "conv_2.weight"], sd[W + "conv_2.bias"], padding=1)
    d = _channel_layer_norm(torch.relu(d), sd[W + "norm_2.gamma"], sd[W + "norm_2.beta"])
    logw = F.conv1d(d * x_mask, sd[W + "proj.weight"], sd[W + "proj.bias"]) * x_mask
    return mu, logw, x_mask


# ----------------------------------------------------------------------------- decoder.py / transformer.py
def time_embedding(sd, t: torch.Tensor, dim: int) -> torch.Tensor:
    """decoder.py:14-29 (sinusoid, scale 1000) then TimestepEmbedding :105-117 (Linear, SiLU, Linear)."""
    E = "decoder.estimator.time_mlp."
    if t.ndim < 1:
        t = t.unsqueeze(0)
    half = dim // 2
    emb = math.log(10000) / (half - 1)
    emb = torch.exp(torch.arange(half).float() * -emb)
    emb = 1000 * t.unsqueeze(1) * emb.unsqueeze(0)
    emb = torch.cat((emb.sin(), emb.cos()), dim=-1)
    emb = F.linear(emb, sd[E + "linear_1.weight"], sd[E + "linear_1.bias"])
    return F.linear(F.silu(emb), sd[E + "linear_2.weight"], sd[E + "linear_2.bias"])


def _block1d(sd, p, x, mask):
    """decoder.py:32-43: conv3(x*mask) -> GroupNorm(8) -> Mish -> *mask (GN statistics include padded frames)."""
    h = F.conv1d(x * mask, sd[p + ".block.0.weight"], sd[p + ".block.0.bias"], padding=1)
    h = F.group_norm(h, 8, sd[p + ".block.1.weight"], sd[p + ".block.1.bias"], eps=1e-5)
    return F.mish(h) * mask


def _resnet(sd, p, x, mask, temb):
    """decoder.py:46-61"""
    h = _block1d(sd, p + ".block1", x, mask)
    h = h + F.linear(F.mish(temb), sd[p + ".mlp.1.weight"], sd[p + ".mlp.1.bias"]).unsqueeze(-1)
    h = _block1d(sd, p + ".block2", h, mask)
    return h + F.conv1d(x * mask, sd[p + ".res_conv.weight"], sd[p + ".res_conv.bias"])


def _diffusers_attention(sd, p, h, mask, heads):
    """diffusers==0.25.0 Attention + AttnProcessor2_0 as called at transformer.py:266-271.

    to_q/k/v: Linear(dim, heads*dim_head, bias=False); the (B, T) float mask is repeated per head, viewed as
    (B, heads, 1, T) and handed to F.scaled_dot_product_attention as a FLOAT attn_mask, i.e. it is ADDED to the
    logits (valid keys +1, padded keys +0) -- no key is ever excluded.  to_out: Linear + bias.
    """
    b, t, _ = h.shape
    q = F.linear(h, sd[p + ".to_q.weight"])
    k = F.linear(h, sd[p + ".to_k.weight"])
    v = F.linear(h, sd[p + ".to_v.weight"])
    hd = q.shape[-1] // heads
    q, k, v = (z.view(b, t, heads, hd).transpose(1, 2) for z in (q, k, v))
    bias = mask.to(q.dtype).repeat_interleave(heads, dim=0).unsqueeze(1)          # prepare_attention_mask
    bias = bias.view(b, heads, -1, bias.shape[-1])
    o = F.scaled_dot_product_attention(q, k, v, attn_mask=bias, dropout_p=0.0, is_causal=False)
    o = o.transpose(1, 2).reshape(b, t, heads * hd)
    return F.linear(o, sd[p + ".to_out.0.weight"], sd[p + ".to_out.0.bias"])


def _snake_beta_ff(sd, p, h):
    """transformer.py:64-80 (SnakeBeta with log-scale alpha/beta) and :126,131-134 (project out)."""
    y = F.linear(h, sd[p + ".net.0.proj.weight"], sd[p + ".net.0.proj.bias"])
    alpha, beta = torch.exp(sd[p + ".net.0.alpha"]), torch.exp(sd[p + ".net.0.beta"])
    y = y + (1.0 / (beta + 0.000000001)) * torch.pow(torch.sin(y * alpha), 2)
    return F.linear(y, sd[p + ".net.2.weight"], sd[p + ".net.2.bias"])


def _transformer(sd, p, x, mask, heads):
    """transformer.py:243-316 at the default config: pre-LN self-attention + pre-LN SnakeBeta feed-forward."""
    n = F.layer_norm(x, (x.shape[-1],), sd[p + ".norm1.weight"], sd[p + ".norm1.bias"], 1e-5)
    x = _diffusers_attention(sd, p + ".attn1", n, mask, heads) + x
    n = F.layer_norm(x, (x.shape[-1],), sd[p + ".norm3.weight"], sd[p + ".norm3.bias"], 1e-5)
    return _snake_beta_ff(sd, p + ".ff", n) + x


def estimator(sd, cfg, x, mask, mu, t, spks):
    """decoder.py:363-443 for channels=(c,c) style U-Nets with n_blocks transformer blocks per level."""
    E = "decoder.estimator."
    temb = time_embedding(sd, t, cfg.dec_in)
    x = torch.cat([x, mu], dim=1)
    if spks is not None:
        x = torch.cat([x, spks.unsqueeze(-1).expand(-1, -1, x.shape[-1])], dim=1)
    nlev = len(cfg.dec_channels)
    hiddens, masks = [], [mask]

    def tblocks(prefix, x, m):
        x = x.transpose(1, 2)
        for j in range(cfg.dec_n_blocks):
            x = _transformer(sd, f"{prefix}.{j}", x, m[:, 0], cfg.dec_heads)
        return x.transpose(1, 2)

    for i in range(nlev):
        m = masks[-1]
        x = _resnet(sd, f"{E}down_blocks.{i}.0", x, m, temb)
        x = tblocks(f"{E}down_blocks.{i}.1", x, m)
        hiddens.append(x)
        if i < nlev - 1:
            x = F.conv1d(x * m, sd[f"{E}down_blocks.{i}.2.conv.weight"], sd[f"{E}down_blocks.{i}.2.conv.bias"],
                         stride=2, padding=1)
        else:
            x = F.conv1d(x * m, sd[f"{E}down_blocks.{i}.2.weight"], sd[f"{E}down_blocks.{i}.2.bias"], padding=1)
        masks.append(m[:, :, ::2])
    masks = masks[:-1]
    m = masks[-1]
    for i in range(cfg.dec_mid_blocks):
        x = _resnet(sd, f"{E}mid_blocks.{i}.0", x, m, temb)
        x = tblocks(f"{E}mid_blocks.{i}.1", x, m)
    for i in range(nlev):
        m = masks.pop()
        x = _resnet(sd, f"{E}up_blocks.{i}.0", torch.cat([x, hiddens.pop()], dim=1), m, temb)
        x = tblocks(f"{E}up_blocks.{i}.1", x, m)
        if i < nlev - 1:
            x = F.conv_transpose1d(x * m, sd[f"{E}up_blocks.{i}.2.conv.weight"], sd[f"{E}up_blocks.{i}.2.conv.bias"],
                                   stride=2, padding=1)
        else:
            x = F.conv1d(x * m, sd[f"{E}up_blocks.{i}.2.weight"], sd[f"{E}up_blocks.{i}.2.bias"], padding=1)
    x = _block1d(sd, E + "final_block", x, m)
    out = F.conv1d(x * m, sd[E + "final_proj.weight"], sd[E + "final_proj.bias"])
    return out * mask


def solve_euler(sd, cfg, x, t_span, mu, mask, spks, return_all=False):
    """flow_matching.py:55-85 (t, dt tracked as float32 0-dim tensors exactly as the reference does)."""
    t, dt = t_span[0], t_span[1] - t_span[0]
    sol = []
    for step in range(1, len(t_span)):
        dphi = estimator(sd, cfg, x, mask, mu, t, spks)
        x = x + dt * dphi
        t = t + dt
        sol.append(x)
        if step < len(t_span) - 1:
            dt = t_span[step + 1] - t
    return sol if return_all else sol[-1]


# ----------------------------------------------------------------------------- matcha_tts.py
@torch.inference_mode()
def synthesise(sd, cfg, x, x_lengths, n_timesteps, temperature=1.0, spks=None, length_scale=1.0, z=None,
               return_steps=False):
    """matcha_tts.py:77-152 with the prior noise injectable (`z`, (B, n_feats, T_pad), before temperature)."""
    spk_emb = F.embedding(spks.long(), sd["spk_emb.weight"]) if cfg.n_spks > 1 else None
    mu_x, logw, x_mask = text_encoder(sd, cfg, x, x_lengths, spk_emb)
    w = torch.exp(logw) * x_mask
    w_ceil = torch.ceil(w) * length_scale
    y_lengths = torch.clamp_min(torch.sum(w_ceil, [1, 2]), 1).long()
    y_max_length = y_lengths.max()
    t_pad = fix_len_compatibility(y_max_length)
    y_mask = sequence_mask(y_lengths, t_pad).unsqueeze(1).to(x_mask.dtype)
    attn_mask = x_mask.unsqueeze(-1) * y_mask.unsqueeze(2)
    attn = generate_path(w_ceil.squeeze(1), attn_mask.squeeze(1)).unsqueeze(1)
    mu_y = torch.matmul(attn.squeeze(1).transpose(1, 2), mu_x.transpose(1, 2)).transpose(1, 2)
    if z is None:
        z = torch.randn_like(mu_y)
    assert z.shape == mu_y.shape, (z.shape, mu_y.shape)
    x0 = z * temperature                                                    # flow_matching.py:51
    t_span = torch.linspace(0, 1, n_timesteps + 1)
    steps = solve_euler(sd, cfg, x0, t_span, mu_y, y_mask, spk_emb, return_all=True)
    dec = steps[-1][:, :, :y_max_length]
    out = {
        "encoder_outputs": mu_y[:, :, :y_max_length],
        "decoder_outputs": dec,
        "attn": attn[:, :, :y_max_length],
        "mel": dec * sd["mel_std"] + sd["mel_mean"],                        # utils/model.py:71-90
        "mel_lengths": y_lengths,
        # intermediates the parity tests look at
        "mu_x": mu_x, "logw": logw, "x_mask": x_mask, "w_ceil": w_ceil, "t_pad": t_pad, "y_mask": y_mask,
        "mu_y": mu_y, "attn_full": attn, "decoder_outputs_full": steps[-1],
    }
    if return_steps:
        out["steps"] = steps
    return out


# ----------------------------------------------------------------------------- matcha_tts.py: training forward
@torch.inference_mode()
def forward_losses(sd, cfg, x, x_lengths, y, y_lengths, spks=None, out_size=None, durations=None, *, t=None, z=None,
                   out_offset=None, prior_loss=True, use_precomputed_durations=False, maximum_path=None):
    """matcha_tts.py:154-245 + CFM.compute_loss (flow_matching.py:87-118) + duration_loss (utils/model.py:44-46) as loss
    VALUES.  The reference's random draws are injectable: `t` (B,) and `z` (like the cut target) -- flow_matching.py:106-108
    draws t = torch.rand([b,1,1]) first, then z = torch.randn_like(x1) -- and `out_offset` (B,) for the segment cut.
    `maximum_path` defaults to the C restatement of the alignment search (oracle/mas_oracle.py).
    -> dict(dur_loss, prior_loss, diff_loss, attn, plus the intermediates the parity tests compare)."""
    if maximum_path is None:
        from . import mas_oracle
        maximum_path = mas_oracle.maximum_path
    spk_emb = F.embedding(spks.long(), sd["spk_emb.weight"]) if cfg.n_spks > 1 else None
    mu_x, logw, x_mask = text_encoder(sd, cfg, x, x_lengths, spk_emb)
    y_max_length = y.shape[-1]
    y_mask = sequence_mask(y_lengths, y_max_length).unsqueeze(1).to(x_mask.dtype)
    attn_mask = x_mask.unsqueeze(-1) * y_mask.unsqueeze(2)
    log_prior = None
    if use_precomputed_durations:
        attn = generate_path(durations.squeeze(1), attn_mask.squeeze(1))
    else:
        const = -0.5 * math.log(2 * math.pi) * cfg.n_feats
        factor = -0.5 * torch.ones(mu_x.shape, dtype=mu_x.dtype)
        y_square = torch.matmul(factor.transpose(1, 2), y ** 2)
        y_mu_double = torch.matmul(2.0 * (factor * mu_x).transpose(1, 2), y)
        mu_square = torch.sum(factor * (mu_x ** 2), 1).unsqueeze(-1)
        log_prior = y_square - y_mu_double + mu_square + const
        attn = maximum_path(log_prior, attn_mask.squeeze(1))
    logw_ = torch.log(1e-8 + torch.sum(attn.unsqueeze(1), -1)) * x_mask
    dur_loss = torch.sum((logw - logw_) ** 2) / torch.sum(x_lengths)
    if out_size is not None:                                               # matcha_tts.py:211-233
        attn_cut = torch.zeros(attn.shape[0], attn.shape[1], out_size, dtype=attn.dtype)
        y_cut = torch.zeros(y.shape[0], cfg.n_feats, out_size, dtype=y.dtype)
        cut_lengths = []
        for i in range(y.shape[0]):
            n = int(out_size + (y_lengths[i] - out_size).clamp(None, 0))
            lo = int(out_offset[i])
            cut_lengths.append(n)
            y_cut[i, :, :n] = y[i, :, lo:lo + n]
            attn_cut[i, :, :n] = attn[i, :, lo:lo + n]
        y_cut_lengths = torch.LongTensor(cut_lengths)
        attn, y, y_mask = attn_cut, y_cut, sequence_mask(y_cut_lengths, out_size).unsqueeze(1).to(y_mask.dtype)
    mu_y = torch.matmul(attn.transpose(1, 2), mu_x.transpose(1, 2)).transpose(1, 2)
    b = mu_y.shape[0]
    t = torch.rand([b, 1, 1]) if t is None else t.reshape(b, 1, 1).to(mu_y.dtype)
    z = torch.randn_like(y) if z is None else z
    sigma_min = cfg.sigma_min
    y_t = (1 - (1 - sigma_min) * t) * z + t * y
    u = y - (1 - sigma_min) * z
    v = estimator(sd, cfg, y_t, y_mask, mu_y, t.squeeze(), spk_emb) if b > 1 else estimator(sd, cfg, y_t, y_mask, mu_y, t.reshape(1), spk_emb)
    diff_loss = F.mse_loss(v, u, reduction="sum") / (torch.sum(y_mask) * u.shape[1])
    if prior_loss:
        pl = torch.sum(0.5 * ((y - mu_y) ** 2 + math.log(2 * math.pi)) * y_mask) / (torch.sum(y_mask) * cfg.n_feats)
    else:
        pl = 0
    return {"dur_loss": dur_loss, "prior_loss": pl, "diff_loss": diff_loss, "attn": attn,
            "log_prior": log_prior, "mu_x": mu_x, "logw": logw, "mu_y": mu_y, "y_t": y_t, "u": u, "v": v, "y_mask": y_mask}
