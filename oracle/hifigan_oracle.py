"""CPU ORACLE (test infrastructure, not product code) for the HiFi-GAN v1 generator and the bias denoiser.

Functional fp32 restatement over a reference-named state_dict (after `remove_weight_norm`, i.e. plain
`weight`/`bias`; `fold_weight_norm` converts the checkpoint form).  Pinned the same way as
oracle/matcha_oracle.py (reference python files executed in the build container -> tests/golden/).
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
Paths cited are relative to /root/reference/Matcha-TTS/matcha/.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

LRELU_SLOPE = 0.1  # hifigan/models.py:11


def get_padding(kernel_size: int, dilation: int = 1) -> int:
    """hifigan/xutils.py:37-38"""
    return int((kernel_size * dilation - dilation) / 2)


def fold_weight_norm(sd: dict) -> dict:
    """torch.nn.utils.weight_norm (dim=0) folded as `remove_weight_norm` does (hifigan/models.py:199-206):
    w = g * v / ||v|| with the norm over every dim but 0."""
    out = {}
    for k, v in sd.items():
        if k.endswith(".weight_g"):
            base = k[: -len(".weight_g")]
            wv = sd[base + ".weight_v"]
            norm = wv.flatten(1).norm(dim=1).reshape(-1, *([1] * (wv.dim() - 1)))
            out[base + ".weight"] = wv * (v / norm)
        elif k.endswith(".weight_v"):
            continue
        else:
            out[k] = v
    return out


def _resblock1(sd, p, x, k, dilations):
    """hifigan/models.py:90-97"""
    for l, d in enumerate(dilations):
        xt = F.leaky_relu(x, LRELU_SLOPE)
        xt = F.conv1d(xt, sd[f"{p}.convs1.{l}.weight"], sd[f"{p}.convs1.{l}.bias"], dilation=d, padding=get_padding(k, d))
        xt = F.leaky_relu(xt, LRELU_SLOPE)
        xt = F.conv1d(xt, sd[f"{p}.convs2.{l}.weight"], sd[f"{p}.convs2.{l}.bias"], padding=get_padding(k, 1))
        x = xt + x
    return x


@torch.inference_mode()
def generator(sd, h, mel: torch.Tensor) -> torch.Tensor:
    """hifigan/models.py:181-197: mel (B, 80, T) -> wav (B, 1, 256*T)."""
    x = F.conv1d(mel, sd["conv_pre.weight"], sd["conv_pre.bias"], padding=3)
    nk = len(h["resblock_kernel_sizes"])
    for i, (u, k) in enumerate(zip(h["upsample_rates"], h["upsample_kernel_sizes"])):
        x = F.leaky_relu(x, LRELU_SLOPE)
        x = F.conv_transpose1d(x, sd[f"ups.{i}.weight"], sd[f"ups.{i}.bias"], stride=u, padding=(k - u) // 2)
        xs = None
        for j in range(nk):
            r = _resblock1(sd, f"resblocks.{i * nk + j}", x, h["resblock_kernel_sizes"][j],
                           h["resblock_dilation_sizes"][j])
            xs = r if xs is None else xs + r
        x = xs / nk
    x = F.leaky_relu(x)          # default slope 0.01 (hifigan/models.py:193)
    x = F.conv1d(x, sd["conv_post.weight"], sd["conv_post.bias"], padding=3)
    return torch.tanh(x)


# ----------------------------------------------------------------------------- hifigan/denoiser.py
def _stft(audio, n_fft=1024, hop=256, win=1024):
    """hifigan/denoiser.py:24-34: centered hann STFT -> (magnitude, phase)."""
    spec = torch.stft(audio, n_fft=n_fft, hop_length=hop, win_length=win, window=torch.hann_window(win),
                      return_complex=True)
    spec = torch.view_as_real(spec)
    return torch.sqrt(spec.pow(2).sum(-1)), torch.atan2(spec[..., -1], spec[..., 0])


@torch.inference_mode()
def denoiser_bias(sd, h) -> torch.Tensor:
    """hifigan/denoiser.py:17-56 with mode="zeros": |STFT(vocoder(zeros(1,80,88)))|[:, :, 0:1]."""
    bias_audio = generator(sd, h, torch.zeros(1, 80, 88)).float().squeeze(0)
    bias_spec, _ = _stft(bias_audio)
    return bias_spec[:, :, 0][:, :, None]


@torch.inference_mode()
def denoise(audio: torch.Tensor, bias_spec: torch.Tensor, strength: float = 0.0005) -> torch.Tensor:
    """hifigan/denoiser.py:58-64: subtract strength*bias from the magnitude, clamp at 0, ISTFT with the phase."""
    mag, ang = _stft(audio)
    mag = torch.clamp(mag - bias_spec * strength, 0.0)
    return torch.istft(torch.complex(mag * torch.cos(ang), mag * torch.sin(ang)), n_fft=1024, hop_length=256,
                       win_length=1024, window=torch.hann_window(1024))


@torch.inference_mode()
def to_waveform(sd, h, mel, bias_spec=None, strength=0.00025):
    """feel_me.py:181-187 / cli.py:121-126: vocoder(mel).clamp(-1,1), optional denoiser, squeeze."""
    audio = generator(sd, h, mel).clamp(-1, 1)
    if bias_spec is not None:
        audio = denoise(audio.squeeze(1) if audio.dim() == 3 else audio, bias_spec, strength)
    return audio.squeeze()
