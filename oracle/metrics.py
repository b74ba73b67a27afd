"""Test infrastructure (like everything under oracle/): error measures between a synthesised mel and the oracle's.

mel_cepstral_distortion: the usual MCD in dB between two log-mel spectrograms (natural log, as Matcha-TTS produces them,
matcha/utils/audio.py:60-80): c = orthonormal DCT-II over the mel axis, MCD = (10 / ln 10) * sqrt(2 * sum_{k=1..K} (c_k - c'_k)^2),
averaged over the valid frames of every utterance (K = 13 coefficients, c_0 excluded)."""
import math

import torch


def _dct_matrix(n: int, k: int) -> torch.Tensor:
    i = torch.arange(n, dtype=torch.float64)
    j = torch.arange(k, dtype=torch.float64)
    m = torch.cos(math.pi / n * (i[None, :] + 0.5) * j[:, None]) * math.sqrt(2.0 / n)
    m[0] *= 1.0 / math.sqrt(2.0)
    return m                                                   # (k, n), orthonormal rows


def mel_cepstral_distortion(mel, ref, lengths=None, n_coef: int = 13) -> float:
    """mel, ref: (B, n_mels, T) log-mel; lengths: (B,) valid frames (default: all T).  Returns the mean MCD in dB."""
    a, b = mel.detach().double().cpu(), ref.detach().double().cpu()
    B, n, T = a.shape
    d = _dct_matrix(n, n_coef + 1)[1:]                          # drop c_0 (energy)
    diff = torch.einsum("kn,bnt->bkt", d, a - b)
    per_frame = (10.0 / math.log(10.0)) * torch.sqrt(2.0 * (diff * diff).sum(1))      # (B, T)
    if lengths is None:
        return float(per_frame.mean())
    lens = torch.as_tensor(lengths).reshape(-1).long().cpu()
    mask = torch.arange(T)[None, :] < lens[:, None]
    return float(per_frame[mask].mean())
