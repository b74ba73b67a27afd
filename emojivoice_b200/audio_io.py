"""Wire / disk formats either side of the hot path: what `matcha/cli.py` reads and writes.

  * `write_wav_pcm24` : 22.05 kHz mono PCM_24 WAV, what `sf.write(path, wav, 22050, "PCM_24")` produces (cli.py:133);
                        `soundfile` is not installed offline, so the RIFF container is written by hand
  * `save_to_folder`  : cli.py:129-135 -- `<name>.npy` (mel) + `<name>.wav`; the spectrogram PNG needs matplotlib, skipped
  * `parse_script`    : the `text|speaker` lines of cli.py:326-330 (file_synthesis_play_only)
  * `parse_emoji_script`: emoji-tagged story lines (hri-demo/storytelling/fairytale_script.txt, demo_story_script.py:177-193)
"""
from __future__ import annotations

import os
import struct

import numpy as np
import torch

from .emoji_frontend import emoji_to_spk


def float_to_pcm24(wav) -> np.ndarray:
    """float [-1, 1] -> int32 holding 24-bit samples: x * 0x7FFFFF, round half to even, clipped (libsndfile's
    normalised-float conversion with clipping on)."""
    x = np.asarray(wav.detach().cpu() if isinstance(wav, torch.Tensor) else wav, dtype=np.float64).reshape(-1)
    return np.clip(np.rint(x * 8388607.0), -8388608, 8388607).astype(np.int32)


def write_wav_pcm24(path, wav, sample_rate: int = 22050) -> str:
    pcm = float_to_pcm24(wav)
    raw = np.empty((pcm.size, 3), dtype=np.uint8)
    raw[:, 0] = pcm & 0xFF
    raw[:, 1] = (pcm >> 8) & 0xFF
    raw[:, 2] = (pcm >> 16) & 0xFF
    data = raw.tobytes()
    fmt = struct.pack("<HHIIHH", 1, 1, sample_rate, sample_rate * 3, 3, 24)      # PCM, mono, byte rate, block align, bits
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 4 + 8 + len(fmt) + 8 + len(data) + (len(data) & 1)) + b"WAVE")
        f.write(b"fmt " + struct.pack("<I", len(fmt)) + fmt)
        f.write(b"data" + struct.pack("<I", len(data)) + data + (b"\x00" if len(data) & 1 else b""))
    return str(path)


def read_wav_pcm24(path):
    """-> (float32 array in [-1, 1), sample_rate); the inverse of write_wav_pcm24 (for tests and round trips)."""
    import wave

    with wave.open(str(path), "rb") as w:
        assert w.getsampwidth() == 3 and w.getnchannels() == 1, "expected mono PCM_24"
        sr, n = w.getframerate(), w.getnframes()
        b = np.frombuffer(w.readframes(n), dtype=np.uint8).reshape(-1, 3).astype(np.int32)
    v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
    v = np.where(v >= 1 << 23, v - (1 << 24), v)
    return (v / 8388607.0).astype(np.float32), sr


def save_to_folder(filename: str, output: dict, folder) -> str:
    """cli.py:129-135: writes `<folder>/<filename>.npy` (mel) and `<folder>/<filename>.wav` (PCM_24, 22.05 kHz)."""
    os.makedirs(folder, exist_ok=True)
    mel = output["mel"]
    np.save(os.path.join(folder, filename), mel.detach().cpu().numpy() if isinstance(mel, torch.Tensor) else np.asarray(mel))
    return os.path.abspath(write_wav_pcm24(os.path.join(folder, f"{filename}.wav"), output["waveform"], 22050))


def parse_script(lines):
    """cli.py:326-330: `text|speaker_id` per line -> [(text, speaker_id)]; blank lines are skipped."""
    out = []
    for ln in lines:
        ln = ln.strip()
        if not ln:
            continue
        text, _, spk = ln.rpartition("|")
        if not text:
            raise ValueError(f"script line without '|speaker': {ln!r}")
        out.append((text.strip(), int(spk)))
    return out


def parse_emoji_script(lines, mapping=None, default: int = 12, order: str = "mapping"):
    """Emoji-tagged story lines -> [(clean_text, speaker_id)] (demo_story_script.py:177-193: the first mapping key found
    in the line picks the voice, default 12 = the neutral speaker; emoji and brackets are stripped)."""
    return [emoji_to_spk(ln.strip(), mapping, default, order) for ln in lines if ln.strip()]
