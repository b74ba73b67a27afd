#!/usr/bin/env python
"""Stage the UNMODIFIED reference files of the synthesis path under baseline/_ref/ (git-ignored, shipped to the GPU box by
gpurun) so that `bench.py --impl reference` and the `cpu_baseline` leg time the reference ITSELF on the box's host cores
(`cpu_baseline.kind = "reference"`), driven through oracle/reference_shim.py (which stubs the third-party imports that are
not installed offline).  Nothing is copied into git history; only the build container holds /root/reference.

    python scripts/install_reference.py [--src /root/reference]

A pip install of the reference (`pip install --no-index ... /root/reference`) is not possible offline: its requirements
(lightning, hydra-core, diffusers==0.25.0, conformer, phonemizer, gradio ...) are not in /opt/wheelhouse and `--no-deps`
leaves a package whose top-level imports fail; the shim route needs none of them.
"""
import argparse
import filecmp
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, "baseline", "_ref")
# the modules `reference_shim.build_matcha / build_hifigan / build_denoiser` end up importing (sys.modules listing)
FILES = [
    "Matcha-TTS/matcha/VERSION",
    "Matcha-TTS/matcha/hifigan/__init__.py", "Matcha-TTS/matcha/hifigan/denoiser.py", "Matcha-TTS/matcha/hifigan/env.py",
    "Matcha-TTS/matcha/hifigan/models.py", "Matcha-TTS/matcha/hifigan/xutils.py", "Matcha-TTS/matcha/hifigan/config.py",
    "Matcha-TTS/matcha/models/__init__.py", "Matcha-TTS/matcha/models/baselightningmodule.py",
    "Matcha-TTS/matcha/models/matcha_tts.py", "Matcha-TTS/matcha/models/components/__init__.py",
    "Matcha-TTS/matcha/models/components/decoder.py", "Matcha-TTS/matcha/models/components/flow_matching.py",
    "Matcha-TTS/matcha/models/components/text_encoder.py", "Matcha-TTS/matcha/models/components/transformer.py",
    "Matcha-TTS/matcha/utils/model.py", "Matcha-TTS/matcha/utils/monotonic_align/__init__.py", "Matcha-TTS/matcha/utils/rich_utils.py", "Matcha-TTS/matcha/utils/utils.py",
]


def install(src="/root/reference", dest=DEST, quiet=False):
    """-> number of files staged (0 when `src` is absent, e.g. on the GPU box, where the staged copy already travelled)."""
    if not os.path.isdir(os.path.join(src, "Matcha-TTS", "matcha")):
        return 0
    n = 0
    for rel in FILES:
        s, d = os.path.join(src, rel), os.path.join(dest, rel)
        if not os.path.exists(s):
            raise FileNotFoundError(s)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if not (os.path.exists(d) and filecmp.cmp(s, d, shallow=False)):
            shutil.copyfile(s, d)
        n += 1
    if not quiet:
        print(f"staged {n} reference files under {dest}")
    return n


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    a = ap.parse_args()
    sys.exit(0 if install(a.src) else 1)
