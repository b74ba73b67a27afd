"""Emoji front-end: find the emoji in a text, map it to a speaker id, strip emoji and brackets.

Restates the inline logic every reference app repeats (feel_me.py:298-312; variant in
hri-demo/storytelling/demo_story_script.py:177-193).  The PyPI `emoji` package the reference uses for
`is_emoji`/`replace_emoji` is not installed offline, so emoji code points are recognised by Unicode block.
"""
from __future__ import annotations

import unicodedata

from .config import EMOJI_MAPPING_FEMALE

_RANGES = (
    (0x1F300, 0x1FAFF),  # pictographs, emoticons, transport, supplemental symbols, symbols-and-pictographs ext-A
    (0x2600, 0x27BF),    # misc symbols + dingbats
    (0x1F000, 0x1F2FF),  # mahjong .. enclosed ideographic supplement
    (0x2B00, 0x2BFF), (0x2300, 0x23FF), (0x2190, 0x21FF), (0x3030, 0x303D), (0x3297, 0x3299),
)
_JOINERS = {0x200D, 0xFE0F, 0x20E3}          # ZWJ, variation selector-16, keycap
_SKIN = range(0x1F3FB, 0x1F400)


def is_emoji(ch: str) -> bool:
    if len(ch) != 1:
        return False
    cp = ord(ch)
    if any(lo <= cp <= hi for lo, hi in _RANGES):
        return unicodedata.category(ch) in ("So", "Sk", "Sm") or cp in _SKIN
    return cp in (0xA9, 0xAE, 0x203C, 0x2049, 0x2122, 0x2139)


def replace_emoji(text: str, repl: str = "") -> str:
    out = []
    for ch in text:
        if is_emoji(ch) or ord(ch) in _JOINERS or ord(ch) in _SKIN:
            out.append(repl)
        else:
            out.append(ch)
    return "".join(out)


def emoji_to_spk(text: str, mapping: dict | None = None, default: int = 0, order: str = "text"):
    """Returns (clean_text, speaker_id).

    order="text"    : feel_me.py:298-308 -- walk the text, collect emoji code points, the first one present in
                      `mapping` wins, else `default` (0 there).
    order="mapping" : demo_story_script.py:177-182 -- walk `mapping` in insertion order, first key contained in the
                      text wins, else `default` (12 there).
    Then strip every emoji and both round brackets (feel_me.py:309-312).
    """
    mapping = EMOJI_MAPPING_FEMALE if mapping is None else mapping
    spk = default
    if order == "text":
        for ch in text:
            if is_emoji(ch) and ch in mapping:
                spk = mapping[ch]
                break
    elif order == "mapping":
        for emote, sid in mapping.items():
            if emote in text:
                spk = sid
                break
    else:
        raise ValueError("order must be 'text' or 'mapping'")
    clean = replace_emoji(text, "").replace(")", "").replace("(", "")
    return clean, int(spk)


def intersperse(lst, item=0):
    """utils/utils.py:131-135: blank token between (and around) the symbol ids."""
    result = [item] * (len(lst) * 2 + 1)
    result[1::2] = lst
    return result
