"""Emoji front-end: find the emoji in a text, map it to a speaker id, strip emoji and brackets.

Restates the inline logic every reference app repeats (feel_me.py:298-312; variant in
hri-demo/storytelling/demo_story_script.py:177-193).  The PyPI `emoji` package the reference uses for
`is_emoji` / `replace_emoji` is not installed offline; its tables come from Unicode's emoji-test.txt, so single code
points are recognised through an embedded copy of the `Emoji` property (emoji-data.txt, Unicode 15.1) without the
entries that are emoji only inside a sequence (digits, # and * in keycaps; regional indicators in flag pairs).
"""
from __future__ import annotations

from bisect import bisect_right

from .config import EMOJI_MAPPING_FEMALE

# (first, last) code points with Emoji=Yes, sorted; keycap bases and regional indicators are handled as sequences below
_EMOJI_RANGES = (
    (0xA9, 0xA9), (0xAE, 0xAE), (0x203C, 0x203C), (0x2049, 0x2049), (0x2122, 0x2122), (0x2139, 0x2139), (0x2194, 0x2199),
    (0x21A9, 0x21AA), (0x231A, 0x231B), (0x2328, 0x2328), (0x23CF, 0x23CF), (0x23E9, 0x23F3), (0x23F8, 0x23FA), (0x24C2, 0x24C2),
    (0x25AA, 0x25AB), (0x25B6, 0x25B6), (0x25C0, 0x25C0), (0x25FB, 0x25FE), (0x2600, 0x2604), (0x260E, 0x260E), (0x2611, 0x2611),
    (0x2614, 0x2615), (0x2618, 0x2618), (0x261D, 0x261D), (0x2620, 0x2620), (0x2622, 0x2623), (0x2626, 0x2626), (0x262A, 0x262A),
    (0x262E, 0x262F), (0x2638, 0x263A), (0x2640, 0x2640), (0x2642, 0x2642), (0x2648, 0x2653), (0x265F, 0x2660), (0x2663, 0x2663),
    (0x2665, 0x2666), (0x2668, 0x2668), (0x267B, 0x267B), (0x267E, 0x267F), (0x2692, 0x2697), (0x2699, 0x2699), (0x269B, 0x269C),
    (0x26A0, 0x26A1), (0x26A7, 0x26A7), (0x26AA, 0x26AB), (0x26B0, 0x26B1), (0x26BD, 0x26BE), (0x26C4, 0x26C5), (0x26C8, 0x26C8),
    (0x26CE, 0x26CF), (0x26D1, 0x26D1), (0x26D3, 0x26D4), (0x26E9, 0x26EA), (0x26F0, 0x26F5), (0x26F7, 0x26FA), (0x26FD, 0x26FD),
    (0x2702, 0x2702), (0x2705, 0x2705), (0x2708, 0x270D), (0x270F, 0x270F), (0x2712, 0x2712), (0x2714, 0x2714), (0x2716, 0x2716),
    (0x271D, 0x271D), (0x2721, 0x2721), (0x2728, 0x2728), (0x2733, 0x2734), (0x2744, 0x2744), (0x2747, 0x2747), (0x274C, 0x274C),
    (0x274E, 0x274E), (0x2753, 0x2755), (0x2757, 0x2757), (0x2763, 0x2764), (0x2795, 0x2797), (0x27A1, 0x27A1), (0x27B0, 0x27B0),
    (0x27BF, 0x27BF), (0x2934, 0x2935), (0x2B05, 0x2B07), (0x2B1B, 0x2B1C), (0x2B50, 0x2B50), (0x2B55, 0x2B55), (0x3030, 0x3030),
    (0x303D, 0x303D), (0x3297, 0x3297), (0x3299, 0x3299), (0x1F004, 0x1F004), (0x1F0CF, 0x1F0CF), (0x1F170, 0x1F171),
    (0x1F17E, 0x1F17F), (0x1F18E, 0x1F18E), (0x1F191, 0x1F19A), (0x1F201, 0x1F202), (0x1F21A, 0x1F21A), (0x1F22F, 0x1F22F),
    (0x1F232, 0x1F23A), (0x1F250, 0x1F251), (0x1F300, 0x1F321), (0x1F324, 0x1F393), (0x1F396, 0x1F397), (0x1F399, 0x1F39B),
    (0x1F39E, 0x1F3F0), (0x1F3F3, 0x1F3F5), (0x1F3F7, 0x1F4FD), (0x1F4FF, 0x1F53D), (0x1F549, 0x1F54E), (0x1F550, 0x1F567),
    (0x1F56F, 0x1F570), (0x1F573, 0x1F57A), (0x1F587, 0x1F587), (0x1F58A, 0x1F58D), (0x1F590, 0x1F590), (0x1F595, 0x1F596),
    (0x1F5A4, 0x1F5A5), (0x1F5A8, 0x1F5A8), (0x1F5B1, 0x1F5B2), (0x1F5BC, 0x1F5BC), (0x1F5C2, 0x1F5C4), (0x1F5D1, 0x1F5D3),
    (0x1F5DC, 0x1F5DE), (0x1F5E1, 0x1F5E1), (0x1F5E3, 0x1F5E3), (0x1F5E8, 0x1F5E8), (0x1F5EF, 0x1F5EF), (0x1F5F3, 0x1F5F3),
    (0x1F5FA, 0x1F64F), (0x1F680, 0x1F6C5), (0x1F6CB, 0x1F6D2), (0x1F6D5, 0x1F6D7), (0x1F6DC, 0x1F6E5), (0x1F6E9, 0x1F6E9),
    (0x1F6EB, 0x1F6EC), (0x1F6F0, 0x1F6F0), (0x1F6F3, 0x1F6FC), (0x1F7E0, 0x1F7EB), (0x1F7F0, 0x1F7F0), (0x1F90C, 0x1F93A),
    (0x1F93C, 0x1F945), (0x1F947, 0x1F9FF), (0x1FA70, 0x1FA7C), (0x1FA80, 0x1FA88), (0x1FA90, 0x1FABD), (0x1FABF, 0x1FAC5),
    (0x1FACE, 0x1FADB), (0x1FAE0, 0x1FAE8), (0x1FAF0, 0x1FAF8),
)
_STARTS = [lo for lo, _ in _EMOJI_RANGES]
_ZWJ, _VS16, _KEYCAP = 0x200D, 0xFE0F, 0x20E3
_KEYCAP_BASES = set(b"0123456789#*")
_SKIN = range(0x1F3FB, 0x1F400)              # emoji modifiers (components; emoji on their own too)
_REGIONAL = range(0x1F1E6, 0x1F200)
_TAGS = range(0xE0020, 0xE0080)              # tag sequences (subdivision flags) end with U+E007F


def _emoji_cp(cp: int) -> bool:
    i = bisect_right(_STARTS, cp) - 1
    return i >= 0 and cp <= _EMOJI_RANGES[i][1]


def is_emoji(ch: str) -> bool:
    """`emoji.is_emoji` for ONE character, as the apps call it (feel_me.py:300): true for a code point that is an emoji by
    itself.  Arrows such as U+2192, maths / technical symbols, digits, a lone regional indicator, ZWJ, VS16 and the
    combining keycap are not."""
    return len(ch) == 1 and _emoji_cp(ord(ch))


def _sequence_end(text: str, i: int) -> int:
    """Index just past the emoji sequence that starts at text[i], or i when none does: base (emoji | flag pair | keycap
    sequence) followed by variation selectors, skin-tone modifiers, tag characters and ZWJ-joined emoji."""
    n = len(text)
    cp = ord(text[i])
    if cp in _REGIONAL:
        if i + 1 < n and ord(text[i + 1]) in _REGIONAL:
            j = i + 2
        else:
            return i
    elif cp < 0x80 and cp in _KEYCAP_BASES:
        j = i + 1
        if j < n and ord(text[j]) == _VS16:
            j += 1
        if j < n and ord(text[j]) == _KEYCAP:
            return j + 1
        return i
    elif _emoji_cp(cp):
        j = i + 1
    else:
        return i
    while j < n:
        c = ord(text[j])
        if c == _VS16 or c in _SKIN or c in _TAGS or c == _KEYCAP:
            j += 1
        elif c == _ZWJ:
            # `emoji.replace_emoji` tokenises with keep_zwj=False: a joiner that follows an emoji goes with it
            k = _sequence_end(text, j + 1) if j + 1 < n else j + 1
            j = k if k > j + 1 else j + 1
        else:
            break
    return j


def replace_emoji(text: str, repl: str = "") -> str:
    """`emoji.replace_emoji(text, repl)`: every emoji SEQUENCE becomes `repl`; joiners / variation selectors / keycap marks
    that do not belong to an emoji stay where they are."""
    out, i, n = [], 0, len(text)
    while i < n:
        j = _sequence_end(text, i)
        if j > i:
            out.append(repl)
            i = j
        else:
            out.append(text[i])
            i += 1
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
