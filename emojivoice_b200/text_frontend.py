"""Symbol inventory and id mapping of the Matcha-TTS text front-end.

Restates the *data format* of `Matcha-TTS/matcha/text/symbols.py:6-17` (pad, punctuation, latin letters, IPA letters in
id order) and the id mapping / sequence helpers of `matcha/text/__init__.py:6-50`, plus `process_text`
(`matcha/cli.py:53-78`).  The grapheme-to-phoneme step (`cleaners.py:248-257`, espeak-ng through `phonemizer`) is not
available offline, so `text_to_sequence` takes the phonemiser as a callable; everything after it -- symbol lookup,
blank interspersing, tensors -- is implemented and tested here.  The table has 198 entries, four of which are
duplicates ("'" x3 more, "-"... see `DUPLICATES`); like the reference's dict comprehension, the LAST index wins.
"""
from __future__ import annotations

import torch

from .emoji_frontend import intersperse

PAD = "_"
PUNCTUATION = ';:,.!?¡¿—…"«»“” '
LETTERS = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz"
LETTERS_IPA = ("ɑɐɒæɓʙβɔɕçɗɖðʤəɘɚɛɜɝɞɟʄɡɠɢʛɦɧħɥʜɨɪʝɭɬɫɮʟɱɯɰŋɳɲɴøɵɸθœɶʘɹɺɾɻʀʁɽʂʃʈʧʉʊʋⱱʌɣɤʍχʎʏʑʐʒʔʡʕʢǀǁǂǃˈˌːˑʼʴʰʱʲʷˠˤ˞↓↑→↗↘'̩'ᵻ'̃'-'̞ᵝʨʦũĩʣʥ%+]\\()[")

SYMBOLS = [PAD] + list(PUNCTUATION) + list(LETTERS) + list(LETTERS_IPA)
SYMBOL_TO_ID = {s: i for i, s in enumerate(SYMBOLS)}          # duplicates: last occurrence wins (text/__init__.py:6)
ID_TO_SYMBOL = dict(enumerate(SYMBOLS))
SPACE_ID = SYMBOLS.index(" ")
DUPLICATES = sorted({s for s in SYMBOLS if SYMBOLS.count(s) > 1})


def cleaned_text_to_sequence(cleaned_text: str) -> list[int]:
    """text/__init__.py:29-38: ids of an already phonemised string; unknown symbols raise KeyError like the reference."""
    return [SYMBOL_TO_ID[ch] for ch in cleaned_text]


def sequence_to_text(sequence) -> str:
    """text/__init__.py:41-47"""
    return "".join(ID_TO_SYMBOL[int(i)] for i in sequence)


def text_to_sequence(text: str, phonemizer=None):
    """text/__init__.py:10-26 with the cleaner chain replaced by `phonemizer(text) -> str` (english_cleaners2 in the
    reference = lowercase, expand abbreviations, espeak-ng IPA, collapse whitespace).  -> (ids, cleaned_text)"""
    cleaned = phonemizer(text) if phonemizer is not None else text
    return cleaned_text_to_sequence(cleaned), cleaned


def process_text(text: str, phonemizer=None, device=None) -> dict:
    """cli.py:53-78: ids with a blank (0) between and around symbols, as a (1, Tx) LongTensor plus its length."""
    ids, cleaned = text_to_sequence(text, phonemizer)
    x = torch.tensor(intersperse(ids, 0), dtype=torch.long, device=device)[None]
    return {"x_orig": text, "x": x, "x_lengths": torch.tensor([x.shape[-1]], dtype=torch.long, device=device),
            "x_phones": sequence_to_text(x[0].tolist()), "cleaned": cleaned}
