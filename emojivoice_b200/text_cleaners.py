"""Rule-based text normalisation in front of the phonemiser.

Restates the deterministic half of `Matcha-TTS/matcha/text/cleaners.py` -- lowercase (:237), abbreviation expansion
(:226-235, tables :76-129), symbol / currency replacements (:212-224, tables :131-210), whitespace collapse (:241) -- and
the cleaner pipelines built from them (:248-318).  The grapheme-to-phoneme step itself (espeak-ng through `phonemizer`,
misaki for Japanese) is a third-party binary that is absent offline: every pipeline takes it as a callable
`g2p(text) -> str`; with `g2p=None` the normalised text is returned as is.

The rule tables are DATA that has to match the reference's character for character (including its quirks: the unescaped
dots in the French / German abbreviation keys, "vergule", the Spanish pipeline that looks up a table that does not
exist); they are pinned by `tests/test_text_cleaners.py` against the reference file imported where it lies.

`normalize_numbers` restates `text/numbers.py:1-71` (not called by any cleaner of the reference, kept for completeness).
Its speller replaces the `inflect` package (absent offline) for the three call forms numbers.py uses; parity for that
function is pinned on hand-checked vectors only.
"""
from __future__ import annotations

import re

_WS = re.compile(r"\s+")

# (key, expansion): compiled as  \b<key>\.  with IGNORECASE; keys are used verbatim as regex source (cleaners.py:76-129)
_ABBREVIATIONS = {
    "en": [("mrs", "misess"), ("ms", "miss"), ("mr", "mister"), ("dr", "doctor"), ("st", "saint"), ("co", "company"),
           ("jr", "junior"), ("maj", "major"), ("gen", "general"), ("drs", "doctors"), ("rev", "reverend"),
           ("lt", "lieutenant"), ("hon", "honorable"), ("sgt", "sergeant"), ("capt", "captain"), ("esq", "esquire"),
           ("ltd", "limited"), ("col", "colonel"), ("ft", "fort")],
    "fr": [("m.", "monsieur"), ("dr", "docteur"), ("st", "saint")],
    "de": [("hr", "herr"), ("fr", "frau"), ("dr", "doktor"), ("prof", "professor"), ("bsp", "beispiel"),
           ("usw", "und so weiter"), ("z", "zu"), ("z.b", "zum beispiel"), ("ca", "zirka"), ("bzw", "beziehungsweise"),
           ("d.h", "das heißt"), ("u.a", "unter anderem"), ("u.u", "unter umständen"), ("u.v.m", "und vieles mehr"),
           ("vgl", "vergleiche")],
}
_ABBREVIATION_RES = {lang: [(re.compile("\\b%s\\." % key, re.IGNORECASE), word) for key, word in table]
                     for lang, table in _ABBREVIATIONS.items()}

_ELLIPSIS_IN, _ELLIPSIS_OUT = (r"\.\.\.", "ELLIPSIS_MARKER"), (r"ELLIPSIS_MARKER", "...")
_DOT_BETWEEN_NON_DIGITS = r"(?<=\D)\.(?=\D)(?!\s)"
# (pattern, replacement[, flags]) in application order (cleaners.py:131-210)
_REPLACEMENTS = {
    "ja": [(r"(?<!\s)\.(?!\s)", " てん"), (r"-(?=\d)", " えん"), (r"%", " パーセント"), (r"@", " アットマーク"),
           (r"\\\\", " バックスラッシュ"), (r"/", " スラッシュ"), (r"\$", " ドル"), (r"€", " ユーロ"), (r"¥", " えん"),
           (r"\+", " プラス"), (r"=", " イコール")],
    "en": [_ELLIPSIS_IN,
           (r"\$(\d+)\.(\d+)", r"\1 dollars and \2 cents"), (r"€(\d+)\.(\d+)", r"\1 euros and \2 cents"),
           (r"¥(\d+)\.(\d+)", r"\1 yen and \2 cents"),
           (_DOT_BETWEEN_NON_DIGITS, " dot ", re.IGNORECASE), (r"(?<=\d)\.(?=\d)(?!\s)", " point "),
           (r"\$(\d+)", r"\1 dollars"), (r"€(\d+)", r"\1 euros"), (r"¥(\d+)", r"\1 yen"),
           _ELLIPSIS_OUT],
    "fr": [_ELLIPSIS_IN, (r"\(", ""), (r"\)", ""),
           (r"(\d+)\.(\d+)\$", r"\1 dollars et \2 centimes"), (r"(\d+)\.(\d+)€", r"\1 euros et \2 centimes"),
           (r"(\d+)\.(\d+)¥", r"\1 yen et \2 centimes"),
           (_DOT_BETWEEN_NON_DIGITS, " point ", re.IGNORECASE), (r"(?<=\d)\,(?=\d)(?!\s)", " vergule "),
           (r"€", " euros"), (r"¥", " yen"), (r"Mme", "madame"), (r"Mlle", "mademoiselle"), (r"=", " égales "),
           (r"/", " slash "), (r"-(?=\d)(?!\s)", "négatif "),
           _ELLIPSIS_OUT],
    "de": [_ELLIPSIS_IN, (r"\(", ""), (r"\)", ""),
           (r"(\d+)\.(\d+)\$", r"\1 Dollar und \2 Cent"), (r"(\d+)\.(\d+)€", r"\1 Euro und \2 Cent"),
           (r"(\d+)\.(\d+)¥", r"\1 Yen und \2 Sen"),
           (_DOT_BETWEEN_NON_DIGITS, " Punkt ", re.IGNORECASE), (r"(?<=\d)\,(?=\d)(?!\s)", " Komma "),
           (r"€", " Euro"), (r"¥", " Yen"), (r"Mme", "Frau"), (r"Mlle", "Fräulein"), (r"=", " gleich "),
           (r"/", " Schrägstrich "), (r"-(?=\d)(?!\s)", "minus "),
           _ELLIPSIS_OUT],
}
_REPLACEMENT_RES = {lang: [(re.compile(rule[0], rule[2] if len(rule) > 2 else 0), rule[1]) for rule in table]
                    for lang, table in _REPLACEMENTS.items()}


def lowercase(text: str) -> str:
    return text.lower()


def collapse_whitespace(text: str) -> str:
    return _WS.sub(" ", text)


def _table(tables: dict, language: str, what: str):
    # the reference leaves its local unbound for a language without a table (cleaners.py:213-221, :227-232): the Spanish
    # pipeline therefore fails with UnboundLocalError; same error type here
    if language not in tables:
        raise UnboundLocalError(f"no {what} table for language {language!r} (the reference fails the same way)")
    return tables[language]


def expand_abbreviations(text: str, language: str) -> str:
    for rx, word in _table(_ABBREVIATION_RES, language, "abbreviation"):
        text = rx.sub(word, text)
    return text


def apply_replacements(text: str, language: str) -> str:
    for rx, repl in _table(_REPLACEMENT_RES, language, "replacement"):
        text = rx.sub(repl, text)
    return text


def basic_cleaners(text: str) -> str:
    """cleaners.py:245: lowercase + whitespace collapse, no transliteration."""
    return collapse_whitespace(lowercase(text))


def _latin_pipeline(text: str, language: str, g2p):
    text = text.encode("utf-8").decode("utf-8")
    text = apply_replacements(expand_abbreviations(lowercase(text), language), language)
    phonemes = g2p(text) if g2p is not None else text
    return collapse_whitespace(phonemes)


def english_cleaners2(text: str, g2p=None) -> str:
    """cleaners.py:248-257; `g2p` stands for espeak-ng en-us (punctuation kept, stress marks, language flags removed)."""
    return _latin_pipeline(text, "en", g2p)


def french_cleaners(text: str, g2p=None) -> str:
    """cleaners.py:259-268"""
    return _latin_pipeline(text, "fr", g2p)


def german_cleaners(text: str, g2p=None) -> str:
    """cleaners.py:270-279"""
    return _latin_pipeline(text, "de", g2p)


def spanish_cleaners(text: str, g2p=None) -> str:
    """cleaners.py:292-301: the reference has no Spanish tables, so this raises UnboundLocalError there and here."""
    return _latin_pipeline(text, "es", g2p)


def japanese_cleaners(text: str, g2p=None) -> str:
    """cleaners.py:281-290: replacements only (no lowercase / abbreviations); `g2p` stands for misaki's JAG2P, whose first
    return value is the phoneme string."""
    text = apply_replacements(text.encode("utf-8").decode("utf-8"), "ja")
    return collapse_whitespace(g2p(text) if g2p is not None else text)


CLEANERS = {f.__name__: f for f in (basic_cleaners, english_cleaners2, french_cleaners, german_cleaners, spanish_cleaners,
                                    japanese_cleaners)}


def clean_text(text: str, cleaner_names, g2p=None) -> str:
    """text/__init__.py:50-56 (_clean_text): run the named cleaners in order; an unknown name raises like the reference."""
    for name in cleaner_names:
        if name not in CLEANERS:
            raise Exception("Unknown cleaner: %s" % name)
        text = CLEANERS[name](text) if name == "basic_cleaners" else CLEANERS[name](text, g2p)
    return text


# ------------------------------------------------------------------------------------------------ numbers.py
_ONES = ["zero", "one", "two", "three", "four", "five", "six", "seven", "eight", "nine", "ten", "eleven", "twelve",
         "thirteen", "fourteen", "fifteen", "sixteen", "seventeen", "eighteen", "nineteen"]
_TENS = ["", "", "twenty", "thirty", "forty", "fifty", "sixty", "seventy", "eighty", "ninety"]
_SCALES = ["", " thousand", " million", " billion", " trillion", " quadrillion", " quintillion"]
_ORDINAL_IRREGULAR = {"one": "first", "two": "second", "three": "third", "five": "fifth", "eight": "eighth", "nine": "ninth",
                      "twelve": "twelfth"}


def _below_100(n: int, zero: str = "zero") -> str:
    if n < 20:
        return zero if n == 0 else _ONES[n]
    return _TENS[n // 10] + ("-" + _ONES[n % 10] if n % 10 else "")


def _below_1000(n: int, andword: str) -> str:
    h, r = divmod(n, 100)
    if h == 0:
        return _below_100(r)
    if r == 0:
        return _ONES[h] + " hundred"
    return _ONES[h] + " hundred " + (andword + " " if andword else "") + _below_100(r)


def number_to_words(num: int, andword: str = "and", zero: str = "zero", group: int = 0) -> str:
    """The subset of inflect.engine().number_to_words numbers.py relies on: cardinal spelling with comma-separated
    thousands groups ("one thousand, two hundred thirty-four"), and group=2 pair reading ("nineteen oh six")."""
    num = int(num)
    if group == 2:
        digits = str(num)
        parts = []
        for i in range(0, len(digits), 2):
            pair = digits[i: i + 2]
            if len(pair) == 1:
                parts.append(zero if pair == "0" else _ONES[int(pair)])
            elif pair[0] == "0":
                parts.append(f"{zero} {zero if pair[1] == '0' else _ONES[int(pair[1])]}")
            else:
                parts.append(_below_100(int(pair)))
        return ", ".join(parts)
    if num == 0:
        return zero
    groups, n = [], num
    while n:
        n, r = divmod(n, 1000)
        groups.append(r)
    words = []
    for i in range(len(groups) - 1, -1, -1):
        if groups[i]:
            # inflect puts the and-word inside a group ("one hundred and five") and also before a final group below 100
            # ("one thousand and five"); numbers.py always passes andword=""
            text = _below_1000(groups[i], andword)
            if andword and i == 0 and len(groups) > 1 and groups[0] < 100:
                text = andword + " " + text
            words.append(text + _SCALES[i])
    out = ", ".join(words)
    return out.replace(", " + andword + " ", " " + andword + " ") if andword else out


def ordinal_words(num: int) -> str:
    """inflect's number_to_words("21st") -> "twenty-first": the cardinal with its last word turned ordinal."""
    card = number_to_words(num, andword="and")
    head, sep, last = card.rpartition(" ")
    pre, dash, tail = last.rpartition("-")
    if tail in _ORDINAL_IRREGULAR:
        tail = _ORDINAL_IRREGULAR[tail]
    elif tail.endswith("y"):
        tail = tail[:-1] + "ieth"
    else:
        tail = tail + "th"
    return head + sep + pre + dash + tail


_comma_number = re.compile(r"([0-9][0-9\,]+[0-9])")
_decimal_number = re.compile(r"([0-9]+\.[0-9]+)")
_pounds = re.compile(r"£([0-9\,]*[0-9]+)")
_dollars = re.compile(r"\$([0-9\.\,]*[0-9]+)")
_ordinal = re.compile(r"[0-9]+(st|nd|rd|th)")
_number = re.compile(r"[0-9]+")


def _spell_dollars(m) -> str:
    amount = m.group(1)
    pieces = amount.split(".")
    if len(pieces) > 2:
        return amount + " dollars"                      # numbers.py:27-28: something like 1.2.3 is left alone
    whole = int(pieces[0]) if pieces[0] else 0
    cents = int(pieces[1]) if len(pieces) > 1 and pieces[1] else 0
    said = []
    if whole:
        said.append("%d %s" % (whole, "dollar" if whole == 1 else "dollars"))
    if cents:
        said.append("%d %s" % (cents, "cent" if cents == 1 else "cents"))
    return ", ".join(said) if said else "zero dollars"


def _spell_number(m) -> str:
    n = int(m.group(0))
    if 1000 < n < 3000:                                  # read as a year (numbers.py:49-58)
        if n == 2000:
            return "two thousand"
        if 2000 < n < 2010:
            return "two thousand " + number_to_words(n % 100)
        if n % 100 == 0:
            return number_to_words(n // 100) + " hundred"
        return number_to_words(n, andword="", zero="oh", group=2).replace(", ", " ")
    return number_to_words(n, andword="")


def normalize_numbers(text: str) -> str:
    """numbers.py:63-70, same order of passes."""
    text = _comma_number.sub(lambda m: m.group(1).replace(",", ""), text)
    text = _pounds.sub(r"\1 pounds", text)
    text = _dollars.sub(_spell_dollars, text)
    text = _decimal_number.sub(lambda m: m.group(1).replace(".", " point "), text)
    text = _ordinal.sub(lambda m: ordinal_words(int(re.match(r"[0-9]+", m.group(0)).group(0))), text)
    text = _number.sub(_spell_number, text)
    return text
