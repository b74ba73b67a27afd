"""Row f3 of SURVEY.md 8f: the rule-based half of the text front-end (`Matcha-TTS/matcha/text/cleaners.py`, `numbers.py`).

Pinning: where /root/reference is present the unmodified cleaners.py is imported where it lies, with its three absent
third-party imports (phonemizer / unidecode / misaki) replaced by stubs whose g2p is the identity, and every pipeline of
`text_cleaners` must return the same string.  A committed fixture (tests/golden/text_cleaners.json, written by that very
comparison when run with EV_WRITE_GOLDEN=1) carries the reference's outputs to boxes without the reference."""
import importlib.util
import json
import os
import sys
import types

import pytest

from emojivoice_b200 import text_cleaners as tc
from oracle import reference_shim as shim

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "text_cleaners.json")

CASES = {
    "en": ["Dr. Smith paid $5.45 for it... Mr. Jones didn't.", "Visit www.example.com at 3.14 o'clock, St. John!", "It costs €20 or ¥300.50, Mrs. Brown",
           "Hello   World\tagain\n", "Capt. Kirk, Lt. Dan and Sgt. Pepper met Gen. Lee, Esq. at Ft. Knox Co. Ltd.", "a.b c.d 1.2 x.5 5.x", "Hon. Rev. Maj. Col. Jr. Drs."],
    "fr": ["M. Dupont (le Dr. Martin) a payé 5.45€ ... Mme Curie et Mlle Dupuis", "3,5 = a/b -4 St. Denis", "Prix: 10.50$ ou 300.20¥ et 20€"],
    "de": ["Hr. Müller und Fr. Schmidt z.B. ca. 5,5 € usw. (vgl. Dr. Prof. Bsp.)", "10.50$ = 9.20€ / -3 d.h. u.a. bzw.", "Mme Mlle ¥ 1.5¥"],
    "ja": ["価格は$5です。3.14 と -5 と 50% a@b.c \\\\ x/y 1+1=2 €10 ¥20"],
}
PIPELINES = {"en": "english_cleaners2", "fr": "french_cleaners", "de": "german_cleaners", "ja": "japanese_cleaners"}


def _load_reference_cleaners():
    """cleaners.py as the reference ships it, imported from its own path; the g2p back ends are identity stubs."""
    class _Backend:
        def __init__(self, *a, **k):
            pass

        def phonemize(self, texts, strip=True, njobs=1):
            return list(texts)

    ph = types.ModuleType("phonemizer")
    ph.backend = types.SimpleNamespace(EspeakBackend=_Backend)
    un = types.ModuleType("unidecode")
    un.unidecode = lambda t: t
    mi = types.ModuleType("misaki")
    mi.ja = types.SimpleNamespace(JAG2P=lambda: (lambda t: (t, None)))
    saved = {k: sys.modules.get(k) for k in ("phonemizer", "unidecode", "misaki")}
    sys.modules.update({"phonemizer": ph, "unidecode": un, "misaki": mi})
    try:
        path = os.path.join(shim.REFERENCE_ROOT, "Matcha-TTS", "matcha", "text", "cleaners.py")
        spec = importlib.util.spec_from_file_location("_ref_cleaners", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


def _ours(lang, text):
    g2p = (lambda t: (t, None)[0]) if lang == "ja" else (lambda t: t)
    return getattr(tc, PIPELINES[lang])(text, g2p)


@pytest.mark.skipif(not shim.available(), reason="/root/reference not present")
def test_pipelines_equal_the_reference_file():
    ref = _load_reference_cleaners()
    golden = {}
    for lang, texts in CASES.items():
        for text in texts:
            want = getattr(ref, PIPELINES[lang])(text)
            assert _ours(lang, text) == want, (lang, text)
            assert tc.expand_abbreviations(text.lower(), lang) == ref.expand_abbreviations(text.lower(), lang) if lang != "ja" else True
            assert tc.apply_replacements(text, lang) == ref.apply_replacements(text, lang)
            golden.setdefault(lang, []).append([text, want])
        assert tc.basic_cleaners(texts[0]) == ref.basic_cleaners(texts[0])
    with pytest.raises(UnboundLocalError):           # the reference's Spanish pipeline looks up tables that do not exist
        ref.spanish_cleaners("hola Sr. Perez")
    with pytest.raises(UnboundLocalError):
        tc.spanish_cleaners("hola Sr. Perez", lambda t: t)
    if os.environ.get("EV_WRITE_GOLDEN"):
        with open(GOLDEN, "w", encoding="utf-8") as f:
            json.dump(golden, f, ensure_ascii=False, indent=1)


def test_pipelines_equal_the_committed_reference_outputs():
    with open(GOLDEN, encoding="utf-8") as f:
        golden = json.load(f)
    assert sorted(golden) == sorted(CASES)
    for lang, pairs in golden.items():
        assert [p[0] for p in pairs] == CASES[lang]
        for text, want in pairs:
            assert _ours(lang, text) == want, (lang, text)


def test_clean_text_dispatch_and_g2p_hook():
    calls = []
    out = tc.clean_text("Dr.  Who", ["english_cleaners2"], lambda t: calls.append(t) or "dˈɑktɚ  hˈuː")
    assert calls == ["doctor  who"] and out == "dˈɑktɚ hˈuː"       # rules first, g2p, then the whitespace collapse (cleaners.py:248-257)
    assert tc.clean_text("A  B", ["basic_cleaners"]) == "a b"
    with pytest.raises(Exception):
        tc.clean_text("x", ["no_such_cleaner"])


def test_normalize_numbers_known_answers():
    """numbers.py needs the `inflect` package, absent offline: hand-checked vectors of its documented behaviour."""
    n = tc.normalize_numbers
    assert n("I have 3 cats") == "I have three cats"
    assert n("12,345 people") == "twelve thousand, three hundred forty-five people"
    assert n("1,234 people") == "twelve thirty-four people"          # 1000 < n < 3000 is read as a year (numbers.py:49-58)
    assert n("In 1984 and 1906 and 2000 and 2005 and 1900") == "In nineteen eighty-four and nineteen oh six and two thousand and two thousand five and nineteen hundred"
    assert n("$1.50 and $2 and $0.01 and $1") == "one dollar, fifty cents and two dollars and one cent and one dollar"
    assert n("£20") == "twenty pounds"
    assert n("3.14") == "three point fourteen"
    assert n("the 1st, 2nd, 3rd, 21st, 100th and 12th") == "the first, second, third, twenty-first, one hundredth and twelfth"
    assert n("1000000") == "one million"
    assert tc.number_to_words(105, andword="and") == "one hundred and five"
    assert tc.number_to_words(3000) == "three thousand"
