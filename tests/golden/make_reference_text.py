"""Generates tests/golden/reference_text.json by EXECUTING the reference's text front-end (/root/reference/tacotron/utils/{text,cleaners,
symbols,cmudict}.py) in this container. Its two third-party imports are absent from the image, so they are stubbed and the fixture
strings are chosen so that neither stub is ever exercised: `unidecode` (identity stub - all strings are ASCII) and `inflect` (a stub that
raises if called - no string holds a digit). What is pinned: the cleaner pipelines on ASCII text (abbreviations, case handling,
whitespace), the `{ARPAbet}` cutting rule, dropped symbols, EOS, and the CMUDict parser. NOT in the fixture: `_` and `~` inside the text.
The reference filters them with `s is not '_'` (text.py:75), an identity test on str literals whose outcome depends on the interpreter
(CPython 3.12 here keeps them and warns; the reference's target 3.6 shares one-character strings and drops them); the product implements
the stated intent - drop - and tests/test_feeders_cpu.py::test_text_front_end checks that by hand.   Run:  python tests/golden/make_reference_text.py"""
import io
import json
import os
import sys
import types

REF = "/root/reference"

TEXTS = [
    "Hello,  World!",
    "Turn left on {HH AW1 S S T AH0 N} Street.",
    "{AY1} think;   therefore {AY1 AE1 M}",
    "Dr. Smith met Mrs. Jones, Lt. Brown and Col. Mustard at St. Ives Co. Ltd.",
    "tabs\tand\nnewlines   collapse",
    "underscore tilde [brackets] #hash <angle> are dropped?",
    "Quotes \"double\" and 'single' (parens): colon; dash - done.",
    "nested {K AE1 T} {D AO1 G} braces } stray",
    "   leading and trailing   ",
    "",
]
CLEANERS = [["english_cleaners"], ["basic_cleaners"], ["transliteration_cleaners"]]
DICT = """;;; comment line
'BOUT  B AW1 T
ABANDON  AH0 B AE1 N D AH0 N
READ  R EH1 D
READ(1)  R IY1 D
BOGUS  B XX1 G
lowercase  L OW1
TOMATO  T AH0 M EY1 T OW2
TOMATO(1)  T AH0 M AA1 T OW2
"""


def main():
    def _no_inflect(*a, **k):
        raise RuntimeError("fixture strings must not contain numbers")
    inflect = types.ModuleType("inflect")
    inflect.engine = lambda: types.SimpleNamespace(number_to_words=_no_inflect)
    unidecode = types.ModuleType("unidecode")
    unidecode.unidecode = lambda s: s
    sys.modules["inflect"], sys.modules["unidecode"] = inflect, unidecode
    sys.path.insert(0, REF)
    from tacotron.utils import cmudict, text
    from tacotron.utils.symbols import symbols
    out = {"symbols": symbols, "valid_symbols": cmudict.valid_symbols, "cases": []}
    for t in TEXTS:
        assert t.isascii()            # digits appear only inside {phones}, which never reach the cleaners (the inflect stub raises if they do)
        for cl in CLEANERS:
            seq = text.text_to_sequence(t, cl)
            out["cases"].append({"text": t, "cleaners": cl, "sequence": seq, "round_trip": text.sequence_to_text(seq)})
    for keep in (True, False):
        d = cmudict.CMUDict(io.StringIO(DICT), keep_ambiguous=keep)
        out["cmudict_keep_%d" % keep] = {"len": len(d), "lookups": {w: d.lookup(w) for w in
                                         ["'bout", "abandon", "read", "bogus", "lowercase", "tomato", "missing"]}}
    out["dict_text"] = DICT
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_text.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote %d cases" % len(out["cases"]))


if __name__ == "__main__":
    main()
