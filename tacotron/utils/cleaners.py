"""Text cleaners (reference tacotron/utils/cleaners.py:1-91). `english_cleaners` follows the reference's pipeline: ASCII folding, number
expansion (tacotron/utils/numbers.py), abbreviation expansion, whitespace collapse - and, like the reference (cleaners.py:87: the
lowercase step is commented out), it KEEPS the case: the symbol table has both cases. ASCII folding: the reference calls `unidecode`,
which is not installable here; accented Latin letters are folded through Unicode NFKD decomposition, the Latin-1 / typographic
characters NFKD has no decomposition for go through the small table below (unidecode's renderings, e.g. the pound sign becomes "PS" -
so, as in the reference, the "pounds" rule of the number normaliser only ever sees text that was not folded), anything else is
dropped (unidecode would transliterate it)."""
import re
import unicodedata

from .numbers import normalize_numbers

_whitespace_re = re.compile(r"\s+")
_abbreviations = [(re.compile(r"\b%s\." % a, re.IGNORECASE), b) for a, b in [
    ("mrs", "misess"), ("mr", "mister"), ("dr", "doctor"), ("st", "saint"), ("co", "company"), ("jr", "junior"), ("maj", "major"),
    ("gen", "general"), ("drs", "doctors"), ("rev", "reverend"), ("lt", "lieutenant"), ("hon", "honorable"), ("sgt", "sergeant"),
    ("capt", "captain"), ("esq", "esquire"), ("ltd", "limited"), ("col", "colonel"), ("ft", "fort")]]


def expand_abbreviations(text):
    for regex, repl in _abbreviations:
        text = regex.sub(repl, text)
    return text


def lowercase(text):
    return text.lower()


def collapse_whitespace(text):
    return _whitespace_re.sub(" ", text)


def basic_cleaners(text):
    return collapse_whitespace(lowercase(text))


_fold = {"£": "PS", "¥": "Y=", "¢": "C/", "€": "EUR", "©": "(c)", "®": "(r)", "«": "<<", "»": ">>", "°": "deg", "±": "+-", "×": "x",
         "÷": "/", "ß": "ss", "æ": "ae", "Æ": "AE", "œ": "oe", "Œ": "OE", "ø": "o", "Ø": "O", "đ": "d", "Đ": "D", "ł": "l", "Ł": "L",
         "þ": "th", "Þ": "Th", "ð": "d", "Ð": "D", "‘": "'", "’": "'", "“": '"', "”": '"', "–": "-", "—": "--", "…": "...", "\u00a0": " "}


def convert_to_ascii(text):
    return "".join(_fold[c] if c in _fold else unicodedata.normalize("NFKD", c).encode("ascii", "ignore").decode("ascii") for c in text)


def expand_numbers(text):
    return normalize_numbers(text)


def transliteration_cleaners(text):
    return collapse_whitespace(lowercase(convert_to_ascii(text)))


def english_cleaners(text):
    return collapse_whitespace(expand_abbreviations(expand_numbers(convert_to_ascii(text))))
