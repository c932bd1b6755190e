"""Text cleaners (reference tacotron/utils/cleaners.py). `basic_cleaners` and `transliteration_cleaners` are complete;
`english_cleaners` lower-cases, collapses whitespace and expands the reference's abbreviation list, but spells numbers
digit-by-digit is NOT attempted: the reference relies on the `inflect` / `unidecode` packages for number expansion and ASCII
folding, which are not available here (digits and non-ASCII characters are simply not in the symbol table and are dropped by
text_to_sequence, exactly as the reference drops unknown symbols)."""
import re

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


def transliteration_cleaners(text):
    return collapse_whitespace(lowercase(text.encode("ascii", "ignore").decode("ascii")))


def english_cleaners(text):
    text = text.encode("ascii", "ignore").decode("ascii")
    return collapse_whitespace(expand_abbreviations(lowercase(text)))
