"""CMU pronouncing dictionary reader (reference tacotron/utils/cmudict.py:1-63): `CMUDict(path_or_file).lookup(word)` -> list of ARPAbet
pronunciations (or None). Entry lines start with a capital letter or an apostrophe, word and phones are separated by two spaces,
alternative pronunciations carry a `(n)` suffix on the word; an entry with a phone outside `valid_symbols` is skipped."""
import re

_vowels = "AA AE AH AO AW AY EH ER EY IH IY OW OY UH UW".split()
_consonants = "B CH D DH F G HH JH K L M N NG P R S SH T TH V W Y Z ZH".split()
# bare vowel + its three stress-marked forms, sorted the way the dictionary lists its phone set
valid_symbols = sorted([v + s for v in _vowels for s in ("", "0", "1", "2")] + _consonants)
_valid = frozenset(valid_symbols)
_alternative = re.compile(r"\([0-9]+\)")


class CMUDict(object):
    def __init__(self, file_or_path, keep_ambiguous=True):
        if isinstance(file_or_path, str):
            with open(file_or_path, encoding="latin-1") as f:
                entries = _parse(f)
        else:
            entries = _parse(file_or_path)
        if not keep_ambiguous:
            entries = {w: p for w, p in entries.items() if len(p) == 1}
        self._entries = entries

    def __len__(self):
        return len(self._entries)

    def lookup(self, word):
        return self._entries.get(word.upper())


def _parse(lines):
    entries = {}
    for line in lines:
        if not line or not ("A" <= line[0] <= "Z" or line[0] == "'"):
            continue
        fields = line.split("  ")
        if len(fields) < 2:
            continue
        phones = fields[1].strip().split(" ")
        if all(p in _valid for p in phones):
            entries.setdefault(_alternative.sub("", fields[0]), []).append(" ".join(phones))
    return entries
