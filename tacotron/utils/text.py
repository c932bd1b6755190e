"""text <-> id sequences (reference tacotron/utils/text.py:16-75): characters outside the symbol table, `_` and `~` are dropped;
the end-of-sequence id 1 is appended. ARPAbet (curly-brace) input is not part of the default symbol set and is not supported."""
from . import cleaners
from .symbols import symbols

_symbol_to_id = {s: i for i, s in enumerate(symbols)}
_id_to_symbol = {i: s for i, s in enumerate(symbols)}


def _clean_text(text, cleaner_names):
    for name in cleaner_names:
        fn = getattr(cleaners, name, None)
        if fn is None:
            raise Exception("Unknown cleaner: %s" % name)
        text = fn(text)
    return text


def text_to_sequence(text, cleaner_names):
    seq = [_symbol_to_id[s] for s in _clean_text(text, cleaner_names) if s in _symbol_to_id and s not in ("_", "~")]
    seq.append(_symbol_to_id["~"])
    return seq


def sequence_to_text(sequence):
    return "".join(_id_to_symbol[i] for i in sequence if i in _id_to_symbol)
