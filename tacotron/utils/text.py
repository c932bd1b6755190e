"""text <-> id sequences (reference tacotron/utils/text.py:14-75). The text is cut at `{...}` groups: what is outside goes through the
cleaners and is mapped character by character; what is inside is read as space-separated ARPAbet phones, looked up as `@PHONE`. Symbols
that are not in the table (with the default table that is every ARPAbet phone, symbols.py:14-17), `_` and `~` are dropped; the
end-of-sequence id 1 is appended. sequence_to_text puts the braces back around runs of phones."""
import re

from . import cleaners
from .symbols import symbols

_symbol_to_id = {s: i for i, s in enumerate(symbols)}
_id_to_symbol = {i: s for i, s in enumerate(symbols)}
_brace_group = re.compile(r"(.*?)\{(.+?)\}(.*)")


def _clean_text(text, cleaner_names):
    for name in cleaner_names:
        fn = getattr(cleaners, name, None)
        if fn is None:
            raise Exception("Unknown cleaner: %s" % name)
        text = fn(text)
    return text


def _ids(syms):
    return [_symbol_to_id[s] for s in syms if s in _symbol_to_id and s not in ("_", "~")]


def text_to_sequence(text, cleaner_names):
    seq = []
    while text:
        m = _brace_group.match(text)
        if m is None:
            seq += _ids(_clean_text(text, cleaner_names))
            break
        head, phones, text = m.groups()
        seq += _ids(_clean_text(head, cleaner_names))
        seq += _ids("@" + p for p in phones.split())
    seq.append(_symbol_to_id["~"])
    return seq


def sequence_to_text(sequence):
    out = []
    for i in sequence:
        s = _id_to_symbol.get(i)
        if s is None:
            continue
        out.append("{%s}" % s[1:] if len(s) > 1 and s[0] == "@" else s)
    return "".join(out).replace("}{", " ")
