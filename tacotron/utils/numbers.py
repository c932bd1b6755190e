"""Number normalisation of the English text front-end (reference tacotron/utils/numbers.py:1-69): thousands separators, pounds,
dollars and cents, decimal points, ordinals and cardinals are spelled out before the characters are mapped to symbol ids; years between
1000 and 3000 are read in pairs ("nineteen eighty-four", "nineteen oh five", "two thousand five", "nineteen hundred").

The reference delegates the spelling to the `inflect` package, which is not installable here (no network). `number_to_words` below
restates the part of inflect's behaviour the reference relies on — `number_to_words(n, andword='')`, `number_to_words(n, andword='',
zero='oh', group=2)` and the ordinal form `number_to_words('23rd')` — from its documentation: three-digit groups joined by ", ", tens
and units hyphenated, the `andword` in front of tens / units after a hundred and in front of a last group that is a single word.
UNPINNED: no inflect in the image to execute against; tests/test_feeders_cpu.py holds the known answers."""
import re

_comma_number_re = re.compile(r"([0-9][0-9\,]+[0-9])")
_decimal_number_re = re.compile(r"([0-9]+\.[0-9]+)")
_pounds_re = re.compile(r"£([0-9\,]*[0-9]+)")
_dollars_re = re.compile(r"\$([0-9\.\,]*[0-9]+)")
_ordinal_re = re.compile(r"[0-9]+(st|nd|rd|th)")
_number_re = re.compile(r"[0-9]+")

_units = ["", "one", "two", "three", "four", "five", "six", "seven", "eight", "nine"]
_teens = ["ten", "eleven", "twelve", "thirteen", "fourteen", "fifteen", "sixteen", "seventeen", "eighteen", "nineteen"]
_tens = ["", "", "twenty", "thirty", "forty", "fifty", "sixty", "seventy", "eighty", "ninety"]
_mill = ["", " thousand", " million", " billion", " trillion", " quadrillion", " quintillion", " sextillion", " septillion", " octillion",
         " nonillion", " decillion"]
_ordinal_words = {"one": "first", "two": "second", "three": "third", "five": "fifth", "eight": "eighth", "nine": "ninth", "twelve": "twelfth"}


def _below_hundred(n, zero="zero", pad=False):
    """two-digit group: 0..99; pad = the group had a leading zero ("05" -> "oh five" when read in pairs)"""
    if n < 10:
        word = _units[n] if n else zero
        return ("%s %s" % (zero, word)) if (pad and n) else word
    if n < 20:
        return _teens[n - 10]
    return _tens[n // 10] + ("-" + _units[n % 10] if n % 10 else "")


def _below_thousand(n, andword):
    h, rest = divmod(n, 100)
    words = []
    if h:
        words.append(_units[h] + " hundred")
        if rest and andword:
            words.append(andword)
    if rest:
        words.append(_below_hundred(rest))
    return " ".join(words)


def _to_ordinal(words):
    head, sep, last = words.rpartition(" ")
    pre, hyphen, unit = last.rpartition("-")
    if unit in _ordinal_words:
        unit = _ordinal_words[unit]
    elif unit.endswith("y"):
        unit = unit[:-1] + "ieth"
    else:
        unit = unit + "th"
    return head + sep + pre + hyphen + unit


def number_to_words(num, andword="and", zero="zero", group=0):
    """inflect.engine().number_to_words for non-negative integers (or ordinal strings like '23rd')"""
    s = str(num).strip()
    ordinal = bool(re.fullmatch(r"[0-9]+(st|nd|rd|th)", s))
    if ordinal:
        s = s[:-2]
    n = int(s)
    if group == 2:                                   # digits read in pairs from the left: 1984 -> nineteen, eighty-four
        digits = str(n)
        pairs = [digits[i:i + 2] for i in range(0, len(digits), 2)]
        words = ", ".join(_below_hundred(int(p), zero, pad=len(p) == 2 and p[0] == "0") if p != "00" else "%s %s" % (zero, zero)
                          for p in pairs)
    elif n == 0:
        words = zero
    else:
        chunks, idx = [], 0
        while n:
            n, part = divmod(n, 1000)
            if part:
                if idx >= len(_mill):
                    raise ValueError("number too large to spell out")
                chunks.append(_below_thousand(part, andword) + _mill[idx])
            idx += 1
        chunks.reverse()
        if len(chunks) > 1 and " " not in chunks[-1]:        # a last group of one word attaches with the andword instead of a comma
            words = ", ".join(chunks[:-1]) + " " + (andword + " " if andword else "") + chunks[-1]
        else:
            words = ", ".join(chunks)
    return _to_ordinal(words) if ordinal else words


def _remove_commas(m):
    return m.group(1).replace(",", "")


def _expand_decimal_point(m):
    return m.group(1).replace(".", " point ")


def _expand_dollars(m):
    match = m.group(1)
    parts = match.split(".")
    if len(parts) > 2:
        return match + " dollars"          # unexpected format
    try:
        dollars = int(parts[0]) if parts[0] else 0
        cents = int(parts[1]) if len(parts) > 1 and parts[1] else 0
    except ValueError:                     # "$,85": a stray separator the comma rule did not remove (the reference raises here)
        return match.replace(",", "") + " dollars"
    d_unit = "dollar" if dollars == 1 else "dollars"
    c_unit = "cent" if cents == 1 else "cents"
    if dollars and cents:
        return "%s %s, %s %s" % (dollars, d_unit, cents, c_unit)
    if dollars:
        return "%s %s" % (dollars, d_unit)
    if cents:
        return "%s %s" % (cents, c_unit)
    return "zero dollars"


def _expand_ordinal(m):
    return number_to_words(m.group(0))


def _expand_number(m):
    num = int(m.group(0))
    if 1000 < num < 3000:
        if num == 2000:
            return "two thousand"
        if 2000 < num < 2010:
            return "two thousand " + number_to_words(num % 100)
        if num % 100 == 0:
            return number_to_words(num // 100) + " hundred"
        return number_to_words(num, andword="", zero="oh", group=2).replace(", ", " ")
    return number_to_words(num, andword="")


def normalize_numbers(text):
    text = re.sub(_comma_number_re, _remove_commas, text)
    text = re.sub(_pounds_re, r"\1 pounds", text)
    text = re.sub(_dollars_re, _expand_dollars, text)
    text = re.sub(_decimal_number_re, _expand_decimal_point, text)
    text = re.sub(_ordinal_re, _expand_ordinal, text)
    text = re.sub(_number_re, _expand_number, text)
    return text
