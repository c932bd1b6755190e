"""Symbol table of the text front-end (reference tacotron/utils/symbols.py:9-17): index 0 is the padding symbol `_`, index 1 the
end-of-sequence symbol `~`, then the 52 ASCII letters and 12 punctuation / space characters = 66 embedding rows."""
_pad = "_"
_eos = "~"
_characters = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz!'\"(),-.:;? "

symbols = [_pad, _eos] + list(_characters)
