"""tacotron/utils/plot.py writes its PNGs without matplotlib: decode them back with zlib and check geometry, orientation and colours."""
import struct
import zlib

import numpy as np

from tacotron.utils import plot


def _read_png(path):
    blob = open(path, "rb").read()
    assert blob[:8] == b"\x89PNG\r\n\x1a\n"
    pos, chunks = 8, []
    while pos < len(blob):
        n, tag = struct.unpack(">I", blob[pos:pos + 4])[0], blob[pos + 4:pos + 8]
        body = blob[pos + 8:pos + 8 + n]
        assert struct.unpack(">I", blob[pos + 8 + n:pos + 12 + n])[0] == zlib.crc32(tag + body) & 0xFFFFFFFF
        chunks.append((tag, body))
        pos += 12 + n
    assert chunks[0][0] == b"IHDR" and chunks[-1][0] == b"IEND"
    w, h, depth, ctype = struct.unpack(">IIBB", chunks[0][1][:10])
    assert (depth, ctype) == (8, 2)
    raw = np.frombuffer(zlib.decompress(b"".join(b for t, b in chunks if t == b"IDAT")), dtype=np.uint8).reshape(h, 1 + 3 * w)
    assert (raw[:, 0] == 0).all()
    text = dict(b.split(b"\x00", 1) for t, b in chunks if t == b"tEXt")
    return raw[:, 1:].reshape(h, w, 3), text


def test_alignment_plot_orientation(tmp_path):
    T_in, T_out = 12, 40
    a = np.zeros((T_in, T_out), dtype=np.float32)
    a[np.minimum(np.arange(T_out) * T_in // T_out, T_in - 1), np.arange(T_out)] = 1.0      # monotonic diagonal
    p = str(tmp_path / "a.png")
    plot.plot_alignment(a, p, title="a b c d e f g", split_title=True, max_len=T_out)
    img, text = _read_png(p)
    assert text[b"Title"] == b"a b c d e\nf g"
    hot, cold = plot._colormap(np.array(1.0)), plot._colormap(np.array(0.0))
    assert (img[-15, 15] == hot).all() and (img[15, 15] == cold).all()          # encoder step 0 / decoder step 0: bottom-left
    assert (img[15, 700] == hot).all() and (img[-15, 700] == cold).all()        # last encoder step at the last decoder steps: top-right


def test_spectrogram_and_wave_plots(tmp_path):
    rng = np.random.default_rng(0)
    pred, target = rng.standard_normal((90, 80)), rng.standard_normal((120, 80))
    p1, p2 = str(tmp_path / "s1.png"), str(tmp_path / "s2.png")
    plot.plot_spectrogram(pred, p1, title="pred only")
    plot.plot_spectrogram(pred, p2, title="both", target_spectrogram=target, max_len=100, auto_aspect=True)
    a, _ = _read_png(p1)
    b, _ = _read_png(p2)
    assert b.shape[1] == a.shape[1] and b.shape[0] == 3 * 10 + 2 * 220
    ramp = np.tile(np.arange(80, dtype=np.float64), (50, 1))                                  # channel index as the value
    plot.plot_spectrogram(ramp, p1)
    a, _ = _read_png(p1)
    assert (a[-11, 100] == plot._colormap(np.array(0.0))).all() and (a[11, 100] == plot._colormap(np.array(1.0))).all()   # channel 0 at the bottom
    w = str(tmp_path / "w.png")
    plot.waveplot(w, np.sin(np.arange(22050) * 0.01), np.zeros(1000), None, title="wave")
    c, text = _read_png(w)
    assert c.shape == (3 * 10 + 2 * 120, 1160 + 20, 3) and text[b"Title"] == b"wave"
    blue = np.array([31, 119, 180])
    assert (c[10 + 59:10 + 61, 500] == blue).any() and (c[10 + 5, 500] == 255).all()                       # silent target: a centre line only
    assert (c[140 + 5, 10:1170] == blue).all(axis=-1).any()                                                # the sine reaches the panel top
