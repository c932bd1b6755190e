"""Alignment / spectrogram plots of the run loops (reference tacotron/utils/plot.py:16-82: plot_alignment, plot_spectrogram; and
wavenet_vocoder/util.py:174-196 waveplot). The reference draws with matplotlib, which this image does not have; these functions keep the
reference's names and arguments and write the PNG themselves: colour-mapped heat maps with a colour bar (nearest-neighbour scaled to the
reference's figure size), panels stacked top to bottom, no axis text - the title goes into the PNG's `Title` text chunk."""
import struct
import zlib

import numpy as np

# five anchors of a viridis-like map, linearly interpolated
_anchors = np.array([[68, 1, 84], [59, 82, 139], [33, 145, 140], [94, 201, 98], [253, 231, 37]], dtype=np.float64)


def split_title_line(title_text, max_words=5):
    words = title_text.split()
    return "\n".join(" ".join(words[i:i + max_words]) for i in range(0, len(words), max_words))


def _colormap(v):
    """v in [0, 1] (any shape) -> uint8 RGB"""
    x = np.clip(np.nan_to_num(v), 0.0, 1.0) * (len(_anchors) - 1)
    lo = np.minimum(x.astype(np.int64), len(_anchors) - 2)
    f = (x - lo)[..., None]
    return np.round(_anchors[lo] * (1.0 - f) + _anchors[lo + 1] * f).astype(np.uint8)


def _resize(img, height, width):
    """nearest-neighbour ('interpolation=none') scaling of a [h, w] array"""
    h, w = img.shape
    rows = np.minimum((np.arange(height) * h) // height, h - 1)
    cols = np.minimum((np.arange(width) * w) // width, w - 1)
    return img[rows][:, cols]


def _panel(data, height, width, bar=24, gap=8):
    """heat map of a 2-D array (row 0 at the top) + a vertical colour bar -> uint8 [height, width, 3]"""
    data = np.asarray(data, dtype=np.float64)
    if data.ndim != 2 or data.size == 0:
        raise ValueError("plot needs a non-empty 2-D array, got shape %s" % (data.shape,))
    lo, hi = float(np.nanmin(data)), float(np.nanmax(data))
    norm = (data - lo) / (hi - lo) if hi > lo else np.zeros_like(data)
    out = np.full((height, width, 3), 255, dtype=np.uint8)
    out[:, :width - bar - gap] = _colormap(_resize(norm, height, width - bar - gap))
    out[:, width - bar:] = _colormap(np.linspace(1.0, 0.0, height))[:, None, :]
    return out


def _chunk(tag, payload):
    return struct.pack(">I", len(payload)) + tag + payload + struct.pack(">I", zlib.crc32(tag + payload) & 0xFFFFFFFF)


def write_png(path, rgb, title=None):
    """uint8 [h, w, 3] -> 8-bit truecolour PNG"""
    rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
    h, w, _ = rgb.shape
    raw = np.concatenate([np.zeros((h, 1), dtype=np.uint8), rgb.reshape(h, w * 3)], axis=1).tobytes()     # filter type 0 on every row
    blob = b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0))
    if title:
        blob += _chunk(b"tEXt", b"Title\x00" + title.encode("latin-1", "replace"))
    blob += _chunk(b"IDAT", zlib.compress(raw, 6)) + _chunk(b"IEND", b"")
    with open(path, "wb") as f:
        f.write(blob)


def _stack(panels, margin=10):
    w = max(p.shape[1] for p in panels) + 2 * margin
    h = sum(p.shape[0] for p in panels) + margin * (len(panels) + 1)
    fig = np.full((h, w, 3), 255, dtype=np.uint8)
    y = margin
    for p in panels:
        fig[y:y + p.shape[0], margin:margin + p.shape[1]] = p
        y += p.shape[0] + margin
    return fig


def plot_alignment(alignment, path, title=None, split_title=False, max_len=None):
    """alignment: [encoder steps, decoder steps]; encoder step 0 at the bottom (origin='lower'), decoder time left to right"""
    alignment = np.asarray(alignment)
    if max_len is not None:
        alignment = alignment[:, :max_len]
    if split_title and title:
        title = split_title_line(title)
    write_png(path, _stack([_panel(alignment[::-1], 440, 760)]), title)


def plot_spectrogram(pred_spectrogram, path, title=None, split_title=False, target_spectrogram=None, max_len=None, auto_aspect=False):
    """spectrograms: [frames, channels]; drawn with time left to right and channel 0 at the bottom (np.rot90 of the reference); the
    target (when given) goes above the prediction. auto_aspect stretches to the panel width, otherwise the pixels stay square-ish."""
    pred_spectrogram = np.asarray(pred_spectrogram)
    if target_spectrogram is not None:
        target_spectrogram = np.asarray(target_spectrogram)
    if max_len is not None:
        pred_spectrogram = pred_spectrogram[:max_len]
        if target_spectrogram is not None:
            target_spectrogram = target_spectrogram[:max_len]
    if split_title and title:
        title = split_title_line(title)
    panels = []
    for s in (target_spectrogram, pred_spectrogram):
        if s is None:
            continue
        img = np.rot90(s)
        width = 960
        height = 220 if auto_aspect else int(min(320, max(40, round((width - 32) * img.shape[0] / max(img.shape[1], 1)))))
        panels.append(_panel(img, height, width))
    write_png(path, _stack(panels), title)


def waveplot(path, y_hat, y_target, hparams, title=None):
    """wavenet_vocoder/util.py:174-196: target (when given) above the prediction, each drawn as its min / max envelope per pixel column"""
    panels = []
    for y in (y_target, y_hat):
        if y is None:
            continue
        y = np.asarray(y, dtype=np.float64).reshape(-1)
        width, height = 1160, 120
        img = np.full((height, width, 3), 255, dtype=np.uint8)
        if y.size:
            edges = np.linspace(0, y.size, width + 1).astype(np.int64)
            scale = max(1.0, float(np.abs(y).max()))
            for x in range(width):
                seg = y[edges[x]:max(edges[x + 1], edges[x] + 1)]
                if seg.size == 0:
                    continue
                top = int(round((1.0 - seg.max() / scale) * 0.5 * (height - 1)))
                bot = int(round((1.0 - seg.min() / scale) * 0.5 * (height - 1)))
                img[top:bot + 1, x] = (31, 119, 180)
        panels.append(img)
    write_png(path, _stack(panels), title)
