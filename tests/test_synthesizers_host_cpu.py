"""Host-side flow of the two Synthesizer classes with the GPU model replaced by a stand-in: what gets cut, clipped, named, written
and plotted (reference tacotron/synthesizer.py:60-233, wavenet_vocoder/synthesizer.py:52-130). The numerics of the models are the
GPU tests' business; this keeps the glue honest on CPU."""
import os

import numpy as np
import pytest
import torch

from hparams import hparams


class _FakeTacotron(object):
    """free-running outputs for B rows, T frames: row b fires its stop token at frame 3 + 2 b (never for the last row)"""

    def __init__(self, hp, T=12):
        self.hp, self.T = hp, T

    def initialize(self, inputs, lens, mel=None, gta=False, **kw):
        B, T_in = inputs.shape
        g = torch.Generator().manual_seed(int(B))
        T = mel.shape[1] if mel is not None else self.T
        self.tower_mel_outputs = [torch.randn(B, T, self.hp.num_mels, generator=g) * 3.0]            # exceeds +-4 somewhere: clip visible
        self.tower_linear_outputs = [torch.randn(B, T, self.hp.num_freq, generator=g) * 3.0]
        self.tower_alignments = [torch.softmax(torch.randn(B, T_in, T, generator=g), dim=1)]
        stop = torch.full((B, T), 0.2)
        for b in range(B - 1):
            stop[b, 3 + 2 * b:] = 0.9
        self.tower_stop_token_prediction = [stop]


def test_tacotron_synthesizer_eval_flow(tmp_path, monkeypatch):
    from datasets import audio
    from tacotron import synthesizer as ts
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    written = {}
    monkeypatch.setattr(audio, "inv_mel_spectrogram", lambda m, hp: np.zeros(275 * (m.shape[1] - 1), dtype=np.float32) + 0.1)
    monkeypatch.setattr(audio, "inv_linear_spectrogram", lambda m, hp: np.zeros(275 * (m.shape[1] - 1), dtype=np.float32) + 0.1)
    monkeypatch.setattr(audio, "save_wav", lambda wav, path, sr: written.setdefault(path, len(wav)))
    hp = hparams.copy()
    s = ts.Synthesizer()
    s._hparams, s.gta, s.model, s._pad, s._target_pad = hp, False, _FakeTacotron(hp), 0, -hp.max_abs_value
    out_dir, log_dir = str(tmp_path / "eval"), str(tmp_path / "logs-eval")
    os.makedirs(out_dir)
    texts = ["Hello there.", "A second, longer sentence {HH AH0}.", "Third."]
    names, speakers = s.synthesize(texts, ["b0", "b1", "b2"], out_dir, log_dir, None)
    assert speakers == ["<no_g>"] * 3 and [os.path.basename(n) for n in names] == ["mel-b0.npy", "mel-b1.npy", "mel-b2.npy"]
    lengths = [np.load(n).shape[0] for n in names]
    assert lengths == [3, 5, 12]                                     # index of the first fired stop; the last row never fires
    for n, b in zip(names, ("b0", "b1", "b2")):
        mel, lin = np.load(n), np.load(os.path.join(out_dir, "linear-%s.npy" % b))
        assert mel.dtype == np.float32 and lin.shape == (mel.shape[0], hp.num_freq)
        assert np.abs(mel).max() <= hp.max_abs_value and np.abs(lin).max() <= hp.max_abs_value and np.abs(mel).max() == hp.max_abs_value
        for f in ("plots/alignment-%s.png", "plots/mel-%s.png", "plots/linear-%s.png"):
            assert os.path.getsize(os.path.join(log_dir, f % b)) > 500
        assert written[os.path.join(log_dir, "wavs", "wav-%s-mel.wav" % b)] == 275 * (mel.shape[0] - 1)
        assert os.path.join(log_dir, "wavs", "wav-%s-linear.wav" % b) in written
    # GTA: teacher forced on the ground-truth mels, cut to THEIR lengths, no previews
    s.gta = True
    gt = []
    for i, n in enumerate((7, 4)):
        p = str(tmp_path / ("gt%d.npy" % i))
        np.save(p, np.zeros((n, hp.num_mels), dtype=np.float32))
        gt.append(p)
    names, _ = s.synthesize(texts[:2], ["g0", "g1"], out_dir, None, gt)
    assert [np.load(n).shape[0] for n in names] == [7, 4]


class _FakeWaveNet(object):
    def __init__(self, hop):
        self.hop = hop

    def initialize(self, y, c, g, input_lengths, **kw):
        B, Tc, _ = c.shape
        self.c = c
        t = torch.arange(Tc * self.hop, dtype=torch.float32)
        self.tower_y_hat = [torch.stack([0.5 * torch.sin(t * (0.01 + 0.01 * b)) for b in range(B)])]


def test_wavenet_synthesizer_flow(tmp_path, monkeypatch):
    from scipy.io import wavfile
    import datasets.audio as audio
    from wavenet_vocoder import synthesizer as ws
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(ws, "melspectrogram", lambda w, hp: np.zeros((hp.num_mels, 1 + len(w) // 275), dtype=np.float32))
    hp = hparams.copy()
    s = ws.Synthesizer()
    s._hparams, s.model = hp, _FakeWaveNet(275)
    mels = [np.random.default_rng(0).uniform(-5, 5, (n, hp.num_mels)).astype(np.float32) for n in (9, 12)]
    out_dir, log_dir = str(tmp_path / "wavs"), str(tmp_path / "plots")
    os.makedirs(out_dir)
    os.makedirs(log_dir)
    names = s.synthesize(mels, None, ["mel-a", "mel-b"], out_dir, log_dir)
    assert [os.path.basename(n) for n in names] == ["wavenet-audio-mel-a.wav", "wavenet-audio-mel-b.wav"]
    c = s.model.c.numpy()
    assert c.shape == (2, 12, hp.num_mels) and c.min() == 0.0 and c.max() == 1.0          # clipped to +-4, padded with -4, mapped to [0, 1]
    assert (c[0, 9:] == 0.0).all()
    for n, frames in zip(names, (9, 12)):
        rate, data = wavfile.read(n)
        assert rate == hp.sample_rate and len(data) == frames * 275 and data.dtype == np.int16 and np.abs(data).max() == 32767
    for b in ("mel-a", "mel-b"):
        assert os.path.getsize(os.path.join(log_dir, "wavenet-waveplot-%s.png" % b)) > 500
        assert os.path.getsize(os.path.join(log_dir, "wavenet-mel-spectrogram-%s.png" % b)) > 500


class _FakeTrainModel(object):
    """stands in for the drop-in Tacotron in the training loop: remembers the last batch's shapes, loss falls with the step"""
    _eng = None

    def __init__(self, hp):
        self.hp, self.calls = hp, 0

    def initialize(self, inputs, input_lengths, mel_targets, stop, linear_targets=None, targets_lengths=None, global_step=0,
                   is_training=False, is_evaluating=False):
        B, T_in = inputs.shape
        T_out = mel_targets.shape[1]
        self.tower_mel_outputs = [mel_targets + 0.1]
        self.tower_alignments = [torch.softmax(torch.randn(B, T_in, T_out), dim=1)]
        self.calls += 1

    def add_loss(self):
        self.before_loss = self.after_loss = self.stop_token_loss = torch.tensor(0.5)
        return torch.tensor(2.0 / self.calls)

    def add_optimizer(self, step):
        return 1e-3


def test_tacotron_training_loop_eval_artefacts(tmp_path, monkeypatch):
    """tacotron/train.py with the model replaced: batches flow from the real feeder, evaluation runs on the held-out batches and writes
    the .npy + .png artefacts, a checkpoint call happens at the interval"""
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    from test_feeders_cpu import _make_dataset
    import t2_checkpoint
    from tacotron import train as tt
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    hp = hparams.copy()
    hp.parse("tacotron_batch_size=4,tacotron_test_size=8,tacotron_test_batches=None")
    base = str(tmp_path)
    _make_dataset(os.path.join(base, "training_data"))
    fake = _FakeTrainModel(hp)
    saved = []
    monkeypatch.setattr(tt, "create_model", lambda name, hparams: fake)
    monkeypatch.setattr(t2_checkpoint, "save", lambda d, name, eng: saved.append(d) or os.path.join(d, name + "-x.npz"))
    monkeypatch.setattr(t2_checkpoint, "latest", lambda d: None)
    from types import SimpleNamespace as NS
    args = NS(base_dir=base, tacotron_input="training_data/train.txt", model="Tacotron", restore=True, tacotron_train_steps=4, eval_interval=2,
              checkpoint_interval=4)
    log_dir = os.path.join(base, "logs-t")
    os.makedirs(log_dir)
    assert tt.train(log_dir, args, hp) == os.path.join(log_dir, "taco_pretrained")
    ev = os.path.join(log_dir, "eval-dir")
    for step in (2, 4):
        a = np.load(os.path.join(ev, "step-%d-eval-align.npy" % step))
        m = np.load(os.path.join(ev, "step-%d-eval-mel-prediction.npy" % step))
        assert a.ndim == 2 and m.shape[1] == 80 and a.shape[1] == m.shape[0]            # alignment [T_in, T_out], mel [T_out, 80]
        assert os.path.getsize(os.path.join(ev, "step-%d-eval-align.png" % step)) > 500
        assert os.path.getsize(os.path.join(ev, "step-%d-eval-mel-spectrogram.png" % step)) > 500
    assert len(saved) == 1 and fake.calls == 4 + 2 * 2


def test_zero_length_outputs_flow_through_both_synthesizers(tmp_path, monkeypatch):
    """an untrained model can fire its stop token on frame 0: the reference's rule gives that row length 0. Files are still written
    (empty mel / linear / wav), previews are skipped, nothing crashes"""
    from scipy.io import wavfile
    from datasets import audio
    from tacotron import synthesizer as ts
    from wavenet_vocoder import synthesizer as ws
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    hp = hparams.copy()

    class Fires(_FakeTacotron):
        def initialize(self, inputs, lens, mel=None, gta=False, **kw):
            _FakeTacotron.initialize(self, inputs, lens, mel, gta, **kw)
            self.tower_stop_token_prediction[0][:] = 0.9
    s = ts.Synthesizer()
    s._hparams, s.gta, s.model, s._pad, s._target_pad = hp, False, Fires(hp), 0, -hp.max_abs_value
    out_dir, log_dir = str(tmp_path / "eval"), str(tmp_path / "logs-eval")
    os.makedirs(out_dir)
    names, _ = s.synthesize(["One.", "Two."], ["a", "b"], out_dir, log_dir, None)
    assert [np.load(n).shape for n in names] == [(0, hp.num_mels)] * 2
    assert np.load(os.path.join(out_dir, "linear-a.npy")).shape == (0, hp.num_freq) and os.listdir(os.path.join(log_dir, "wavs")) == []
    w = ws.Synthesizer()
    w._hparams, w.model = hp, None                                   # the model must not even be called
    wav_dir = str(tmp_path / "wavs")
    os.makedirs(wav_dir)
    out = w.synthesize([np.load(n) for n in names], None, ["mel-a", "mel-b"], wav_dir, str(tmp_path))
    for p in out:
        rate, data = wavfile.read(p)
        assert rate == hp.sample_rate and len(data) == 0
