"""The reference's command-line workflow end to end on a toy corpus: preprocess.py -> train.py --model Tacotron-2
(Tacotron training, GTA synthesis, WaveNet training, with the state_log hand-over) -> synthesize.py --model Tacotron-2 --mode eval.
Everything runs through the shipped CLIs in subprocesses with --hparams overrides, as a user would."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

HP = ("predict_linear=False,enc_conv_channels=256,embedding_dim=256,encoder_lstm_units=128,decoder_lstm_units=256,postnet_channels=256,"
      "prenet_layers=[128,128],attention_dim=128,tacotron_batch_size=4,tacotron_test_size=4,tacotron_test_batches=None,max_iters=60,"
      "tacotron_synthesis_batch_size=4,input_type=mulaw-quantize,quantize_channels=256,out_channels=256,layers=4,stacks=2,"
      "residual_channels=128,gate_channels=256,skip_out_channels=128,wavenet_batch_size=2,wavenet_test_size=2,wavenet_test_batches=None,"
      "max_time_steps=4400,wavenet_synthesis_batch_size=2,trim_silence=False,train_with_GTA=True")


def _run(args, cwd):
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable] + args, cwd=cwd, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, "%s\nSTDOUT:\n%s\nSTDERR:\n%s" % (" ".join(args), r.stdout[-3000:], r.stderr[-3000:])
    return r.stdout


def test_cli_workflow(tmp_path):
    from scipy.io import wavfile
    base = str(tmp_path)
    ds = os.path.join(base, "LJSpeech-1.1")
    os.makedirs(os.path.join(ds, "wavs"))
    rng = np.random.default_rng(0)
    rows = []
    for i in range(12):
        n = int(rng.integers(9000, 16000))
        t = np.arange(n) / 22050.0
        w = 0.4 * np.sin(2 * np.pi * (200 + 40 * i) * t * (1 + 0.3 * t)) + 0.02 * rng.standard_normal(n)
        wavfile.write(os.path.join(ds, "wavs", "LJ%03d.wav" % i), 22050, (w * 32767).astype(np.int16))
        rows.append("LJ%03d|Utterance number %d.|utterance number %s, a test sentence." % (i, i, "one two three four five".split()[i % 5]))
    open(os.path.join(ds, "metadata.csv"), "w").write("\n".join(rows) + "\n")

    out = _run([os.path.join(ROOT, "preprocess.py"), "--base_dir", base, "--hparams", HP], base)
    assert "Write 12 utterances" in out
    td = os.path.join(base, "training_data")
    meta = [l.strip().split("|") for l in open(os.path.join(td, "train.txt"))]
    assert len(meta) == 12
    for m in meta[:3]:
        a, mel, lin = (np.load(os.path.join(td, d, f)) for d, f in (("audio", m[0]), ("mels", m[1]), ("linear", m[2])))
        assert a.dtype == np.int16 and mel.dtype == np.float32 and mel.shape == (int(m[4]), 80) and lin.shape == (int(m[4]), 1025)
        assert len(a) == int(m[3]) == int(m[4]) * 275 and a.min() >= 0 and a.max() <= 255 and np.abs(mel).max() <= 4.0

    common = ["--base_dir", base, "--hparams", HP, "--name", "toy", "--input_dir", td, "--checkpoint_interval", "6", "--eval_interval", "6"]
    out = _run([os.path.join(ROOT, "train.py"), "--model", "Tacotron-2", "--tacotron_train_steps", "6", "--wavenet_train_steps", "6"] + common, base)
    log_dir = os.path.join(base, "logs-toy")
    assert open(os.path.join(log_dir, "state_log")).read().startswith("1|1|1|")
    assert os.path.isfile(os.path.join(log_dir, "taco_pretrained", "tacotron_model.ckpt-6.npz"))
    assert os.path.isfile(os.path.join(log_dir, "wave_pretrained", "wavenet_model.ckpt-6.npz"))
    gta_map = [l.strip().split("|") for l in open(os.path.join(base, "tacotron_output", "gta", "map.txt"))]
    assert len(gta_map) == 12 and all(os.path.isfile(os.path.join(base, r[2])) for r in gta_map)     # paths relative to the run directory, as in the reference
    g = np.load(os.path.join(base, gta_map[0][2]))
    assert g.shape == np.load(gta_map[0][1]).shape                      # GTA mels are frame-aligned with the ground truth
    # restart: everything is marked done, a second invocation resumes nothing and says so
    out = _run([os.path.join(ROOT, "train.py"), "--model", "Tacotron-2", "--tacotron_train_steps", "6", "--wavenet_train_steps", "6"] + common, base)
    assert "TRAINING IS ALREADY COMPLETE" in out
    # resume of a single model from its checkpoint continues the step counter
    out = _run([os.path.join(ROOT, "train.py"), "--model", "Tacotron", "--tacotron_train_steps", "8"] + common, base)
    assert "Loading checkpoint" in out and os.path.isfile(os.path.join(log_dir, "taco_pretrained", "tacotron_model.ckpt-8.npz"))

    txt = os.path.join(base, "sentences.txt")
    open(txt, "w").write("A short test.\nAnother one, please.\n")
    _run([os.path.join(ROOT, "synthesize.py"), "--model", "Tacotron-2", "--mode", "eval", "--name", "toy", "--hparams", HP, "--text_list", txt,
          "--mels_dir", "tacotron_output/eval/"], base)
    wavs = [f for f in os.listdir(os.path.join(base, "wavenet_output", "wavs")) if f.endswith(".wav")]
    assert len(wavs) == 2
    for w in wavs:      # whole hops; a row whose stop token fires on the first frame has length 0 (synthesizer.py:254-257 keeps the INDEX)
        rate, data = wavfile.read(os.path.join(base, "wavenet_output", "wavs", w))
        assert rate == 22050 and data.dtype == np.int16 and len(data) % 275 == 0


def test_cli_tacotron_with_the_default_linear_head(tmp_path):
    """`predict_linear=True` is the reference's default (hparams.py:175): train.py --model Tacotron with the CBHG post-processing net and
    linear targets from training_data/linear, then synthesize.py --mode eval writes the linear spectrogram and its Griffin-Lim inversion."""
    from scipy.io import wavfile
    base = str(tmp_path)
    ds = os.path.join(base, "LJSpeech-1.1")
    os.makedirs(os.path.join(ds, "wavs"))
    rng = np.random.default_rng(1)
    rows = []
    for i in range(8):
        n = int(rng.integers(9000, 14000))
        t = np.arange(n) / 22050.0
        w = 0.4 * np.sin(2 * np.pi * (220 + 30 * i) * t) + 0.02 * rng.standard_normal(n)
        wavfile.write(os.path.join(ds, "wavs", "LJ%03d.wav" % i), 22050, (w * 32767).astype(np.int16))
        rows.append("LJ%03d|Sentence %d.|sentence number %d of the toy corpus." % (i, i, i))
    open(os.path.join(ds, "metadata.csv"), "w").write("\n".join(rows) + "\n")
    hp = HP.replace("predict_linear=False,", "")          # the stock default: predict_linear=True
    _run([os.path.join(ROOT, "preprocess.py"), "--base_dir", base, "--hparams", hp], base)
    td = os.path.join(base, "training_data")
    common = ["--base_dir", base, "--hparams", hp, "--name", "lin", "--input_dir", td, "--checkpoint_interval", "4", "--eval_interval", "4"]
    out = _run([os.path.join(ROOT, "train.py"), "--model", "Tacotron", "--tacotron_train_steps", "4"] + common, base)
    assert os.path.isfile(os.path.join(base, "logs-lin", "taco_pretrained", "tacotron_model.ckpt-4.npz"))
    txt = os.path.join(base, "sentences.txt")
    open(txt, "w").write("A short test.\n")
    _run([os.path.join(ROOT, "synthesize.py"), "--model", "Tacotron", "--mode", "eval", "--name", "lin", "--hparams", hp, "--text_list", txt], base)
    ev = os.path.join(base, "tacotron_output", "eval")
    lin = np.load(os.path.join(ev, "linear-batch_0_sentence_0.npy"))
    mel = np.load(os.path.join(ev, "mel-batch_0_sentence_0.npy"))
    assert lin.shape == (mel.shape[0], 1025) and np.isfinite(lin).all() and np.abs(lin).max() <= 4.1
    preview = os.path.join(base, "tacotron_output", "logs-eval", "wavs", "wav-batch_0_sentence_0-linear.wav")
    if mel.shape[0] >= 2:       # previews need two frames; the length is the index of the first fired stop token (synthesizer.py:254-257)
        rate, data = wavfile.read(preview)
        assert rate == 22050 and len(data) == 275 * (mel.shape[0] - 1)
    else:
        assert not os.path.exists(preview)
