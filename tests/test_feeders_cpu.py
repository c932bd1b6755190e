"""Feeders and on-disk formats on synthetic data, no GPU: the batch contract of tacotron/feeder.py:198-256 and
wavenet_vocoder/feeder.py:295-428 (padding values, token targets, hop-aligned crops, [0, 1] conditioning, deterministic split),
plus the padding helpers against the vectors produced by executing the reference's own feeder code."""
import os

import numpy as np
import pytest

from hparams import hparams

R = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_exec.npz"))


def _make_dataset(root, n=48, hop=275, seed=0):
    rng = np.random.default_rng(seed)
    for d in ("audio", "mels", "linear"):
        os.makedirs(os.path.join(root, d), exist_ok=True)
    rows = []
    for i in range(1, n + 1):
        frames = int(rng.integers(20, 90))
        np.save(os.path.join(root, "audio", "audio-%d.npy" % i), rng.integers(0, 256, frames * hop).astype(np.int16))
        np.save(os.path.join(root, "mels", "mel-%d.npy" % i), rng.uniform(-4.5, 4.5, (frames, 80)).astype(np.float32))
        np.save(os.path.join(root, "linear", "linear-%d.npy" % i), rng.uniform(-4.5, 4.5, (frames, 1025)).astype(np.float32))
        text = "".join(rng.choice(list("abcdefghij klmnop, qrstu!"), int(rng.integers(8, 40))))
        rows.append("audio-%d.npy|mel-%d.npy|linear-%d.npy|%d|%d|%s" % (i, i, i, frames * hop, frames, text))
    path = os.path.join(root, "train.txt")
    open(path, "w", encoding="utf-8").write("\n".join(rows) + "\n")
    return path


def test_padding_helpers_match_reference_executed_vectors():
    from tacotron import feeder as tf_
    from wavenet_vocoder import feeder as wf
    assert np.array_equal(tf_.pad_input(np.arange(1, 8, dtype=np.int32), 10, 0), R["feeder_pad_input"])
    assert np.array_equal(tf_.pad_target(R["feeder_target_in"], 9, float(R["taco_target_pad"])), R["feeder_pad_target"])
    assert np.array_equal(tf_.pad_token_target(np.zeros(7, dtype=np.float32), 9, 1.0), R["feeder_pad_token_target"])
    assert [tf_._round_up(n, 3) for n in range(8)] == R["feeder_round_up"].tolist()
    assert [tf_._round_down(n, 3) for n in range(8)] == R["feeder_round_down"].tolist()
    got = [[wf._ensure_divisible(n, 275, True), wf._ensure_divisible(n, 275, False)] for n in (274, 275, 276, 8000, 12000)]
    assert got == R["wn_ensure_divisible"].tolist()


def test_text_front_end():
    from tacotron.utils.symbols import symbols
    from tacotron.utils.text import sequence_to_text, text_to_sequence
    assert len(symbols) == 66 and symbols[0] == "_" and symbols[1] == "~"
    seq = text_to_sequence("Hello,  World_~ 9!", ["english_cleaners"])
    # case is kept (the reference's english_cleaners does not lowercase, cleaners.py:87), numbers are spelled out, pad / eos symbols
    # inside the text are dropped, EOS is appended
    assert seq[-1] == 1 and 0 not in seq and sequence_to_text(seq) == "Hello, World nine!~"
    assert sequence_to_text(text_to_sequence("Hello,  World 9!", ["basic_cleaners"])) == "hello, world !~"
    from tacotron.utils.cleaners import english_cleaners
    from tacotron.utils.numbers import number_to_words
    known = {"In 1984 he paid $5.50 for 3 books.": "In nineteen eighty-four he paid five dollars, fifty cents for three books.",
             "1905, 2000, 2005, 1900, 2015, 1010": "nineteen oh five, two thousand, two thousand five, nineteen hundred, twenty fifteen, ten ten",
             "It costs £1,000 or $1.": "It costs PSone thousand or one dollar.",       # unidecode turns the pound sign into "PS" before the rule
             "“Quoted” — naïve straße…": "\"Quoted\" -- naive strasse...",
             "Pi is 3.14; the 22nd of May, 101st.": "Pi is three point fourteen; the twenty-second of May, one hundred and first.",
             "Dr. Smith and Mr. Jones at Café Noël": "doctor Smith and mister Jones at Cafe Noel",
             "12,345 and 1,000,000": "twelve thousand, three hundred forty-five and one million"}
    for text, want in known.items():
        assert english_cleaners(text) == want, (text, english_cleaners(text))
    assert number_to_words(1001, andword="") == "one thousand one" and number_to_words(1001) == "one thousand and one"
    assert number_to_words(1234) == "one thousand, two hundred and thirty-four" and number_to_words("12th") == "twelfth"


def test_tacotron_feeder_batches(tmp_path):
    from tacotron.feeder import Feeder
    hp = hparams.copy()
    hp.parse("tacotron_batch_size=4,tacotron_test_size=8,tacotron_test_batches=None")
    path = _make_dataset(str(tmp_path))
    f = Feeder(path, hp)
    assert len(f._test_meta) == 8 and len(f._train_meta) == 40 and f.test_steps == 2
    f2 = Feeder(path, hp)
    assert [m[0] for m in f2._test_meta] == [m[0] for m in f._test_meta]                      # deterministic split (random_state)
    group = f.train_group()
    assert len(group) == 32
    for b in group[:6] + f.test_batches():
        B, T_in = b["inputs"].shape
        assert B == 4 and b["mel_targets"].shape[0] == 4 and b["mel_targets"].shape[2] == 80
        To = b["mel_targets"].shape[1]
        assert To == b["targets_lengths"].max() and b["token_targets"].shape[1] == To          # (len - 1 zeros) + 1, rounded to r (feeder.py:240-243)
        for i in range(B):
            n, L = int(b["targets_lengths"][i]), int(b["input_lengths"][i])
            assert (b["inputs"][i, L:] == 0).all() and b["inputs"][i, L - 1] == 1               # zero padding after the EOS id
            assert (b["mel_targets"][i, n:] == -hp.max_abs_value).all()                         # symmetric mels: pad with -max_abs_value
            if hp.predict_linear:       # the reference default: linear targets of the post-processing net, padded like the mels
                assert b["linear_targets"].shape == (4, b["mel_targets"].shape[1], hp.num_freq)
                assert (b["linear_targets"][i, n:] == -hp.max_abs_value).all() and (b["linear_targets"][i, :n] != -hp.max_abs_value).any()
            assert (b["token_targets"][i, :n - 1] == 0).all() and (b["token_targets"][i, n - 1:] == 1).all()
    # batches of a group hold utterances of similar length (sorted in groups of 32 batches, then shuffled)
    spans = [int(b["targets_lengths"].max() - b["targets_lengths"].min()) for b in group]
    assert np.mean(spans) < 12
    r0, r1 = Feeder(path, hp, rank=0, world_size=2), Feeder(path, hp, rank=1, world_size=2)
    assert len(r0.train_group()) == 16 and len(r1.train_group()) == 16


def test_tacotron_feeder_thread(tmp_path):
    from tacotron.feeder import Feeder
    hp = hparams.copy()
    hp.parse("tacotron_batch_size=4,tacotron_test_size=8,tacotron_test_batches=None")
    f = Feeder(_make_dataset(str(tmp_path)), hp).start()
    try:
        b = [f.next_batch(timeout=60) for _ in range(3)]
        assert all(x["inputs"].dtype.is_floating_point is False and x["mel_targets"].shape[0] == 4 for x in b)
    finally:
        f.stop()


def test_wavenet_feeder_batches(tmp_path):
    from wavenet_vocoder.feeder import Feeder
    hp = hparams.copy()
    hp.parse("input_type=mulaw-quantize,quantize_channels=256,out_channels=256,wavenet_batch_size=4,wavenet_test_size=8,"
             "wavenet_test_batches=None,max_time_steps=8000,train_with_GTA=False")
    root = str(tmp_path)
    path = _make_dataset(root)
    # the preprocessor's train.txt names files relative to audio/ and mels/: the WaveNet feeder joins base_dir + row entries, which is
    # what the GTA map.txt (full paths) and wavenet_preprocess maps provide; mimic a map with relative paths
    rows = [l.strip().split("|") for l in open(path, encoding="utf-8")]
    mp = os.path.join(root, "map.txt")
    open(mp, "w").write("\n".join("audio/%s|mels/%s|mels/%s|<no_g>|%s" % (r[0], r[1], r[1], r[5]) for r in rows) + "\n")
    f = Feeder(mp, root, hp)
    assert f.test_steps == 2 and len(f._train_meta) == 40
    hop = 275
    limit = 8000 - 8000 % hop
    for b in f.train_group()[:8] + f.test_batches():
        x, c, L = b["inputs"], b["local_condition_features"], b["input_lengths"]
        assert x.dtype == np.int32 and x.shape[0] == 4 and b["targets"].shape == x.shape + (1,)
        assert x.shape[1] == L.max() and (L % hop == 0).all() and L.max() <= limit            # hop-aligned crops below max_time_steps
        assert c.shape == (4, 80, x.shape[1] // hop) and c.min() >= 0.0 and c.max() <= 1.0      # clip +-4 -> [0, 1] (feeder.py:323-335)
        for i in range(4):
            assert (x[i, L[i]:] == 0).all() and (c[i, :, L[i] // hop:] == 0).all()             # audio zero-padded, mels padded with the range minimum -> 0


def test_wavenet_feeder_crop_and_conditioning_match_reference_executed_vectors():
    """the reference's OWN _adjust_time_resolution / _prepare_local_conditions (wavenet_vocoder/feeder.py:319-401) executed with
    np.random.seed(2024): the same random hop-aligned crops (np.random.randint stream) and the same clipped + [0, 1]-rescaled conditioning"""
    from wavenet_vocoder.feeder import Feeder
    hp = hparams.copy()
    f = Feeder.__new__(Feeder)
    f._hparams, f.local_condition = hp, True
    f._rng = np.random.RandomState(2024)
    assert f._limit_time() == int(R["wnf_max_time_steps"])
    n = len(R["wnf_frames"])
    batch = [(R["wnf_x%d" % i], R["wnf_c%d" % i], len(R["wnf_x%d" % i])) for i in range(n)]
    crops = [f._crop(x, c) for x, c, _ in batch]
    for i, (x, c) in enumerate(crops):
        assert np.array_equal(x, R["wnf_crop_x%d" % i]) and np.array_equal(c, R["wnf_crop_c%d" % i]), i
    f._rng = np.random.RandomState(2024)
    out = f.prepare_batch(batch)
    assert np.array_equal(out["local_condition_features"], R["wnf_local_conditions"])
    assert out["local_condition_features"].min() >= 0.0 and out["local_condition_features"].max() <= 1.0
    assert out["inputs"].shape == (n, 11000) and (out["input_lengths"] == 11000).all()


def test_tacotron_feeder_batch_matches_reference_executed_vectors():
    """one whole batch through the reference's OWN Feeder._prepare_batch (tacotron/feeder.py:198-229; its in-batch np.random.shuffle
    replayed): inputs, lengths, mel / stop-token / linear targets with the reference's padding values, bit for bit"""
    from tacotron.feeder import Feeder
    hp = hparams.copy()
    f = Feeder.__new__(Feeder)
    f._hparams, f._pad, f._token_pad, f._target_pad = hp, 0, 1.0, float(R["taco_target_pad"])
    ex = [(R["tf_in%d" % i], R["tf_mel%d" % i], np.zeros(len(R["tf_mel%d" % i]) - 1, dtype=np.float32), R["tf_lin%d" % i], len(R["tf_mel%d" % i]))
          for i in range(4)]
    np.random.RandomState(99).shuffle(ex)
    b = f.prepare_batch(ex)
    for ours, theirs in (("inputs", "inputs"), ("input_lengths", "input_lengths"), ("mel_targets", "mel_targets"), ("token_targets", "token_targets"),
                         ("linear_targets", "linear_targets"), ("targets_lengths", "targets_lengths")):
        assert np.array_equal(b[ours], R["tf_batch_" + theirs]), ours
    assert R["tf_batch_split_infos"].tolist() == [[b["inputs"].shape[1], b["mel_targets"].shape[1], b["token_targets"].shape[1], b["linear_targets"].shape[1]]]


def test_wavenet_preprocessor_layout_on_cpu(tmp_path, monkeypatch):
    """datasets/wavenet_preprocessor.py (reference :11-154) end to end with the three GPU-backed audio calls replaced by the oracle's
    restatements (the kernels have their own GPU parity tests): file naming, dtypes, hop alignment, map.txt rows, and the rows feed the
    WaveNet feeder"""
    from scipy.io import wavfile
    from oracle import audio as oa
    from datasets import audio, wavenet_preprocessor as wp
    import wavenet_vocoder.util as wu
    monkeypatch.setattr(audio, "melspectrogram", lambda w, hp: oa.melspectrogram(np.asarray(w, dtype=np.float32), hp).astype(np.float32))
    monkeypatch.setattr(audio, "preemphasis", lambda w, k, p=True: oa.preemphasis(w, k, p))
    monkeypatch.setattr(wp, "mulaw_quantize", lambda x, mu=256: oa.mulaw_quantize(np.asarray(x, dtype=np.float32)))
    src = tmp_path / "wavs"
    src.mkdir()
    rng = np.random.default_rng(5)
    for i in range(9):
        n = int(rng.integers(6000, 9000))
        w = 0.3 * np.sin(np.arange(n) * (0.03 + 0.004 * i)) + 0.01 * rng.standard_normal(n)
        wavfile.write(str(src / ("utt%02d.wav" % i)), 22050, (w * 32767).astype(np.int16))
    hp = hparams.copy()
    hp.parse("input_type=mulaw-quantize,quantize_channels=256,out_channels=256,trim_silence=False,wavenet_batch_size=2,wavenet_test_size=2,"
             "wavenet_test_batches=None,max_time_steps=5500,train_with_GTA=False")
    mel_dir, wav_dir = tmp_path / "out" / "mels", tmp_path / "out" / "audio"
    mel_dir.mkdir(parents=True)
    wav_dir.mkdir(parents=True)
    rows = wp.build_from_path(hp, str(src), str(mel_dir), str(wav_dir))
    assert len(rows) == 9
    for a_path, m_path, m2, g, steps, frames in rows:
        a, m = np.load(a_path), np.load(m_path)
        assert m2 == m_path and g == "<no_g>" and a.dtype == np.int16 and m.dtype == np.float32
        assert m.shape == (frames, 80) and len(a) == steps == frames * 275 and 0 <= a.min() and a.max() <= 255
    import wavenet_preprocess
    wavenet_preprocess.write_metadata(rows, str(tmp_path / "out"), hp)
    from wavenet_vocoder.feeder import Feeder
    f = Feeder(str(tmp_path / "out" / "map.txt"), "", hp)
    b = f.train_group()[0]
    assert b["inputs"].shape[0] == 2 and b["inputs"].shape[1] <= 5500 and b["inputs"].shape[1] % 275 == 0
    assert b["local_condition_features"].shape[1:] == (80, b["inputs"].shape[1] // 275)


def test_text_front_end_matches_reference_executed_vectors():
    """tests/golden/reference_text.json is written by executing the reference's tacotron/utils/text.py + cleaners.py + cmudict.py
    (make_reference_text.py): cleaner pipelines on ASCII text, the {ARPAbet} cutting rule, dropped symbols, EOS, the dictionary parser"""
    import io
    import json
    from tacotron.utils import cmudict
    from tacotron.utils.symbols import symbols
    from tacotron.utils.text import sequence_to_text, text_to_sequence
    with open(os.path.join(os.path.dirname(__file__), "golden", "reference_text.json")) as f:
        R = json.load(f)
    assert symbols == R["symbols"] and cmudict.valid_symbols == R["valid_symbols"]
    assert len(R["cases"]) >= 30
    for c in R["cases"]:
        seq = text_to_sequence(c["text"], c["cleaners"])
        assert seq == c["sequence"], (c["text"], c["cleaners"])
        assert sequence_to_text(seq) == c["round_trip"]
    for keep in (1, 0):
        d = cmudict.CMUDict(io.StringIO(R["dict_text"]), keep_ambiguous=bool(keep))
        want = R["cmudict_keep_%d" % keep]
        assert len(d) == want["len"]
        for w, p in want["lookups"].items():
            assert d.lookup(w) == p, w


def test_paper_hparams_is_a_superset_of_the_defaults():
    import paper_hparams
    base, paper = hparams.values(), paper_hparams.hparams.values()
    assert set(base) | {"upsample_conditional_features"} == set(paper)
    changed = {k for k in base if base[k] != paper[k]}
    assert {"layers", "stacks", "residual_channels", "gate_channels", "skip_out_channels", "out_channels", "upsample_type",
            "upsample_scales", "predict_linear"} <= changed and len(changed) == 23
    assert (paper["layers"], paper["stacks"], paper["out_channels"], paper["upsample_scales"]) == (24, 4, 30, [5, 5, 11])
    assert int(np.prod(paper["upsample_scales"])) == paper["hop_size"]
    assert hparams.layers == 20 and hparams.predict_linear is True           # the defaults were not touched by the import


def test_synthesizer_output_lengths_match_reference_executed_vectors():
    """tacotron/synthesizer.py:254-257 executed (reference_exec.npz section L): the length of a synthesized row is the INDEX of the first
    rounded stop prediction of 1 - the frame the stop fires on is dropped; 0.5 rounds to 0 (half to even)."""
    import ast
    src = open(os.path.join(os.path.dirname(os.path.dirname(__file__)), "tacotron", "synthesizer.py")).read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "get_output_lengths")
    ns = {"np": np}
    exec(compile(ast.Module([fn], []), "synthesizer.py", "exec"), ns)          # the module itself imports torch + the CUDA binding
    assert ns["get_output_lengths"](R["synth_stop_rows"]) == R["synth_output_lengths"].tolist()
    assert R["synth_output_lengths"][:4].tolist() == [12, 0, 11, 12]
    from tacotron.feeder import pad_input, pad_target
    assert pad_input(np.arange(1, 6, dtype=np.int32), 8, 0).tolist() == R["synth_pad_input"].tolist()
    assert np.array_equal(pad_target(np.ones((3, 2), dtype=np.float32), 5, -4.1), R["synth_pad_target"])


@pytest.mark.parametrize("itype,qc", [("mulaw-quantize", 256), ("mulaw", 256), ("raw", 65536)])
def test_preprocessors_match_the_executed_reference(tmp_path, monkeypatch, itype, qc):
    """reference_exec.npz section M: the reference's two `_process_utterance` functions executed on one synthetic utterance per input
    type (librosa primitives substituted by the oracle restatements). The product preprocessors, with their GPU-backed calls replaced by
    the same restatements, write the same files: names by basename, rows, lengths, silence cut, padding, dtypes, values."""
    from scipy.io import wavfile
    from oracle import audio as oa
    from datasets import audio, preprocessor as pre, wavenet_preprocessor as wpre
    f32 = lambda w: np.asarray(w, dtype=np.float32)
    monkeypatch.setattr(audio, "melspectrogram", lambda w, hp: oa.melspectrogram(f32(w), hp).astype(np.float32))
    monkeypatch.setattr(audio, "linearspectrogram", lambda w, hp: oa.linearspectrogram(f32(w), hp).astype(np.float32))
    monkeypatch.setattr(audio, "preemphasis", lambda w, k, p=True: oa.preemphasis(w, k, p))
    for mod in (pre, wpre):
        monkeypatch.setattr(mod, "mulaw_quantize", lambda x, mu=256: oa.mulaw_quantize(f32(x)))
        monkeypatch.setattr(mod, "mulaw", lambda x, mu=256: oa.mulaw(f32(x)))
    wav_path = str(tmp_path / "utt.wav")
    wavfile.write(wav_path, 22050, R["pre_wav_i16"])
    hp = hparams.copy()
    hp.parse("trim_silence=False,input_type=%s,quantize_channels=%d" % (itype, qc))
    tag = itype.replace("-", "_")
    d = str(tmp_path / "taco")
    os.makedirs(d)
    row = pre._process_utterance(d, d, d, "utt", wav_path, "some text", hp)
    assert [str(x) for x in row] == R["pre_%s_row" % tag].tolist()
    a, mel, lin = (np.load(os.path.join(d, row[i])) for i in range(3))
    ref_a = R["pre_%s_audio" % tag]
    assert a.dtype == ref_a.dtype and a.shape == ref_a.shape
    if itype == "mulaw-quantize":
        assert np.array_equal(a, ref_a)
    else:
        assert np.abs(a - ref_a).max() <= 1e-6
    assert mel.dtype == np.float32 and np.abs(mel - R["pre_%s_mel" % tag]).max() <= 2e-4
    assert np.abs(lin[:, ::16] - R["pre_%s_linear_cols" % tag]).max() <= 2e-4
    d = str(tmp_path / "wn")
    os.makedirs(d)
    row = wpre._process_utterance(d, d, "utt", wav_path, hp)
    assert [os.path.basename(str(x)) for x in row] == R["wpre_%s_row" % tag].tolist()
    a, mel = np.load(row[0]), np.load(row[1])
    ref_a = R["wpre_%s_audio" % tag]
    assert a.dtype == ref_a.dtype and a.shape == ref_a.shape and (np.array_equal(a, ref_a) if itype == "mulaw-quantize" else np.abs(a - ref_a).max() <= 1e-6)
    assert np.abs(mel - R["wpre_%s_mel" % tag]).max() <= 2e-4


def test_stock_configurations_pass_the_engine_config_checks():
    """hparams.py and paper_hparams.py as shipped are inside the supported matrix of both engines; flags that would silently change
    the arithmetic are rejected by name"""
    import importlib
    import paper_hparams
    taco, wn = importlib.import_module("tacotron-2_b200.tacotron"), importlib.import_module("tacotron-2_b200.wavenet")
    for hp in (hparams, paper_hparams.hparams):
        assert taco.unsupported_hparams(hp) == [] and wn.unsupported_hparams(hp) == []
    for name, value in (("tacotron_natural_eval", True), ("wavenet_natural_eval", True), ("synthesis_constraint", True), ("wavenet_synth_debug", True), ("outputs_per_step", 2), ("smoothing", True),
                        ("wavenet_weight_normalization", True), ("upsample_type", "Resize"), ("tacotron_teacher_forcing_mode", "scheduled")):
        hp = hparams.copy()
        setattr(hp, name, value)
        bad = taco.unsupported_hparams(hp) + wn.unsupported_hparams(hp)
        assert len(bad) == 1 and bad[0].startswith(name + "="), (name, bad)


def test_number_normaliser_never_raises_and_leaves_no_digits():
    from hypothesis import given, settings, strategies as st
    from tacotron.utils.cleaners import english_cleaners

    @settings(max_examples=400, deadline=None)
    @given(st.text(alphabet="0123456789$£.,-stndrh aA{}%", max_size=30))
    def run(s):
        assert not any(c.isdigit() for c in english_cleaners(s))
    run()
