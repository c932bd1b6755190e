"""CPU-side checks of the C-ABI: the in-tree library loads, exports every function include/t2b200.h declares, reports
errors through return codes (no GPU compute is launched here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "t2b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(t2_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from t2_import import t2
    lib = t2.lib.load()
    names = _declared()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.t2_abi_version() == 2


def test_errors_are_return_codes_with_messages():
    from t2_import import t2
    lib = t2.lib.load()
    lib.t2_last_error.restype = ctypes.c_char_p
    cfg = t2.wavenet.WnConfig()          # all zeros: invalid
    sz = t2.wavenet.WnSizes()
    rc = lib.t2_wn_sizes(ctypes.byref(cfg), ctypes.byref(sz))
    assert rc < 0 and len(lib.t2_last_error()) > 0
    with pytest.raises(t2.lib.T2Error):
        t2.lib.check(rc)


def test_layout_queries_match_the_oracle_parameter_tables():
    """host logic without a GPU: the C++ layout enumerates the same (name, shape) list as the oracle"""
    from hparams import hparams, paper_hparams
    from oracle import tacotron as ot
    from oracle import wavenet as ow
    from t2_import import t2
    lib = t2.lib.load()
    hp = paper_hparams()
    hp.parse("input_type=mulaw-quantize,quantize_channels=256,out_channels=256,upsample_type=SubPixel,upsample_scales=[11,25]")
    cfg = t2.wavenet.make_config(hp, 2, 275 * 4)
    sz = t2.wavenet.WnSizes()
    t2.lib.check(lib.t2_wn_sizes(ctypes.byref(cfg), ctypes.byref(sz)))
    name = ctypes.create_string_buffer(160)
    off, nd, shp = ctypes.c_longlong(), ctypes.c_int(), (ctypes.c_int * 4)()
    got = []
    for i in range(sz.n_tensors):
        t2.lib.check(lib.t2_wn_param_info(ctypes.byref(cfg), i, name, 160, ctypes.byref(off), ctypes.byref(nd), shp))
        got.append((name.value.decode(), tuple(shp[k] for k in range(nd.value))))
    assert got == [(k, tuple(v)) for k, v in ow.param_shapes(hp).items()]
    hp2 = hparams.copy()
    hp2.set_hparam("predict_linear", False)
    tc = t2.tacotron.make_config(hp2, 4, 40, 80)
    n, pb, wb, nt, tr = ctypes.c_longlong(), ctypes.c_longlong(), ctypes.c_longlong(), ctypes.c_int(), ctypes.c_int()
    t2.lib.check(lib.t2_taco_sizes(ctypes.byref(tc), ctypes.byref(n), ctypes.byref(pb), ctypes.byref(wb), ctypes.byref(nt)))
    got = []
    for i in range(nt.value):
        t2.lib.check(lib.t2_taco_param_info(ctypes.byref(tc), i, name, 160, ctypes.byref(off), ctypes.byref(nd), shp, ctypes.byref(tr)))
        got.append((name.value.decode(), tuple(shp[k] for k in range(nd.value)), bool(tr.value)))
    assert got == [(k, tuple(v), ot.is_trainable(k)) for k, v in ot.param_shapes(hp2).items()]


def test_product_path_has_no_oracle_import():
    """the shipped package must never route through the CPU oracle"""
    pkg = os.path.join(ROOT, "tacotron-2_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            assert not re.search(r"^\s*(from|import)\s+oracle", open(os.path.join(pkg, f)).read(), re.M), f
    assert not re.search(r"^\s*(from|import)\s+oracle", open(os.path.join(ROOT, "datasets", "audio.py")).read(), re.M)


def test_product_initialiser_matches_reference_init_rules():
    """tacotron-2_b200/init.py (product side): same NN_init upsampling kernels as the oracle, glorot limits, unit batch-norm"""
    import math

    import torch
    from hparams import hparams, paper_hparams
    from oracle import tacotron as ot
    from oracle import wavenet as ow
    from t2_import import t2
    for hp in (paper_hparams(), hparams.copy()):
        tens = [(k, 0, tuple(v)) for k, v in ow.param_shapes(hp).items()]
        a, b = t2.init.wavenet_variables(hp, tens, 7), ow.init_params(hp, seed=7)
        for k in b:
            assert a[k].shape == b[k].shape
            if k.endswith("bias"):
                assert a[k].abs().max() == 0
            elif "upsampling" in k:
                assert torch.equal(a[k], b[k]), k
    k = "residual_block_causal_conv_ResidualConv1DGLU_3/kernel"
    kk = [n for n in a if n.endswith("kernel") and a[n].dim() == 3 and a[n].shape[0] == 3][0]
    kw, cin, cout = a[kk].shape
    lim = math.sqrt(6.0 / (kw * cin + kw * cout))
    assert a[kk].abs().max() <= lim and a[kk].abs().max() > 0.95 * lim
    hp = hparams.copy()
    hp.set_hparam("predict_linear", False)
    tens = [(n, 0, tuple(v), ot.is_trainable(n)) for n, v in ot.param_shapes(hp).items()]
    p = t2.init.tacotron_variables(hp, tens, 3)
    assert all((p[n] == 1).all() for n in p if n.endswith(("gamma", "moving_variance")))
    assert all((p[n] == 0).all() for n in p if n.endswith(("beta", "moving_mean", "bias")))
    e = p["inputs_embedding"]
    assert e.abs().max() <= math.sqrt(6.0 / sum(e.shape))


def test_reference_python_surface_validates_like_the_reference():
    """create_model / initialize argument checks fire before any GPU work (tacotron.py:41-54, wavenet models/__init__.py:6-9)"""
    import torch
    from hparams import hparams
    from tacotron.models import create_model as create_taco
    from wavenet_vocoder.models import create_model as create_wn
    from wavenet_vocoder import util
    with pytest.raises(Exception, match="Unknown model"):
        create_taco("Tacotron3", hparams)
    with pytest.raises(Exception, match="Unknow model"):
        hp = hparams.copy()
        hp.parse("input_type=raw,out_channels=30")
        create_wn("WaveRNN", hp)
    bad = hparams.copy()
    bad.parse("input_type=mulaw-quantize,quantize_channels=256,out_channels=30")
    with pytest.raises(RuntimeError):
        create_wn("WaveNet", bad)
    m = create_taco("Tacotron", hparams)
    ids, lens = torch.zeros(2, 5, dtype=torch.int32), torch.tensor([5, 4])
    mel, stop = torch.zeros(2, 7, hparams.num_mels), torch.zeros(2, 7)
    with pytest.raises(ValueError):
        m.initialize(ids, lens, stop_token_targets=stop)                    # stop targets without mel targets
    with pytest.raises(ValueError):
        m.initialize(ids, lens, mel_targets=mel)                            # mel targets without stop targets
    with pytest.raises(ValueError):
        m.initialize(ids, lens, mel, stop, is_training=True)                # predict_linear=True (default) without linear targets
    hp_mel = hparams.copy()
    hp_mel.set_hparam("predict_linear", False)
    with pytest.raises(RuntimeError):
        create_taco("Tacotron", hp_mel).initialize(ids, lens, mel, stop, is_training=True, is_evaluating=True)
    with pytest.raises(ValueError):
        m.initialize(ids, lens, mel, gta=True, linear_targets=mel)
    assert util.is_mulaw_quantize("mulaw-quantize") and util.is_scalar_input("raw") and not util.is_raw("mulaw")
    with pytest.raises(AssertionError):
        util.is_mulaw("pcm")
    mask = util.sequence_mask([3, 1], max_len=4, expand=False)
    assert mask.tolist() == [[1, 1, 1, 0], [1, 0, 0, 0]] and util.sequence_mask([2, 1]).shape == (2, 2, 1)


def test_ctypes_struct_mirrors_match_the_library():
    """every POD struct that crosses the C-ABI has the same size in the Python mirror (field order is checked by the GPU tests)"""
    import ctypes
    from t2_import import t2
    lib = t2.lib.load()
    lib.t2_struct_size.argtypes = [ctypes.c_char_p]
    pairs = {"t2_wn_config_t": t2.wavenet.WnConfig, "t2_wn_sizes_t": t2.wavenet.WnSizes, "t2_taco_config_t": t2.tacotron.TacoConfig,
             "t2_cbhg_config_t": t2.tacotron.CbhgConfig, "t2_audio_config_t": t2.audio.AudioConfig}
    for name, mirror in pairs.items():
        assert lib.t2_struct_size(name.encode()) == ctypes.sizeof(mirror), name
    assert lib.t2_struct_size(b"nope") == -1
