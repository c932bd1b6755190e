"""The oracle against the reference's OWN Tacotron graph code, executed: tests/golden/reference_graph.npz holds what
`Tacotron.initialize()` + `add_loss()` of /root/reference/tacotron/models/tacotron.py produce when they run (with modules.py,
attention.py, Architecture_wrappers.py, helpers.py, custom_decoder.py underneath) on the TF-1 stand-in of tests/golden/tf_shim*.py
(make_reference_graph_vectors.py; the layer primitives there restate TF 1.x, the composition above them is the reference's).
Checked here, on CPU:
  * the variable names / shapes the reference's scopes produce == t2_tf_bundle.tacotron_tf_name over the engine's parameter table
    (the TF-checkpoint name map), and the regularisation filter picks the same variables;
  * oracle.forward / loss_fn / synthesize / linear_head reproduce the executed reference in training (all dropout / zoneout masks
    injected), masked-loss training, evaluation, GTA and free-running synthesis (stop rule and max_iters);
  * d loss / d variable through the oracle == autograd through the executed reference graph, for every trainable variable."""
import os

import numpy as np
import pytest
import torch

import t2_tf_bundle
from hparams import hparams
from oracle import tacotron as ot

PATH = os.path.join(os.path.dirname(__file__), "golden", "reference_graph.npz")
TOL = 2e-5


@pytest.fixture(scope="module")
def R():
    return np.load(PATH)


def _hp(R, **kw):
    hp = hparams.copy()
    for k, v in zip(R["small_hparams_keys"], R["small_hparams_values"]):
        setattr(hp, str(k), eval(str(v)))
    for k, v in kw.items():
        setattr(hp, k, v)
    return hp


def _params(R, drop=()):
    """fixture variables (reference scope names) -> the oracle's parameter dict, through the checkpoint name map"""
    out = {}
    for name in R["var_names"]:
        eng = t2_tf_bundle.engine_name("Tacotron_model/" + str(name))
        assert eng is not None, name
        if not any(d in eng for d in drop):
            out[eng] = torch.from_numpy(R["var/" + str(name)]).clone()
    return out


def _inputs(R):
    t = lambda k, dt=None: torch.from_numpy(R[k]).to(dt) if dt else torch.from_numpy(R[k])
    return (t("inputs", torch.int64), t("input_lengths", torch.int64), t("mel_targets"), t("stop_targets"), t("linear_targets"),
            t("targets_lengths", torch.int64))


def _masks(R, tag, training, hp):
    m = {"prenet_drop": [torch.from_numpy(R["%s_mask_prenet_drop_%d" % (tag, i)]) for i in range(len(hp.prenet_layers))]}
    if training:
        for i in range(hp.enc_conv_num_layers):
            m[("enc_drop", i)] = torch.from_numpy(R["%s_mask_enc_drop_%d" % (tag, i)])
        for i in range(hp.postnet_num_layers):
            m[("post_drop", i)] = torch.from_numpy(R["%s_mask_post_drop_%d" % (tag, i)])
        ez, dz = {}, {}
        for d in ("fw", "bw"):
            for s in "ch":
                a = torch.from_numpy(R["%s_mask_enc_zone_%s_%s" % (tag, d, s)])
                for t in range(a.shape[0]):
                    ez[(d, s, t)] = a[t]
        for l in (1, 2):
            for s in "ch":
                a = torch.from_numpy(R["%s_mask_dec_zone_%d_%s" % (tag, l, s)])
                for t in range(a.shape[0]):
                    dz[(l, s, t)] = a[t]
        m["enc_zone"], m["dec_zone"] = ez, dz
    return m


def _close(a, b, tol=TOL):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    err = np.abs(a - b).max() if a.size else 0.0
    assert err <= tol * max(1.0, np.abs(b).max()), err
    return err


def _check_outputs(R, tag, out, stop_is_logit):
    _close(out["decoder_output"].detach(), R[tag + "_decoder_output"])
    _close(out["mel_outputs"].detach(), R[tag + "_mel_outputs"])
    _close(out["alignments"].detach().transpose(1, 2), R[tag + "_alignments"])                 # reference layout [B, T_in, T_out]
    _close(out["stop_logits" if stop_is_logit else "stop_token_prediction"].detach(), R[tag + "_stop_token_prediction"])


def test_variable_names_of_the_executed_reference_graph_match_the_checkpoint_name_map(R):
    hp = _hp(R, predict_linear=True)
    got = {"Tacotron_model/" + str(n): tuple(R["var/" + str(n)].shape) for n in R["var_names"]}
    want = {t2_tf_bundle.tacotron_tf_name(k): tuple(v) for k, v in ot.param_shapes(hp).items()}
    assert set(got) == set(want), (sorted(set(got) - set(want))[:5], sorted(set(want) - set(got))[:5])
    assert got == want
    # trainable flags and the regularisation filter (tacotron.py:343-345 runs on the TF names, the oracle's on the engine names)
    for n, tr in zip(R["var_names"], R["var_trainable"]):
        tf_name = "Tacotron_model/" + str(n)
        eng = t2_tf_bundle.engine_name(tf_name)
        assert ot.is_trainable(eng) == bool(tr)
        ref_reg = bool(tr) and not any(s in tf_name for s in ("bias", "Bias", "_projection", "inputs_embedding", "RNN", "LSTM"))
        assert ot.is_regularized(eng) == ref_reg, tf_name


def test_training_graph_outputs_losses_and_gradients(R):
    hp = _hp(R, predict_linear=True, mask_decoder=False)
    params = {k: v.requires_grad_(ot.is_trainable(k)) for k, v in _params(R).items()}
    ids, in_len, mel, stop, lin, tgt_len = _inputs(R)
    out = ot.forward(params, ids, in_len, mel, hp, training=True, masks=_masks(R, "train", True, hp))
    _check_outputs(R, "train", out, True)
    _close(out["linear_outputs"].detach(), R["train_linear_outputs"])
    total, parts = ot.loss_fn(out, mel, stop, params, hp, tgt_len, lin)
    for k, ref in (("before", "before_loss"), ("after", "after_loss"), ("stop", "stop_token_loss"), ("reg", "regularization_loss"),
                   ("linear", "linear_loss")):
        assert abs(float(parts[k].detach()) - float(R["train_" + ref])) <= 1e-5 * max(1e-3, abs(float(R["train_" + ref]))), k
    assert abs(float(total.detach()) - float(R["train_loss"])) <= 1e-5 * abs(float(R["train_loss"]))
    total.backward()
    worst = 0.0
    # a conv bias in front of a training-mode batch norm has an exactly-zero gradient (rounding noise on both sides): errors are
    # measured against max(|reference gradient|, 1e-3 x the largest gradient entry of the whole model)
    floor = 1e-3 * max(np.abs(R[k]).max() for k in R.files if k.startswith("grad/"))
    for name in R["var_names"]:
        eng = t2_tf_bundle.engine_name("Tacotron_model/" + str(name))
        if not ot.is_trainable(eng):
            continue
        ref = R["grad/" + str(name)]
        g = params[eng].grad
        g = np.zeros_like(ref) if g is None else g.numpy()
        scale = max(np.abs(ref).max(), floor)
        err = np.abs(g - ref).max() / scale
        worst = max(worst, err)
        assert err <= 2e-4, (eng, err)
    assert worst > 0.0


def test_training_graph_with_masked_losses(R):
    hp = _hp(R, predict_linear=False, mask_decoder=True)
    params = _params(R, drop=("CBHG", "cbhg"))
    ids, in_len, mel, stop, lin, tgt_len = _inputs(R)
    out = ot.forward(params, ids, in_len, mel, hp, training=True, masks=_masks(R, "train_md", True, hp))
    _check_outputs(R, "train_md", out, True)
    total, parts = ot.loss_fn(out, mel, stop, params, hp, tgt_len)
    for k, ref in (("before", "before_loss"), ("after", "after_loss"), ("stop", "stop_token_loss"), ("reg", "regularization_loss")):
        assert abs(float(parts[k]) - float(R["train_md_" + ref])) <= 1e-5 * max(1e-3, abs(float(R["train_md_" + ref]))), k
    assert abs(float(total) - float(R["train_md_loss"])) <= 1e-5 * abs(float(R["train_md_loss"]))
    assert float(R["train_md_linear_loss"]) == 0.0


def test_evaluation_and_gta_graphs(R):
    hp = _hp(R, predict_linear=True, mask_decoder=False)
    params = _params(R)
    ids, in_len, mel, stop, lin, tgt_len = _inputs(R)
    out = ot.forward(params, ids, in_len, mel, hp, training=False, masks=_masks(R, "eval", False, hp))
    _check_outputs(R, "eval", out, True)                                   # evaluation keeps the stop LOGITS (modules.py:340-341)
    _close(out["linear_outputs"], R["eval_linear_outputs"])
    total, parts = ot.loss_fn(out, mel, stop, params, hp, tgt_len, lin)
    assert abs(float(total) - float(R["eval_loss"])) <= 1e-5 * abs(float(R["eval_loss"]))
    out = ot.forward(params, ids, in_len, mel, hp, training=False, masks=_masks(R, "gta", False, hp))
    out["stop_token_prediction"] = torch.sigmoid(out["stop_logits"])      # GTA is a synthesis mode: sigmoid applied
    _check_outputs(R, "gta", out, False)


@pytest.mark.parametrize("tag", ["synth", "synth_stop"])
def test_free_running_synthesis_and_stop_rule(R, tag):
    hp = _hp(R, predict_linear=True)
    params = _params(R)
    if tag == "synth_stop":                       # the generator negated the stop projection and shifted its bias (see its comments)
        params["stop_token_projection/kernel"] = -params["stop_token_projection/kernel"]
        params["stop_token_projection/bias"] = -params["stop_token_projection/bias"] + float(R["synth_stop_bias_shift"])
    ids, in_len = _inputs(R)[:2]
    pm = [torch.from_numpy(R["%s_mask_prenet_drop_%d" % (tag, i)]) for i in range(len(hp.prenet_layers))]
    steps = pm[0].shape[1]
    assert steps == (hp.max_iters if tag == "synth" else 4)
    out = ot.synthesize(params, ids, in_len, hp, prenet_masks=[[m[:, t] for m in pm] for t in range(steps)] + [None] * 4)
    assert out["mel_outputs"].shape[1] == steps                            # same number of decoder steps as the executed reference
    _check_outputs(R, tag, out, False)
    _close(ot.linear_head(out["mel_outputs"], params, hp, False), R[tag + "_linear_outputs"])


def test_one_optimizer_step_of_the_executed_reference(R):
    """tacotron.py:371-463 executed: LR schedule at step 60000, clip_by_global_norm(1.0), Adam; and the batch-norm moving averages its
    UPDATE_OPS dependency advances. oracle.train_step + adam_step land on the same variables; moving statistics follow
    0.99 old + 0.01 batch with the (biased) batch moments the oracle reports."""
    hp = _hp(R, predict_linear=True, mask_decoder=False)
    params = _params(R)
    ids, in_len, mel, stop, lin, tgt_len = _inputs(R)
    step = int(R["train_global_step"])
    assert abs(ot.learning_rate(hp, step) - float(R["train_learning_rate"])) <= 1e-6 * float(R["train_learning_rate"])
    assert hp.tacotron_final_learning_rate < ot.learning_rate(hp, step) < hp.tacotron_initial_learning_rate      # inside the decay
    loss, grads, out, parts = ot.train_step(params, ids, in_len, mel, stop, hp, masks=_masks(R, "train", True, hp), targets_lengths=tgt_len,
                                            linear_targets=lin)
    gn = float(torch.sqrt(sum((g * g).sum() for g in grads.values())))
    assert gn > 1.0                                                                        # the clip is active in this fixture
    new = {k: v.clone() for k, v in params.items()}
    ot.adam_step(new, grads, {}, hp, step)
    lr = float(R["train_learning_rate"])
    stats = {}
    ot.forward(params, ids, in_len, mel, hp, training=True, masks=_masks(R, "train", True, hp), stats_out=stats)
    n_moving = 0
    for name in R["var_names"]:
        eng = t2_tf_bundle.engine_name("Tacotron_model/" + str(name))
        ref = R["train_new/" + str(name)]
        if ot.is_trainable(eng):
            delta_ref = ref - R["var/" + str(name)]
            delta = new[eng].numpy() - params[eng].numpy()
            # Adam's first step is lr * g / (|g| + eps): measured against the learning rate (elements whose gradient is rounding noise
            # around eps = 1e-6 move by an arbitrary fraction of lr on both sides), plus fp32 spacing of the parameter itself
            assert np.abs(delta - delta_ref).max() <= 5e-3 * lr + 2e-7 * np.abs(ref).max(), eng
        else:
            prefix, leaf = eng.rsplit("/", 1)
            mean, var = stats[prefix + "/"]
            want = 0.99 * params[eng] + 0.01 * (mean if leaf == "moving_mean" else var)
            assert np.abs(want.numpy() - ref).max() <= 1e-6, eng
            n_moving += 1
    assert n_moving == 2 * (hp.enc_conv_num_layers + hp.postnet_num_layers + hp.cbhg_kernels + 2)


def test_error_contract_of_the_drop_in_models_matches_the_executed_reference(R):
    """tests/golden/reference_errors.json: exception class + message of every call the reference rejects (tacotron.py:41-54 and both
    create_model functions), recorded while executing it. The drop-in surface raises the same ones - before touching the GPU, so on CPU."""
    import json
    from tacotron.models import create_model as taco_create
    from wavenet_vocoder.models import create_model as wn_create
    with open(os.path.join(os.path.dirname(__file__), "golden", "reference_errors.json")) as f:
        want = json.load(f)
    ids, in_len, mel, stop, lin, tgt_len = _inputs(R)
    hp = _hp(R, predict_linear=True, mask_decoder=True)
    base = dict(mel_targets=mel, stop_token_targets=stop, linear_targets=lin, targets_lengths=tgt_len)
    cases = {
        "stop_without_mel": dict(base, mel_targets=None, is_training=True),
        "mel_without_stop": dict(base, stop_token_targets=None, is_training=True),
        "linear_missing_in_training": dict(base, linear_targets=None, is_training=True),
        "linear_given_in_gta": dict(base, gta=True),
        "mask_without_lengths": dict(base, targets_lengths=None, is_training=True),
        "training_and_evaluating": dict(base, is_training=True, is_evaluating=True),
    }
    got = {}

    def record(key, fn):
        try:
            fn()
            got[key] = None
        except Exception as e:                                       # noqa: BLE001
            got[key] = [type(e).__name__, str(e)]
    for key, kw in cases.items():
        record("tacotron_initialize/" + key, lambda kw=kw: taco_create("Tacotron", hp).initialize(ids, in_len, **kw))
    record("tacotron_create_model/unknown", lambda: taco_create("Tacotron-3", hp))
    wh = hparams.copy()
    wh.input_type, wh.quantize_channels, wh.out_channels = "mulaw-quantize", 256, 30
    record("wavenet_create_model/out_channels_mismatch", lambda: wn_create("WaveNet", wh))
    wh.out_channels = 256
    record("wavenet_create_model/unknown", lambda: wn_create("WaveRNN", wh))
    assert got == want


def test_training_graph_with_asymmetric_mels_and_scaled_regulariser(R):
    """symmetric_mels=False moves the output clip to [0 - lower_bound_decay, max_abs_value] (tacotron.py:89,176,199,216) and
    tacotron_scale_regularization=True divides the L2 weight by max_abs_value (:334-338); executed reference vs oracle"""
    hp = _hp(R, predict_linear=True, mask_decoder=False, symmetric_mels=False, tacotron_scale_regularization=True)
    params = _params(R)
    ids, in_len, mel, stop, lin, tgt_len = _inputs(R)
    out = ot.forward(params, ids, in_len, mel, hp, training=True, masks=_masks(R, "train_asym", True, hp))
    _check_outputs(R, "train_asym", out, True)
    _close(out["linear_outputs"], R["train_asym_linear_outputs"])
    lo = torch.tensor(-hp.lower_bound_decay)                                 # float32(-0.1)
    assert out["decoder_output"].min() == lo and float((out["decoder_output"] == lo).float().mean()) > 0.05
    total, parts = ot.loss_fn(out, mel, stop, params, hp, tgt_len, lin)
    for k, ref in (("before", "before_loss"), ("after", "after_loss"), ("stop", "stop_token_loss"), ("reg", "regularization_loss"),
                   ("linear", "linear_loss")):
        assert abs(float(parts[k]) - float(R["train_asym_" + ref])) <= 1e-5 * max(1e-3, abs(float(R["train_asym_" + ref]))), k
    assert abs(float(R["train_asym_regularization_loss"]) / float(R["train_regularization_loss"]) - 1.0 / hp.max_abs_value) < 1e-4
    # the product config folds the asymmetric lower bound into the two fields its kernels clip with
    import importlib
    taco = importlib.import_module("tacotron-2_b200.tacotron")
    assert abs((-hp.max_abs_value - taco._decay_field(hp)) - (0.0 - hp.lower_bound_decay)) < 1e-7
    assert taco._decay_field(_hp(R)) == hp.lower_bound_decay


def test_stand_in_layer_primitives_against_torch_kernels():
    """tests/golden/tf_shim_graph.py --selfcheck: the Conv1D / Conv2D / Conv2DTranspose / BatchNormalization / max-pool / LSTMCell stand-ins
    the reference's code was executed on, against torch.nn.functional's own kernels (run in a subprocess: the stand-in installs a fake
    `tensorflow` module)."""
    import subprocess
    import sys
    golden = os.path.join(os.path.dirname(__file__), "golden")
    r = subprocess.run([sys.executable, "tf_shim_graph.py", "--selfcheck"], cwd=golden, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "selfcheck ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_fine_tuning_optimizer_step_of_the_executed_reference(R):
    """tacotron.py:401 executed with tacotron_fine_tuning=True: the embedding and every `encoder_*` variable get no gradient and no Adam
    update, and the global-norm clip sees the remaining gradients only. oracle.adam_step follows; the product freezes the same leading
    range of its flat buffer (tacotron-2_b200/tacotron.py optimizer_step)."""
    hp = _hp(R, predict_linear=True, mask_decoder=False, tacotron_fine_tuning=True)
    params = _params(R)
    ids, in_len, mel, stop, lin, tgt_len = _inputs(R)
    loss, grads, _, _ = ot.train_step(params, ids, in_len, mel, stop, hp, masks=_masks(R, "train", True, hp), targets_lengths=tgt_len,
                                      linear_targets=lin)
    new = {k: v.clone() for k, v in params.items()}
    lr = ot.adam_step(new, grads, {}, hp, int(R["train_global_step"]))
    frozen = updated = 0
    for name in R["var_names"]:
        eng = t2_tf_bundle.engine_name("Tacotron_model/" + str(name))
        if not ot.is_trainable(eng):
            continue
        key = "train_ft_new/" + str(name)
        if eng.startswith(("inputs_embedding", "encoder_")):
            assert key not in R.files and torch.equal(new[eng], params[eng]), eng
            frozen += 1
        else:
            delta_ref = R[key] - R["var/" + str(name)]
            delta = new[eng].numpy() - params[eng].numpy()
            assert np.abs(delta - delta_ref).max() <= 5e-3 * lr + 2e-7 * np.abs(R[key]).max(), eng
            updated += 1
    assert frozen == 1 + 4 * hp.enc_conv_num_layers + 4 and updated > 50
    # the clip really differs between the two modes in this fixture (so the norm's membership is exercised)
    k = "inference/postnet_projection/projection_postnet_projection/kernel"
    assert np.abs(R["train_ft_new/" + k] - R["train_new/" + k]).max() > 0
    # the product's frozen range = the leading tensors of its parameter table
    names = list(ot.param_shapes(hp))
    first_free = next(i for i, n in enumerate(names) if not n.startswith(("inputs_embedding", "encoder_")))
    assert all(n.startswith(("inputs_embedding", "encoder_")) for n in names[:first_free]) and not any(
        n.startswith(("inputs_embedding", "encoder_")) or "encoder_" in n or "inputs_embedding" in n for n in names[first_free:])


@pytest.mark.parametrize("tag,flags", [("train_smooth", dict(smoothing=True, cumulative_weights=False)), ("train_nomask", dict(mask_encoder=False))])
def test_attention_variants_of_the_executed_reference(R, tag, flags):
    """attention.py:72-92 (smoothing normalisation), :220-224 (state = last alignments when cumulative_weights is off), :140-141
    (mask_encoder off: neither memory nor scores masked). The product rejects these flags; the oracle carries them for later kernels."""
    hp = _hp(R, predict_linear=False, mask_decoder=False, **flags)
    params = _params(R, drop=("CBHG", "cbhg"))
    ids, in_len, mel, stop, lin, tgt_len = _inputs(R)
    out = ot.forward(params, ids, in_len, mel, hp, training=True, masks=_masks(R, tag, True, hp))
    _check_outputs(R, tag, out, True)
    a = out["alignments"]
    assert torch.allclose(a.sum(-1), torch.ones_like(a.sum(-1)), atol=1e-5)
    pad = (torch.arange(a.shape[2])[None, :] >= in_len[:, None])                       # padded encoder positions
    if tag == "train_nomask":
        assert float(a[1][:, pad[1]].min()) > 0                                         # attention mass leaks onto the padding
    else:
        assert float(a[1][:, pad[1]].abs().max()) == 0
    base = ot.forward(params, ids, in_len, mel, _hp(R, predict_linear=False, mask_decoder=False), training=True, masks=_masks(R, tag, True, hp))
    assert float((base["alignments"] - a).abs().max()) > 1e-3                          # the flags do change the result


def _params_r2(R):
    params = _params(R, drop=("CBHG", "cbhg", "linear_transform_projection", "stop_token_projection"))
    for k in R.files:
        if k.startswith("r2_var/"):
            params[t2_tf_bundle.engine_name("Tacotron_model/" + k[len("r2_var/"):])] = torch.from_numpy(R[k]).clone()
    return params


def test_reduction_factor_two_of_the_executed_reference(R):
    """outputs_per_step = 2 (tacotron.py:141-143,176-177; helpers.py:77,113-124,40-50): r frames per decoder step, the LAST frame of a
    group is what gets fed (teacher forcing and free running), stop tokens come r per step. Rejected by the product; oracle only."""
    hp = _hp(R, predict_linear=False, mask_decoder=False, outputs_per_step=2)
    params = _params_r2(R)
    assert params["linear_transform_projection/kernel"].shape[1] == 2 * hp.num_mels and params["stop_token_projection/kernel"].shape[1] == 2
    ids, in_len, mel, stop, lin, tgt_len = _inputs(R)
    out = ot.forward(params, ids, in_len, mel, hp, training=True, masks=_masks(R, "train_r2", True, hp))
    assert out["alignments"].shape[1] == mel.shape[1] // 2
    _check_outputs(R, "train_r2", out, True)
    total, parts = ot.loss_fn(out, mel, stop, params, hp, tgt_len)
    assert abs(float(total) - float(R["train_r2_loss"])) <= 1e-5 * abs(float(R["train_r2_loss"]))
    params["stop_token_projection/bias"] = torch.from_numpy(R["r2_synth_stop_bias"]).clone()
    pm = [torch.from_numpy(R["synth_r2_mask_prenet_drop_%d" % i]) for i in range(len(hp.prenet_layers))]
    steps = pm[0].shape[1]
    syn = ot.synthesize(params, ids, in_len, hp, prenet_masks=[[m[:, t] for m in pm] for t in range(steps)])
    assert steps == hp.max_iters and syn["mel_outputs"].shape[1] == 2 * steps
    _check_outputs(R, "synth_r2", syn, False)


def test_scheduled_teacher_forcing_of_the_executed_reference(R):
    """helpers.py:86-128,135-169: ratio = cosine decay of the initial ratio (0.5 at the fixture's global step), ONE uniform draw per
    decoder step for the whole batch chooses between the ground-truth frame and the model's own last frame, and the loss back-propagates
    through the fed-back frames. Rejected by the product; oracle only."""
    hp = _hp(R, predict_linear=False, mask_decoder=False, tacotron_teacher_forcing_mode="scheduled")
    step = int(R["train_sched_global_step"])
    ratio = ot.teacher_forcing_ratio(hp, step)
    assert abs(ratio - float(R["train_sched_ratio"])) < 1e-6 and abs(ratio - 0.5) < 1e-6
    assert ot.teacher_forcing_ratio(hp, 5000) == 1.0 and abs(ot.teacher_forcing_ratio(hp, 10 ** 6)) < 1e-7 and ot.teacher_forcing_ratio(hp, step, gta=True) == 1.0
    params = {k: v.requires_grad_(ot.is_trainable(k)) for k, v in _params(R, drop=("CBHG", "cbhg")).items()}
    ids, in_len, mel, stop, lin, tgt_len = _inputs(R)
    draws = torch.from_numpy(R["train_sched_tf_draws"])
    out = ot.forward(params, ids, in_len, mel, hp, training=True, masks=_masks(R, "train_sched", True, hp), tf_ratio=ratio, tf_draws=draws)
    _check_outputs(R, "train_sched", out, True)
    total, _ = ot.loss_fn(out, mel, stop, params, hp, tgt_len)
    assert abs(float(total.detach()) - float(R["train_sched_loss"])) <= 1e-5 * abs(float(R["train_sched_loss"]))
    total.backward()
    for eng, key in (("inputs_embedding", "train_sched_grad_embedding"), ("decoder_prenet/dense_1/kernel", "train_sched_grad_prenet")):
        ref = R[key]
        assert np.abs(params[eng].grad.numpy() - ref).max() <= 2e-4 * np.abs(ref).max(), eng
    forced = ot.forward(params, ids, in_len, mel, hp, training=True, masks=_masks(R, "train_sched", True, hp))
    assert float((forced["decoder_output"] - out["decoder_output"]).detach().abs().max()) > 1e-3      # feeding predictions changes the result


@pytest.mark.parametrize("tag,kind,win", [("synth_window", "window", 3), ("synth_mono", "monotonic", 2)])
def test_synthesis_attention_constraints_of_the_executed_reference(R, tag, kind, win):
    """attention.py:201-214 with synthesis_constraint on: energies outside the allowed span around the previous argmax are replaced by
    -2^32 + 1 before the (masked) softmax. Rejected by the product; oracle only."""
    hp = _hp(R, predict_linear=True, synthesis_constraint=True, synthesis_constraint_type=kind, attention_win_size=win)
    params = _params(R)
    ids, in_len = _inputs(R)[:2]
    pm = [torch.from_numpy(R["%s_mask_prenet_drop_%d" % (tag, i)]) for i in range(len(hp.prenet_layers))]
    steps = pm[0].shape[1]
    out = ot.synthesize(params, ids, in_len, hp, prenet_masks=[[m[:, t] for m in pm] for t in range(steps)] + [None] * 4)
    assert out["mel_outputs"].shape[1] == steps
    _check_outputs(R, tag, out, False)
    free = ot.synthesize(params, ids, in_len, _hp(R, predict_linear=True), prenet_masks=[[m[:, t] for m in pm] for t in range(steps)] + [None] * 4)
    n = min(free["alignments"].shape[1], steps)
    assert float((free["alignments"][:, :n] - out["alignments"][:, :n]).abs().max()) > 1e-3      # the constraint bites in this fixture
    a = out["alignments"]                                                                            # [B, steps, T_in]
    assert float(a[:, 1:].sum(-1).min()) > 0.999
