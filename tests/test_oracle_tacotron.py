"""Pins the Tacotron oracle against the invariants the reference implies."""
import numpy as np
import torch

from hparams import hparams
from oracle import tacotron as ot


def small_hp(**kw):
    hp = hparams.copy()
    hp.parse("embedding_dim=16,enc_conv_channels=16,encoder_lstm_units=8,attention_dim=8,attention_filters=4,"
             "attention_kernel=[5],prenet_layers=[12,12],decoder_lstm_units=16,postnet_channels=16,num_mels=6,"
             "enc_conv_num_layers=2,postnet_num_layers=3,predict_linear=False")
    for k, v in kw.items():
        hp.set_hparam(k, v)
    return hp


def _batch(hp, B=3, T_in=9, T_out=7, seed=0):
    g = torch.Generator().manual_seed(seed)
    inputs = torch.randint(2, 66, (B, T_in), generator=g)
    lens = torch.tensor([T_in, T_in - 3, T_in - 5][:B])
    for b in range(B):
        inputs[b, lens[b]:] = 0
    mel = torch.randn(B, T_out, hp.num_mels, generator=g).clamp(-4, 4)
    stop = torch.zeros(B, T_out)
    stop[:, -2:] = 1
    return inputs, lens, mel, stop


def test_param_count_matches_survey():
    hp = hparams.copy()
    hp.set_hparam("predict_linear", False)
    n = sum(int(np.prod(s)) for k, s in ot.param_shapes(hp).items() if ot.is_trainable(k))
    assert abs(n / 1e6 - 27.19) < 0.02
    reg = [k for k in ot.param_shapes(hp) if ot.is_regularized(k)]
    assert "attention/memory_layer/kernel" in reg and "decoder_prenet/dense_1/kernel" in reg
    assert "encoder_convolutions/conv_layer_1/gamma" in reg and "postnet_convolutions/conv_layer_5/kernel" in reg
    assert not any(("LSTM" in k) or ("_projection" in k) or ("bias" in k) or ("embedding" in k) for k in reg)


def test_alignments_are_masked_distributions_and_deterministic_without_noise():
    hp = small_hp(tacotron_dropout_rate=0.0, tacotron_zoneout_rate=0.0)
    p = ot.init_params(hp, seed=1, random_bias=True)
    inputs, lens, mel, stop = _batch(hp)
    a = ot.forward(p, inputs, lens, mel, hp)
    b = ot.forward(p, inputs, lens, mel, hp)
    assert torch.equal(a["mel_outputs"], b["mel_outputs"])
    al = a["alignments"]
    assert torch.allclose(al.sum(-1), torch.ones_like(al.sum(-1)), atol=1e-5)
    for i in range(3):
        assert al[i, :, lens[i]:].abs().sum() == 0          # scores past the input length are -inf
    assert a["decoder_output"].shape == mel.shape and a["stop_logits"].shape == stop.shape
    assert a["decoder_output"].max() <= 4.0 and a["decoder_output"].min() >= -4.1


def test_bilstm_ignores_padding_values_beyond_length():
    hp = small_hp(tacotron_dropout_rate=0.0, tacotron_zoneout_rate=0.0)
    p = ot.init_params(hp, seed=2)
    x = torch.randn(2, 6, 16)
    lens = torch.tensor([6, 3])
    y1 = ot.encoder_rnn(x, lens, p, hp, True)
    x2 = x.clone()
    x2[1, 3:] = 99.0
    y2 = ot.encoder_rnn(x2, lens, p, hp, True)
    assert torch.allclose(y1, y2)
    assert y1[1, 3:].abs().max() == 0


def test_zoneout_and_losses():
    prev, new = torch.zeros(4, 5), torch.ones(4, 5)
    assert torch.allclose(ot.zoneout(prev, new, 0.1, False), torch.full((4, 5), 0.9))
    m = torch.tensor([[1., 0, 1, 0, 1]] * 4)
    assert torch.equal(ot.zoneout(prev, new, 0.1, True, m), m)
    hp = small_hp(tacotron_dropout_rate=0.0, tacotron_zoneout_rate=0.0)
    p = ot.init_params(hp, seed=3)
    inputs, lens, mel, stop = _batch(hp)
    loss, grads, out, parts = ot.train_step(p, inputs, lens, mel, stop, hp)
    assert np.isfinite(loss.item()) and all(torch.isfinite(g).all() for g in grads.values())
    assert abs(parts["stop"].item() - np.log(2)) < 0.3
    assert grads["attention/attention_variable_projection"].abs().sum() > 0
    st = {}
    p2 = {k: v.clone() for k, v in p.items()}
    ot.adam_step(p2, grads, st, hp, 0)
    assert ot.learning_rate(hp, 0) == 1e-3 and abs(ot.learning_rate(hp, 40000 + 18000) - 5e-4) < 1e-9
    assert ot.learning_rate(hp, 10 ** 7) == 1e-4


def test_free_running_equals_teacher_forcing_on_its_own_outputs():
    """TacoTestHelper feeds frame t-1 back; TacoTrainingHelper with targets := those frames must reproduce the run
    (is_training = False on both sides: same batch-norm statistics and zoneout blend)."""
    hp = small_hp(tacotron_zoneout_rate=0.1, tacotron_dropout_rate=0.0)
    params = ot.init_params(hp, seed=5, random_bias=True)
    params["stop_token_projection/bias"] = torch.full((1,), -6.0)
    g = torch.Generator().manual_seed(5)
    inputs = torch.randint(2, 66, (2, 17), generator=g)
    lens = torch.tensor([17, 11])
    syn = ot.synthesize(params, inputs, lens, hp, max_iters=9)
    assert syn["mel_outputs"].shape == (2, 9, hp.num_mels)
    assert syn["decoder_output"].abs().max() < hp.max_abs_value          # nothing clipped: raw frame == decoder output
    tf_ = ot.forward(params, inputs, lens, syn["decoder_output"], hp, training=False)
    assert (tf_["decoder_output"] - syn["decoder_output"]).abs().max() < 1e-5
    assert (tf_["mel_outputs"] - syn["mel_outputs"]).abs().max() < 1e-5
    assert (torch.sigmoid(tf_["stop_logits"]) - syn["stop_token_prediction"]).abs().max() < 1e-6
    # stop rule: a +6 bias finishes at the first step and keeps that step's frame
    params["stop_token_projection/bias"] = torch.full((1,), 6.0)
    assert ot.synthesize(params, inputs, lens, hp, max_iters=9)["mel_outputs"].shape[1] == 1
