"""The reference's model-level Python surface (`create_model` / `initialize` / `add_loss` / `add_optimizer`) driven like
wavenet_vocoder/train.py:168-303 and tacotron/train.py:190-331 drive it: the loss must fall on a fixed batch."""
import pytest
import torch

from hparams import hparams

pytestmark = pytest.mark.gpu


def test_wavenet_create_model_train_eval_synthesize():
    from wavenet_vocoder.models import create_model
    from wavenet_vocoder.util import mulaw_quantize
    hp = hparams.copy()
    hp.parse("layers=4,stacks=2,residual_channels=128,gate_channels=256,skip_out_channels=128,upsample_scales=[4,4],hop_size=16,"
             "input_type=mulaw-quantize,quantize_channels=256,out_channels=256,wavenet_dropout=0.05")
    with pytest.raises(RuntimeError):
        bad = hp.copy()
        bad.set_hparam("out_channels", 30)
        create_model("WaveNet", bad)
    model = create_model("WaveNet", hp)
    g = torch.Generator().manual_seed(0)
    B, T = 2, 512
    wav = (torch.sin(torch.arange(T) * 0.05)[None] * 0.5 + 0.02 * torch.randn(B, T, generator=g)).clamp(-1, 1)
    idx = torch.from_numpy(mulaw_quantize(wav.numpy())).cuda()
    x = torch.nn.functional.one_hot(idx.long(), 256).float().transpose(1, 2)        # feeder layout [B, 256, T]
    c = torch.rand(B, 80, T // 16, generator=g).cuda()
    lengths = torch.tensor([T, T - 40]).cuda()
    losses = []
    for step in range(30):
        model.initialize(idx.unsqueeze(-1), c, None, lengths, x=x)
        losses.append(float(model.add_loss()))
        lr = model.add_optimizer(step)
    assert lr == pytest.approx(hp.wavenet_learning_rate * hp.wavenet_decay_rate ** (29 / hp.wavenet_decay_steps))
    assert losses[-1] < 0.9 * losses[0], losses
    model.initialize(idx, c, None, lengths)                                         # evaluation: y only
    assert float(model.add_loss()) < losses[0] and model.is_evaluating
    # synthesis (wavenet.py:408-456): c is [batch, frames, cin]; the length is Tc * hop whatever synthesis_length says; the
    # published y_hat is the DECODED waveform (inv_mulaw_quantize of the sampled indices), float in [-1, 1]
    model.initialize(None, c[:1, :, :4].transpose(1, 2).contiguous(), None, None, synthesis_length=999)
    out = model.tower_y_hat[0]
    assert out.shape == (1, 64) and out.dtype == torch.float32 and float(out.min()) >= -1.0 and float(out.max()) <= 1.0
    with pytest.raises(ValueError):
        model.initialize(None, c[:1, :, :4], None, None)                            # channels-first c is rejected like the reference does
    # batches of other lengths (the reference feeder pads each batch to its own maximum): lengths that are not a multiple of the
    # 8-hop bucket are padded internally, results are sliced back, old engines are evicted (unbounded engine cache)
    hp2 = hp.copy()
    hp2.add_hparam("keep_logits", True)
    m2 = create_model("WaveNet", hp2)
    for T2 in (496, 256, 384, 128, 496):
        m2.initialize(idx[:, :T2].unsqueeze(-1), c[:, :, :T2 // 16], None, torch.tensor([T2, T2 - 9]).cuda(), x=x[:, :, :T2])
        assert m2.tower_y_hat[0].shape == (B, 256, T2) and m2.tower_upsampled_local_features[0].shape[1] == T2
        l_a = float(m2.add_loss())
        m2.add_optimizer()
        assert l_a == l_a and len(m2._engines) <= 3
    # padding to the bucket does not change the loss: T = 496 (padded to 512 inside the drop-in) vs the exact-size engine of the
    # product API, same variables, dropout off
    from t2_import import t2
    hp3 = hp.copy()
    hp3.set_hparam("wavenet_dropout", 0.0)
    eng = t2.wavenet.WaveNet(hp3, B, 496)
    eng.init_variables(seed=3)
    m3 = create_model("WaveNet", hp3)
    m3.load_variables(eng.export_params())
    len496 = torch.tensor([496, 400]).int().cuda()
    eng.forward(idx[:, :496].int().contiguous(), c[:, :, :31].contiguous(), idx[:, :496].int().contiguous(), len496)
    m3.initialize(idx[:, :496].unsqueeze(-1), c[:, :, :31], None, len496, x=x[:, :, :496])
    torch.cuda.synchronize()
    assert abs(float(m3.add_loss()) - eng.loss_value()) < 2e-5


def test_tacotron_create_model_train_gta_synthesize():
    from tacotron.models import create_model
    hp = hparams.copy()
    hp.parse("predict_linear=False,enc_conv_channels=256,embedding_dim=256,encoder_lstm_units=128,decoder_lstm_units=256,"
             "postnet_channels=256,prenet_layers=[128,128],attention_dim=128,max_iters=40,tacotron_initial_learning_rate=0.01,tacotron_decay_learning_rate=False")
    model = create_model("Tacotron", hp)
    g = torch.Generator().manual_seed(0)
    B, T_in, T_out = 4, 20, 32
    inputs = torch.randint(2, 66, (B, T_in), generator=g).cuda()
    lens = torch.tensor([20, 18, 15, 11]).cuda()
    mel = (torch.randn(B, T_out, 80, generator=g) * 0.5 - 1).clamp(-4, 4).cuda()
    stop = torch.zeros(B, T_out).cuda()
    stop[:, -2:] = 1
    with pytest.raises(ValueError):
        model.initialize(inputs, lens, mel_targets=mel)                              # targets without stop tokens
    losses = []
    for step in range(25):
        model.initialize(inputs, lens, mel, stop, global_step=step, is_training=True)
        losses.append(float(model.add_loss()))
        model.add_optimizer(step)
    assert losses[-1] < 0.9 * losses[0], losses
    assert model.tower_mel_outputs[0].shape == (B, T_out, 80) and model.tower_alignments[0].shape == (B, T_in, T_out)
    model.initialize(inputs, lens, mel, gta=True)                                    # GTA: teacher forced, inference statistics
    assert model.tower_mel_outputs[0].shape == (B, T_out, 80)
    model.initialize(inputs, lens)                                                   # free-running synthesis
    T = model.tower_mel_outputs[0].shape[1]
    assert 1 <= T <= 40 and model.tower_stop_token_prediction[0].shape == (B, T)
    assert float(model.tower_stop_token_prediction[0].min()) >= 0 and float(model.tower_stop_token_prediction[0].max()) <= 1
