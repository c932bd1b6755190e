"""ORACLE (test infrastructure, not product code): fp32 PyTorch-CPU restatement of the reference's Tacotron-2
mel-spectrogram predictor (training / evaluation / GTA / free-running graphs) and of the CBHG post-processing net +
linear head (predict_linear = True: tacotron.py:203-219, modules.py:4-78,457-485).

Only tests/, __graft_entry__.smoke() and bench / tools CPU-baseline legs may import this.

Follows, function by function:
  Tacotron.initialize / add_loss / add_optimizer          tacotron/models/tacotron.py:104-200, 273-369, 371-463
  conv1d, EncoderConvolutions, EncoderRNN, ZoneoutLSTMCell  tacotron/models/modules.py:379-391, 145-174, 177-217, 81-142
  Prenet, DecoderRNN, FrameProjection, StopProjection      tacotron/models/modules.py:220-342
  Postnet                                                  tacotron/models/modules.py:345-376
  LocationSensitiveAttention, _location_sensitive_score    tacotron/models/attention.py:38-70, 95-226
  TacotronDecoderCell.__call__ (step order)                tacotron/models/Architecture_wrappers.py:169-213
  TacoTrainingHelper (go frame, teacher forcing)           tacotron/models/helpers.py:62-128
TF-layer semantics (LSTMCell gate order i,j,f,o + forget_bias 1, BahdanauAttention memory / score masking, batch-norm
eps 1e-3 with biased batch variance, 'same' conv padding, tf.losses.mean_squared_error) are restated from the public
TF 1.x definitions.

PINNING, two levels.
(1) Elementary-op code (tests/golden/make_reference_vectors.py, tests/test_reference_pinned.py): MaskedMSE, MaskedSigmoidCrossEntropy,
MaskedLinearLoss, sequence_mask (modules.py:400-485), the location-sensitive score and the smoothing normalisation
(attention.py:38-92), the learning-rate schedule (tacotron.py:439-463), the feeder padding helpers - the reference's source runs AS IS.
(2) The whole graph (tests/golden/make_reference_graph_vectors.py, tests/test_reference_graph.py): the reference's
`Tacotron.initialize()` + `add_loss()` are EXECUTED - tacotron.py, modules.py, attention.py, Architecture_wrappers.py, helpers.py,
custom_decoder.py unchanged - on a stand-in for the tf.layers / rnn_cell / seq2seq classes they compose (tests/golden/tf_shim_graph.py),
in training (every dropout / zoneout mask recorded and injected here), masked-loss training, evaluation, GTA, free-running
synthesis, and one add_optimizer step (LR schedule, global-norm clip, Adam, batch-norm moving averages == adam_step here). forward / loss_fn / synthesize / linear_head reproduce the executed reference to <= 4e-6 (outputs), 1e-7 (loss terms) and
1e-5 relative (d loss / d variable, all 102 trainable variables), and the variable names the reference's scopes generate equal
t2_tf_bundle.tacotron_tf_name over the parameter table. This pins the reference's COMPOSITION: layer order, scopes, activation /
batch-norm / dropout placement, the zoneout wrapper and its un-zoned output, decoder-cell wiring, helpers, stop rule, CBHG, loss terms,
regularisation filter.
Variants the CUDA path rejects are carried here too and pinned the same way, so that their kernels have a checker when they are
written: outputs_per_step > 1, scheduled teacher forcing (cosine-decayed ratio, one draw per step, gradients through fed-back frames),
smoothing normalisation, non-cumulative attention state, un-masked encoder memory, the synthesis window / monotonic constraints.
STILL A RESTATEMENT: the primitives under that composition (Dense, Conv1D 'same', BatchNormalization eps 1e-3 / biased variance,
LSTMCell i,j,f,o + forget_bias 1, GRUCell, dynamic_rnn length handling, BahdanauAttention memory / score masking, dynamic_decode) - in
the stand-in as in this file they follow the public TF 1.x definitions; TensorFlow itself cannot run here.
Known answers on top: parameter counts 27.19 M / 29.02 M with the CBHG head, alignments are a masked probability distribution,
zero-length-padding invariance of the encoder (tests/test_oracle_tacotron.py).
"""
import math

import torch
import torch.nn.functional as F

N_SYMBOLS = 66  # tacotron/utils/symbols.py:9-17


def param_shapes(hp):
    E, C = hp.embedding_dim, hp.enc_conv_channels
    H, D, A = hp.encoder_lstm_units, hp.decoder_lstm_units, hp.attention_dim
    M = hp.num_mels * hp.outputs_per_step
    P = hp.postnet_channels
    k_enc, k_post, k_att = hp.enc_conv_kernel_size[0], hp.postnet_kernel_size[0], hp.attention_kernel[0]
    sh = {"inputs_embedding": (N_SYMBOLS, E)}
    cin = E
    for i in range(hp.enc_conv_num_layers):
        p = "encoder_convolutions/conv_layer_%d/" % (i + 1)
        sh[p + "kernel"] = (k_enc, cin, C)
        sh[p + "bias"] = (C,)
        for n in ("gamma", "beta", "moving_mean", "moving_variance"):
            sh[p + n] = (C,)
        cin = C
    for d in ("fw", "bw"):
        sh["encoder_LSTM/%s/kernel" % d] = (C + H, 4 * H)
        sh["encoder_LSTM/%s/bias" % d] = (4 * H,)
    sh["attention/memory_layer/kernel"] = (2 * H, A)
    sh["attention/query_layer/kernel"] = (D, A)
    sh["attention/location_features_convolution/kernel"] = (k_att, 1, hp.attention_filters)
    sh["attention/location_features_convolution/bias"] = (hp.attention_filters,)
    sh["attention/location_features_layer/kernel"] = (hp.attention_filters, A)
    sh["attention/attention_variable_projection"] = (A,)
    sh["attention/attention_bias"] = (A,)
    pin = hp.num_mels
    for i, n in enumerate(hp.prenet_layers):
        sh["decoder_prenet/dense_%d/kernel" % (i + 1)] = (pin, n)
        sh["decoder_prenet/dense_%d/bias" % (i + 1)] = (n,)
        pin = n
    lin = pin + 2 * H
    for i in range(hp.decoder_layers):
        sh["decoder_LSTM/cell_%d/kernel" % (i + 1)] = (lin + D, 4 * D)
        sh["decoder_LSTM/cell_%d/bias" % (i + 1)] = (4 * D,)
        lin = D
    sh["linear_transform_projection/kernel"] = (D + 2 * H, M * hp.outputs_per_step)        # r frames per decoder step (tacotron.py:141)
    sh["linear_transform_projection/bias"] = (M * hp.outputs_per_step,)
    sh["stop_token_projection/kernel"] = (D + 2 * H, hp.outputs_per_step)
    sh["stop_token_projection/bias"] = (hp.outputs_per_step,)
    cin = hp.num_mels
    for i in range(hp.postnet_num_layers):
        p = "postnet_convolutions/conv_layer_%d/" % (i + 1)
        sh[p + "kernel"] = (k_post, cin, P)
        sh[p + "bias"] = (P,)
        for n in ("gamma", "beta", "moving_mean", "moving_variance"):
            sh[p + n] = (P,)
        cin = P
    sh["postnet_projection/kernel"] = (P, hp.num_mels)
    sh["postnet_projection/bias"] = (hp.num_mels,)
    if hp.predict_linear:                       # CBHG post-processing net + linear projection (tacotron.py:203-219, modules.py:19-78)
        cc, hu, ru = hp.cbhg_conv_channels, hp.cbhg_highway_units, hp.cbhg_rnn_units

        def conv(prefix, k, ci, co):
            sh[prefix + "kernel"] = (k, ci, co)
            sh[prefix + "bias"] = (co,)
            for n in ("gamma", "beta", "moving_mean", "moving_variance"):
                sh[prefix + n] = (co,)
        for k in range(1, hp.cbhg_kernels + 1):
            conv("CBHG_postnet/conv_bank/conv1d_%d/" % k, k, hp.num_mels, cc)
        conv("CBHG_postnet/proj1/", hp.cbhg_projection_kernel_size, hp.cbhg_kernels * cc, hp.cbhg_projection)
        conv("CBHG_postnet/proj2/", hp.cbhg_projection_kernel_size, hp.cbhg_projection, hp.num_mels)
        if hp.num_mels != hu:                   # modules.py:62-63
            sh["CBHG_postnet/dense/kernel"] = (hp.num_mels, hu)
            sh["CBHG_postnet/dense/bias"] = (hu,)
        for i in range(hp.cbhg_highwaynet_layers):
            for n in ("H", "T"):
                sh["CBHG_postnet/highwaynet_%d/%s/kernel" % (i + 1, n)] = (hu, hu)
                sh["CBHG_postnet/highwaynet_%d/%s/bias" % (i + 1, n)] = (hu,)
        for d in ("forward", "backward"):       # tf.nn.rnn_cell.GRUCell: gates kernel [in + n, 2n], candidate kernel [in + n, n]
            sh["CBHG_postnet/%s_RNN/gates/kernel" % d] = (hu + ru, 2 * ru)
            sh["CBHG_postnet/%s_RNN/gates/bias" % d] = (2 * ru,)
            sh["CBHG_postnet/%s_RNN/candidate/kernel" % d] = (hu + ru, ru)
            sh["CBHG_postnet/%s_RNN/candidate/bias" % d] = (ru,)
        sh["cbhg_linear_specs_projection/kernel"] = (2 * ru, hp.num_freq)
        sh["cbhg_linear_specs_projection/bias"] = (hp.num_freq,)
    return sh


NON_TRAINABLE = ("moving_mean", "moving_variance")


def is_trainable(name):
    return not name.endswith(NON_TRAINABLE)


def is_regularized(name):
    """tacotron.py:343-345: every trainable variable whose name has none of these substrings."""
    if not is_trainable(name):
        return False
    return not any(s in name for s in ("bias", "Bias", "_projection", "inputs_embedding", "RNN", "LSTM"))


def init_params(hp, seed=None, random_bias=False):
    gen = torch.Generator().manual_seed(hp.tacotron_random_seed if seed is None else seed)
    params = {}
    for name, shape in param_shapes(hp).items():
        if name.endswith("gamma") or name.endswith("moving_variance"):
            params[name] = torch.ones(shape)
        elif name.endswith(("beta", "moving_mean")):
            params[name] = torch.zeros(shape)
        elif name.endswith("bias"):
            params[name] = torch.randn(shape, generator=gen) * 0.1 if random_bias else torch.zeros(shape)
            if not random_bias and name.endswith("RNN/gates/bias"):
                params[name] = torch.ones(shape)            # GRUCell gate bias initialiser 1.0
            if not random_bias and "/T/bias" in name:
                params[name] = -torch.ones(shape)           # HighwayNet T gate bias -1 (modules.py:10)
        else:
            if len(shape) == 1:
                fan_in = fan_out = shape[0]
            elif len(shape) == 2:
                fan_in, fan_out = shape
            else:
                fan_in, fan_out = shape[0] * shape[1], shape[0] * shape[2]
            lim = math.sqrt(6.0 / (fan_in + fan_out))
            params[name] = (torch.rand(shape, generator=gen) * 2 - 1) * lim
    return params


# ---------------------------------------------------------------------------------------------------------
# layers
# ---------------------------------------------------------------------------------------------------------
def _clip_low(hp):
    """tacotron.py:89,176,199: T2_output_range[0] - lower_bound_decay, the range starting at -max_abs_value (symmetric mels) or 0"""
    return (-hp.max_abs_value if hp.symmetric_mels else 0.0) - hp.lower_bound_decay


def _reg_weight(hp):
    """tacotron.py:334-338"""
    if getattr(hp, "tacotron_scale_regularization", False):
        return hp.tacotron_reg_weight * (1.0 / (2 * hp.max_abs_value) if hp.symmetric_mels else 1.0 / hp.max_abs_value)
    return hp.tacotron_reg_weight


def conv_block(x, params, prefix, activation, training, drop_rate, drop_mask=None, stats_out=None):
    """modules.py:379-391 with batch_norm_position='after': conv('same') -> activation -> BN -> dropout.
    x [B, T, Cin] channels-last."""
    k = params[prefix + "kernel"]  # [kw, in, out]
    kw = k.shape[0]
    xp = F.pad(x.transpose(1, 2), ((kw - 1) // 2, kw - 1 - (kw - 1) // 2))     # 'same': the extra pad of an even kernel goes right
    y = F.conv1d(xp, k.permute(2, 1, 0).contiguous(), params[prefix + "bias"]).transpose(1, 2)
    if activation == "relu":
        y = F.relu(y)
    elif activation == "tanh":
        y = torch.tanh(y)
    if training:
        mean = y.mean(dim=(0, 1))
        var = y.var(dim=(0, 1), unbiased=False)
        if stats_out is not None:
            stats_out[prefix] = (mean.detach(), var.detach())
    else:
        mean, var = params[prefix + "moving_mean"], params[prefix + "moving_variance"]
    y = (y - mean) / torch.sqrt(var + 1e-3) * params[prefix + "gamma"] + params[prefix + "beta"]
    if training and drop_rate > 0:
        if drop_mask is None:
            drop_mask = (torch.rand_like(y) >= drop_rate).float() / (1 - drop_rate)
        y = y * drop_mask
    return y


def lstm_cell(x, c, h, kernel, bias):
    """tf.nn.rnn_cell.LSTMCell: z = [x, h] W + b; i, j, f, o; forget_bias = 1."""
    z = torch.cat([x, h], dim=-1) @ kernel + bias
    i, j, f, o = z.chunk(4, dim=-1)
    new_c = torch.sigmoid(f + 1.0) * c + torch.sigmoid(i) * torch.tanh(j)
    new_h = torch.sigmoid(o) * torch.tanh(new_c)
    return new_c, new_h


def zoneout(prev, new, rate, training, mask=None):
    """modules.py:133-138."""
    if training:
        if rate == 0:
            return new
        if mask is None:
            mask = (torch.rand_like(new) >= rate).float()
        return mask * (new - prev) + prev
    return (1 - rate) * new + rate * prev


def encoder_rnn(x, lengths, params, hp, training, zmasks=None):
    """bidirectional_dynamic_rnn of ZoneoutLSTMCell with sequence_length (modules.py:207-217): past the length the
    output is zero and the state is carried; the backward direction walks the reversed (per length) sequence."""
    B, T, _ = x.shape
    H = hp.encoder_lstm_units
    z = hp.tacotron_zoneout_rate
    outs = []
    for d in ("fw", "bw"):
        K, b = params["encoder_LSTM/%s/kernel" % d], params["encoder_LSTM/%s/bias" % d]
        c = torch.zeros(B, H)
        h = torch.zeros(B, H)
        out = [None] * T
        order = range(T) if d == "fw" else range(T - 1, -1, -1)
        for t in order:
            live = (t < lengths).float().unsqueeze(-1)
            nc, nh = lstm_cell(x[:, t], c, h, K, b)
            mc = zmasks[(d, "c", t)] if zmasks else None
            mh = zmasks[(d, "h", t)] if zmasks else None
            zc = zoneout(c, nc, z, training, mc)
            zh = zoneout(h, nh, z, training, mh)
            c = live * zc + (1 - live) * c
            h = live * zh + (1 - live) * h
            out[t] = live * nh
        outs.append(torch.stack(out, dim=1))
    return torch.cat(outs, dim=-1)


def prenet(x, params, hp, drop_masks=None):
    """modules.py:240-251: dropout is ALWAYS on."""
    for i in range(len(hp.prenet_layers)):
        x = F.relu(x @ params["decoder_prenet/dense_%d/kernel" % (i + 1)] + params["decoder_prenet/dense_%d/bias" % (i + 1)])
        r = hp.tacotron_dropout_rate
        if r > 0:
            m = drop_masks[i] if drop_masks else (torch.rand_like(x) >= r).float() / (1 - r)
            x = x * m
    return x


def location_sensitive_score(W_query, W_fil, W_keys, v_a, b_a):
    """attention.py:38-70: energy = sum_units v_a * tanh(W_keys + W_query + W_fil + b_a). [B,1,A], [B,T,A], [B,T,A] -> [B,T]."""
    return (v_a * torch.tanh(W_keys + W_query + W_fil + b_a)).sum(-1)


def _constraint_mask(prev_max, T_in, win, kind):
    """attention.py:201-214: positions the synthesis constraint forbids, from the previous step's argmax. 'monotonic': before the
    previous maximum or more than `win` ahead of it; 'window': outside a window of `win` positions centred on it."""
    pos = torch.arange(T_in)[None, :]
    if kind == "monotonic":
        before = pos < prev_max[:, None]
        after = pos >= T_in - (T_in - win - prev_max)[:, None]                     # reversed sequence_mask(Tx - win - prev_max)
    else:
        assert kind == "window"
        before = pos < (prev_max - (win // 2 + (win % 2 != 0)))[:, None]
        after = pos >= T_in - (T_in - win // 2 - prev_max)[:, None]
    return before | after


def attention_step(query, cum, keys, values, mask, params, smoothing=False, constraint=None):
    """attention.py:169-226 + _compute_attention :10-35. query [B, D]; cum [B, T_in] = the attention state (cumulated alignments, or
    the previous alignments with cumulative_weights=False); mask None = mask_encoder False; returns (context, alignments)."""
    pq = (query @ params["attention/query_layer/kernel"]).unsqueeze(1)              # [B, 1, A]
    kf = params["attention/location_features_convolution/kernel"]                      # [31, 1, 32]
    f = F.conv1d(cum.unsqueeze(1), kf.permute(2, 1, 0).contiguous(), params["attention/location_features_convolution/bias"],
                 padding=(kf.shape[0] - 1) // 2).transpose(1, 2)                       # [B, T_in, 32]
    pl = f @ params["attention/location_features_layer/kernel"]                        # [B, T_in, A]
    e = location_sensitive_score(pq, pl, keys, params["attention/attention_variable_projection"],
                                 params["attention/attention_bias"])                   # [B, T_in]
    if constraint is not None:                                                         # (prev_max [B] int, win, kind), synthesis only
        e = torch.where(_constraint_mask(constraint[0], e.shape[1], constraint[1], constraint[2]), torch.full_like(e, float(-2 ** 32 + 1)), e)
    if mask is not None:
        e = torch.where(mask > 0, e, torch.full_like(e, -float("inf")))                # _maybe_mask_score
    if smoothing:                                                                      # attention.py:72-92
        a = torch.sigmoid(e) / torch.sigmoid(e).sum(dim=-1, keepdim=True)
    else:
        a = torch.softmax(e, dim=-1)
    ctx = torch.bmm(a.unsqueeze(1), values).squeeze(1)
    return ctx, a


def teacher_forcing_ratio(hp, global_step, training=True, evaluating=False, gta=False):
    """helpers.py:86-100,135-169: GTA always feeds the ground truth; evaluation with tacotron_natural_eval always feeds predictions;
    'scheduled' mode = init ratio until start_decay, then tf.train.cosine_decay of it towards alpha * init over decay_steps."""
    if gta:
        return 1.0
    if evaluating and hp.tacotron_natural_eval:
        return 0.0
    if hp.tacotron_teacher_forcing_mode != "scheduled" or not training:
        return float(hp.tacotron_teacher_forcing_ratio)
    init = hp.tacotron_teacher_forcing_init_ratio
    if hp.tacotron_teacher_forcing_final_ratio is not None:
        alpha = float(hp.tacotron_teacher_forcing_final_ratio / init)
    else:
        alpha = hp.tacotron_teacher_forcing_decay_alpha
    if global_step < hp.tacotron_teacher_forcing_start_decay:
        return float(init)
    step = min(global_step - hp.tacotron_teacher_forcing_start_decay, hp.tacotron_teacher_forcing_decay_steps)
    cosine = 0.5 * (1.0 + math.cos(math.pi * step / hp.tacotron_teacher_forcing_decay_steps))
    return float(init * ((1.0 - alpha) * cosine + alpha))


def forward(params, inputs, input_lengths, mel_targets, hp, training=True, masks=None, stats_out=None, tf_ratio=1.0, tf_draws=None):
    """Training / GTA graph with teacher forcing ratio 1. inputs [B, T_in] int64; mel_targets [B, T_out, num_mels].
    Returns dict(decoder_output, mel_outputs, stop_logits, alignments [B, T_out, T_in])."""
    masks = masks or {}
    B, T_in = inputs.shape
    T_out = mel_targets.shape[1]
    D = hp.decoder_lstm_units
    zr = hp.tacotron_zoneout_rate
    x = params["inputs_embedding"][inputs]
    for i in range(hp.enc_conv_num_layers):
        x = conv_block(x, params, "encoder_convolutions/conv_layer_%d/" % (i + 1), "relu", training,
                       hp.tacotron_dropout_rate, masks.get(("enc_drop", i)), stats_out)
    memory = encoder_rnn(x, input_lengths, params, hp, training, masks.get("enc_zone"))
    mask = (torch.arange(T_in)[None, :] < input_lengths[:, None]).float()
    if not getattr(hp, "mask_encoder", True):                                # attention.py:140-141: neither memory nor scores are masked
        mask = None
    values = memory * mask.unsqueeze(-1) if mask is not None else memory     # BahdanauAttention memory masking
    keys = values @ params["attention/memory_layer/kernel"]
    # TacoTrainingHelper (helpers.py:62-128): r = outputs_per_step frames per decoder step; step t consumes the go frame (t = 0) or the
    # LAST target frame of group t - 1 (targets[:, r-1::r], teacher forcing)
    r = hp.outputs_per_step
    assert T_out % r == 0, "the feeder pads targets to a multiple of outputs_per_step (feeder.py:240-243)"
    fed = mel_targets[:, r - 1::r, :]
    # per step ONE uniform draw for the whole batch decides between the ground truth and the model's own last frame
    # (helpers.py:121-124 tf.cond(random_uniform([]) < ratio, ...)); tf_draws[t] is the draw made at the END of step t
    forced = [True] * (T_out // r) if tf_ratio >= 1.0 and tf_draws is None else [bool(float(u) < tf_ratio) for u in tf_draws]
    pmask = masks.get("prenet_drop")
    c1 = torch.zeros(B, D); h1 = torch.zeros(B, D); c2 = torch.zeros(B, D); h2 = torch.zeros(B, D)
    ctx = torch.zeros(B, values.shape[-1])
    cum = torch.zeros(B, T_in)
    K1, b1 = params["decoder_LSTM/cell_1/kernel"], params["decoder_LSTM/cell_1/bias"]
    K2, b2 = params["decoder_LSTM/cell_2/kernel"], params["decoder_LSTM/cell_2/bias"]
    frames, stops, aligns = [], [], []
    zm = masks.get("dec_zone")
    frame_in = torch.zeros(B, hp.num_mels)                                    # go frame
    for t in range(T_out // r):
        pre_t = prenet(frame_in, params, hp, [m[:, t] for m in pmask] if pmask else None)
        nc1, nh1 = lstm_cell(torch.cat([pre_t, ctx], dim=-1), c1, h1, K1, b1)
        c1n = zoneout(c1, nc1, zr, training, zm[(1, "c", t)] if zm else None)
        h1n = zoneout(h1, nh1, zr, training, zm[(1, "h", t)] if zm else None)
        nc2, nh2 = lstm_cell(nh1, c2, h2, K2, b2)                            # layer 2 sees the UN-zoned output
        c2n = zoneout(c2, nc2, zr, training, zm[(2, "c", t)] if zm else None)
        h2n = zoneout(h2, nh2, zr, training, zm[(2, "h", t)] if zm else None)
        c1, h1, c2, h2 = c1n, h1n, c2n, h2n
        ctx, a = attention_step(nh2, cum, keys, values, mask, params, getattr(hp, "smoothing", False))
        cum = cum + a if getattr(hp, "cumulative_weights", True) else a      # attention.py:220-224
        pin = torch.cat([nh2, ctx], dim=-1)
        frames.append(pin @ params["linear_transform_projection/kernel"] + params["linear_transform_projection/bias"])
        frame_in = fed[:, t] if forced[t] else frames[-1][:, -hp.num_mels:]   # no stop_gradient: the loss back-propagates through it
        stops.append(pin @ params["stop_token_projection/kernel"] + params["stop_token_projection/bias"])
        aligns.append(a)
    decoder_output = torch.stack(frames, dim=1).reshape(B, -1, hp.num_mels)     # [B, T_out / r, r M] -> [B, T_out, M] (tacotron.py:176)
    stop_logits = torch.stack(stops, dim=1).reshape(B, -1)
    if hp.clip_outputs:
        decoder_output = torch.clamp(decoder_output, _clip_low(hp), hp.max_abs_value)
    y = decoder_output
    for i in range(hp.postnet_num_layers):
        act = "tanh" if i < hp.postnet_num_layers - 1 else None
        y = conv_block(y, params, "postnet_convolutions/conv_layer_%d/" % (i + 1), act, training,
                       hp.tacotron_dropout_rate, masks.get(("post_drop", i)), stats_out)
    residual = y @ params["postnet_projection/kernel"] + params["postnet_projection/bias"]
    mel_outputs = decoder_output + residual
    if hp.clip_outputs:
        mel_outputs = torch.clamp(mel_outputs, _clip_low(hp), hp.max_abs_value)
    out = {"decoder_output": decoder_output, "mel_outputs": mel_outputs, "stop_logits": stop_logits,
           "alignments": torch.stack(aligns, dim=1)}
    if hp.predict_linear and "cbhg_linear_specs_projection/kernel" in params:
        out["linear_outputs"] = linear_head(mel_outputs, params, hp, training, stats_out)
    return out


def gru_cell(x, h, gk, gb, ck, cb):
    """tf.nn.rnn_cell.GRUCell: r, u = split(sigmoid([x, h] Wg + bg)); c = tanh([x, r * h] Wc + bc); h' = u h + (1 - u) c."""
    r, u = torch.sigmoid(torch.cat([x, h], dim=-1) @ gk + gb).chunk(2, dim=-1)
    c = torch.tanh(torch.cat([x, r * h], dim=-1) @ ck + cb)
    return u * h + (1 - u) * c


def cbhg(x, params, hp, training, stats_out=None):
    """CBHG (modules.py:19-78) as the post-processing net (`CBHG_postnet`, tacotron.py:203-210): conv bank K = 1..8 (ReLU, BN, no
    dropout) -> max-pool(2, stride 1, 'same') -> two projection convs (ReLU / linear, BN) -> + input -> dense to the highway width ->
    4 highway layers -> bidirectional GRU over the whole padded sequence (input_lengths = None). x [B, T, num_mels] -> [B, T, 2 n]."""
    P = "CBHG_postnet/"
    bank = torch.cat([conv_block(x, params, P + "conv_bank/conv1d_%d/" % k, "relu", training, 0.0, None, stats_out)
                      for k in range(1, hp.cbhg_kernels + 1)], dim=-1)
    ps = hp.cbhg_pool_size
    # tf.layers.max_pooling1d(pool_size, strides=1, padding='same'): pad (ps - 1) // 2 left, the rest right, with -inf
    mp = F.max_pool1d(F.pad(bank.transpose(1, 2), ((ps - 1) // 2, ps - 1 - (ps - 1) // 2), value=float("-inf")), ps, stride=1).transpose(1, 2)
    p1 = conv_block(mp, params, P + "proj1/", "relu", training, 0.0, None, stats_out)
    p2 = conv_block(p1, params, P + "proj2/", None, training, 0.0, None, stats_out)
    h = p2 + x
    if P + "dense/kernel" in params:
        h = h @ params[P + "dense/kernel"] + params[P + "dense/bias"]
    for i in range(hp.cbhg_highwaynet_layers):
        q = P + "highwaynet_%d/" % (i + 1)
        Hh = F.relu(h @ params[q + "H/kernel"] + params[q + "H/bias"])
        Tt = torch.sigmoid(h @ params[q + "T/kernel"] + params[q + "T/bias"])
        h = Hh * Tt + h * (1.0 - Tt)
    B, T, _ = h.shape
    n = hp.cbhg_rnn_units
    outs = []
    for d, order in (("forward", range(T)), ("backward", range(T - 1, -1, -1))):
        q = P + d + "_RNN/"
        st = torch.zeros(B, n)
        seq = [None] * T
        for t in order:
            st = gru_cell(h[:, t], st, params[q + "gates/kernel"], params[q + "gates/bias"], params[q + "candidate/kernel"], params[q + "candidate/bias"])
            seq[t] = st
        outs.append(torch.stack(seq, dim=1))
    return torch.cat(outs, dim=-1)


def linear_head(mel_outputs, params, hp, training, stats_out=None):
    """tacotron.py:203-219: CBHG(mel_outputs) -> Dense(num_freq) -> clip"""
    y = cbhg(mel_outputs, params, hp, training, stats_out)
    lin = y @ params["cbhg_linear_specs_projection/kernel"] + params["cbhg_linear_specs_projection/bias"]
    if hp.clip_outputs:
        lin = torch.clamp(lin, _clip_low(hp), hp.max_abs_value)
    return lin


def linear_loss(linear_targets, linear_outputs, hp, targets_lengths=None):
    """tacotron.py:323-330 (plain) / MaskedLinearLoss (modules.py:457-485): L1 with half of the weight on the bins below 2 kHz."""
    l1 = (linear_targets - linear_outputs).abs()
    n_prio = int(2000 / (hp.sample_rate * 0.5) * hp.num_freq)
    if hp.mask_decoder:
        mask = (torch.arange(linear_targets.shape[1])[None, :] < targets_lengths[:, None]).float().unsqueeze(-1) * torch.ones_like(linear_targets)
        return 0.5 * (l1 * mask).sum() / mask.sum() + 0.5 * (l1 * mask)[:, :, :n_prio].sum() / mask.sum()
    return 0.5 * l1.mean() + 0.5 * l1[:, :, :n_prio].mean()


def synthesize(params, inputs, input_lengths, hp, max_iters=None, prenet_masks=None):
    """Free-running inference graph (tacotron.py:150-200 with is_training = is_evaluating = gta = False):
    TacoTestHelper (helpers.py:6-59) feeds the last predicted frame back (raw, before clipping) and finishes after the
    first step where round(sigmoid(stop)) is 1 for EVERY batch row - in any of the r outputs of the step (stop_at_any) or in all of them - or at
    hparams.max_iters; batch-norm uses moving statistics, zoneout its deterministic blend, prenet dropout stays on
    (prenet_masks[t] = list of per-layer masks to inject; None with rate 0). Stop output is the sigmoid (modules.py:340-342)."""
    B, T_in = inputs.shape
    D = hp.decoder_lstm_units
    zr = hp.tacotron_zoneout_rate
    max_iters = hp.max_iters if max_iters is None else max_iters
    x = params["inputs_embedding"][inputs]
    for i in range(hp.enc_conv_num_layers):
        x = conv_block(x, params, "encoder_convolutions/conv_layer_%d/" % (i + 1), "relu", False, hp.tacotron_dropout_rate)
    memory = encoder_rnn(x, input_lengths, params, hp, False)
    mask = (torch.arange(T_in)[None, :] < input_lengths[:, None]).float()
    if not getattr(hp, "mask_encoder", True):
        mask = None
    values = memory * mask.unsqueeze(-1) if mask is not None else memory
    keys = values @ params["attention/memory_layer/kernel"]
    c1 = torch.zeros(B, D); h1 = torch.zeros(B, D); c2 = torch.zeros(B, D); h2 = torch.zeros(B, D)
    ctx = torch.zeros(B, values.shape[-1])
    cum = torch.zeros(B, T_in)
    K1, b1 = params["decoder_LSTM/cell_1/kernel"], params["decoder_LSTM/cell_1/bias"]
    K2, b2 = params["decoder_LSTM/cell_2/kernel"], params["decoder_LSTM/cell_2/bias"]
    frame = torch.zeros(B, hp.num_mels)                                           # go frame
    prev_max = torch.zeros(B, dtype=torch.int64)                                  # Architecture_wrappers.py:166 max_attentions
    frames, stops, aligns = [], [], []
    for t in range(max_iters):
        pre = prenet(frame[:, -hp.num_mels:], params, hp, prenet_masks[t] if prenet_masks else None)      # the LAST of the r frames is fed back
        nc1, nh1 = lstm_cell(torch.cat([pre, ctx], dim=-1), c1, h1, K1, b1)
        c1n, h1n = zoneout(c1, nc1, zr, False), zoneout(h1, nh1, zr, False)
        nc2, nh2 = lstm_cell(nh1, c2, h2, K2, b2)
        c2n, h2n = zoneout(c2, nc2, zr, False), zoneout(h2, nh2, zr, False)
        c1, h1, c2, h2 = c1n, h1n, c2n, h2n
        cons = (prev_max, hp.attention_win_size, hp.synthesis_constraint_type) if getattr(hp, "synthesis_constraint", False) else None
        ctx, a = attention_step(nh2, cum, keys, values, mask, params, getattr(hp, "smoothing", False), cons)
        prev_max = a.argmax(dim=-1)
        cum = cum + a if getattr(hp, "cumulative_weights", True) else a
        pin = torch.cat([nh2, ctx], dim=-1)
        frame = pin @ params["linear_transform_projection/kernel"] + params["linear_transform_projection/bias"]
        stop = torch.sigmoid(pin @ params["stop_token_projection/kernel"] + params["stop_token_projection/bias"])
        frames.append(frame); stops.append(stop); aligns.append(a)
        fired = torch.round(stop).bool()                              # [B, r]; helpers.py:40-46: all rows, then any (stop_at_any) / all of the r
        if bool(fired.all(dim=0).any() if hp.stop_at_any else fired.all()):
            break
    decoder_output = torch.stack(frames, dim=1).reshape(B, -1, hp.num_mels)
    if hp.clip_outputs:
        decoder_output = torch.clamp(decoder_output, _clip_low(hp), hp.max_abs_value)
    y = decoder_output
    for i in range(hp.postnet_num_layers):
        act = "tanh" if i < hp.postnet_num_layers - 1 else None
        y = conv_block(y, params, "postnet_convolutions/conv_layer_%d/" % (i + 1), act, False, hp.tacotron_dropout_rate)
    mel_outputs = decoder_output + y @ params["postnet_projection/kernel"] + params["postnet_projection/bias"]
    if hp.clip_outputs:
        mel_outputs = torch.clamp(mel_outputs, _clip_low(hp), hp.max_abs_value)
    return {"decoder_output": decoder_output, "mel_outputs": mel_outputs, "stop_token_prediction": torch.stack(stops, dim=1).reshape(B, -1),
            "alignments": torch.stack(aligns, dim=1)}


def masked_mse(targets, outputs, targets_lengths):
    """MaskedMSE (modules.py:412-431): tf.losses.mean_squared_error(weights = mask) = sum(mask (t - o)^2) / count_nonzero(mask)."""
    mask = (torch.arange(targets.shape[1])[None, :] < targets_lengths[:, None]).float().unsqueeze(-1) * torch.ones_like(targets)
    return (mask * (targets - outputs) ** 2).sum() / torch.count_nonzero(mask).float()


def masked_sigmoid_cross_entropy(targets, outputs, targets_lengths, pos_weight):
    """MaskedSigmoidCrossEntropy (modules.py:433-455): weighted CE, masked, divided by the number of NON-ZERO masked terms."""
    mask = (torch.arange(targets.shape[1])[None, :] < targets_lengths[:, None]).float()
    losses = (1 - targets) * outputs + (1 + (pos_weight - 1) * targets) * (torch.log1p(torch.exp(-outputs.abs())) + torch.clamp(-outputs, min=0))
    masked = losses * mask
    return masked.sum() / torch.count_nonzero(masked).float()


def loss_fn(out, mel_targets, stop_targets, params, hp, targets_lengths=None, linear_targets=None):
    """tacotron.py:297-354: plain means over padded tensors (mask_decoder=False) or the masked variants, + L2 regulariser
    (+ the linear-spectrogram L1 when the CBHG head is on)."""
    if "linear_outputs" in out and linear_targets is not None:
        total, parts = loss_fn({k: v for k, v in out.items() if k != "linear_outputs"}, mel_targets, stop_targets, params, hp, targets_lengths)
        lin = linear_loss(linear_targets, out["linear_outputs"], hp, targets_lengths)
        parts["linear"] = lin
        return total + lin, parts
    if hp.mask_decoder:
        before = masked_mse(mel_targets, out["decoder_output"], targets_lengths)
        after = masked_mse(mel_targets, out["mel_outputs"], targets_lengths)
        stop = masked_sigmoid_cross_entropy(stop_targets, out["stop_logits"], targets_lengths, hp.cross_entropy_pos_weight)
        reg = sum((v * v).sum() / 2 for k, v in params.items() if is_regularized(k)) * _reg_weight(hp)
        return before + after + stop + reg, {"before": before, "after": after, "stop": stop, "reg": reg}
    before = F.mse_loss(out["decoder_output"], mel_targets)
    after = F.mse_loss(out["mel_outputs"], mel_targets)
    stop = F.binary_cross_entropy_with_logits(out["stop_logits"], stop_targets)
    reg = sum((v * v).sum() / 2 for k, v in params.items() if is_regularized(k)) * _reg_weight(hp)
    return before + after + stop + reg, {"before": before, "after": after, "stop": stop, "reg": reg}


def learning_rate(hp, global_step):
    """tacotron.py:439-463."""
    if not hp.tacotron_decay_learning_rate:
        return hp.tacotron_initial_learning_rate
    lr = hp.tacotron_initial_learning_rate * hp.tacotron_decay_rate ** (
        (global_step - hp.tacotron_start_decay) / hp.tacotron_decay_steps)
    return min(max(lr, hp.tacotron_final_learning_rate), hp.tacotron_initial_learning_rate)


def train_step(params, inputs, input_lengths, mel_targets, stop_targets, hp, masks=None, targets_lengths=None, linear_targets=None):
    ps = {k: (v.clone().requires_grad_(True) if is_trainable(k) else v.clone()) for k, v in params.items()}
    out = forward(ps, inputs, input_lengths, mel_targets, hp, True, masks)
    loss, parts = loss_fn(out, mel_targets, stop_targets, ps, hp, targets_lengths, linear_targets)
    names = [k for k in ps if is_trainable(k)]
    gr = torch.autograd.grad(loss, [ps[k] for k in names], allow_unused=True)
    grads = {k: (g if g is not None else torch.zeros_like(ps[k])) for k, g in zip(names, gr)}
    return loss.detach(), grads, {k: v.detach() for k, v in out.items()}, {k: v.detach() for k, v in parts.items()}


def adam_step(params, grads, state, hp, global_step):
    """tacotron.py:393, 429-437: clip_by_global_norm(1.0) then Adam."""
    lr = learning_rate(hp, global_step)
    b1, b2, eps = hp.tacotron_adam_beta1, hp.tacotron_adam_beta2, hp.tacotron_adam_epsilon
    state["t"] = state.get("t", 0) + 1
    t = state["t"]
    lr_t = lr * math.sqrt(1 - b2 ** t) / (1 - b1 ** t)
    if getattr(hp, "tacotron_fine_tuning", False):          # tacotron.py:401: no gradient (and no Adam update) for the embedding and the encoder
        grads = {k: g for k, g in grads.items() if not ("inputs_embedding" in k or "encoder_" in k)}
    if hp.tacotron_clip_gradients:
        gn = torch.sqrt(sum((g * g).sum() for g in grads.values()))
        scale = 1.0 / max(gn.item(), 1.0)
        grads = {k: g * scale for k, g in grads.items()}
    for k, g in grads.items():
        m = state.setdefault("m", {}).setdefault(k, torch.zeros_like(g))
        v = state.setdefault("v", {}).setdefault(k, torch.zeros_like(g))
        m.mul_(b1).add_(g, alpha=1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        params[k] = params[k] - lr_t * m / (v.sqrt() + eps)
    return lr
