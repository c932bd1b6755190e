"""Generates tests/golden/reference_graph.npz by EXECUTING the reference's own Tacotron graph-construction code
(/root/reference/tacotron/models/tacotron.py `Tacotron.initialize` + `add_loss`, with modules.py, attention.py,
Architecture_wrappers.py, helpers.py and custom_decoder.py underneath) on the TF-1 stand-in of tf_shim.py + tf_shim_graph.py.

  python tests/golden/make_reference_graph_vectors.py        # needs /root/reference; only the committed .npz travels

Scenarios (small widths so that the fixture stays a few hundred KB; every hparam not listed keeps the reference's default):
  train     is_training=True, predict_linear=True, mask_decoder=False: all dropout / zoneout paths on; outputs, the five loss terms,
            d loss / d variable for every trainable variable (autograd through the executed reference graph) and every variable
            after one `add_optimizer` step (LR schedule, global-norm clip, Adam, batch-norm moving averages)
  train_md  is_training=True, predict_linear=False, mask_decoder=True: the masked losses inside the whole graph
  train_smooth / train_nomask  smoothing normalisation + non-cumulative attention state; un-masked encoder memory
  train_sched  scheduled teacher forcing (cosine-decayed ratio, per-step batch-wide draw, gradients through the fed-back frames)
  train_r2 / synth_r2  outputs_per_step = 2
  train_asym  symmetric_mels=False + tacotron_scale_regularization=True: the other output-clipping range and the scaled regulariser
  eval      is_evaluating=True: teacher forced, inference statistics, zoneout blend, prenet dropout still on
  gta       gta=True: as eval without the post-processing net
  synth     free running (TacoTestHelper), stop rule, max_iters; synth_window / synth_mono add the synthesis attention constraints
The variables are created by the reference code's own tf.get_variable / layer calls (values drawn here from a seeded generator) and
are stored under the names the reference's scopes give them - tests/test_reference_graph.py checks those names against
t2_tf_bundle.tacotron_tf_name, i.e. the checkpoint name map is pinned by the reference's code, not by a reading of it.
Dropout / zoneout masks are recorded in execution order and stored in the oracle's convention.

Honesty: the layer primitives under the reference's code (Dense, Conv1D, BatchNormalization, LSTMCell, GRUCell, dynamic_rnn,
BahdanauAttention, dynamic_decode) are tf_shim_graph.py's restatement of the TF 1.x definitions; what is executed unchanged is the
reference's composition of them. See tf_shim_graph.py's header."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"

SMALL = dict(
    embedding_dim=16, enc_conv_num_layers=3, enc_conv_kernel_size=5, enc_conv_channels=16, encoder_lstm_units=8,
    attention_dim=12, attention_filters=4, attention_kernel=7, prenet_layers=[16, 16], decoder_layers=2, decoder_lstm_units=24,
    num_mels=10, num_freq=33, postnet_num_layers=5, postnet_kernel_size=5, postnet_channels=16,
    cbhg_kernels=4, cbhg_conv_channels=8, cbhg_pool_size=2, cbhg_projection=16, cbhg_projection_kernel_size=3,
    cbhg_highwaynet_layers=2, cbhg_highway_units=12, cbhg_rnn_units=6,
    outputs_per_step=1, tacotron_num_gpus=1, split_on_cpu=True, max_iters=9,
)


def main():
    assert os.path.isdir(REF), "the reference tree is needed to (re)generate these fixtures"
    sys.path.insert(0, HERE)
    import tf_shim
    import tf_shim_graph as G
    G.install()
    sys.path.insert(0, REF)
    import hparams as ref_hparams_mod
    rhp = ref_hparams_mod.hparams
    from tacotron.models.tacotron import Tacotron

    for k, v in SMALL.items():
        assert hasattr(rhp, k), k
        setattr(rhp, k, v)
    out = {"small_hparams_keys": np.array(sorted(SMALL)), "small_hparams_values": np.array([repr(SMALL[k]) for k in sorted(SMALL)])}

    g = torch.Generator().manual_seed(77)
    B, T_in, T_out = 3, 7, 6
    in_len = torch.tensor([7, 5, 4], dtype=torch.int32)
    tgt_len = torch.tensor([6, 4, 5], dtype=torch.int32)
    inputs = torch.randint(2, 66, (B, T_in), generator=g, dtype=torch.int32)
    for b in range(B):
        inputs[b, in_len[b]:] = 0
    mel = torch.randn(B, T_out, rhp.num_mels, generator=g)
    lin = torch.randn(B, T_out, rhp.num_freq, generator=g)
    stop = (torch.arange(T_out)[None, :] >= (tgt_len[:, None] - 1)).float()
    split_infos = np.array([[T_in, T_out * rhp.num_mels, T_out, T_out * rhp.num_freq]], dtype=np.int32)
    out.update(inputs=inputs.numpy(), input_lengths=in_len.numpy(), targets_lengths=tgt_len.numpy(), mel_targets=mel.numpy(),
               linear_targets=lin.numpy(), stop_targets=stop.numpy())
    Tt = tf_shim.T

    def run(tag, variables, seed, allow_create=False, **kw):
        """one execution of the reference's initialize(); returns the model and the recorded random masks"""
        G.reset(seed=seed, variables=variables, allow_create=allow_create)
        model = Tacotron(rhp)
        args = dict(mel_targets=Tt(mel.clone()), stop_token_targets=Tt(stop.clone()), targets_lengths=Tt(tgt_len.clone()),
                    split_infos=split_infos)
        args.update(kw)
        for k in [k for k, v in args.items() if v is None]:
            del args[k]
        model.initialize(Tt(inputs.clone()), Tt(in_len.clone()), **args)
        return model, list(G.S.drops)

    def names_of(drops):
        return [(scope, kind, tuple(m.shape)) for scope, kind, m in drops]

    def masks_to_oracle(tag, drops, training, T_steps, with_post=True):
        """execution order of the recorded draws (see the scenario list) -> the oracle's mask dict, flattened into arrays"""
        q = list(drops)
        H, D = rhp.encoder_lstm_units, rhp.decoder_lstm_units
        keep = 1.0 - rhp.tacotron_dropout_rate

        def pop(scope_part, kind, shape):
            scope, k, m = q.pop(0)
            assert scope_part in scope and k == kind and tuple(m.shape) == tuple(shape), (tag, scope, k, tuple(m.shape), scope_part, shape)
            return m
        if training:
            for i in range(rhp.enc_conv_num_layers):
                out["%s_mask_enc_drop_%d" % (tag, i)] = (pop("encoder_convolutions", "layers.dropout", (B, T_in, rhp.enc_conv_channels)) / keep).numpy()
            for d in ("fw", "bw"):
                c = torch.stack([torch.zeros(B, H)] * T_in)
                h = torch.stack([torch.zeros(B, H)] * T_in)
                for tau in range(T_in):
                    mc = pop("bidirectional_rnn/" + d, "nn.dropout", (B, H))
                    mh = pop("bidirectional_rnn/" + d, "nn.dropout", (B, H))
                    for b in range(B):
                        # loop step tau of the (per-length reversed) backward pass is original time len - 1 - tau
                        t = tau if d == "fw" else int(in_len[b]) - 1 - tau
                        if 0 <= t < T_in and tau < int(in_len[b]):
                            c[t, b], h[t, b] = mc[b], mh[b]
                out["%s_mask_enc_zone_%s_c" % (tag, d)], out["%s_mask_enc_zone_%s_h" % (tag, d)] = c.numpy(), h.numpy()
        pre = [[], []]
        zone = {(l, s): [] for l in (1, 2) for s in "ch"}
        for t in range(T_steps):
            for i, n in enumerate(rhp.prenet_layers):
                pre[i].append(pop("decoder_prenet", "layers.dropout", (B, n)) / keep)
            if training:
                for l in (1, 2):
                    for s in "ch":
                        zone[(l, s)].append(pop("decoder_LSTM", "nn.dropout", (B, D)))
        for i in range(len(rhp.prenet_layers)):
            out["%s_mask_prenet_drop_%d" % (tag, i)] = torch.stack(pre[i], dim=1).numpy()               # [B, T, n]
        if training:
            for (l, s), v in zone.items():
                out["%s_mask_dec_zone_%d_%s" % (tag, l, s)] = torch.stack(v).numpy()                      # [T, B, D]
            for i in range(rhp.postnet_num_layers):
                out["%s_mask_post_drop_%d" % (tag, i)] = (pop("postnet_convolutions", "layers.dropout", (B, T_steps * rhp.outputs_per_step,
                                                                                                       rhp.postnet_channels)) / keep).numpy()
        assert not q, (tag, names_of(q))

    def save_outputs(tag, model, linear):
        out[tag + "_decoder_output"] = model.tower_decoder_output[0].detach().numpy()
        out[tag + "_mel_outputs"] = model.tower_mel_outputs[0].detach().numpy()
        out[tag + "_alignments"] = model.tower_alignments[0].detach().numpy()                               # [B, T_in, T_out]
        out[tag + "_stop_token_prediction"] = model.tower_stop_token_prediction[0].detach().numpy()
        if linear:
            out[tag + "_linear_outputs"] = model.tower_linear_outputs[0].detach().numpy()

    def save_losses(tag, model):
        for k in ("before_loss", "after_loss", "stop_token_loss", "regularization_loss", "linear_loss", "loss"):
            out["%s_%s" % (tag, k)] = np.asarray(float(getattr(model, k)), dtype=np.float64)

    # ---- train: creates the variables -------------------------------------------------------------------------------------------
    rhp.predict_linear, rhp.mask_decoder = True, False
    model, drops = run("train", None, 1, linear_targets=Tt(lin.clone()), is_training=True, global_step=Tt(torch.tensor(0)))
    variables = {k: v.detach().clone() for k, v in G.S.vars.items()}
    out["var_names"] = np.array(list(variables))
    out["var_trainable"] = np.array([bool(v.requires_grad) for v in G.S.vars.values()])
    for k, v in variables.items():
        out["var/" + k] = v.numpy()
    masks_to_oracle("train", drops, True, T_out)
    save_outputs("train", model, True)
    model.add_loss()
    save_losses("train", model)
    # one optimizer step (tacotron.py:371-437): learning-rate schedule at global step 60000, clip_by_global_norm(1.0), Adam; plus the
    # batch-norm moving-average updates it runs under (UPDATE_OPS)
    model.add_optimizer(Tt(torch.tensor(60000)))
    out["train_global_step"] = np.asarray(60000)
    out["train_learning_rate"] = np.asarray(float(model.learning_rate), dtype=np.float64)
    for k, v in model.optimize.new_values.items():
        out["train_new/" + k] = v.numpy()
    for k, v in G.S.updates.items():
        out["train_new/" + k] = v.numpy()
    # the same step with tacotron_fine_tuning: gradients only for the variables without 'inputs_embedding' / 'encoder_' in their names
    # (tacotron.py:401); the global norm of the clip is taken over those only
    rhp.tacotron_fine_tuning = True
    model.add_optimizer(Tt(torch.tensor(60000)))
    rhp.tacotron_fine_tuning = False
    for k, v in model.optimize.new_values.items():
        out["train_ft_new/" + k] = v.numpy()
    model.loss.backward()
    for k, v in G.S.vars.items():
        if v.requires_grad:
            out["grad/" + k] = (v.grad if v.grad is not None else torch.zeros_like(v)).detach().numpy()
    n_params = sum(int(v.numel()) for v in G.S.vars.values() if v.requires_grad)
    print("train: %d variables (%d trainable parameters), loss %.6f" % (len(variables), n_params, float(model.loss)))

    # ---- train with masked losses, no linear head --------------------------------------------------------------------------------
    rhp.predict_linear, rhp.mask_decoder = False, True
    no_cbhg = {k: v for k, v in variables.items() if "CBHG" not in k and "cbhg" not in k}
    model, drops = run("train_md", no_cbhg, 2, is_training=True, global_step=Tt(torch.tensor(0)))
    masks_to_oracle("train_md", drops, True, T_out)
    save_outputs("train_md", model, False)
    model.add_loss()
    save_losses("train_md", model)

    # ---- training with the non-default range flags: asymmetric mels (outputs clipped to [0 - decay, max]) + scaled regulariser ---------
    rhp.predict_linear, rhp.mask_decoder, rhp.symmetric_mels, rhp.tacotron_scale_regularization = True, False, False, True
    model, drops = run("train_asym", variables, 6, linear_targets=Tt(lin.clone()), is_training=True, global_step=Tt(torch.tensor(0)))
    masks_to_oracle("train_asym", drops, True, T_out)
    save_outputs("train_asym", model, True)
    model.add_loss()
    save_losses("train_asym", model)
    assert float((model.tower_decoder_output[0] == -rhp.lower_bound_decay).float().mean()) > 0.05       # the lower clip is active
    rhp.symmetric_mels, rhp.tacotron_scale_regularization = True, False

    # ---- attention variants the product rejects (kept for the oracle: the kernels of a later round are checked against it) --------------
    rhp.predict_linear, rhp.mask_decoder = False, False
    rhp.smoothing, rhp.cumulative_weights = True, False
    model, drops = run("train_smooth", no_cbhg, 7, is_training=True, global_step=Tt(torch.tensor(0)))
    masks_to_oracle("train_smooth", drops, True, T_out)
    save_outputs("train_smooth", model, False)
    rhp.smoothing, rhp.cumulative_weights, rhp.mask_encoder = False, True, False
    model, drops = run("train_nomask", no_cbhg, 8, is_training=True, global_step=Tt(torch.tensor(0)))
    masks_to_oracle("train_nomask", drops, True, T_out)
    save_outputs("train_nomask", model, False)
    rhp.mask_encoder = True

    # ---- scheduled teacher forcing (rejected by the product; oracle only): cosine-decayed ratio, one draw per step for the whole batch ---
    rhp.tacotron_teacher_forcing_mode = "scheduled"
    model, drops = run("train_sched", no_cbhg, 11, is_training=True, global_step=Tt(torch.tensor(30000)))
    draws = [u for kind, u in G.S.uniforms if kind == "random_uniform"]
    assert len(draws) == T_out and all(u.dim() == 0 for u in draws)
    out["train_sched_tf_draws"] = torch.stack(draws).numpy()
    out["train_sched_ratio"] = np.asarray(float(model.ratio), dtype=np.float64)
    out["train_sched_global_step"] = np.asarray(30000)
    assert 0 < float((torch.stack(draws) < model.ratio).float().mean()) < 1          # both branches taken in this fixture
    masks_to_oracle("train_sched", drops, True, T_out)
    save_outputs("train_sched", model, False)
    model.add_loss()
    save_losses("train_sched", model)
    model.loss.backward()
    out["train_sched_grad_embedding"] = G.S.vars["inference/inputs_embedding"].grad.numpy()
    out["train_sched_grad_prenet"] = G.S.vars["inference/decoder/decoder_prenet/dense_1/kernel"].grad.numpy()
    rhp.tacotron_teacher_forcing_mode = "constant"

    # ---- reduction factor r = 2 (rejected by the product; oracle only): two frames per decoder step, the last one fed back ----------------
    rhp.outputs_per_step = 2
    keep = {k: v for k, v in no_cbhg.items() if "linear_transform_projection" not in k and "stop_token_projection" not in k}
    model, drops = run("train_r2", keep, 9, allow_create=True, is_training=True, global_step=Tt(torch.tensor(0)))
    r2_vars = {k: v.detach().clone() for k, v in G.S.vars.items()}
    for k in r2_vars:
        if k not in keep:
            out["r2_var/" + k] = r2_vars[k].numpy()
    masks_to_oracle("train_r2", drops, True, T_out // 2)
    save_outputs("train_r2", model, False)
    model.add_loss()
    save_losses("train_r2", model)
    sb = [k for k in r2_vars if k.endswith("stop_token_projection/projection_stop_token_projection/bias")][0]
    r2_synth = dict(r2_vars)
    r2_synth[sb] = r2_vars[sb] - 5.0            # keep the stop token quiet: the fed-back frames run for max_iters steps
    out["r2_synth_stop_bias"] = r2_synth[sb].numpy()
    model, drops = run("synth_r2", r2_synth, 10, mel_targets=None, stop_token_targets=None, targets_lengths=None)
    masks_to_oracle("synth_r2", drops, False, int(model.tower_mel_outputs[0].shape[1]) // 2)
    save_outputs("synth_r2", model, False)
    print("synth_r2: %d frames" % int(model.tower_mel_outputs[0].shape[1]))
    rhp.outputs_per_step = 1

    # ---- eval / GTA ------------------------------------------------------------------------------------------------------------------
    rhp.predict_linear, rhp.mask_decoder = True, False
    model, drops = run("eval", variables, 3, linear_targets=Tt(lin.clone()), is_evaluating=True)
    masks_to_oracle("eval", drops, False, T_out)
    save_outputs("eval", model, True)
    model.add_loss()
    save_losses("eval", model)
    model, drops = run("gta", variables, 4, stop_token_targets=None, targets_lengths=None, gta=True)
    masks_to_oracle("gta", drops, False, T_out)
    save_outputs("gta", model, False)

    # ---- free-running synthesis ------------------------------------------------------------------------------------------------------
    model, drops = run("synth", variables, 5, mel_targets=None, stop_token_targets=None, targets_lengths=None)
    steps = int(model.tower_mel_outputs[0].shape[1])
    masks_to_oracle("synth", drops, False, steps)
    save_outputs("synth", model, True)
    print("synth: %d decoder steps (max_iters %d), stop predictions %s" % (steps, rhp.max_iters,
          np.round(out["synth_stop_token_prediction"][:, -1], 3)))
    # synthesis with the attention constraints (rejected by the product; oracle only): scores outside a window around / ahead of the
    # previous step's argmax are pushed to -2^32 + 1 (attention.py:201-214)
    for tag_c, kind, win in (("synth_window", "window", 3), ("synth_mono", "monotonic", 2)):
        rhp.synthesis_constraint, rhp.synthesis_constraint_type, rhp.attention_win_size = True, kind, win
        model, drops = run(tag_c, variables, 12, mel_targets=None, stop_token_targets=None, targets_lengths=None)
        masks_to_oracle(tag_c, drops, False, int(model.tower_mel_outputs[0].shape[1]))
        save_outputs(tag_c, model, True)
    rhp.synthesis_constraint, rhp.synthesis_constraint_type, rhp.attention_win_size = False, "window", 7

    # a second synthesis in which the stop rule (not max_iters) ends the loop: the stop logits above fall with time, so the stop
    # projection is negated (rising logits) and its bias shifted to put every row's crossing of 0.5 between step index 2 and 3
    biased = dict(variables)
    sk = [k for k in variables if "stop_token_projection/projection_stop_token_projection/" in k]
    assert len(sk) == 2
    neg = -torch.logit(torch.as_tensor(out["synth_stop_token_prediction"]).clamp(1e-6, 1 - 1e-6))          # [B, steps]
    low = neg.min(0).values
    assert float(low[3]) > float(low[:3].max())
    shift = -0.5 * (float(low[3]) + float(low[:3].max()))
    for k in sk:
        biased[k] = -variables[k] + (shift if k.endswith("bias") else 0.0)
    model, drops = run("synth_stop", biased, 5, mel_targets=None, stop_token_targets=None, targets_lengths=None)
    steps2 = int(model.tower_mel_outputs[0].shape[1])
    masks_to_oracle("synth_stop", drops, False, steps2)
    save_outputs("synth_stop", model, True)
    out["synth_stop_bias_shift"] = np.asarray(shift, dtype=np.float64)
    print("synth_stop: %d decoder steps" % steps2)

    # ---- error contract (tacotron.py:41-54, models/__init__.py of both packages): exception class + message of every rejected call -----
    import json
    from tacotron.models import create_model as taco_create
    from wavenet_vocoder.models import create_model as wn_create
    errors = {}

    def record(key, fn):
        try:
            fn()
            errors[key] = None
        except Exception as e:                                      # noqa: BLE001 - the class is what gets recorded
            errors[key] = [type(e).__name__, str(e)]
    rhp.predict_linear, rhp.mask_decoder = True, True
    base = dict(mel_targets=Tt(mel.clone()), stop_token_targets=Tt(stop.clone()), linear_targets=Tt(lin.clone()),
                targets_lengths=Tt(tgt_len.clone()), split_infos=split_infos)
    cases = {
        "stop_without_mel": dict(base, mel_targets=None, is_training=True),
        "mel_without_stop": dict(base, stop_token_targets=None, is_training=True),
        "linear_missing_in_training": dict(base, linear_targets=None, is_training=True),
        "linear_given_in_gta": dict(base, gta=True),
        "mask_without_lengths": dict(base, targets_lengths=None, is_training=True),
        "training_and_evaluating": dict(base, is_training=True, is_evaluating=True),
    }
    for key, kw in cases.items():
        kw = {k: v for k, v in kw.items() if v is not None}
        G.reset(seed=9, variables=None)
        record("tacotron_initialize/" + key, lambda kw=kw: Tacotron(rhp).initialize(Tt(inputs.clone()), Tt(in_len.clone()), **kw))
    record("tacotron_create_model/unknown", lambda: taco_create("Tacotron-3", rhp))
    rhp.input_type, rhp.quantize_channels, rhp.out_channels = "mulaw-quantize", 256, 30
    record("wavenet_create_model/out_channels_mismatch", lambda: wn_create("WaveNet", rhp))
    rhp.out_channels = 256
    record("wavenet_create_model/unknown", lambda: wn_create("WaveRNN", rhp))
    assert all(v is not None for v in errors.values()), errors
    with open(os.path.join(HERE, "reference_errors.json"), "w") as f:
        json.dump(errors, f, indent=1, sort_keys=True)

    path = os.path.join(HERE, "reference_graph.npz")
    np.savez_compressed(path, **out)
    print("wrote %s: %d arrays, %.1f KB" % (path, len(out), os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
