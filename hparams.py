"""Hyper-parameters — same names and default values as the reference's hparams.py:5-374, served by a
TensorFlow-free HParams object (attribute access, .values(), .parse("a=b,c=[1,2]"), .set_hparam) so that
`--hparams` overrides and every `hparams.<name>` read of the reference keep working unchanged.
`paper_hparams()` applies the deltas of the reference's paper_hparams.py.
"""
import ast
import math
import re


class HParams(object):
    def __init__(self, **kw):
        object.__setattr__(self, "_names", [])
        for k, v in kw.items():
            self.add_hparam(k, v)

    def add_hparam(self, name, value):
        if name in self._names:
            raise ValueError("hyperparameter %s already exists" % name)
        self._names.append(name)
        object.__setattr__(self, name, value)

    def set_hparam(self, name, value):
        if name not in self._names:
            raise KeyError("unknown hyperparameter %s" % name)
        old = getattr(self, name)
        object.__setattr__(self, name, _coerce(value, old, name))

    def __setattr__(self, name, value):
        if name in self._names:
            self.set_hparam(name, value)
        else:
            self.add_hparam(name, value)

    def values(self):
        return {k: getattr(self, k) for k in self._names}

    def __contains__(self, name):
        return name in self._names

    def copy(self):
        return HParams(**{k: (list(v) if isinstance(v, list) else v) for k, v in self.values().items()})

    def parse(self, spec):
        """Comma-separated name=value overrides; list values in brackets: upsample_scales=[11,25]."""
        if not spec:
            return self
        for m in re.finditer(r"\s*([A-Za-z_]\w*)\s*=\s*(\[[^\]]*\]|\"[^\"]*\"|'[^']*'|[^,]*)\s*(?:,|$)", spec):
            name, raw = m.group(1), m.group(2).strip()
            if name not in self._names:
                raise ValueError("Unknown hyperparameter: %s" % name)
            self.set_hparam(name, _parse_literal(raw))
        return self

    def to_json(self):
        import json
        return json.dumps(self.values(), sort_keys=True, default=str)


def _parse_literal(raw):
    low = raw.lower()
    if low in ("true", "false"):
        return low == "true"
    if low == "none":
        return None
    try:
        return ast.literal_eval(raw)
    except (ValueError, SyntaxError):
        return raw


def _coerce(value, old, name):
    if old is None or value is None:
        return value
    if isinstance(old, bool):
        if isinstance(value, str):
            return value.lower() == "true"
        return bool(value)
    if isinstance(old, int) and not isinstance(old, bool):
        if isinstance(value, float) and value != int(value):
            raise ValueError("hparam %s expects an int, got %r" % (name, value))
        return int(value)
    if isinstance(old, float):
        return float(value)
    if isinstance(old, (list, tuple)):
        if not isinstance(value, (list, tuple)):
            value = [value]
        return type(old)(value) if isinstance(old, tuple) else list(value)
    return value


_SENTENCES = [
    "Scientists at the CERN laboratory say they have discovered a new particle.",
    "There's a way to measure the acute emotional intelligence that has never gone out of style.",
    "President Trump met with other leaders at the Group of 20 conference.",
    "The Senate's bill to repeal and replace the Affordable Care Act is now imperiled.",
    "Generative adversarial network or variational auto-encoder.",
    "Basilar membrane and otolaryngology are not auto-correlations.",
    "He has read the whole thing.",
    "He reads books.",
    "He thought it was time to present the present.",
    "Thisss isrealy awhsome.",
    "The big brown fox jumps over the lazy dog.",
    "Did the big brown fox jump over the lazy dog?",
    "Peter Piper picked a peck of pickled peppers. How many pickled peppers did Peter Piper pick?",
    "She sells sea-shells on the sea-shore. The shells she sells are sea-shells I'm sure.",
    "Tajima Airport serves Toyooka.",
    "Thank you so much for your support!",
]

_DEFAULTS = dict(
    cleaners="english_cleaners",
    # hardware setup
    tacotron_num_gpus=1, wavenet_num_gpus=1, split_on_cpu=True,
    # audio
    num_mels=80, num_freq=1025, rescale=True, rescaling_max=0.999, clip_mels_length=True,
    max_mel_frames=900, use_lws=False, silence_threshold=2, n_fft=2048, hop_size=275, win_size=1100,
    sample_rate=22050, frame_shift_ms=None, magnitude_power=2., trim_silence=True, trim_fft_size=2048,
    trim_hop_size=512, trim_top_db=40, signal_normalization=True, allow_clipping_in_normalization=True,
    symmetric_mels=True, max_abs_value=4., normalize_for_wavenet=True, clip_for_wavenet=True,
    wavenet_pad_sides=1, preemphasize=True, preemphasis=0.97, min_level_db=-100, ref_level_db=20,
    fmin=55, fmax=7600, power=1.5, griffin_lim_iters=60, GL_on_GPU=True,
    # tacotron
    outputs_per_step=1, stop_at_any=True, batch_norm_position="after", clip_outputs=True,
    lower_bound_decay=0.1, embedding_dim=512, enc_conv_num_layers=3, enc_conv_kernel_size=(5,),
    enc_conv_channels=512, encoder_lstm_units=256, smoothing=False, attention_dim=128,
    attention_filters=32, attention_kernel=(31,), cumulative_weights=True, synthesis_constraint=False,
    synthesis_constraint_type="window", attention_win_size=7, prenet_layers=[256, 256], decoder_layers=2,
    decoder_lstm_units=1024, max_iters=10000, postnet_num_layers=5, postnet_kernel_size=(5,),
    postnet_channels=512, cbhg_kernels=8, cbhg_conv_channels=128, cbhg_pool_size=2, cbhg_projection=256,
    cbhg_projection_kernel_size=3, cbhg_highwaynet_layers=4, cbhg_highway_units=128, cbhg_rnn_units=128,
    mask_encoder=True, mask_decoder=False, cross_entropy_pos_weight=1, predict_linear=True,
    # wavenet
    input_type="raw", quantize_channels=2 ** 16, use_bias=True, legacy=True, residual_legacy=True,
    log_scale_min=float(math.log(1e-14)), log_scale_min_gauss=float(math.log(1e-7)), cdf_loss=False,
    out_channels=2, layers=20, stacks=2, residual_channels=128, gate_channels=256, skip_out_channels=128,
    kernel_size=3, cin_channels=80, upsample_type="SubPixel", upsample_activation="Relu",
    upsample_scales=[11, 25], freq_axis_kernel_size=3, leaky_alpha=0.4, NN_init=True, NN_scaler=0.3,
    gin_channels=-1, use_speaker_embedding=True, n_speakers=5, speakers_path=None,
    speakers=["speaker0", "speaker1", "speaker2", "speaker3", "speaker4"],
    # tacotron training
    tacotron_random_seed=5339, tacotron_data_random_state=1234, tacotron_swap_with_cpu=False,
    tacotron_batch_size=32, tacotron_synthesis_batch_size=1, tacotron_test_size=0.05,
    tacotron_test_batches=None, tacotron_decay_learning_rate=True, tacotron_start_decay=40000,
    tacotron_decay_steps=18000, tacotron_decay_rate=0.5, tacotron_initial_learning_rate=1e-3,
    tacotron_final_learning_rate=1e-4, tacotron_adam_beta1=0.9, tacotron_adam_beta2=0.999,
    tacotron_adam_epsilon=1e-6, tacotron_reg_weight=1e-6, tacotron_scale_regularization=False,
    tacotron_zoneout_rate=0.1, tacotron_dropout_rate=0.5, tacotron_clip_gradients=True,
    tacotron_natural_eval=False, tacotron_teacher_forcing_mode="constant",
    tacotron_teacher_forcing_ratio=1., tacotron_teacher_forcing_init_ratio=1.,
    tacotron_teacher_forcing_final_ratio=0., tacotron_teacher_forcing_start_decay=10000,
    tacotron_teacher_forcing_decay_steps=40000, tacotron_teacher_forcing_decay_alpha=None,
    tacotron_fine_tuning=False,
    # wavenet training
    wavenet_random_seed=5339, wavenet_data_random_state=1234, wavenet_swap_with_cpu=False,
    wavenet_batch_size=8, wavenet_synthesis_batch_size=10 * 2, wavenet_test_size=None,
    wavenet_test_batches=1, wavenet_lr_schedule="exponential", wavenet_learning_rate=1e-3,
    wavenet_warmup=float(4000), wavenet_decay_rate=0.5, wavenet_decay_steps=200000,
    wavenet_adam_beta1=0.9, wavenet_adam_beta2=0.999, wavenet_adam_epsilon=1e-6,
    wavenet_clip_gradients=True, wavenet_ema_decay=0.9999, wavenet_weight_normalization=False,
    wavenet_init_scale=1., wavenet_dropout=0.05, wavenet_gradient_max_norm=100.0,
    wavenet_gradient_max_value=5.0, max_time_sec=None, max_time_steps=11000, wavenet_natural_eval=False,
    train_with_GTA=True,
    sentences=_SENTENCES,
    wavenet_synth_debug=False, wavenet_debug_wavs=["training_data/audio/audio-LJ001-0008.npy"],
    wavenet_debug_mels=["training_data/mels/mel-LJ001-0008.npy"],
)

hparams = HParams(**_DEFAULTS)


def paper_hparams():
    """The reference's paper_hparams.py expressed as deltas over the defaults (paper_hparams.py:13-204)."""
    hp = hparams.copy()
    for k, v in dict(
            max_mel_frames=1000, trim_top_db=45, preemphasize=False, fmin=75, predict_linear=False, legacy=False,
            residual_legacy=False, log_scale_min_gauss=float(math.log(9.1188196 * 1e-4)), cdf_loss=True,
            out_channels=30, layers=24, stacks=4, residual_channels=256, gate_channels=512,
            skip_out_channels=256, upsample_type="2D", upsample_scales=[5, 5, 11], NN_scaler=0.1,
            tacotron_decay_steps=24500, tacotron_final_learning_rate=1e-5, tacotron_reg_weight=1e-7,
            wavenet_learning_rate=1e-4).items():
        hp.set_hparam(k, v)
    return hp


def hparams_debug_string():
    values = hparams.values()
    hp = ["  %s: %s" % (name, values[name]) for name in sorted(values) if name != "sentences"]
    return "Hyperparameters:\n" + "\n".join(hp)


# default evaluation sentences of `synthesize.py --mode eval` when no --text_list is given (the reference ships its own list in
# hparams.py; any plain-English lines work)
sentences = [
    "The quick brown fox jumps over the lazy dog.",
    "Speech synthesis turns written text into an audible waveform.",
    "A vocoder conditioned on mel spectrograms generates one sample at a time.",
    "Attention aligns every decoder step with the characters it is reading.",
    "Does the quality of the voice depend on the size of the training set?",
]
