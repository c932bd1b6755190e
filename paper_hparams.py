"""The reference's second stock configuration (paper_hparams.py:1-376: "close to the paper" - 24-layer / 4-stack WaveNet with R256 / G512 /
S256, 30-channel MoL head, ConvTranspose2D conditioning upsampling, no linear head). Written as the difference to hparams.py so that the
two cannot drift apart: `from paper_hparams import hparams` has every key of hparams.py plus `upsample_conditional_features`.
These are the widths BASELINE.json's Cfg-2 / Cfg-4 time (`bench.py --workload wavenet_mol`)."""
from hparams import hparams as _base, hparams_debug_string as _debug_string

_paper = dict(
    # audio
    max_mel_frames=1000, trim_top_db=45, preemphasize=False, fmin=75,
    # Tacotron
    predict_linear=False, tacotron_decay_steps=24500, tacotron_final_learning_rate=1e-5, tacotron_reg_weight=1e-7,
    # WaveNet
    legacy=False, residual_legacy=False, log_scale_min_gauss=-7.000000006091266, cdf_loss=True, out_channels=10 * 3,
    layers=24, stacks=4, residual_channels=256, gate_channels=512, skip_out_channels=256,
    upsample_type="2D", upsample_scales=[5, 5, 11], NN_scaler=0.1, wavenet_learning_rate=1e-4,
)

hparams = _base.copy()
for _k, _v in _paper.items():
    setattr(hparams, _k, _v)
hparams.add_hparam("upsample_conditional_features", True)
# the paper configuration's evaluation list swaps two of the default sentences (paper_hparams.py:356-357)
hparams.sentences = [{"The big brown fox jumps over the lazy dog.": "Punctuation sensitivity, is working.",
                      "Did the big brown fox jump over the lazy dog?": "Punctuation sensitivity is working."}.get(s, s) for s in hparams.sentences]


def hparams_debug_string():
    values = hparams.values()
    return "Hyperparameters:\n" + "\n".join("  %s: %s" % (k, values[k]) for k in sorted(values) if k != "sentences")
