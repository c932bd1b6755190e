"""python tools/griffin_lim_synthesis_tool.py --mels_dir tacotron_output/eval --out_dir gl_wavs [--linear] [--hparams a=b,...]
The reference ships this as a notebook (griffin_lim_synthesis_tool.ipynb): invert every mel-*.npy (or linear-*.npy with --linear) of a
folder to a waveform with Griffin-Lim - here the GPU kernels behind datasets/audio.py inv_mel_spectrogram / inv_linear_spectrogram."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from datasets import audio  # noqa: E402
from hparams import hparams  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mels_dir", required=True)
    ap.add_argument("--out_dir", default="gl_wavs")
    ap.add_argument("--linear", action="store_true", help="invert linear-*.npy files instead of mel-*.npy")
    ap.add_argument("--hparams", default="")
    args = ap.parse_args()
    hp = hparams.copy().parse(args.hparams)
    os.makedirs(args.out_dir, exist_ok=True)
    prefix = "linear-" if args.linear else "mel-"
    files = sorted(f for f in os.listdir(args.mels_dir) if f.startswith(prefix) and f.endswith(".npy"))
    for f in files:
        spec = np.load(os.path.join(args.mels_dir, f))                 # [frames, channels]
        if len(spec) < 2:
            continue
        wav = (audio.inv_linear_spectrogram if args.linear else audio.inv_mel_spectrogram)(spec.T, hp)
        audio.save_wav(wav, os.path.join(args.out_dir, f.replace(".npy", ".wav")), hp.sample_rate)
    print("inverted %d spectrograms into %s" % (len(files), args.out_dir))


if __name__ == "__main__":
    main()
