"""python wavenet_preprocess.py [--base_dir D] [--input_dir LJSpeech-1.1/wavs] [--output tacotron_output/gta/] [--hparams a=b,...]
Same flags as the reference's wavenet_preprocess.py:36-52: prepares a folder of wavs for WaveNet-only training (ground-truth mels
as the conditioning): <base_dir>/<output>/{audio,mels}/*.npy + map.txt. Train with `train.py --model WaveNet` and `train_with_GTA=False`."""
import argparse
import os
from multiprocessing import cpu_count

from hparams import hparams


def write_metadata(metadata, out_dir, hp):
    with open(os.path.join(out_dir, "map.txt"), "w", encoding="utf-8") as f:
        for m in metadata:
            f.write("|".join(str(x) for x in m) + "\n")
    timesteps = sum(int(m[4]) for m in metadata)
    print("Write %d utterances, %d audio timesteps, (%.2f hours)" % (len(metadata), timesteps, timesteps / hp.sample_rate / 3600))
    if metadata:
        print("Max mel frames length: %d" % max(int(m[5]) for m in metadata))
        print("Max audio timesteps length: %d" % max(int(m[4]) for m in metadata))


def run_preprocess(args, hp):
    from datasets import wavenet_preprocessor
    out_dir = os.path.join(args.base_dir, args.output)
    mel_dir, wav_dir = os.path.join(out_dir, "mels"), os.path.join(out_dir, "audio")
    os.makedirs(mel_dir, exist_ok=True)
    os.makedirs(wav_dir, exist_ok=True)
    try:
        from tqdm import tqdm
    except ImportError:
        tqdm = lambda x: x
    write_metadata(wavenet_preprocessor.build_from_path(hp, args.input_dir, mel_dir, wav_dir, args.n_jobs, tqdm=tqdm), out_dir, hp)


def main():
    print("initializing preprocessing..")
    parser = argparse.ArgumentParser()
    parser.add_argument("--base_dir", default="")
    parser.add_argument("--hparams", default="", help="Hyperparameter overrides as a comma-separated list of name=value pairs")
    parser.add_argument("--input_dir", default="LJSpeech-1.1/wavs")
    parser.add_argument("--output", default="tacotron_output/gta/")
    parser.add_argument("--n_jobs", type=int, default=cpu_count())
    args = parser.parse_args()
    run_preprocess(args, hparams.copy().parse(args.hparams))


if __name__ == "__main__":
    main()
