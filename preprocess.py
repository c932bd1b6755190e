"""python preprocess.py [--base_dir D] [--dataset LJSpeech-1.1] [--output training_data] [--hparams a=b,...]
Same flags as the reference's preprocess.py:86-106. Writes <base_dir>/<output>/{audio,mels,linear}/*.npy and train.txt."""
import argparse
import os
from multiprocessing import cpu_count

from hparams import hparams


def write_metadata(metadata, out_dir, hp):
    with open(os.path.join(out_dir, "train.txt"), "w", encoding="utf-8") as f:
        for m in metadata:
            f.write("|".join([str(x) for x in m]) + "\n")
    mel_frames = sum(int(m[4]) for m in metadata)
    timesteps = sum(int(m[3]) for m in metadata)
    hours = timesteps / hp.sample_rate / 3600
    print("Write %d utterances, %d mel frames, %d audio timesteps, (%.2f hours)" % (len(metadata), mel_frames, timesteps, hours))
    if metadata:
        print("Max input length (text chars): %d" % max(len(m[5]) for m in metadata))
        print("Max mel frames length: %d" % max(int(m[4]) for m in metadata))
        print("Max audio timesteps length: %d" % max(m[3] for m in metadata))


def norm_data(args):
    supported = ["LJSpeech-1.0", "LJSpeech-1.1", "M-AILABS"]
    if args.dataset not in supported:
        raise ValueError("dataset value entered %s does not belong to supported datasets: %s" % (args.dataset, supported))
    if args.dataset.startswith("LJSpeech"):
        return [os.path.join(args.base_dir, args.dataset)]
    languages = ["en_US", "en_UK", "fr_FR", "it_IT", "de_DE", "es_ES", "ru_RU", "uk_UK", "pl_PL", "nl_NL", "pt_PT", "fi_FI", "se_SE", "tr_TR", "ar_SA"]
    if args.language not in languages:
        raise ValueError("Please enter a supported language to use from M-AILABS dataset! \n%s" % languages)
    if args.voice not in ("female", "male", "mix"):
        raise ValueError("Please enter a supported voice option to use from M-AILABS dataset! \n%s" % ["female", "male", "mix"])
    path = os.path.join(args.base_dir, args.language, "by_book", args.voice)
    readers = [e for e in os.listdir(path) if os.path.isdir(os.path.join(path, e))]
    if args.reader not in readers:
        raise ValueError("Please enter a valid reader for your language and voice settings! \n%s" % readers)
    path = os.path.join(path, args.reader)
    books = [e for e in os.listdir(path) if os.path.isdir(os.path.join(path, e))]
    if args.merge_books == "True":
        return [os.path.join(path, b) for b in books]
    if args.book not in books:
        raise ValueError("Please enter a valid book for your reader settings! \n%s" % books)
    return [os.path.join(path, args.book)]


def run_preprocess(args, hp):
    from datasets import preprocessor
    out_dir = os.path.join(args.base_dir, args.output)
    mel_dir, wav_dir, lin_dir = (os.path.join(out_dir, d) for d in ("mels", "audio", "linear"))
    for d in (mel_dir, wav_dir, lin_dir):
        os.makedirs(d, exist_ok=True)
    try:
        from tqdm import tqdm
    except ImportError:
        tqdm = lambda x: x
    metadata = preprocessor.build_from_path(hp, norm_data(args), mel_dir, lin_dir, wav_dir, args.n_jobs, tqdm=tqdm)
    write_metadata(metadata, out_dir, hp)


def main():
    print("initializing preprocessing..")
    parser = argparse.ArgumentParser()
    parser.add_argument("--base_dir", default="")
    parser.add_argument("--hparams", default="", help="Hyperparameter overrides as a comma-separated list of name=value pairs")
    parser.add_argument("--dataset", default="LJSpeech-1.1")
    parser.add_argument("--language", default="en_US")
    parser.add_argument("--voice", default="female")
    parser.add_argument("--reader", default="mary_ann")
    parser.add_argument("--merge_books", default="False")
    parser.add_argument("--book", default="northandsouth")
    parser.add_argument("--output", default="training_data")
    parser.add_argument("--n_jobs", type=int, default=cpu_count())
    args = parser.parse_args()
    assert args.merge_books in ("False", "True")
    run_preprocess(args, hparams.copy().parse(args.hparams))


if __name__ == "__main__":
    main()
