"""Tacotron synthesis drivers (reference tacotron/synthesize.py:42-136): `eval` mode synthesises a list of sentences into
<output>/eval/mel-*.npy + map.txt (rows text|mel|speaker), `synthesis` mode runs the whole training set (GTA = teacher forced)
into <output>/gta|natural/ and writes the map.txt the WaveNet feeder reads (rows wav|mel|gta_mel|speaker_id|text)."""
import os

import t2_checkpoint
from infolog import log
from tacotron.synthesizer import Synthesizer


def _checkpoint(checkpoint):
    path = t2_checkpoint.latest(checkpoint)
    if path is None:
        raise RuntimeError("Failed to load checkpoint at %s" % checkpoint)
    log("loaded model at %s" % path)
    return path


def run_live(args, checkpoint_path, hparams):
    """tacotron/synthesize.py:18-38: read lines from stdin, synthesize each one, Griffin-Lim it into temp.wav and play it"""
    synth = Synthesizer()
    synth.load(checkpoint_path, hparams)
    greetings = "Hello, Welcome to the Live testing tool. Please type a message and I will try to read it!"
    log(greetings)
    synth.synthesize([greetings], None, None, None, None)
    while True:
        try:
            synth.synthesize([input()], None, None, None, None)
        except (KeyboardInterrupt, EOFError):
            leave = "Thank you for testing our features. see you soon."
            log(leave)
            synth.synthesize([leave], None, None, None, None)
            break


def run_eval(args, checkpoint_path, output_dir, hparams, sentences):
    eval_dir = os.path.join(output_dir, "eval")
    log_dir = os.path.join(output_dir, "logs-eval")          # tacotron/synthesize.py:38-45: Griffin-Lim previews go to logs-eval/wavs
    os.makedirs(eval_dir, exist_ok=True)
    os.makedirs(os.path.join(log_dir, "wavs"), exist_ok=True)
    synth = Synthesizer()
    synth.load(checkpoint_path, hparams)
    n = hparams.tacotron_synthesis_batch_size
    batches = [sentences[i:i + n] for i in range(0, len(sentences), n)]
    log("Starting Synthesis")
    with open(os.path.join(eval_dir, "map.txt"), "w", encoding="utf-8") as f:
        for i, texts in enumerate(batches):
            basenames = ["batch_%d_sentence_%d" % (i, j) for j in range(len(texts))]
            mel_filenames, speaker_ids = synth.synthesize(texts, basenames, eval_dir, log_dir, None)
            for elems in zip(texts, mel_filenames, speaker_ids):
                f.write("|".join(str(x) for x in elems) + "\n")
    log("synthesized mel spectrograms at %s" % eval_dir)
    return eval_dir


def run_synthesis(args, checkpoint_path, output_dir, hparams):
    gta = args.GTA == "True"
    synth_dir = os.path.join(output_dir, "gta" if gta else "natural")
    os.makedirs(synth_dir, exist_ok=True)
    with open(os.path.join(args.input_dir, "train.txt"), encoding="utf-8") as f:
        metadata = [line.strip().split("|") for line in f if line.strip()]
    hours = sum(int(x[4]) for x in metadata) * hparams.hop_size / hparams.sample_rate / 3600
    log("Loaded metadata for %d examples (%.2f hours)" % (len(metadata), hours))
    synth = Synthesizer()
    synth.load(checkpoint_path, hparams, gta=gta)
    n = hparams.tacotron_synthesis_batch_size
    mel_dir, wav_dir = os.path.join(args.input_dir, "mels"), os.path.join(args.input_dir, "audio")
    log("Starting Synthesis")
    with open(os.path.join(synth_dir, "map.txt"), "w", encoding="utf-8") as f:
        for i in range(0, len(metadata), n):
            meta = metadata[i:i + n]
            texts = [m[5] for m in meta]
            mel_filenames = [os.path.join(mel_dir, m[1]) for m in meta]
            wav_filenames = [os.path.join(wav_dir, m[0]) for m in meta]
            basenames = [os.path.basename(m).replace(".npy", "").replace("mel-", "") for m in mel_filenames]
            out_names, speaker_ids = synth.synthesize(texts, basenames, synth_dir, None, mel_filenames)
            for elems in zip(wav_filenames, mel_filenames, out_names, speaker_ids, texts):
                f.write("|".join(str(x) for x in elems) + "\n")
    log("synthesized mel spectrograms at %s" % synth_dir)
    return os.path.join(synth_dir, "map.txt")


def tacotron_synthesize(args, hparams, checkpoint, sentences=None):
    output_dir = "tacotron_" + args.output_dir
    checkpoint_path = _checkpoint(checkpoint)
    if args.mode == "eval":
        return run_eval(args, checkpoint_path, output_dir, hparams, sentences)
    if args.mode == "synthesis":
        return run_synthesis(args, checkpoint_path, output_dir, hparams)
    return run_live(args, checkpoint_path, hparams)
