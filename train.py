"""python train.py --model Tacotron|WaveNet|Tacotron-2 [--base_dir D] [--hparams a=b,...] [--tacotron_train_steps N] ...
Same flags and sequencing as the reference's train.py:17-134: Tacotron-2 = train Tacotron -> GTA synthesis -> train WaveNet, with
the progress state kept in <log_dir>/state_log so that a restart resumes at the right stage. Launch under torchrun for
data-parallel training (one process per GPU; the reference's *_num_gpus towers)."""
import argparse
import os
from time import sleep

import infolog
from hparams import hparams as default_hparams
from infolog import log
from tacotron.synthesize import tacotron_synthesize
from tacotron.train import tacotron_train
from wavenet_vocoder.train import wavenet_train


def save_seq(file, sequence, input_path):
    """1 for 'step done', 0 otherwise: Tacotron | GTA | WaveNet | input path (train.py:17-20)"""
    with open(file, "w") as f:
        f.write("|".join([str(int(s)) for s in sequence] + [input_path]))


def read_seq(file):
    if os.path.isfile(file):
        with open(file, "r") as f:
            seq = f.read().split("|")
        return [bool(int(s)) for s in seq[:-1]], seq[-1]
    return [0, 0, 0], ""


def prepare_run(args):
    modified_hp = default_hparams.copy().parse(args.hparams)
    run_name = args.name or args.model
    log_dir = os.path.join(args.base_dir, "logs-%s" % run_name)
    os.makedirs(log_dir, exist_ok=True)
    infolog.init(os.path.join(log_dir, "Terminal_train_log"), run_name, args.slack_url)
    return log_dir, modified_hp


def train(args, log_dir, hparams):
    state_file = os.path.join(log_dir, "state_log")
    (taco_state, GTA_state, wave_state), input_path = read_seq(state_file)
    if not taco_state:
        log("\n#############################################################\nTacotron Train\n###########################################################\n")
        checkpoint = tacotron_train(args, log_dir, hparams)
        if checkpoint is None:
            raise RuntimeError("Error occured while training Tacotron, Exiting!")
        taco_state = 1
        save_seq(state_file, [taco_state, GTA_state, wave_state], input_path)
    else:
        checkpoint = os.path.join(log_dir, "taco_pretrained/")
    if not GTA_state:
        log("\n#############################################################\nTacotron GTA Synthesis\n###########################################################\n")
        args.mode, args.GTA = "synthesis", "True"
        input_path = tacotron_synthesize(args, hparams, checkpoint)
        GTA_state = 1
        save_seq(state_file, [taco_state, GTA_state, wave_state], input_path)
    if input_path == "" or input_path is None:
        raise RuntimeError("input_path has an unpleasant value -> %s" % input_path)
    if not wave_state:
        log("\n#############################################################\nWavenet Train\n###########################################################\n")
        checkpoint = wavenet_train(args, log_dir, hparams, input_path)
        if checkpoint is None:
            raise RuntimeError("Error occured while training Wavenet, Exiting!")
        wave_state = 1
        save_seq(state_file, [taco_state, GTA_state, wave_state], input_path)
    if wave_state and GTA_state and taco_state:
        log("TRAINING IS ALREADY COMPLETE!!")


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--base_dir", default="")
    parser.add_argument("--hparams", default="", help="Hyperparameter overrides as a comma-separated list of name=value pairs")
    parser.add_argument("--tacotron_input", default="training_data/train.txt")
    parser.add_argument("--wavenet_input", default="tacotron_output/gta/map.txt")
    parser.add_argument("--name", help="Name of logging directory.")
    parser.add_argument("--model", default="Tacotron-2")
    parser.add_argument("--input_dir", default="training_data", help="folder to contain inputs sentences/targets")
    parser.add_argument("--output_dir", default="output", help="folder to contain synthesized mel spectrograms")
    parser.add_argument("--mode", default="synthesis", help="mode for synthesis of tacotron after training")
    parser.add_argument("--GTA", default="True", help="Ground truth aligned synthesis, defaults to True, only considered in Tacotron synthesis mode")
    parser.add_argument("--restore", type=lambda s: str(s).lower() not in ("false", "0", ""), default=True, help="Set this to False to do a fresh training")
    parser.add_argument("--summary_interval", type=int, default=250, help="Steps between running summary ops")
    parser.add_argument("--embedding_interval", type=int, default=5000, help="Steps between updating embeddings projection visualization")
    parser.add_argument("--checkpoint_interval", type=int, default=2500, help="Steps between writing checkpoints")
    parser.add_argument("--eval_interval", type=int, default=5000, help="Steps between eval on test data")
    parser.add_argument("--tacotron_train_steps", type=int, default=100000, help="total number of tacotron training steps")
    parser.add_argument("--wavenet_train_steps", type=int, default=500000, help="total number of wavenet training steps")
    parser.add_argument("--tf_log_level", type=int, default=1, help="(ignored: there is no TensorFlow here)")
    parser.add_argument("--slack_url", default=None, help="slack webhook notification destination link")
    args = parser.parse_args()
    accepted_models = ["Tacotron", "WaveNet", "Tacotron-2"]
    if args.model not in accepted_models:
        raise ValueError("please enter a valid model to train: %s" % accepted_models)
    log_dir, hparams = prepare_run(args)
    if args.model == "Tacotron":
        tacotron_train(args, log_dir, hparams)
    elif args.model == "WaveNet":
        wavenet_train(args, log_dir, hparams, args.wavenet_input)
    else:
        train(args, log_dir, hparams)


if __name__ == "__main__":
    main()
