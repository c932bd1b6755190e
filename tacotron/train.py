"""Tacotron training loop (reference tacotron/train.py:136-398): feeder thread -> create_model('Tacotron') ->
initialize / add_loss / add_optimizer per batch, periodic evaluation on the held-out batches, checkpoints under
<log_dir>/taco_pretrained/tacotron_model.ckpt-<step>.npz. One process per GPU (torchrun) replaces the towers."""
import os
import time

import numpy as np
import torch

import infolog
import t2_checkpoint
from tacotron.feeder import Feeder
from tacotron.models import create_model
from tacotron.utils import plot

log = infolog.log


class ValueWindow(object):
    def __init__(self, window_size=100):
        self._window_size, self._values = window_size, []

    def append(self, x):
        self._values = self._values[-(self._window_size - 1):] + [x]

    @property
    def average(self):
        return sum(self._values) / max(1, len(self._values))


def _dist():
    import torch.distributed as dist
    if "RANK" in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1 and not dist.is_initialized():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    return (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)


def _cuda(batch):
    return {k: v.cuda(non_blocking=True) for k, v in batch.items()}


def _run_model(model, b, step, is_training, is_evaluating=False):
    stop = b["token_targets"]
    model.initialize(b["inputs"], b["input_lengths"], b["mel_targets"], stop, linear_targets=b.get("linear_targets"),
                     targets_lengths=b["targets_lengths"], global_step=step, is_training=is_training, is_evaluating=is_evaluating)
    return model.add_loss()


def train(log_dir, args, hparams):
    rank, world = _dist()
    save_dir = os.path.join(log_dir, "taco_pretrained")
    eval_dir = os.path.join(log_dir, "eval-dir")
    os.makedirs(save_dir, exist_ok=True)
    os.makedirs(eval_dir, exist_ok=True)
    input_path = os.path.join(args.base_dir, args.tacotron_input)
    log("Checkpoint path: %s" % os.path.join(save_dir, "tacotron_model.ckpt"))
    log("Loading training data from: %s" % input_path)
    log("Using model: %s" % args.model)
    torch.manual_seed(hparams.tacotron_random_seed)
    feeder = Feeder(input_path, hparams, rank=rank, world_size=world)
    log("Loaded metadata for %d examples (%.2f hours)" % (len(feeder._metadata), feeder.hours))
    model = create_model("Tacotron", hparams)
    step = 0
    if args.restore:
        path = t2_checkpoint.latest(save_dir)
        if path:
            log("Loading checkpoint %s" % path)
            variables, state = t2_checkpoint.load(path)
            model.load_variables(variables)
            model._restore_state = state
            step = state["global_step"]
        else:
            log("No model to load at %s" % save_dir)
    else:
        log("Starting new training!")
    feeder.start()
    time_window, loss_window = ValueWindow(100), ValueWindow(100)
    log("Tacotron training set to a maximum of %d steps" % args.tacotron_train_steps)
    try:
        while step < args.tacotron_train_steps:
            t0 = time.time()
            b = _cuda(feeder.next_batch())
            loss = _run_model(model, b, step, True)
            state = getattr(model, "_restore_state", None)
            if state is not None:                       # optimizer moments of the restored run, once the first engine exists
                t2_checkpoint.restore_engine(model._eng, model._eng.export_params(), state)
                model._restore_state = None
            model.add_optimizer(step)
            step += 1
            loss = float(loss)
            time_window.append(time.time() - t0)
            loss_window.append(loss)
            log("Step %7d [%.3f sec/step, loss=%.5f, avg_loss=%.5f]" % (step, time_window.average, loss, loss_window.average),
                end="\r" if step % 10 else "\n")
            if np.isnan(loss) or loss > 100.0:
                log("Loss exploded to %.5f at step %d" % (loss, step))
                raise Exception("Loss exploded")
            if step % args.eval_interval == 0 and feeder.test_steps > 0:
                log("\nRunning evaluation at step %d" % step)
                losses = []
                for tb in feeder.test_batches():
                    tb = _cuda(Feeder.to_tensors(tb))
                    losses.append([float(_run_model(model, tb, step, False, True)), float(model.before_loss), float(model.after_loss),
                                   float(model.stop_token_loss)])
                m = np.mean(np.asarray(losses), axis=0)
                log("Eval loss for global step %d: %.3f (before %.3f, after %.3f, stop %.3f)" % (step, m[0], m[1], m[2], m[3]))
                np.save(os.path.join(eval_dir, "step-%d-eval-mel-prediction.npy" % step), model.tower_mel_outputs[0][0].cpu().numpy())
                np.save(os.path.join(eval_dir, "step-%d-eval-align.npy" % step), model.tower_alignments[0][0].cpu().numpy())
                if rank == 0:           # plots of the last held-out batch's first utterance (tacotron/train.py:296-311)
                    n = int(tb["targets_lengths"][0])
                    title = "Tacotron, step=%d, loss=%.5f" % (step, m[0])
                    plot.plot_alignment(model.tower_alignments[0][0].cpu().numpy(), os.path.join(eval_dir, "step-%d-eval-align.png" % step),
                                        title=title, max_len=n)
                    plot.plot_spectrogram(model.tower_mel_outputs[0][0].cpu().numpy(), os.path.join(eval_dir, "step-%d-eval-mel-spectrogram.png"
                                          % step), title=title, target_spectrogram=tb["mel_targets"][0].cpu().numpy(), max_len=n)
            if (step % args.checkpoint_interval == 0 or step == args.tacotron_train_steps) and rank == 0:
                path = t2_checkpoint.save(save_dir, "tacotron_model.ckpt", model._eng)
                log("\nSaving Model at step %d: %s" % (step, path))
        log("Tacotron training complete after %d global steps!" % args.tacotron_train_steps)
        return save_dir
    finally:
        feeder.stop()


def tacotron_train(args, log_dir, hparams):
    return train(log_dir, args, hparams)
