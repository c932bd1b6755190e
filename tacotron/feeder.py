"""Tacotron feeder: reads `training_data/train.txt` (rows `audio|mel|linear|time_steps|mel_frames|text`,
datasets/preprocessor.py), tokenises the text, loads the mel targets and produces padded batches on a background thread.

Semantics kept from the reference's tacotron/feeder.py: deterministic train / test split (sklearn train_test_split with
`tacotron_data_random_state`, test set rounded down to whole batches, :36-52), groups of 32 batches sorted by output length and
then shuffled (:152-169), input ids padded with 0, mel targets padded with -max_abs_value (symmetric) or 0 (:60-69), stop-token
targets 0 for the first len-1 frames then 1-padding up to a multiple of outputs_per_step of (max_len + 1) (:130,240-252).
What differs by design: no TF queue / placeholders - `next_batch()` returns pinned torch tensors (one process per GPU replaces
the towers, so there is no time-axis concatenation and no split_infos); with torch.distributed each rank reads every
world_size-th batch of a group."""
import os
import queue
import threading

import numpy as np
import torch

from tacotron.utils.text import text_to_sequence

_batches_per_group = 32


def _round_up(x, multiple):
    r = x % multiple
    return x if r == 0 else x + multiple - r


def _round_down(x, multiple):
    return x - x % multiple


def pad_input(x, length, pad=0):
    return np.pad(x, (0, length - x.shape[0]), mode="constant", constant_values=pad)


def pad_target(t, length, pad):
    return np.pad(t, [(0, length - t.shape[0]), (0, 0)], mode="constant", constant_values=pad)


def pad_token_target(t, length, pad=1.0):
    return np.pad(t, (0, length - t.shape[0]), mode="constant", constant_values=pad)


class Feeder(object):
    def __init__(self, metadata_filename, hparams, rank=0, world_size=1, seed=None, prefetch=8):
        self._hparams = hparams
        self._cleaner_names = [x.strip() for x in hparams.cleaners.split(",")]
        self._mel_dir = os.path.join(os.path.dirname(metadata_filename), "mels")
        self._linear_dir = os.path.join(os.path.dirname(metadata_filename), "linear")
        with open(metadata_filename, encoding="utf-8") as f:
            self._metadata = [line.strip().split("|") for line in f if line.strip()]
        self.hours = sum(int(x[4]) for x in self._metadata) * hparams.hop_size / hparams.sample_rate / 3600.0
        from sklearn.model_selection import train_test_split
        bs = hparams.tacotron_batch_size
        test_size = hparams.tacotron_test_size if hparams.tacotron_test_size is not None else hparams.tacotron_test_batches * bs
        test_size = int(test_size) if test_size >= 1 else float(test_size)     # sklearn: an int counts examples, a float is a fraction
        idx = np.arange(len(self._metadata))
        train_idx, test_idx = train_test_split(idx, test_size=test_size, random_state=hparams.tacotron_data_random_state)
        keep = _round_down(len(test_idx), bs)
        train_idx = np.concatenate([train_idx, test_idx[keep:]])
        test_idx = test_idx[:keep]
        self._train_meta = [self._metadata[i] for i in train_idx]
        self._test_meta = [self._metadata[i] for i in test_idx]
        self.test_steps = len(self._test_meta) // bs
        self._pad = 0
        self._target_pad = -hparams.max_abs_value if hparams.symmetric_mels else 0.0
        self._token_pad = 1.0
        self._rank, self._world = rank, world_size
        self._rng = np.random.RandomState(hparams.tacotron_random_seed if seed is None else seed)
        self._train_offset = 0
        self._queue = queue.Queue(maxsize=prefetch)
        self._stop = threading.Event()
        self._thread = None

    # ---- examples --------------------------------------------------------------------------------------------------
    def _load(self, meta):
        ids = np.asarray(text_to_sequence(meta[5], self._cleaner_names), dtype=np.int32)
        mel = np.load(os.path.join(self._mel_dir, meta[1]))
        token = np.zeros(len(mel) - 1, dtype=np.float32)
        # linear-spectrogram target of the post-processing net (tacotron/feeder.py:136-139 loads it unconditionally; here only when used)
        linear = np.load(os.path.join(self._linear_dir, meta[2])) if getattr(self._hparams, "predict_linear", False) else None
        return ids, mel, token, linear, len(mel)

    def _next_example(self):
        if self._train_offset >= len(self._train_meta):
            self._train_offset = 0
            self._rng.shuffle(self._train_meta)
        meta = self._train_meta[self._train_offset]
        self._train_offset += 1
        return self._load(meta)

    def prepare_batch(self, batch):
        """list of (ids, mel [frames, num_mels], token [frames - 1], linear [frames, num_freq] | None, frames) -> dict of numpy arrays"""
        r = self._hparams.outputs_per_step
        in_len = max(len(x[0]) for x in batch)
        mel_len = _round_up(max(len(x[1]) for x in batch), r)
        tok_len = _round_up(max(len(x[2]) for x in batch) + 1, r)
        return {"inputs": np.stack([pad_input(x[0], in_len, self._pad) for x in batch]).astype(np.int32),
                "input_lengths": np.asarray([len(x[0]) for x in batch], dtype=np.int32),
                "mel_targets": np.stack([pad_target(x[1], mel_len, self._target_pad) for x in batch]).astype(np.float32),
                "token_targets": np.stack([pad_token_target(x[2], tok_len, self._token_pad) for x in batch]).astype(np.float32),
                "targets_lengths": np.asarray([x[-1] for x in batch], dtype=np.int32),
                **({"linear_targets": np.stack([pad_target(x[3], mel_len, self._target_pad) for x in batch]).astype(np.float32)}
                   if batch[0][3] is not None else {})}

    def train_group(self):
        n = self._hparams.tacotron_batch_size
        examples = [self._next_example() for _ in range(n * _batches_per_group)]
        examples.sort(key=lambda x: x[-1])
        batches = [examples[i:i + n] for i in range(0, len(examples), n)]
        self._rng.shuffle(batches)
        return [self.prepare_batch(b) for b in batches[self._rank::self._world]]

    def test_batches(self):
        n = self._hparams.tacotron_batch_size
        examples = sorted((self._load(m) for m in self._test_meta), key=lambda x: x[-1])
        return [self.prepare_batch(examples[i:i + n]) for i in range(0, len(examples), n)]

    # ---- background thread (reference: one feeder thread per queue, tacotron/feeder.py:111-119) -----------------------
    @staticmethod
    def to_tensors(batch, pin=True):
        out = {k: torch.from_numpy(v) for k, v in batch.items()}
        if pin and torch.cuda.is_available():
            out = {k: v.pin_memory() for k, v in out.items()}
        return out

    def _run(self):
        while not self._stop.is_set():
            for b in self.train_group():
                t = self.to_tensors(b)
                while not self._stop.is_set():
                    try:
                        self._queue.put(t, timeout=0.2)
                        break
                    except queue.Full:
                        continue
                if self._stop.is_set():
                    return

    def start(self):
        self._thread = threading.Thread(target=self._run, name="tacotron-feeder", daemon=True)
        self._thread.start()
        return self

    def next_batch(self, timeout=600):
        if self._thread is None:
            self.start()
        return self._queue.get(timeout=timeout)

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=5)
