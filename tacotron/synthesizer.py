"""Tacotron Synthesizer (reference tacotron/synthesizer.py): load a checkpoint, then `synthesize(texts, basenames, out_dir, log_dir,
mel_filenames)` writes `mel-<basename>.npy` ([frames, num_mels] float32). GTA mode teacher-forces on the ground-truth mels
(mel_filenames given); natural mode runs the free-running decoder until every row's stop token fires. Griffin-Lim previews and
plots of the reference's eval mode are not produced (Griffin-Lim is outside the hot path)."""
import os

import numpy as np
import torch

import t2_checkpoint
from tacotron.feeder import pad_input, pad_target
from tacotron.models import create_model
from tacotron.utils.text import text_to_sequence


class Synthesizer(object):
    def load(self, checkpoint_path, hparams, gta=False, model_name="Tacotron"):
        self._hparams, self.gta = hparams, gta
        self.model = create_model(model_name, hparams)
        variables, _ = t2_checkpoint.load(checkpoint_path)
        self.model.load_variables(variables)
        self._pad = 0
        self._target_pad = -hparams.max_abs_value if hparams.symmetric_mels else 0.0

    def synthesize(self, texts, basenames, out_dir, log_dir, mel_filenames):
        hp = self._hparams
        cleaners = [x.strip() for x in hp.cleaners.split(",")]
        seqs = [np.asarray(text_to_sequence(t, cleaners), dtype=np.int32) for t in texts]
        lens = torch.tensor([len(s) for s in seqs], dtype=torch.int32).cuda()
        T_in = max(len(s) for s in seqs)
        inputs = torch.from_numpy(np.stack([pad_input(s, T_in, self._pad) for s in seqs])).cuda()
        if self.gta:
            targets = [np.load(f) for f in mel_filenames]
            target_lengths = [len(t) for t in targets]
            T_out = max(target_lengths)
            mel = torch.from_numpy(np.stack([pad_target(t, T_out, self._target_pad) for t in targets]).astype(np.float32)).cuda()
            self.model.initialize(inputs, lens, mel, gta=True)
            mels = self.model.tower_mel_outputs[0].cpu().numpy()
            mels = [m[:n] for m, n in zip(mels, target_lengths)]          # take off the batch-wise padding (synthesizer.py:167)
        else:
            self.model.initialize(inputs, lens)
            mels = self.model.tower_mel_outputs[0].cpu().numpy()
            stop = self.model.tower_stop_token_prediction[0].cpu().numpy()
            # cut each row at its own first <stop> (synthesizer.py:170-176 _get_output_lengths)
            cut = [int(np.argmax(np.round(s) > 0)) + 1 if (np.round(s) > 0).any() else len(s) for s in stop]
            mels = [m[:n] for m, n in zip(mels, cut)]
        mels = [np.clip(m, -hp.max_abs_value - hp.lower_bound_decay if hp.symmetric_mels else 0.0, hp.max_abs_value) for m in mels]
        names = []
        for m, b in zip(mels, basenames):
            path = os.path.join(out_dir, "mel-%s.npy" % b)
            np.save(path, m.astype(np.float32), allow_pickle=False)
            names.append(path)
        return names, ["<no_g>"] * len(names)
