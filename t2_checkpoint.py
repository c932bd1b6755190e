"""Checkpoint files of the drop-in models: one `.npz` per save holding every variable under its TensorFlow-style name (the names /
layouts of SURVEY.md Appendix B, so that a converter from / to a TF bundle is a rename), the Adam moments (`adam_m/<name>`,
`adam_v/<name>`), the EMA shadow (`ema/<name>`, WaveNet) and `global_step`; `<dir>/checkpoint` names the latest file like
tf.train.get_checkpoint_state does (tacotron/train.py:205-215, wavenet_vocoder/train.py:262-276)."""
import os

import numpy as np


def save(save_dir, prefix, eng, keep=20):
    os.makedirs(save_dir, exist_ok=True)
    step = int(eng.global_step)
    path = os.path.join(save_dir, "%s-%d.npz" % (prefix, step))
    out = {"global_step": np.asarray(step, dtype=np.int64)}
    for k, v in eng.export_params().items():
        out[k] = v.numpy()
    for tag, buf in (("adam_m", getattr(eng, "m", None)), ("adam_v", getattr(eng, "v", None)), ("ema", getattr(eng, "ema", None))):
        if buf is not None:
            for k, v in eng.unflatten(buf).items():
                out["%s/%s" % (tag, k)] = v.numpy()
    np.savez(path, **out)
    with open(os.path.join(save_dir, "checkpoint"), "w") as f:
        f.write(os.path.basename(path) + "\n")
    files = sorted((f for f in os.listdir(save_dir) if f.startswith(prefix + "-") and f.endswith(".npz")),
                   key=lambda f: int(f[len(prefix) + 1:-4]))
    for f in files[:-keep]:
        os.remove(os.path.join(save_dir, f))
    return path


def latest(save_dir):
    p = os.path.join(save_dir, "checkpoint")
    if not os.path.isfile(p):
        return None
    name = open(p).read().strip()
    path = os.path.join(save_dir, name)
    return path if os.path.isfile(path) else None


def load(path):
    """-> (variables {name: tensor}, state {'global_step', 'adam_m', 'adam_v', 'ema'})"""
    import torch
    z = np.load(path)
    variables, state = {}, {"global_step": int(z["global_step"]), "adam_m": {}, "adam_v": {}, "ema": {}}
    for k in z.files:
        if k == "global_step":
            continue
        head = k.split("/", 1)[0]
        if head in ("adam_m", "adam_v", "ema"):
            state[head][k.split("/", 1)[1]] = torch.from_numpy(z[k])
        else:
            variables[k] = torch.from_numpy(z[k])
    return variables, state


def restore_engine(eng, variables, state):
    """variables + optimizer state into a product engine (t2.wavenet.WaveNet / t2.tacotron.Tacotron)"""
    import torch
    eng.load_params(variables)
    eng.global_step = state["global_step"]

    def flat(d):
        buf = torch.zeros(eng.n_params, dtype=torch.float32)
        for t in eng.tensors:
            name, off, shape = t[0], t[1], t[2]
            if name in d:
                buf[off:off + d[name].numel()] = d[name].reshape(-1).float()
        return buf.to(eng.device)
    if state["adam_m"]:
        eng.m, eng.v = flat(state["adam_m"]), flat(state["adam_v"])
    if state["ema"] and hasattr(eng, "ema"):
        eng.ema = flat(state["ema"])
