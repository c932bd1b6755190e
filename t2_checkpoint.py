"""Checkpoint files of the drop-in models: one `.npz` per save holding every variable under its TensorFlow-style name (the names /
layouts of SURVEY.md Appendix B, so that a converter from / to a TF bundle is a rename), the Adam moments (`adam_m/<name>`,
`adam_v/<name>`), the EMA shadow (`ema/<name>`, WaveNet) and `global_step`; `<dir>/checkpoint` names the latest file like
tf.train.get_checkpoint_state does (tacotron/train.py:205-215, wavenet_vocoder/train.py:262-276).

With `fmt="tf"` (or T2_CHECKPOINT_FORMAT=tf in the environment) the same content is written as a TensorFlow checkpoint-V2 bundle
under the reference graph's variable names (`<prefix>-<step>.index` / `.data-00000-of-00001`, t2_tf_bundle.py) and `load` /
`latest` read either kind, so checkpoints saved by the reference's tf.train.Saver restore here and vice versa."""
import os

import numpy as np

import t2_tf_bundle


def _format(fmt):
    fmt = fmt or os.environ.get("T2_CHECKPOINT_FORMAT", "npz")
    if fmt not in ("npz", "tf"):
        raise ValueError("checkpoint format must be 'npz' or 'tf'")
    return fmt


def save(save_dir, prefix, eng, keep=20, fmt=None):
    os.makedirs(save_dir, exist_ok=True)
    step = int(eng.global_step)
    if _format(fmt) == "tf":
        model = "Tacotron" if prefix.startswith("tacotron") else "WaveNet"
        path = os.path.join(save_dir, "%s-%d" % (prefix, step))
        t2_tf_bundle.export_tf(path, model, eng)
        idx = sorted((f[:-len(".index")] for f in os.listdir(save_dir) if f.startswith(prefix + "-") and f.endswith(".index")),
                     key=lambda f: int(f[len(prefix) + 1:]))
        for old in idx[:-keep]:
            for ext in (".index", ".data-00000-of-00001"):
                if os.path.isfile(os.path.join(save_dir, old + ext)):
                    os.remove(os.path.join(save_dir, old + ext))
        t2_tf_bundle.write_checkpoint_state(save_dir, os.path.basename(path), idx[-keep:])
        return path
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
    name = t2_tf_bundle.read_checkpoint_state(save_dir)          # tf.train.Saver's CheckpointState text proto
    if name is not None:
        path = name if os.path.isabs(name) else os.path.join(save_dir, name)
        return path if os.path.isfile(path + ".index") else None
    name = open(p).read().strip()
    path = os.path.join(save_dir, name)
    return path if os.path.isfile(path) else None


def load(path):
    """-> (variables {name: tensor}, state {'global_step', 'adam_m', 'adam_v', 'ema'})"""
    import torch
    if not path.endswith(".npz") and os.path.isfile(path + ".index"):
        variables, state = t2_tf_bundle.load_as_engine_dicts(path)
        as_t = lambda d: {k: torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)) for k, v in d.items()}
        return as_t(variables), {"global_step": state["global_step"], "adam_m": as_t(state["adam_m"]),
                                 "adam_v": as_t(state["adam_v"]), "ema": as_t(state["ema"])}
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
