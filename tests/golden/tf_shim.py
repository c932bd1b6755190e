"""TEST INFRASTRUCTURE — a minimal stand-in for `import tensorflow as tf` (TF 1.x graph API) backed by torch-CPU, so that
the reference's OWN Python source files can be imported and their pure tensor-math functions EXECUTED in this container
(TensorFlow 1.x has no cp312 wheel and there is no network). Used only by tests/golden/make_reference_vectors.py to
generate reference-executed fixtures; nothing in the product or in the GPU tests imports it.

Each implemented op restates the documented TF 1.x semantics of one elementary op (elementwise math, reductions, masks,
losses' closed forms — SURVEY.md Appendix A cites the TF definitions). Everything NOT implemented resolves to a stub that
lets `import` succeed (class stubs for CamelCase names so they can be base classes) and raises when called.
Random draws are INJECTED: tf.random_uniform / Normal.sample pop pre-seeded tensors from `inject(...)`.
Variables are injected too: tf.get_variable(name, ...) returns `variables[name]`."""
import contextlib
import math
import sys
import types

import numpy as np
import torch

float32, float64, int32, int64, bool_ = torch.float32, torch.float64, torch.int32, torch.int64, torch.bool
_uniform_queue, _normal_queue, variables = [], [], {}


def inject(uniform=(), normal=(), vars=None):
    _uniform_queue[:] = list(uniform)
    _normal_queue[:] = list(normal)
    variables.clear()
    variables.update(vars or {})


class _Stub(types.ModuleType):
    """attribute access never fails: CamelCase -> an empty class (usable as a base class), otherwise a sub-stub"""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        obj = type(name, (object,), {"__init__": _kw_init}) if name[0].isupper() else type(self)(self.__name__ + "." + name)
        setattr(self, name, obj)
        return obj

    def __call__(self, *a, **k):
        raise NotImplementedError("tf_shim: %s is not implemented" % self.__name__)


def _kw_init(self, *a, **k):
    self.__dict__.update(k)       # e.g. tf.contrib.training.HParams(**values) keeps the reference's own hparam values


class _NoopStub(_Stub):
    """import-only stand-in for plotting / text-normalisation packages: calls are accepted and ignored"""

    def __call__(self, *a, **k):
        return _NoopStub(self.__name__ + "()")


class Dim(int):
    value = property(lambda self: int(self))


class Shape(tuple):
    def as_list(self):
        return [int(d) for d in self]


class T(torch.Tensor):
    """torch tensor that also answers the TF shape protocol (x.get_shape(), x.shape[-1].value)"""

    @staticmethod
    def __new__(cls, data):
        return torch.Tensor._make_subclass(cls, torch.as_tensor(data))

    def get_shape(self):
        return Shape(Dim(d) for d in torch.Tensor.size(self))

    shape = property(get_shape)

    def __getitem__(self, idx):
        """numpy / TF style reversed slices (x[:, ::-1]), which torch does not index"""
        items = idx if isinstance(idx, tuple) else (idx,)
        if not any(isinstance(i, slice) and i.step == -1 and i.start is None and i.stop is None for i in items):
            return torch.Tensor.__getitem__(self, idx)
        plain = tuple(slice(None) if (isinstance(i, slice) and i.step == -1) else i for i in items)
        out = torch.Tensor.__getitem__(self, plain)
        dims, d = [], 0
        for i in items:
            if isinstance(i, slice):
                if i.step == -1:
                    dims.append(d)
                d += 1
            elif i is None:
                d += 1
        return torch.flip(out, dims)


def _t(x, like=None):
    if isinstance(x, torch.Tensor):
        return x
    dt = like.dtype if (like is not None and isinstance(x, (int, float)) and like.is_floating_point()) else None
    return torch.as_tensor(x, dtype=dt)


def _ax(axis):
    if isinstance(axis, (list, tuple)):
        return tuple(int(a) for a in axis)
    return axis if axis is None else int(axis)


def _shape_list(shape):
    return [int(s) for s in shape]


def build():
    tf = _Stub("tensorflow")
    tf.float32, tf.float64, tf.int32, tf.int64, tf.bool = float32, float64, int32, int64, bool_
    tf.shape = lambda x, **k: torch.tensor(list(x.shape), dtype=torch.int64)
    tf.rank = lambda x: x.dim()
    tf.mod = lambda a, b: a % b

    def assert_equal(a, b, **k):
        a, b = torch.as_tensor(a), torch.as_tensor(b)
        assert torch.equal(a.to(torch.int64), b.to(torch.int64)), "tf.assert_equal failed: %s vs %s" % (a, b)
    tf.assert_equal = assert_equal
    tf.control_dependencies = lambda deps: contextlib.nullcontext()
    tf.convert_to_tensor = lambda x, dtype=None, **k: torch.as_tensor(x, dtype=dtype)
    tf.cast = lambda x, dtype: torch.trunc(x).to(dtype) if (not dtype.is_floating_point and x.is_floating_point()) else x.to(dtype)
    tf.transpose = lambda x, perm=None: x.permute(*perm) if perm is not None else x.t()
    tf.maximum = lambda a, b: torch.maximum(_t(a, _t(b)) if not isinstance(a, torch.Tensor) else a, _t(b, a if isinstance(a, torch.Tensor) else None))
    tf.minimum = lambda a, b: torch.minimum(_t(a, _t(b)) if not isinstance(a, torch.Tensor) else a, _t(b, a if isinstance(a, torch.Tensor) else None))
    tf.ones = lambda shape, dtype=float32, **k: torch.ones(_shape_list(shape), dtype=dtype)
    tf.zeros = lambda shape, dtype=float32, **k: torch.zeros(_shape_list(shape), dtype=dtype)
    tf.ones_like, tf.zeros_like = torch.ones_like, torch.zeros_like
    for name in ("exp", "log", "abs", "square", "tanh", "sign", "log1p", "sqrt", "sigmoid", "floor", "ceil", "round"):
        setattr(tf, name, getattr(torch, name))
    tf.pow = lambda a, b: torch.pow(_t(a), _t(b))
    tf.where = lambda c, a, b: torch.where(c, a, b)
    tf.equal = lambda a, b: torch.as_tensor(a) == torch.as_tensor(b)
    tf.cond = lambda pred, true_fn, false_fn: true_fn() if bool(torch.as_tensor(pred).all()) else false_fn()
    tf.squeeze = lambda x, axis=None: x.squeeze() if axis is None else x.squeeze(_ax(axis)[0] if isinstance(_ax(axis), tuple) else _ax(axis))
    tf.expand_dims = lambda x, axis: x.unsqueeze(_ax(axis)[0] if isinstance(_ax(axis), tuple) else _ax(axis))
    tf.concat = lambda xs, axis: torch.cat(list(xs), dim=int(axis))
    tf.reshape = lambda x, shape: x.reshape(_shape_list(shape))

    def _reduce(fn):
        def f(x, axis=None, keepdims=False, **k):
            ax = _ax(axis)
            if ax is None:
                return fn(x)
            r = fn(x, dim=ax, keepdim=keepdims)
            return r[0] if isinstance(r, tuple) else r
        return f
    tf.reduce_sum, tf.reduce_mean = _reduce(torch.sum), _reduce(torch.mean)
    tf.reduce_max = _reduce(lambda x, **k: torch.amax(x, **k) if k else torch.max(x))
    tf.reduce_min = _reduce(lambda x, **k: torch.amin(x, **k) if k else torch.min(x))
    tf.reduce_logsumexp = _reduce(lambda x, **k: torch.logsumexp(x, **k) if k else torch.logsumexp(x.reshape(-1), 0))
    tf.count_nonzero = lambda x, dtype=int64, **k: (x != 0).sum().to(dtype)
    tf.argmax = lambda x, axis=None, **k: torch.argmax(x, dim=_ax(axis))
    tf.one_hot = lambda idx, depth, dtype=float32, **k: torch.nn.functional.one_hot(idx.long(), int(depth)).to(dtype)
    tf.sequence_mask = lambda lengths, maxlen=None, dtype=bool_, **k: (
        torch.arange(int(maxlen if maxlen is not None else lengths.max()))[None, :] < torch.as_tensor(lengths)[:, None]).to(dtype)

    def random_uniform(shape, minval=0.0, maxval=1.0, dtype=float32, **k):
        u = _uniform_queue.pop(0)
        assert list(u.shape) == _shape_list(shape) and float(u.min()) >= minval and float(u.max()) <= maxval
        return u
    tf.random_uniform = random_uniform

    def get_variable(name, shape=None, dtype=float32, **k):
        v = variables[name]
        assert shape is None or list(v.shape) == _shape_list(shape), (name, v.shape, shape)
        return v
    tf.get_variable = get_variable

    nn = tf.nn
    nn.sigmoid, nn.tanh, nn.relu = torch.sigmoid, torch.tanh, torch.relu
    nn.softplus = lambda x: torch.logaddexp(x, torch.zeros_like(x))          # log(exp(x) + 1), overflow-safe as in TF
    nn.log_softmax = lambda x, axis=-1: torch.log_softmax(x, dim=int(axis))
    nn.softmax = lambda x, axis=-1: torch.softmax(x, dim=int(axis))
    nn.softmax_cross_entropy_with_logits_v2 = lambda logits=None, labels=None, **k: -(labels * torch.log_softmax(logits, -1)).sum(-1)
    nn.sigmoid_cross_entropy_with_logits = lambda labels=None, logits=None, **k: (
        torch.clamp(logits, min=0) - logits * labels + torch.log1p(torch.exp(-logits.abs())))
    # tf.nn.weighted_cross_entropy_with_logits: (1 - z) x + (1 + (q - 1) z) (log1p(exp(-|x|)) + max(-x, 0))
    nn.weighted_cross_entropy_with_logits = lambda targets=None, logits=None, pos_weight=1.0, **k: (
        (1 - targets) * logits + (1 + (pos_weight - 1) * targets) * (torch.log1p(torch.exp(-logits.abs())) + torch.clamp(-logits, min=0)))

    def mean_squared_error(labels=None, predictions=None, weights=1.0, **k):
        # tf.losses default reduction SUM_BY_NONZERO_WEIGHTS: sum(w (l - p)^2) / count_nonzero(broadcast w)
        w = torch.as_tensor(weights, dtype=predictions.dtype) * torch.ones_like(predictions)
        return (w * (labels - predictions) ** 2).sum() / (w != 0).sum().to(predictions.dtype)
    tf.losses.mean_squared_error = mean_squared_error

    class Normal(object):                                        # tf.contrib.distributions.Normal
        def __init__(self, loc=None, scale=None, **k):
            self.loc, self.scale = loc, scale

        def cdf(self, x):                                        # special_math.ndtr((x - loc) / scale)
            return torch.special.ndtr((x - self.loc) / self.scale)

        def sample(self):
            n = _normal_queue.pop(0)
            assert n.shape == self.loc.shape
            return self.loc + self.scale * n
    tf.contrib.distributions.Normal = Normal
    # tf.train.exponential_decay (non-staircase): lr * rate ^ (step / decay_steps), float32 like the TF op
    tf.train.exponential_decay = lambda learning_rate, global_step, decay_steps, decay_rate, staircase=False, name=None: (
        torch.as_tensor(learning_rate, dtype=torch.float32) * torch.pow(torch.as_tensor(decay_rate, dtype=torch.float32),
                                                                          torch.as_tensor(global_step).to(torch.float32) / float(decay_steps)))
    tf.contrib.layers.xavier_initializer = lambda *a, **k: None
    tf.zeros_initializer = lambda *a, **k: None
    return tf


def install(extra_stubs=("librosa", "librosa.filters", "librosa.display", "librosa.effects", "librosa.core", "matplotlib",
                         "matplotlib.pyplot", "keras", "keras.utils", "lws", "unidecode", "inflect", "tqdm")):
    """Put the shim (and import-only stubs of the reference's other missing dependencies) into sys.modules."""
    tf = build()
    sys.modules["tensorflow"] = tf
    for sub in ("contrib", "contrib.seq2seq", "contrib.rnn", "contrib.framework", "python", "python.framework", "python.ops",
                "python.layers", "python.util", "contrib.seq2seq.python", "contrib.seq2seq.python.ops", "python.framework.ops",
                "python.framework.tensor_shape", "python.ops.array_ops", "python.ops.rnn_cell_impl", "python.ops.math_ops",
                "python.ops.nn_ops", "python.ops.variable_scope", "python.ops.check_ops", "python.ops.control_flow_ops",
                "python.ops.tensor_array_ops", "python.layers.core", "python.util.nest", "contrib.seq2seq.python.ops.attention_wrapper",
                "contrib.seq2seq.python.ops.decoder", "contrib.seq2seq.python.ops.helper", "python.ops.functional_ops",
                "python.framework.dtypes", "python.layers.base"):
        mod = tf
        for part in sub.split("."):
            mod = getattr(mod, part)
        sys.modules["tensorflow." + sub] = mod
    for name in extra_stubs:
        if name not in sys.modules:
            sys.modules[name] = _NoopStub(name)
            if "." in name:
                parent, child = name.rsplit(".", 1)
                setattr(sys.modules[parent], child, sys.modules[name])
    torch.Tensor.get_shape = lambda self: Shape(Dim(d) for d in self.size())   # TF static-shape protocol (this process only)
    if not hasattr(np, "int"):
        np.int = int          # removed in numpy 1.24; the reference (numpy 1.14) spells astype(np.int)
    return tf
