"""TEST INFRASTRUCTURE - the graph-building half of the TF-1 stand-in (tf_shim.py holds the elementary ops). With it the reference's
own model files (tacotron/models/{tacotron,modules,attention,Architecture_wrappers,helpers,custom_decoder}.py) can be imported and
`Tacotron.initialize()` / `add_loss()` EXECUTED eagerly on torch-CPU tensors: variable scopes with TensorFlow's naming and
uniquification rules, tf.get_variable backed by a recording store, the tf.layers / tf.nn.rnn_cell / tf.contrib.seq2seq classes the
reference composes (Dense, Conv1D, BatchNormalization, Dropout, MaxPooling1D, LSTMCell, GRUCell, MultiRNNCell, dynamic_rnn,
BahdanauAttention, dynamic_decode, TensorArray).

What this pins and what it does not: the LAYERS below restate the documented TF 1.x semantics - they are
this repo's reading of TensorFlow, like the oracle's. What is the reference's own and gets executed unchanged is everything
above the layers: which layers exist under which variable scope and name, their order, activation / batch-norm / dropout
placement, the zoneout wrapper, the encoder / decoder cell wiring (prenet -> LSTM stack -> attention -> projections, input feeding,
cumulative alignments), the helpers' teacher forcing and stop rule, the post-net, the CBHG block, the loss terms and the
regularisation filter. Random draws (dropout / zoneout masks) are generated here from a seeded generator and RECORDED so that the
oracle can be run with exactly the same masks.

The WaveNet side (wavenet_vocoder/models/{wavenet,modules,mixture,gaussian}.py) runs on the same stand-in: tf.layers.Conv1D / Conv2D /
Conv2DTranspose with build() / call() / data_format / dilation_rate as the reference's keras wrappers use them, tf.keras.layers.Wrapper,
tf.while_loop, tf.TensorArray, tf.multinomial (inverse CDF, draw recorded), tf.pad, tf.batch_to_space_nd, tf.image.resize_images.

Used only by tests/golden/make_reference_graph_vectors.py and make_reference_wavenet_graph_vectors.py."""
import collections
import contextlib
import math
import re
import types

import numpy as np
import torch

import tf_shim
from tf_shim import T

S = types.SimpleNamespace(scope=[""], opened={}, layer_uid={}, vars=collections.OrderedDict(), gen=None, drops=[], create=True, inits={}, uniforms=[], updates=collections.OrderedDict(), assigned={})


def reset(seed=0, variables=None, allow_create=False):
    """new graph: empty scope stack / name counters; `variables` (name -> tensor) are reused instead of created when given (missing
    ones are an error unless allow_create)"""
    S.scope[:] = [""]
    S.opened.clear()
    S.layer_uid.clear()
    S.vars.clear()
    S.drops[:] = []
    S.inits.clear()
    S.uniforms[:] = []
    S.updates.clear()
    S.assigned.clear()
    S.gen = torch.Generator().manual_seed(seed)
    S.create = variables is None or allow_create
    for k, v in (variables or {}).items():
        S.vars[k] = _as_var(k, torch.as_tensor(v, dtype=torch.float32).clone())


# ---- variable scopes -------------------------------------------------------------------------------------------------------------
class VarScope(object):
    def __init__(self, name):
        self.name = name


def _join(a, b):
    return b if not a else a + "/" + b


@contextlib.contextmanager
def variable_scope(name_or_scope=None, default_name=None, values=None, reuse=None, **kw):
    """tf.variable_scope: an explicit name nests under the current scope as is; a scope object re-enters that scope; with
    name_or_scope=None the default_name is made unique among the scopes opened so far (variable_scope.py
    _get_unique_variable_scope: prefix, prefix_1, prefix_2, ...)"""
    cur = S.scope[-1]
    if isinstance(name_or_scope, VarScope):
        full = name_or_scope.name
    elif name_or_scope is None:
        base = _join(cur, default_name)
        full, i = base, 0
        while S.opened.get(full, 0) > 0:
            i += 1
            full = "%s_%d" % (base, i)
    else:
        full = _join(cur, name_or_scope)
    S.opened[full] = S.opened.get(full, 0) + 1
    S.scope.append(full)
    try:
        yield VarScope(full)
    finally:
        S.scope.pop()


def loop_snapshot():
    """A while_loop body is BUILT once in TensorFlow; run eagerly it is executed once per step. Restoring the name counters to their
    value before the first iteration makes every iteration resolve the same scope / variable names."""
    return dict(S.opened), dict(S.layer_uid)


def loop_restore(snap):
    S.opened.clear()
    S.opened.update(snap[0])
    S.layer_uid.clear()
    S.layer_uid.update(snap[1])


class V(T):
    """a variable: a tensor that answers `.name` like tf.Variable ('<scope>/<name>:0')"""
    name = property(lambda self: getattr(self, "var_name", "") + ":0")


def _as_var(full, value):
    v = V(value)
    v.requires_grad_(not full.endswith(("moving_mean", "moving_variance")))
    v.var_name = full
    return v


def _new_value(full, shape):
    leaf = full.rsplit("/", 1)[-1]
    g = S.gen
    shape = [int(s) for s in shape]
    if leaf == "gamma":
        return 1.0 + 0.2 * torch.randn(shape, generator=g)
    if leaf == "moving_variance":
        return 0.5 + torch.rand(shape, generator=g)
    if leaf in ("beta", "bias", "moving_mean", "attention_bias"):
        return 0.2 * torch.randn(shape, generator=g)
    fan_in = int(np.prod(shape[:-1])) if len(shape) > 1 else 4
    return torch.randn(shape, generator=g) / math.sqrt(fan_in)


def get_variable(name, shape=None, dtype=torch.float32, initializer=None, trainable=True, **kw):
    full = _join(S.scope[-1], name)
    if full in S.vars:
        v = S.vars[full]
        assert shape is None or list(v.shape) == [int(s) for s in shape], (full, tuple(v.shape), shape)
        return v
    if isinstance(getattr(initializer, "value", None), np.ndarray) and initializer.value.ndim >= 2:
        S.inits[full] = np.array(initializer.value)      # NN_init kernels (modules.py:642-654,761-770): kept for a known-answer test
    if full in S.vars:
        return S.vars[full]
    assert S.create, "variable %s is not in the injected set" % full
    v = _as_var(full, _new_value(full, shape))
    S.vars[full] = v
    return v


def trainable_variables():
    return [v for v in S.vars.values() if v.requires_grad]


# ---- random draws ----------------------------------------------------------------------------------------------------------------
def _dropout(x, keep_prob, kind):
    """tf.nn.dropout: x / keep_prob * floor(keep_prob + U[0, 1)); keep_prob == 1 returns x"""
    keep_prob = float(keep_prob)
    if keep_prob >= 1.0:
        return x
    mask = torch.floor(keep_prob + torch.rand(tuple(x.shape), generator=S.gen))
    S.drops.append((S.scope[-1], kind, mask))
    return x / keep_prob * mask


# ---- layers ----------------------------------------------------------------------------------------------------------------------
def _snake(name):
    return re.sub(r"(?<=[a-z0-9])(?=[A-Z])|(?<=[A-Z])(?=[A-Z][a-z])", "_", name).lower().replace("conv1_d", "conv1d").replace("pooling1_d", "pooling1d")


def _unique_layer_name(base):
    key = (S.scope[-1], base)
    n = S.layer_uid.get(key, 0)
    S.layer_uid[key] = n + 1
    return base if n == 0 else "%s_%d" % (base, n)


class Layer(object):
    """tf.layers.Layer naming: the variable scope is fixed at the first call - variable_scope(None, default_name=<name>) under the
    scope current at that moment (functional layers with an explicit name: variable_scope(<name>), no uniquification)"""

    def __init__(self, name=None, _scope=None, **kw):
        self._name_arg, self._explicit_scope, self._scope = name, _scope, None
        self._base_name = name or _snake(type(self).__name__)
        self.built = False

    def _enter(self):
        if getattr(self, "_scope", None) is None:
            if getattr(self, "_explicit_scope", None):
                cm = variable_scope(self._explicit_scope)
            else:
                cm = variable_scope(None, default_name=getattr(self, "_name_arg", None) or _unique_layer_name(_snake(type(self).__name__)))
            with cm as sc:
                self._scope = sc
        return variable_scope(self._scope)

    def __call__(self, inputs, *args, **kwargs):
        kwargs.pop("scope", None)
        if hasattr(self, "build") and not getattr(self, "built", False):
            self.build(inputs.shape)
        with self._enter():
            return self.call(inputs, *args, **kwargs)


def _tuple(v, n):
    return tuple(int(x) for x in v) if isinstance(v, (tuple, list)) else (int(v),) * n


class TensorShape(list):
    def __init__(self, dims=()):
        list.__init__(self, [None if d is None else int(d) for d in dims])

    def as_list(self):
        return list(self)


class ConstantInitializer(object):
    def __init__(self, value=0, dtype=None, **kw):
        self.value = value


def _act(fn, y):
    return y if fn is None else fn(y)


class Dense(Layer):
    def __init__(self, units, activation=None, use_bias=True, name=None, **kw):
        Layer.__init__(self, name=name, _scope=kw.get("_scope"))
        self.units, self.activation, self.use_bias = int(units), activation, use_bias

    def call(self, x):
        y = x @ get_variable("kernel", [x.shape[-1], self.units])
        if self.use_bias:
            y = y + get_variable("bias", [self.units])
        return _act(self.activation, y)


class _Conv(Layer):
    """tf.layers.Conv1D / Conv2D / Conv2DTranspose constructor surface; variables are made in build() (kernel, then bias) under the
    layer's scope - also when build() is called directly (the reference's keras wrappers do that: add_weight sets the scope)"""
    rank = 1

    def __init__(self, filters, kernel_size, strides=1, padding="valid", data_format="channels_last", dilation_rate=1, activation=None,
                 use_bias=True, kernel_initializer=None, bias_initializer=None, name=None, **kw):
        Layer.__init__(self, name=name, _scope=kw.get("_scope"))
        self.filters, self.kernel_size, self.strides = int(filters), _tuple(kernel_size, self.rank), _tuple(strides, self.rank)
        self.padding, self.data_format, self.dilation_rate = padding.lower(), data_format, _tuple(dilation_rate, self.rank)
        self.activation, self.use_bias, self.kernel_initializer = activation, use_bias, kernel_initializer
        self.kernel = self.bias = None

    def _cin(self, input_shape):
        return int(input_shape[1] if self.data_format == "channels_first" else input_shape[-1])

    def _kernel_shape(self, cin):
        return list(self.kernel_size) + [cin, self.filters]

    def build(self, input_shape):
        with self._enter():
            self.kernel = get_variable("kernel", self._kernel_shape(self._cin(input_shape)), initializer=self.kernel_initializer)
            self.bias = get_variable("bias", [self.filters]) if self.use_bias else None
        self.built = True

    def _finish(self, y, channel_axis):
        if self.use_bias:
            shape = [1] * y.dim()
            shape[channel_axis] = self.filters
            y = y + self.bias.reshape(shape)
        return _act(self.activation, y)


def _same_pad(x, axis, kernel_extent):
    """TF SAME for stride 1: extent - 1 zeros in total, total // 2 of them in front"""
    total = kernel_extent - 1
    if total <= 0:
        return x
    shape = list(x.shape)
    shape[axis] = total // 2
    front = x.new_zeros(shape)
    shape[axis] = total - total // 2
    return torch.cat([front, x, x.new_zeros(shape)], dim=axis)


class Conv1D(_Conv):
    """cross-correlation over time with dilation; 'valid' or 'same'; channels_last [B, T, C] or channels_first [B, C, T]"""
    rank = 1

    def call(self, x):
        assert self.strides == (1,)
        if self.data_format == "channels_first":
            x = x.transpose(1, 2)
        kw, d = self.kernel_size[0], self.dilation_rate[0]
        extent = (kw - 1) * d + 1
        if self.padding == "same":
            x = _same_pad(x, 1, extent)
        windows = x.unfold(1, extent, 1)[..., ::d]                           # [B, T', Cin, kw]
        y = torch.einsum("btck,kcf->btf", windows, self.kernel)
        if self.data_format == "channels_first":
            return self._finish(y.transpose(1, 2), 1)
        return self._finish(y, -1)


class Conv2D(_Conv):
    """stride-1 cross-correlation, channels_last [B, H, W, C], kernel [kh, kw, in, out]"""
    rank = 2

    def call(self, x):
        assert self.strides == (1, 1) and self.dilation_rate == (1, 1) and self.data_format == "channels_last"
        kh, kw = self.kernel_size
        if self.padding == "same":
            x = _same_pad(_same_pad(x, 1, kh), 2, kw)
        windows = x.unfold(1, kh, 1).unfold(2, kw, 1)                        # [B, H', W', Cin, kh, kw]
        return self._finish(torch.einsum("bhwcij,ijcf->bhwf", windows, self.kernel), -1)


class Conv2DTranspose(_Conv):
    """tf.layers.Conv2DTranspose, channels_first [B, C, H, W], kernel [kh, kw, out, in]: the gradient of a strided SAME convolution -
    input pixel (y, x) adds x * kernel[i, j] at output (y sy + i - pad_top, x sx + j - pad_left); with 'same' the output is
    [H sy, W sx] and pad = max(k - s, 0) // 2 per axis"""
    rank = 2

    def _kernel_shape(self, cin):
        return list(self.kernel_size) + [self.filters, cin]

    def call(self, x):
        assert self.data_format == "channels_first" and self.padding == "same" and self.dilation_rate == (1, 1)
        B, C, H, W = [int(v) for v in x.shape]
        (kh, kw), (sy, sx) = self.kernel_size, self.strides
        full = x.new_zeros(B, self.filters, (H - 1) * sy + kh, (W - 1) * sx + kw)
        for i in range(kh):
            for j in range(kw):
                contrib = torch.einsum("bchw,oc->bohw", x, self.kernel[i, j])
                full[:, :, i:i + (H - 1) * sy + 1:sy, j:j + (W - 1) * sx + 1:sx] += contrib
        top, left = max(kh - sy, 0) // 2, max(kw - sx, 0) // 2
        return self._finish(full[:, :, top:top + H * sy, left:left + W * sx], 1)


class Wrapper(object):
    """tf.keras.layers.Wrapper: holds `layer`; __call__ builds once (with the input's shape) and then calls `call`. keras' base
    Layer.__call__ opens a NAME scope only, so nothing here touches the variable scopes."""

    def __init__(self, layer, name=None, **kw):
        self.layer, self._name, self.built = layer, name, False

    def build(self, input_shape=None):
        self.built = True

    def _track_checkpointable(self, *a, **kw):
        pass

    def __call__(self, inputs, *args, **kwargs):
        if not self.built:
            self.build(inputs.shape)
        return self.call(inputs, *args, **kwargs)


class BatchNormalization(Layer):
    """tf.layers.batch_normalization defaults: axis -1, epsilon 1e-3; training: moments of the batch over every other axis (biased
    variance); inference: the moving statistics. (The moving-average update ops live in UPDATE_OPS; they do not change the outputs.)"""

    def call(self, x, training=False):
        C = x.shape[-1]
        gamma, beta = get_variable("gamma", [C]), get_variable("beta", [C])
        mm, mv = get_variable("moving_mean", [C], trainable=False), get_variable("moving_variance", [C], trainable=False)
        if training:
            flat = x.reshape(-1, int(C))
            mean = flat.mean(0)
            var = ((flat - mean) ** 2).mean(0)
            # UPDATE_OPS (momentum 0.99; the non-fused path of tf.layers - the fused kernel needs 4-D inputs - feeds the same biased
            # batch variance into the moving average): moving <- moving * 0.99 + batch * 0.01
            S.updates[mm.var_name] = (mm * 0.99 + mean * 0.01).detach()
            S.updates[mv.var_name] = (mv * 0.99 + var * 0.01).detach()
        else:
            mean, var = mm, mv
        return (x - mean) * torch.rsqrt(var + 1e-3) * gamma + beta


class MaxPooling1D(Layer):
    pass


def max_pooling1d(x, pool_size, strides, padding="valid", **kw):
    assert strides == 1 and padding == "same"
    total = int(pool_size) - 1
    neg = x.new_full((x.shape[0], 1, x.shape[2]), -float("inf"))
    xp = torch.cat([neg] * (total // 2) + [x] + [neg] * (total - total // 2), dim=1)
    return xp.unfold(1, int(pool_size), 1).amax(-1)


def layers_dropout(x, rate=0.5, training=False, name=None, **kw):
    return _dropout(x, 1.0 - float(rate), "layers.dropout") if training else x


LSTMStateTuple = collections.namedtuple("LSTMStateTuple", ("c", "h"))


def map_structure(fn, *structs):
    s0 = structs[0]
    if isinstance(s0, tuple) and hasattr(s0, "_fields"):
        return type(s0)(*[map_structure(fn, *[s[i] for s in structs]) for i in range(len(s0))])
    if isinstance(s0, (tuple, list)):
        return type(s0)(map_structure(fn, *[s[i] for s in structs]) for i in range(len(s0)))
    return fn(*structs)


def _zero_state_tensors(state_size, batch_size, dtype):
    return map_structure(lambda n: T(torch.zeros(int(batch_size), int(n), dtype=dtype)), state_size)


class RNNCell(Layer):
    def __init__(self, *a, **kw):
        Layer.__init__(self, name=kw.get("name"))

    def zero_state(self, batch_size, dtype):
        return _zero_state_tensors(self.state_size, batch_size, dtype)

    def __call__(self, inputs, state, scope=None):
        assert scope is None
        return Layer.__call__(self, inputs, state)


class LSTMCell(RNNCell):
    """tf.nn.rnn_cell.LSTMCell without peepholes / projection: [x, h] W + b -> i, j, f, o; c' = c sigmoid(f + forget_bias) + sigmoid(i)
    tanh(j); h' = tanh(c') sigmoid(o); forget_bias = 1"""

    def __init__(self, num_units, state_is_tuple=True, name=None, **kw):
        RNNCell.__init__(self, name=name)
        self._num_units, self._num_proj = int(num_units), None
        assert state_is_tuple

    state_size = property(lambda self: LSTMStateTuple(self._num_units, self._num_units))
    output_size = property(lambda self: self._num_units)

    def call(self, x, state):
        c, h = state
        n = self._num_units
        z = torch.cat([x, h], dim=1) @ get_variable("kernel", [x.shape[-1] + n, 4 * n]) + get_variable("bias", [4 * n])
        i, j, f, o = z[:, :n], z[:, n:2 * n], z[:, 2 * n:3 * n], z[:, 3 * n:]
        new_c = c * torch.sigmoid(f + 1.0) + torch.sigmoid(i) * torch.tanh(j)
        new_h = torch.tanh(new_c) * torch.sigmoid(o)
        return new_h, LSTMStateTuple(new_c, new_h)


class GRUCell(RNNCell):
    """tf.nn.rnn_cell.GRUCell: [r, u] = sigmoid([x, h] W_gates + b_gates); c = tanh([x, r h] W_cand + b_cand); h' = u h + (1 - u) c"""

    def __init__(self, num_units, name=None, **kw):
        RNNCell.__init__(self, name=name)
        self._num_units = int(num_units)

    state_size = property(lambda self: self._num_units)
    output_size = property(lambda self: self._num_units)

    def call(self, x, h):
        n = self._num_units
        d = x.shape[-1] + n
        g = torch.sigmoid(torch.cat([x, h], 1) @ get_variable("gates/kernel", [d, 2 * n]) + get_variable("gates/bias", [2 * n]))
        r, u = g[:, :n], g[:, n:]
        c = torch.tanh(torch.cat([x, r * h], 1) @ get_variable("candidate/kernel", [d, n]) + get_variable("candidate/bias", [n]))
        new_h = u * h + (1 - u) * c
        return new_h, new_h


class MultiRNNCell(RNNCell):
    def __init__(self, cells, state_is_tuple=True):
        RNNCell.__init__(self)
        self._cells = list(cells)
        assert state_is_tuple

    state_size = property(lambda self: tuple(c.state_size for c in self._cells))
    output_size = property(lambda self: self._cells[-1].output_size)

    def zero_state(self, batch_size, dtype):
        return tuple(c.zero_state(batch_size, dtype) for c in self._cells)

    def call(self, x, state):
        new_states = []
        for i, cell in enumerate(self._cells):
            with variable_scope("cell_%d" % i):
                x, ns = cell(x, state[i])
            new_states.append(ns)
        return x, tuple(new_states)


def _reverse_sequence(x, lengths):
    if lengths is None:
        return x.flip(1)
    out = x.clone()
    for b in range(x.shape[0]):
        n = int(lengths[b])
        out[b, :n] = x[b, :n].flip(0)
    return out


def _dynamic_rnn(cell, x, lengths, scope):
    """tf.nn.dynamic_rnn with sequence_length: past a row's length the output is zero and the state is copied through"""
    with variable_scope(scope):
        state = cell.zero_state(x.shape[0], torch.float32)
        outs, snap = [], None
        for t in range(int(x.shape[1])):
            if snap is None:
                snap = loop_snapshot()
            else:
                loop_restore(snap)
            out, new_state = cell(x[:, t], state)
            if lengths is not None:
                live = (t < torch.as_tensor(lengths).reshape(-1)).unsqueeze(1)
                out = torch.where(live, out, torch.zeros_like(out))
                state = map_structure(lambda n, o: torch.where(live, n, o), new_state, state)
            else:
                state = new_state
            outs.append(out)
        return torch.stack(outs, dim=1), state


def bidirectional_dynamic_rnn(cell_fw, cell_bw, inputs, sequence_length=None, dtype=None, swap_memory=False, scope=None, **kw):
    with variable_scope(scope or "bidirectional_rnn"):
        with variable_scope("fw") as fw_scope:
            out_fw, st_fw = _dynamic_rnn(cell_fw, inputs, sequence_length, fw_scope)
        with variable_scope("bw") as bw_scope:
            rev = _reverse_sequence(inputs, sequence_length)
            out_bw, st_bw = _dynamic_rnn(cell_bw, rev, sequence_length, bw_scope)
            out_bw = _reverse_sequence(out_bw, sequence_length)
    return (out_fw, out_bw), (st_fw, st_bw)


# ---- seq2seq ---------------------------------------------------------------------------------------------------------------------
class BahdanauAttention(object):
    """tf.contrib.seq2seq.BahdanauAttention as far as the reference's subclass uses it (attention_wrapper.py _BaseAttentionMechanism):
    memory masked by its lengths (_prepare_memory), keys = memory_layer(values) built right here, query_layer built at the first
    step, probability_fn = softmax over scores whose padded positions are set to -inf (_maybe_mask_score)"""

    def __init__(self, num_units, memory, memory_sequence_length=None, normalize=False, probability_fn=None, score_mask_value=None,
                 dtype=None, name="BahdanauAttention"):
        fn = probability_fn if probability_fn is not None else (lambda score: torch.softmax(score, dim=-1))
        Tm = int(memory.shape[1])
        if memory_sequence_length is not None:
            seq_mask = torch.arange(Tm)[None, :] < torch.as_tensor(memory_sequence_length).reshape(-1, 1)
            memory = memory * seq_mask.unsqueeze(-1).to(memory.dtype)
            mask_value = -float("inf") if score_mask_value is None else score_mask_value
            self._probability_fn = lambda score, prev: fn(torch.where(seq_mask, score, torch.full_like(score, mask_value)))
        else:
            self._probability_fn = lambda score, prev: fn(score)
        self.query_layer = Dense(num_units, name="query_layer", use_bias=False)
        self.memory_layer = Dense(num_units, name="memory_layer", use_bias=False)
        self._values = memory
        self._keys = self.memory_layer(memory)
        self._batch_size, self._alignments_size = int(memory.shape[0]), Tm

    values = property(lambda self: self._values)
    keys = property(lambda self: self._keys)
    batch_size = property(lambda self: self._batch_size)
    alignments_size = property(lambda self: self._alignments_size)

    def initial_alignments(self, batch_size, dtype):
        return T(torch.zeros(int(batch_size), self._alignments_size, dtype=dtype))

    initial_state = initial_alignments


class TensorArray(object):
    def __init__(self, dtype=None, size=0, dynamic_size=True, **kw):
        self._items = []

    def write(self, index, value):
        assert int(index) == len(self._items)
        new = TensorArray()
        new._items = self._items + [value]
        return new

    def stack(self):
        return torch.stack(self._items, dim=0)


class Helper(object):
    pass


class Decoder(object):
    pass


def dynamic_decode(decoder, output_time_major=False, impute_finished=False, maximum_iterations=None, swap_memory=False, scope=None, **kw):
    """tf.contrib.seq2seq.dynamic_decode (decoder.py): variable scope 'decoder'; loop while not all(finished); a step's `finished`
    is OR-ed with the running flags and with time + 1 >= maximum_iterations; outputs are stacked over time and returned batch-major"""
    assert not impute_finished and not output_time_major
    with variable_scope(scope, default_name="decoder"):
        finished, inputs, state = decoder.initialize()
        finished = torch.as_tensor(finished).clone()
        lengths = torch.zeros_like(finished, dtype=torch.int64)
        time, steps, snap = 0, [], None
        while not bool(finished.all()):
            if snap is None:
                snap = loop_snapshot()
            else:
                loop_restore(snap)
            outputs, state, inputs, dec_finished = decoder.step(time, inputs, state)
            nxt = torch.logical_or(torch.as_tensor(dec_finished), finished)
            if maximum_iterations is not None:
                nxt = torch.logical_or(nxt, torch.as_tensor(time + 1 >= int(maximum_iterations)))
            lengths = torch.where(finished, lengths, torch.full_like(lengths, time + 1))
            steps.append(outputs)
            finished, time = nxt, time + 1
        stacked = type(steps[0])(*[torch.stack([torch.as_tensor(s[i]) for s in steps], dim=1) for i in range(len(steps[0]))])
    return stacked, state, lengths


# ---- tf.train --------------------------------------------------------------------------------------------------------------------
class AdamOptimizer(object):
    """tf.train.AdamOptimizer: slots m, v start at zero, beta powers at beta1 / beta2, so the first apply_gradients uses t = 1:
    lr_t = lr sqrt(1 - beta2^t) / (1 - beta1^t); m <- b1 m + (1 - b1) g; v <- b2 v + (1 - b2) g^2; var <- var - lr_t m / (sqrt(v) + eps).
    apply_gradients does not assign in place: the new values land in `self.new_values` (name -> tensor) for the generator to store."""

    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8, **kw):
        self.lr, self.b1, self.b2, self.eps, self.t = learning_rate, float(beta1), float(beta2), float(epsilon), 0
        self.m, self.v, self.new_values = {}, {}, collections.OrderedDict()

    def compute_gradients(self, loss, var_list=None, **kw):
        variables = list(var_list) if var_list is not None else trainable_variables()
        grads = torch.autograd.grad(loss, variables, retain_graph=True, allow_unused=True)
        return list(zip(grads, variables))

    def apply_gradients(self, grads_and_vars, global_step=None, **kw):
        self.t += 1
        lr = float(torch.as_tensor(self.lr))
        lr_t = lr * math.sqrt(1 - self.b2 ** self.t) / (1 - self.b1 ** self.t)
        for g, var in grads_and_vars:
            if g is None:
                continue
            k = var.var_name
            g = g.detach()
            self.m[k] = self.b1 * self.m.get(k, torch.zeros_like(g)) + (1 - self.b1) * g
            self.v[k] = self.b2 * self.v.get(k, torch.zeros_like(g)) + (1 - self.b2) * g * g
            self.new_values[k] = var.detach() - lr_t * self.m[k] / (torch.sqrt(self.v[k]) + self.eps)
            S.assigned[k] = self.new_values[k]
        return self


class ExponentialMovingAverage(object):
    """tf.train.ExponentialMovingAverage(decay) without num_updates: the shadow starts at the variable's value when apply() builds
    it and moves by shadow -= (1 - decay) (shadow - variable); apply() runs after the optimizer (the reference puts it under a control
    dependency on the Adam op, wavenet.py:611-613), so `variable` is the value Adam just assigned (S.assigned)"""

    def __init__(self, decay=0.999, **kw):
        self.decay, self.shadow = float(decay), collections.OrderedDict()

    def apply(self, var_list=None):
        for var in var_list:
            k = var.var_name
            start = var.detach()
            self.shadow[k] = start - (1 - self.decay) * (start - S.assigned.get(k, start))
        return self


def clip_by_global_norm(t_list, clip_norm, **kw):
    """t * clip_norm / max(global_norm, clip_norm)"""
    gn = torch.sqrt(sum((t.detach() ** 2).sum() for t in t_list if t is not None))
    scale = float(clip_norm) / max(float(gn), float(clip_norm))
    return [None if t is None else t * scale for t in t_list], gn


def clip_by_norm(t, clip_norm, **kw):
    n = torch.sqrt((t * t).sum())
    return t * float(clip_norm) / torch.clamp(n, min=float(clip_norm))


# ---- install ---------------------------------------------------------------------------------------------------------------------
def install():
    """tf_shim.install() + the graph-building surface; returns the `tensorflow` stand-in"""
    tf = tf_shim.install()
    wrap = lambda f: (lambda *a, **k: T(f(*a, **k)))
    tf.variable_scope, tf.get_variable, tf.trainable_variables = variable_scope, get_variable, trainable_variables
    tf.name_scope = lambda *a, **k: contextlib.nullcontext()
    tf.device = lambda *a, **k: contextlib.nullcontext()
    tf.train.replica_device_setter = lambda *a, **k: None
    tf.train.AdamOptimizer, tf.train.ExponentialMovingAverage = AdamOptimizer, ExponentialMovingAverage

    def cosine_decay(learning_rate, global_step, decay_steps, alpha=0.0, name=None):
        """tf.train.cosine_decay: step = min(global_step, decay_steps); lr ((1 - alpha) 0.5 (1 + cos(pi step / decay_steps)) + alpha)"""
        step = min(float(torch.as_tensor(global_step)), float(decay_steps))
        return T(torch.tensor(float(learning_rate) * ((1.0 - alpha) * 0.5 * (1.0 + math.cos(math.pi * step / float(decay_steps))) + alpha),
                              dtype=torch.float32))
    tf.train.cosine_decay = cosine_decay
    tf.clip_by_global_norm, tf.clip_by_norm = clip_by_global_norm, clip_by_norm
    tf.clip_by_value = lambda t, lo, hi, **k: torch.clamp(t, float(lo), float(hi))
    tf.get_collection = lambda key, *a, **k: list(S.updates.items())
    tf.GraphKeys.UPDATE_OPS = "update_ops"
    tf.identity = lambda x, name=None: x
    tf.zeros, tf.ones = wrap(tf.zeros), wrap(tf.ones)
    keep = lambda t: t if isinstance(t, T) else T(t)          # T(...) makes a new leaf: never re-wrap a tensor that is already in the graph
    tf.convert_to_tensor = lambda x, dtype=None, **k: keep(x if isinstance(x, torch.Tensor) and dtype in (None, x.dtype) else torch.as_tensor(x, dtype=dtype))
    tf.tile = lambda x, multiples: keep((x if isinstance(x, torch.Tensor) else torch.as_tensor(x)).repeat(*[int(m) for m in multiples]))
    tf.less = lambda a, b: torch.as_tensor(a) < torch.as_tensor(b)
    tf.logical_or = torch.logical_or
    tf.reduce_all = lambda x, axis=None, **k: torch.as_tensor(x).all() if axis is None else torch.as_tensor(x).all(dim=int(axis))
    tf.reduce_any = lambda x, axis=None, **k: torch.as_tensor(x).any() if axis is None else torch.as_tensor(x).any(dim=int(axis))
    tf.split = lambda value, num_or_size_splits, axis=0, **k: list(torch.chunk(value, int(num_or_size_splits), dim=int(axis)))
    tf.add_n = lambda xs: sum(xs[1:], xs[0])
    tf.concat = lambda values=None, axis=0, **k: torch.cat(list(values), dim=int(axis))
    tf.matmul = torch.matmul
    base_uniform = tf.random_uniform

    def random_uniform(shape, minval=0.0, maxval=1.0, dtype=torch.float32, **k):
        if tf_shim._uniform_queue:
            return base_uniform(shape, minval, maxval, dtype)
        u = T(minval + (maxval - minval) * torch.rand([int(s) for s in shape], generator=S.gen))
        S.uniforms.append(("random_uniform", u))
        return u
    tf.random_uniform = random_uniform

    def multinomial(logits, num_samples, **k):
        """one categorical draw per row by inverse CDF over softmax(logits) - tf.multinomial takes LOGITS; the uniform is recorded"""
        assert int(num_samples) == 1
        u = torch.rand(int(logits.shape[0]), 1, generator=S.gen)
        S.uniforms.append(("multinomial", u))
        cdf = torch.cumsum(torch.softmax(logits, dim=-1), dim=-1)
        return (cdf < u).sum(-1, keepdim=True).clamp(max=int(logits.shape[-1]) - 1)
    tf.multinomial = multinomial
    tf.Print = lambda x, data=None, **k: x
    tf.TensorArray = TensorArray
    tf.one_hot = lambda indices, depth, dtype=torch.float32, **k: torch.nn.functional.one_hot(torch.as_tensor(indices).long(), int(depth)).to(dtype)

    def while_loop(cond, body, loop_vars, **k):
        state, snap = list(loop_vars), None
        while bool(torch.as_tensor(cond(*state))):
            if snap is None:
                snap = loop_snapshot()
            else:
                loop_restore(snap)
            state = list(body(*state))
        return state
    tf.while_loop = while_loop
    base_reduce_max = tf.reduce_max
    tf.reduce_max = lambda x, *a, **k: base_reduce_max(x if isinstance(x, torch.Tensor) else torch.as_tensor([int(v) for v in x]), *a, **k)
    base_sequence_mask = tf.sequence_mask
    tf.sequence_mask = lambda lengths, maxlen=None, dtype=torch.bool, **k: T(base_sequence_mask(
        lengths if isinstance(lengths, torch.Tensor) else torch.as_tensor([int(v) for v in lengths]), maxlen, dtype))

    def py_func(func, inp, Tout, **k):
        out = func(*[np.asarray(torch.as_tensor(x).detach().cpu().numpy()) for x in inp])
        return [T(torch.as_tensor(np.ascontiguousarray(o))) for o in out]
    tf.py_func = py_func

    nn = tf.nn
    nn.relu = lambda x, name=None: torch.relu(x)
    nn.leaky_relu = lambda x, alpha=0.2, name=None: torch.where(x >= 0, x, alpha * x)
    nn.bias_add = lambda x, b, **k: x + b
    base_normal = tf.contrib.distributions.Normal

    class Normal(base_normal):
        def sample(self):
            if tf_shim._normal_queue:
                return base_normal.sample(self)
            n = torch.randn(tuple(self.loc.shape), generator=S.gen)
            S.uniforms.append(("normal", n))
            return self.loc + self.scale * n
    tf.contrib.distributions.Normal = Normal
    nn.embedding_lookup = lambda table, ids, **k: table[torch.as_tensor(ids).long()]
    nn.l2_loss = lambda v: (v * v).sum() / 2
    nn.dropout = lambda x, keep_prob, **k: _dropout(x, keep_prob, "nn.dropout")
    nn.bidirectional_dynamic_rnn = bidirectional_dynamic_rnn
    cells = nn.rnn_cell
    cells.RNNCell, cells.LSTMCell, cells.GRUCell, cells.LSTMStateTuple, cells.MultiRNNCell = RNNCell, LSTMCell, GRUCell, LSTMStateTuple, MultiRNNCell
    tf.contrib.rnn.RNNCell, tf.contrib.rnn.MultiRNNCell, tf.contrib.rnn.LSTMStateTuple = RNNCell, MultiRNNCell, LSTMStateTuple

    L = tf.layers
    L.Layer, L.Dense, L.Conv1D = Layer, Dense, Conv1D
    L.dense = lambda inputs, units, activation=None, use_bias=True, name=None, **k: Dense(units, activation, use_bias, name=name, _scope=name)(inputs)
    L.conv1d = lambda inputs, filters, kernel_size, padding="valid", activation=None, use_bias=True, name=None, **k: Conv1D(
        filters, kernel_size, padding=padding, activation=activation, use_bias=use_bias, name=name, _scope=name)(inputs)
    L.Conv2D, L.Conv2DTranspose = Conv2D, Conv2DTranspose
    tf.keras.layers.Wrapper = Wrapper
    tf.TensorShape = TensorShape
    tf.constant_initializer = ConstantInitializer
    tf.truncated_normal_initializer = lambda *a, **k: None
    tf.constant = lambda v, *a, **k: v

    def pad(x, paddings, **k):
        for axis, (before, after) in enumerate([[int(a) for a in p] for p in paddings]):
            if before or after:
                shape = list(x.shape)
                shape[axis] = before
                front = x.new_zeros(shape)
                shape[axis] = after
                x = torch.cat([front, x, x.new_zeros(shape)], dim=axis)
        return x
    tf.pad = pad

    def batch_to_space_nd(x, block_shape, crops, **k):
        """one block dimension, no crops: batch index = j * (batch / b) + n  ->  out[n, i * b + j]"""
        (b,) = [int(v) for v in block_shape]
        assert [list(c) for c in crops] == [[0, 0]]
        shape = [int(v) for v in x.shape]
        y = x.reshape([b, shape[0] // b] + shape[1:])
        y = y.permute([1, 2, 0] + list(range(3, y.dim())))
        return y.reshape([shape[0] // b, shape[1] * b] + shape[2:])
    tf.batch_to_space_nd = batch_to_space_nd

    def resize_images(images, size, method=0, **k):
        """method 1 = nearest neighbour, align_corners False: out[i] = in[floor(i * in / out)]"""
        assert method == 1
        H, W = int(images.shape[1]), int(images.shape[2])
        oh, ow = int(size[0]), int(size[1])
        rows = (torch.arange(oh) * H) // oh
        cols = (torch.arange(ow) * W) // ow
        return images[:, rows][:, :, cols]
    tf.image.resize_images = resize_images
    L.batch_normalization = lambda inputs, training=False, name=None, **k: BatchNormalization(name=name, _scope=name)(inputs, training=training)
    L.dropout = layers_dropout
    L.max_pooling1d = max_pooling1d

    s2s = tf.contrib.seq2seq
    s2s.dynamic_decode, s2s.Helper, s2s.BahdanauAttention = dynamic_decode, Helper, BahdanauAttention
    s2s.python.ops.attention_wrapper.BahdanauAttention = BahdanauAttention
    s2s.python.ops.helper.Helper = Helper
    s2s.python.ops.decoder.Decoder = Decoder

    py = tf.python
    py.layers.base.Layer = Layer
    ao = py.ops.array_ops
    ao.expand_dims, ao.squeeze, ao.concat, ao.shape, ao.identity = tf.expand_dims, tf.squeeze, tf.concat, tf.shape, tf.identity
    ao.zeros = tf.zeros
    py.ops.math_ops.matmul = torch.matmul
    py.ops.variable_scope.variable_scope = variable_scope
    py.ops.check_ops.assert_equal = lambda *a, **k: None
    py.ops.rnn_cell_impl._zero_state_tensors = _zero_state_tensors
    py.ops.rnn_cell_impl.assert_like_rnncell = lambda *a, **k: None
    py.ops.tensor_array_ops.TensorArray = TensorArray
    py.framework.ops.name_scope = lambda *a, **k: contextlib.nullcontext()
    py.framework.ops.control_dependencies = lambda *a, **k: contextlib.nullcontext()
    py.util.nest.map_structure = map_structure
    return tf


def selfcheck():
    """The layer stand-ins against torch.nn.functional's own kernels (an implementation none of this repo's code shares): convolutions
    (VALID / SAME incl. even kernels and dilation, 2-D, transposed 2-D with SAME cropping), batch norm, max pool, LSTM cell (torch's
    gate order i, f, g, o re-ordered to TF's i, j, f, o with forget_bias folded into the bias)."""
    import torch.nn.functional as F
    install()
    g = torch.Generator().manual_seed(3)
    rnd = lambda *s: torch.randn(*s, generator=g)
    worst = {}

    def check(name, a, b):
        worst[name] = float((a - b).abs().max())
        assert worst[name] < 5e-5, (name, worst[name])
    reset(seed=1)
    x = rnd(2, 11, 5)
    for kw, d, pad in ((3, 1, "same"), (4, 1, "same"), (2, 1, "same"), (3, 2, "valid"), (5, 3, "same")):
        with variable_scope("c%d_%d_%s" % (kw, d, pad)):
            layer = Conv1D(7, kw, padding=pad, dilation_rate=d)
            y = layer(x)
        w = layer.kernel.detach().permute(2, 1, 0)                      # [out, in, kw]
        xi = x.transpose(1, 2)
        if pad == "same":
            total = (kw - 1) * d
            xi = F.pad(xi, (total // 2, total - total // 2))
        check("conv1d k%d d%d %s" % (kw, d, pad), y.detach(), F.conv1d(xi, w, layer.bias.detach(), dilation=d).transpose(1, 2))
    x4 = rnd(2, 6, 9, 3)
    with variable_scope("c2d"):
        layer = Conv2D(4, (3, 3), padding="same")
        y = layer(x4)
    check("conv2d same", y.detach(), F.conv2d(x4.permute(0, 3, 1, 2), layer.kernel.detach().permute(3, 2, 0, 1), layer.bias.detach(),
                                              padding=1).permute(0, 2, 3, 1))
    xt = rnd(2, 1, 6, 5)
    for kh, kw, s in ((3, 4, 4), (3, 2, 2), (3, 5, 5)):
        with variable_scope("ct_%d_%d" % (kw, s)):
            layer = Conv2DTranspose(1, (kh, kw), strides=(1, s), padding="same", data_format="channels_first")
            y = layer(xt)
        ref = F.conv_transpose2d(xt, layer.kernel.detach().permute(3, 2, 0, 1), layer.bias.detach(), stride=(1, s), padding=(kh // 2, 0))
        check("conv2d_transpose k%d s%d" % (kw, s), y.detach(), ref)
    xb = rnd(3, 7, 6)
    with variable_scope("bn"):
        bn = BatchNormalization()
        y_train, y_inf = bn(xb, training=True), bn(xb, training=False)
    v = {k.rsplit("/", 1)[1]: t.detach() for k, t in S.vars.items() if k.startswith("bn/")}
    flat = xb.reshape(-1, 6)
    check("batch_norm training", y_train.detach().reshape(-1, 6), F.batch_norm(flat, None, None, v["gamma"], v["beta"], True, 0.0, 1e-3))
    check("batch_norm inference", y_inf.detach().reshape(-1, 6), F.batch_norm(flat, v["moving_mean"], v["moving_variance"], v["gamma"], v["beta"], False, 0.0, 1e-3))
    check("max_pool same", max_pooling1d(xb, 2, 1, "same"), F.max_pool1d(F.pad(xb.transpose(1, 2), (0, 1), value=-float("inf")), 2, 1).transpose(1, 2))
    n, xin = 5, rnd(4, 3)
    with variable_scope("lstm"):
        cell = LSTMCell(n, name="cell")
        c0, h0 = rnd(4, n), rnd(4, n)
        out, (c1, h1) = cell(xin, LSTMStateTuple(c0, h0))
    K, b = S.vars["lstm/cell/kernel"].detach(), S.vars["lstm/cell/bias"].detach()
    i, j, f, o = (K[:, k * n:(k + 1) * n] for k in range(4))
    bi, bj, bf, bo = (b[k * n:(k + 1) * n] for k in range(4))
    Wt = torch.cat([i, f, j, o], dim=1).t()                              # torch order: input, forget, cell (g), output
    bt = torch.cat([bi, bf + 1.0, bj, bo])
    h_ref, c_ref = torch._VF.lstm_cell(xin, (h0, c0), Wt[:, :3].contiguous(), Wt[:, 3:].contiguous(), bt, torch.zeros_like(bt))
    check("lstm_cell h", h1.detach(), h_ref)
    check("lstm_cell c", c1.detach(), c_ref)
    return worst


if __name__ == "__main__":
    import sys
    if "--selfcheck" in sys.argv:
        for k, v in selfcheck().items():
            print("%-32s %.2e" % (k, v))
        print("selfcheck ok")
