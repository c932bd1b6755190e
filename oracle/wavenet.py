"""ORACLE (test infrastructure, not product code): fp32 PyTorch-CPU restatement of the reference's WaveNet vocoder.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.

Follows, function by function:
  WaveNet.__init__ / step / incremental / add_loss / add_optimizer   wavenet_vocoder/models/wavenet.py:89-208,650-721,724-911,476-519,522-628
  CausalConv1D, Conv1D1x1, ResidualConv1DGLU                           wavenet_vocoder/models/modules.py:184-333,336-389,392-521
  SubPixelConvolution, ConvTranspose2D                                 wavenet_vocoder/models/modules.py:539-654,736-770
  MaskedCrossEntropyLoss, DiscretizedMixtureLogisticLoss               wavenet_vocoder/models/modules.py:781-817
  discretized_mix_logistic_loss, sample_from_discretized_mix_logistic  wavenet_vocoder/models/mixture.py:18-107
Parameters are kept in the reference's TensorFlow variable layouts (conv kernels [kw, in, out]) under the names of
SURVEY.md Appendix B so that a parameter dict is interchangeable with the CUDA model (tacotron-2_b200/wavenet.py).

PINNING, two levels.
(1) Elementary-op code (tests/golden/make_reference_vectors.py, tests/test_reference_pinned.py): discretized_mix_logistic_loss /
sample_from_discretized_mix_logistic (mixture.py), the Gaussian loss and sampler (gaussian.py), MaskedCrossEntropyLoss /
DiscretizedMixtureLogisticLoss / GaussianMaximumLikelihoodEstimation with their masks and normalisers (modules.py:781-852,
wavenet.py:476-519), the learning-rate schedules (wavenet.py:615-633), the mu-law tensor path - the reference's source runs AS IS.
(2) The whole graph (tests/golden/make_reference_wavenet_graph_vectors.py, tests/test_reference_wavenet_graph.py): the reference's
`WaveNet.__init__` / `initialize` / `step` / `incremental` / `add_loss` are EXECUTED - wavenet.py and modules.py (CausalConv1D,
Conv1D1x1, ResidualConv1DGLU, SubPixelConvolution, ConvTranspose2D, NearestNeighborUpsample) unchanged - on a stand-in for
tf.layers.Conv1D / Conv2D / Conv2DTranspose and the keras Wrapper they are built from (tests/golden/tf_shim_graph.py): the training
graph in three configurations (mu-law CE + SubPixel, MoL + ConvTranspose2D, Gaussian + NearestNeighbor; dropout masks recorded and
injected), the evaluation branch (teacher-forced incremental pass through the convolution queues + eval loss) and the free-running
synthesis branch (every categorical / mixture / logistic draw recorded and injected), and one add_optimizer step (LR schedule,
per-tensor clip_by_norm + clip_by_value, Adam, EMA of the updated variables == adam_step here). step / upsample / loss_fn / incremental
reproduce the executed reference: outputs <= 2e-5, losses 1e-5, d loss / d variable 2e-4 relative for every variable, the sampled
waveforms sample by sample; incremental == parallel forward on the same inputs; the variable names the reference's scopes generate
equal t2_tf_bundle.wavenet_tf_name over the parameter table; the NN_init kernels its `_init_kernel` methods produce equal
_upsample_init_kernel here and the product initialiser.
Variants the CUDA path rejects are carried here too and pinned the same way (their kernels then have a checker): global (speaker)
conditioning through gc_embedding + residual_block_gin_conv, the ResizeConvolution and ConvTranspose1D upsamplers with their NN_init kernels.
STILL A RESTATEMENT: the convolution primitives under that composition (VALID / SAME cross-correlation with dilation, Conv2DTranspose
'same' arithmetic) and tf.train (Adam, EMA, glorot init) follow the public TF 1.x definitions in the stand-in
as in this file; TensorFlow itself cannot run here. Known answers on top: receptive_field_size, mulaw_quantize(0) == 127
(tests/test_oracle_wavenet.py).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

SQRT_HALF = float(np.sqrt(0.5))


def receptive_field_size(total_layers, num_cycles, kernel_size, dilation=lambda x: 2 ** x):
    assert total_layers % num_cycles == 0
    layers_per_cycle = total_layers // num_cycles
    dilations = [dilation(i % layers_per_cycle) for i in range(total_layers)]
    return (kernel_size - 1) * sum(dilations) + 1


def is_mulaw_quantize(s):
    assert s in ("mulaw-quantize", "mulaw", "raw")
    return s == "mulaw-quantize"


def is_scalar_input(s):
    return not is_mulaw_quantize(s)


def dilation_of(hp, layer):
    return 2 ** (layer % (hp.layers // hp.stacks))


# ---------------------------------------------------------------------------------------------------------
# parameters
# ---------------------------------------------------------------------------------------------------------
def param_shapes(hp):
    """Ordered {name: shape} in TF variable layout."""
    R, G, S, C = hp.residual_channels, hp.gate_channels, hp.skip_out_channels, hp.cin_channels
    cin = 1 if is_scalar_input(hp.input_type) else hp.quantize_channels
    sh = {}
    sh["input_convolution/kernel"] = (1, cin, R)
    sh["input_convolution/bias"] = (R,)
    for l in range(hp.layers):
        p = "ResidualConv1DGLU_%d/" % l
        sh[p + "residual_block_causal_conv/kernel"] = (hp.kernel_size, R, G)
        sh[p + "residual_block_causal_conv/bias"] = (G,)
        if C > 0:
            sh[p + "residual_block_cin_conv/kernel"] = (1, C, G)
            sh[p + "residual_block_cin_conv/bias"] = (G,)
        if getattr(hp, "gin_channels", -1) > 0:
            sh[p + "residual_block_gin_conv/kernel"] = (1, hp.gin_channels, G)
            sh[p + "residual_block_gin_conv/bias"] = (G,)
        sh[p + "residual_block_skip_conv/kernel"] = (1, G // 2, S)
        sh[p + "residual_block_skip_conv/bias"] = (S,)
        sh[p + "residual_block_out_conv/kernel"] = (1, G // 2, R)
        sh[p + "residual_block_out_conv/bias"] = (R,)
    sh["final_convolution_1/kernel"] = (1, S, S)
    sh["final_convolution_1/bias"] = (S,)
    sh["final_convolution_2/kernel"] = (1, S, hp.out_channels)
    sh["final_convolution_2/bias"] = (hp.out_channels,)
    if getattr(hp, "gin_channels", -1) > 0 and hp.use_speaker_embedding:      # modules.py:12-21; created outside the `inference` scope
        sh["gc_embedding"] = (hp.n_speakers, hp.gin_channels)
    if C > 0 and hp.upsample_type != "NearestNeighbor":      # NearestNeighborUpsample has no variables (modules.py:524-536)
        for i, s in enumerate(hp.upsample_scales):
            p = "local_conditioning_upsampling_%d/" % (i + 1)
            if hp.upsample_type == "SubPixel":
                sh[p + "kernel"] = (hp.freq_axis_kernel_size, 3, 1, s)
                sh[p + "bias"] = (s,)
            elif hp.upsample_type == "2D":
                sh[p + "kernel"] = (hp.freq_axis_kernel_size, s, 1, 1)
                sh[p + "bias"] = (1,)
            elif hp.upsample_type == "Resize":                           # ResizeConvolution (modules.py:657-693): NN resize, then this conv
                sh[p + "kernel"] = (hp.freq_axis_kernel_size, s, 1, 1)
                sh[p + "bias"] = (1,)
            elif hp.upsample_type == "1D":                               # ConvTranspose1D (modules.py:695-733): mixes the cin channels
                sh[p + "kernel"] = (1, s, C, C)
                sh[p + "bias"] = (C,)
            else:
                raise NotImplementedError("unknown upsample_type %s" % hp.upsample_type)
    return sh


def _glorot(shape, gen):
    if len(shape) == 3:
        fan_in, fan_out = shape[0] * shape[1], shape[0] * shape[2]
    elif len(shape) == 4:
        rf = shape[0] * shape[1]
        fan_in, fan_out = rf * shape[2], rf * shape[3]
    else:
        fan_in, fan_out = shape[0], shape[-1]
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    return (torch.rand(shape, generator=gen, dtype=torch.float32) * 2 - 1) * lim


def _upsample_init_kernel(hp, i, s):
    """NN_init kernels: SubPixel._init_kernel (modules.py:642-654), ConvTranspose2D._init_kernel (:761-770)."""
    n = len(hp.upsample_scales)
    if hp.upsample_type == "SubPixel":
        ks = (hp.freq_axis_kernel_size, 3)
        overlap = ks[1] // s
        k = np.zeros(ks, dtype=np.float32)
        ii = ks[0] // 2
        js = [ks[1] // 2 - 1, ks[1] // 2] if ks[1] % 2 == 0 else [ks[1] // 2]
        for j in js:
            k[ii, j] = 1. / max(overlap, 1.) if ks[1] % 2 == 0 else 1.
        k = np.tile(k[:, :, None, None], [1, 1, 1, s])
        return torch.from_numpy(k * hp.NN_scaler ** (1 / n))
    if hp.upsample_type == "Resize":                                      # modules.py:683-693
        ks = (hp.freq_axis_kernel_size, s)
        overlap = ks[1] // s
        k = np.zeros(ks, dtype=np.float32)
        js = [ks[1] // 2 - 1, ks[1] // 2] if ks[1] % 2 == 0 else [ks[1] // 2]
        for j in js:
            k[ks[0] // 2, j] = 1. / max(overlap, 1.) if ks[1] % 2 == 0 else 1.
        return torch.from_numpy((k * hp.NN_scaler ** (1 / n))[:, :, None, None])
    if hp.upsample_type == "1D":                                          # modules.py:723-733: identity over channels on every tap
        C = hp.cin_channels
        k = np.tile(np.eye(C, dtype=np.float32).reshape(1, 1, C, C), [1, s, 1, 1])
        k = k / max(float(s // s), 1.) if s % 2 == 0 else k
        return torch.from_numpy(k * hp.NN_scaler ** (1 / n))
    ks = (hp.freq_axis_kernel_size, s)
    overlap = ks[1] // s
    k = np.zeros(ks, dtype=np.float32)
    ii = ks[0] // 2
    for j in range(ks[1]):
        k[ii, j] = 1. / max(overlap, 1.) if ks[1] % 2 == 0 else 1.
    return torch.from_numpy((k * hp.NN_scaler ** (1 / n))[:, :, None, None])


def init_params(hp, seed=None, random_bias=False):
    """glorot-uniform kernels, zero biases (TF defaults; modules.py:195-196); NN_init for the upsampling net."""
    gen = torch.Generator().manual_seed(hp.wavenet_random_seed if seed is None else seed)
    params = {}
    for name, shape in param_shapes(hp).items():
        if name.endswith("bias"):
            params[name] = (torch.randn(shape, generator=gen) * 0.1) if random_bias else torch.zeros(shape)
        elif name.startswith("local_conditioning_upsampling") and hp.NN_init:
            i = int(name.split("_")[-1].split("/")[0]) - 1
            params[name] = _upsample_init_kernel(hp, i, hp.upsample_scales[i])
            _ = _glorot(shape, gen)  # keep the generator stream independent of NN_init
        else:
            params[name] = _glorot(shape, gen)
    return params


# ---------------------------------------------------------------------------------------------------------
# layers
# ---------------------------------------------------------------------------------------------------------
def conv1x1(x, kernel, bias):
    """Conv1D1x1 channels_first: x [B, Cin, T], TF kernel [1, Cin, Cout]."""
    return F.conv1d(x, kernel[0].t().unsqueeze(-1), bias)


def causal_conv(x, kernel, bias, dilation):
    """CausalConv1D.call parallel path (modules.py:305-325): left pad (k-1)*d, VALID cross-correlation.
    y[t] = sum_j W[j] x[t - (k-1-j) d]; TF kernel [kw, in, out]."""
    kw = kernel.shape[0]
    w = kernel.permute(2, 1, 0).contiguous()  # [out, in, kw]
    return F.conv1d(F.pad(x, ((kw - 1) * dilation, 0)), w, bias, dilation=dilation)


def upsample(c, params, hp):
    """c [B, cin, Tc] -> [B, cin, T] through the learnable upsampling net + ReLU (wavenet.py:680-702), or the non-learnable
    nearest-neighbour resize by the hop size (wavenet.py:165-167, modules.py:524-536: no activation follows it)."""
    if hp.upsample_type == "NearestNeighbor":
        return c.repeat_interleave(hp.hop_size, dim=-1)
    for i, s in enumerate(hp.upsample_scales):
        k = params["local_conditioning_upsampling_%d/kernel" % (i + 1)]
        b = params["local_conditioning_upsampling_%d/bias" % (i + 1)]
        x = c.unsqueeze(1)  # [B, 1, H=cin, W]
        if hp.upsample_type == "SubPixel":
            w = k.permute(3, 2, 0, 1).contiguous()  # [s, 1, kh, kw]
            y = F.conv2d(x, w, b, padding=(k.shape[0] // 2, k.shape[1] // 2))  # [B, s, H, W]
            B_, _, H, W = y.shape
            c = y.permute(0, 2, 3, 1).reshape(B_, H, W * s)  # periodic shuffle: out[.., w*s + k] = y[k, .., w]
        elif hp.upsample_type == "Resize":  # nearest-neighbour repeat by s along time, then a (kh, s) 'same' convolution, 1 -> 1 channel
            xr = x.repeat_interleave(s, dim=-1)
            kh, kw = k.shape[0], k.shape[1]
            xr = F.pad(xr, ((kw - 1) // 2, kw - 1 - (kw - 1) // 2, (kh - 1) // 2, kh - 1 - (kh - 1) // 2))
            c = F.conv2d(xr, k.permute(3, 2, 0, 1).contiguous(), b).squeeze(1)
        elif hp.upsample_type == "1D":      # transposed convolution over time that mixes channels, kernel s = stride s
            c = F.conv_transpose1d(c, k[0].permute(2, 1, 0).contiguous(), b, stride=s)
        else:  # '2D' ConvTranspose2D, kernel (kh, s), strides (1, s), 'same'
            w = k.permute(3, 2, 0, 1).contiguous()  # TF [kh, kw, out, in] -> torch [in, out, kh, kw]
            y = F.conv_transpose2d(x, w, b, stride=(1, s), padding=(k.shape[0] // 2, 0))
            c = y.squeeze(1)
        if hp.upsample_activation == "Relu":
            c = F.relu(c)
        elif hp.upsample_activation == "LeakyRelu":
            c = F.leaky_relu(c, hp.leaky_alpha)
    return c


def residual_block(x, c, params, hp, l, dropout_mask=None, gcond=None):
    """ResidualConv1DGLU.step, parallel mode (modules.py:471-521). Returns (x_out, skip)."""
    p = "ResidualConv1DGLU_%d/" % l
    residual = x
    if dropout_mask is not None:
        x = x * dropout_mask  # already scaled by 1/keep
    g = causal_conv(x, params[p + "residual_block_causal_conv/kernel"],
                    params[p + "residual_block_causal_conv/bias"], dilation_of(hp, l))
    a, b = g.chunk(2, dim=1)
    if c is not None:
        cc = conv1x1(c, params[p + "residual_block_cin_conv/kernel"], params[p + "residual_block_cin_conv/bias"])
        ca, cb = cc.chunk(2, dim=1)
        a, b = a + ca, b + cb
    if gcond is not None:                                                # [B, gin, 1]: the same vector at every time step (modules.py:503-508)
        gg = conv1x1(gcond, params[p + "residual_block_gin_conv/kernel"], params[p + "residual_block_gin_conv/bias"])
        ga, gb = gg.chunk(2, dim=1)
        a, b = a + ga, b + gb
    z = torch.tanh(a) * torch.sigmoid(b)
    s = conv1x1(z, params[p + "residual_block_skip_conv/kernel"], params[p + "residual_block_skip_conv/bias"])
    o = conv1x1(z, params[p + "residual_block_out_conv/kernel"], params[p + "residual_block_out_conv/bias"])
    x_out = (o + residual) * SQRT_HALF if hp.residual_legacy else (o + residual)
    return x_out, s


def step(x, c, params, hp, dropout_masks=None, c_is_upsampled=False, g=None):
    """WaveNet.step (wavenet.py:650-721): x [B, Cin, T] (one-hot float or scalar), c [B, cin, Tc] -> [B, out, T]. g: speaker ids [B, 1]
    (embedded through gc_embedding, wavenet.py:669-678) or None."""
    if g is not None:
        g = params["gc_embedding"][g.reshape(-1).long()].unsqueeze(-1)     # [B, gin, 1]
    if c is not None and not c_is_upsampled:
        c = upsample(c, params, hp)
        assert c.shape[-1] == x.shape[-1]
    h = conv1x1(x, params["input_convolution/kernel"], params["input_convolution/bias"])
    skips = None
    for l in range(hp.layers):
        h, s = residual_block(h, c, params, hp, l, None if dropout_masks is None else dropout_masks[l], g)
        if skips is None:
            skips = s
        else:
            skips = skips + s
            if hp.legacy:
                skips = skips * SQRT_HALF
    y = F.relu(skips)
    y = conv1x1(y, params["final_convolution_1/kernel"], params["final_convolution_1/bias"])
    y = F.relu(y)
    y = conv1x1(y, params["final_convolution_2/kernel"], params["final_convolution_2/bias"])
    return y


# ---------------------------------------------------------------------------------------------------------
# bf16 storage emulation of the CUDA data path (tests only)
# ---------------------------------------------------------------------------------------------------------
# The product path keeps activations, activation-gradients and GEMM operands in bf16 (fp32 accumulate). A handful
# of ReLU / gate units whose pre-activation is within bf16 rounding of zero then flip relative to the fp32 graph,
# which dominates gradient differences at random init (the gradient is a random-walk sum there). `step_sim`
# restates WaveNet.step with a bf16 round at every point where the CUDA path stores bf16 so the backward arithmetic
# can be pinned tightly; the plain fp32 `step` stays the parity reference for losses and logits.
def _r(x):
    return x.to(torch.bfloat16).to(torch.float32)


class _QF(torch.autograd.Function):  # round forward value, pass gradient through
    @staticmethod
    def forward(ctx, x):
        return _r(x)

    @staticmethod
    def backward(ctx, g):
        return g


class _QB(torch.autograd.Function):  # identity forward, round the gradient (bf16 gradient storage)
    @staticmethod
    def forward(ctx, x):
        return x.clone()

    @staticmethod
    def backward(ctx, g):
        return _r(g)


class _QFB(torch.autograd.Function):  # both
    @staticmethod
    def forward(ctx, x):
        return _r(x)

    @staticmethod
    def backward(ctx, g):
        return _r(g)


class _GateSim(torch.autograd.Function):
    """z = bf16(tanh(a) * sigmoid(b)); backward uses the bf16-stashed tanh / sigmoid and stores (da, db) in bf16."""
    @staticmethod
    def forward(ctx, a, b):
        ta, sb = torch.tanh(a), torch.sigmoid(b)
        ctx.save_for_backward(_r(ta), _r(sb))
        return _r(ta * sb)

    @staticmethod
    def backward(ctx, dz):
        ta, sb = ctx.saved_tensors
        return _r(dz * (1 - ta * ta) * sb), _r(dz * ta * sb * (1 - sb))


def step_sim(x, c, params, hp):
    """WaveNet.step with bf16 storage emulation (dropout off). Same signature / result layout as `step`."""
    qf, qb, qfb = _QF.apply, _QB.apply, _QFB.apply
    W = lambda name: qf(params[name])  # noqa: E731  (bf16 GEMM operand copy of an fp32 master)
    if c is not None:
        c = qf(upsample(c, params, hp))
    h = conv1x1(x, params["input_convolution/kernel"], params["input_convolution/bias"])  # fp32 gather / axpy
    h = qfb(h)
    zs = []
    for l in range(hp.layers):
        p = "ResidualConv1DGLU_%d/" % l
        g = causal_conv(h, W(p + "residual_block_causal_conv/kernel"), None, dilation_of(hp, l))
        g = g + (params[p + "residual_block_causal_conv/bias"] + params[p + "residual_block_cin_conv/bias"])[None, :, None]
        g = g + conv1x1(c, W(p + "residual_block_cin_conv/kernel"), None)
        a, b = g.chunk(2, dim=1)
        z = _GateSim.apply(a, b)
        zs.append(z)
        if l + 1 < hp.layers:
            o = conv1x1(z, W(p + "residual_block_out_conv/kernel"), params[p + "residual_block_out_conv/bias"])
            h = qfb((o + h) * SQRT_HALF if hp.residual_legacy else (o + h))
    L = hp.layers
    skips = 0
    for l in range(L):
        p = "ResidualConv1DGLU_%d/" % l
        e = (L - 1 if l == 0 else L - l) if hp.legacy else 0
        sc = SQRT_HALF ** e
        skips = skips + conv1x1(zs[l], qf(params[p + "residual_block_skip_conv/kernel"] * sc),
                                params[p + "residual_block_skip_conv/bias"] * sc)
    y = qf(F.relu(qb(skips)))
    y = conv1x1(y, W("final_convolution_1/kernel"), params["final_convolution_1/bias"])
    y = qf(F.relu(qb(y)))
    y = conv1x1(y, W("final_convolution_2/kernel"), params["final_convolution_2/bias"])
    return y


def train_step_sim(params, x, c, y, lengths, hp):
    ps = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    y_hat = step_sim(x, c, ps, hp)
    loss = loss_fn(y_hat, y, lengths, hp)
    used = [k for k in ps]
    gr = torch.autograd.grad(loss, [ps[k] for k in used], allow_unused=True)
    grads = {k: (g if g is not None else torch.zeros_like(ps[k])) for k, g in zip(used, gr)}
    return loss.detach(), grads, y_hat.detach()


# ---------------------------------------------------------------------------------------------------------
# losses
# ---------------------------------------------------------------------------------------------------------
def sequence_mask(lengths, max_len):
    return (torch.arange(max_len)[None, :] < lengths[:, None]).float()


def masked_cross_entropy(y_hat, y, lengths):
    """wavenet.py:488 + modules.py:781-798. y_hat [B, Q, T] logits, y [B, T] int64 -> scalar."""
    T = y_hat.shape[-1]
    logits = y_hat.transpose(1, 2)[:, :-1, :]
    targets = y[:, 1:]
    mask = sequence_mask(lengths, T)[:, 1:]
    losses = F.cross_entropy(logits.reshape(-1, logits.shape[-1]), targets.reshape(-1), reduction="none").view_as(mask)
    masked = losses * mask
    return masked.sum() / torch.count_nonzero(masked).float()


def discretized_mix_logistic_loss(y_hat, y, num_classes=256, log_scale_min=-7.0, reduce=True):
    """mixture.py:18-74. y_hat [B, 3*nm, T], y [B, T, 1] -> [B, T, 1] (reduce=False)."""
    nr_mix = y_hat.shape[1] // 3
    y_hat = y_hat.transpose(1, 2)
    logit_probs = y_hat[:, :, :nr_mix]
    means = y_hat[:, :, nr_mix:2 * nr_mix]
    log_scales = torch.clamp(y_hat[:, :, 2 * nr_mix:3 * nr_mix], min=log_scale_min)
    y = y * torch.ones(1, 1, nr_mix)
    centered_y = y - means
    inv_stdv = torch.exp(-log_scales)
    plus_in = inv_stdv * (centered_y + 1. / (num_classes - 1))
    cdf_plus = torch.sigmoid(plus_in)
    min_in = inv_stdv * (centered_y - 1. / (num_classes - 1))
    cdf_min = torch.sigmoid(min_in)
    log_cdf_plus = plus_in - F.softplus(plus_in)
    log_one_minus_cdf_min = -F.softplus(min_in)
    cdf_delta = cdf_plus - cdf_min
    mid_in = inv_stdv * centered_y
    log_pdf_mid = mid_in - log_scales - 2. * F.softplus(mid_in)
    log_probs = torch.where(y < -0.999, log_cdf_plus,
                            torch.where(y > 0.999, log_one_minus_cdf_min,
                                        torch.where(cdf_delta > 1e-5,
                                                    torch.log(torch.clamp(cdf_delta, min=1e-12)),
                                                    log_pdf_mid - np.log((num_classes - 1) / 2))))
    log_probs = log_probs + F.log_softmax(logit_probs, -1)
    if reduce:
        return -torch.sum(torch.logsumexp(log_probs, -1))
    return -torch.logsumexp(log_probs, -1).unsqueeze(-1)


def masked_mol_loss(y_hat, y, lengths, hp):
    """wavenet.py:494 + modules.py:800-817. y_hat [B, 3nm, T], y [B, T] float."""
    T = y_hat.shape[-1]
    mask = sequence_mask(lengths, T)[:, 1:].unsqueeze(-1)
    losses = discretized_mix_logistic_loss(y_hat[:, :, :-1], y[:, 1:].unsqueeze(-1),
                                           num_classes=hp.quantize_channels, log_scale_min=hp.log_scale_min,
                                           reduce=False)
    return (losses * mask).sum() / mask.sum()


def gaussian_maximum_likelihood_estimation_loss(y_hat, y, log_scale_min_gauss, num_classes, use_cdf=True, reduce=True):
    """gaussian.py:5-37. y_hat [B, 2, T] (mean, log_scale), y [B, T, 1] -> [B, T, 1] (reduce=False)."""
    y_hat = y_hat.transpose(1, 2)
    mean = y_hat[:, :, 0]
    log_scale = torch.clamp(y_hat[:, :, 1], min=log_scale_min_gauss)
    y = y.squeeze(-1)
    if use_cdf:
        scale = torch.exp(log_scale)
        cdf_plus = torch.special.ndtr((y + 1. / (num_classes - 1) - mean) / scale)
        cdf_min = torch.special.ndtr((y - 1. / (num_classes - 1) - mean) / scale)
        log_prob = torch.log(torch.clamp(cdf_plus - cdf_min, min=1e-12))
    else:
        log_prob = -0.5 * (np.log(2. * np.pi) + 2. * log_scale + (y - mean) ** 2 * torch.exp(-2. * log_scale))
    if reduce:
        return -log_prob.sum()
    return -log_prob.unsqueeze(-1)


def masked_gaussian_loss(y_hat, y, lengths, hp):
    """wavenet.py:491-492 + modules.py:819-836. y_hat [B, 2, T], y [B, T] float."""
    T = y_hat.shape[-1]
    mask = sequence_mask(lengths, T)[:, 1:].unsqueeze(-1)
    losses = gaussian_maximum_likelihood_estimation_loss(y_hat[:, :, :-1], y[:, 1:].unsqueeze(-1), hp.log_scale_min_gauss,
                                                         hp.quantize_channels, use_cdf=hp.cdf_loss, reduce=False)
    return (losses * mask).sum() / mask.sum()


def sample_from_gaussian(y, log_scale_min_gauss, normal):
    """gaussian.py:39-52 with the standard-normal draw injected. y [B, 2, T] -> [B, T]."""
    y = y.transpose(1, 2)
    mean = y[:, :, 0]
    scale = torch.exp(torch.clamp(y[:, :, 1], min=log_scale_min_gauss))
    return torch.clamp(mean + scale * normal, -1., 1.)


def loss_fn(y_hat, y, lengths, hp):
    if is_mulaw_quantize(hp.input_type):
        return masked_cross_entropy(y_hat, y, lengths)
    if hp.out_channels == 2:
        return masked_gaussian_loss(y_hat, y, lengths, hp)
    return masked_mol_loss(y_hat, y, lengths, hp)


def sample_from_discretized_mix_logistic(y, log_scale_min, u_mix, u_logistic):
    """mixture.py:76-107 with the two uniform draws injected (u in [1e-5, 1-1e-5]). y [B, 3nm, T] -> [B, T]."""
    nr_mix = y.shape[1] // 3
    y = y.transpose(1, 2)
    logit_probs = y[:, :, :nr_mix]
    temp = logit_probs - torch.log(-torch.log(u_mix))
    argmax = temp.argmax(-1)
    one_hot = F.one_hot(argmax, nr_mix).float()
    means = (y[:, :, nr_mix:2 * nr_mix] * one_hot).sum(-1)
    log_scales = torch.clamp((y[:, :, 2 * nr_mix:3 * nr_mix] * one_hot).sum(-1), min=log_scale_min)
    x = means + torch.exp(log_scales) * (torch.log(u_logistic) - torch.log(1 - u_logistic))
    return torch.clamp(x, -1., 1.)


# ---------------------------------------------------------------------------------------------------------
# incremental (Fast-WaveNet) forward — wavenet.py:724-911, modules.py:273-303
# ---------------------------------------------------------------------------------------------------------
def incremental(initial_input, c, params, hp, time_length, test_inputs=None, u_mix=None, u_logistic=None,
                u_cat=None, c_is_upsampled=False, softmax=False, normal=None):
    """initial_input [B, 1, Cin]; c [B, cin, Tc]; test_inputs [B, T, Cin] (teacher forcing) or None.
    Returns (outputs [B, T, Cin-like], raw network outputs [B, T, out])."""
    B = initial_input.shape[0]
    if c is not None:
        cu = c if c_is_upsampled else upsample(c, params, hp)
        cu = cu.transpose(1, 2)  # [B, T, cin]
    R = hp.residual_channels
    kw = hp.kernel_size
    queues = [torch.zeros(B, kw + (kw - 1) * (dilation_of(hp, l) - 1), R) for l in range(hp.layers)]
    lin_w = [params["ResidualConv1DGLU_%d/residual_block_causal_conv/kernel" % l].reshape(-1, hp.gate_channels)
             for l in range(hp.layers)]
    cur = initial_input
    outs, raws = [], []
    for t in range(time_length):
        ct = cu[:, t, :] if c is not None else None
        x = cur[:, -1, :] @ params["input_convolution/kernel"][0] + params["input_convolution/bias"]
        skips = None
        for l in range(hp.layers):
            p = "ResidualConv1DGLU_%d/" % l
            d = dilation_of(hp, l)
            residual = x
            q = torch.cat([queues[l][:, 1:, :], x.unsqueeze(1)], dim=1)
            queues[l] = q
            taps = q[:, 0::d, :] if d > 1 else q
            g = taps.reshape(B, -1) @ lin_w[l] + params[p + "residual_block_causal_conv/bias"]
            a, b = g.chunk(2, dim=-1)
            if ct is not None:
                cc = ct @ params[p + "residual_block_cin_conv/kernel"][0] + params[p + "residual_block_cin_conv/bias"]
                ca, cb = cc.chunk(2, dim=-1)
                a, b = a + ca, b + cb
            z = torch.tanh(a) * torch.sigmoid(b)
            s = z @ params[p + "residual_block_skip_conv/kernel"][0] + params[p + "residual_block_skip_conv/bias"]
            o = z @ params[p + "residual_block_out_conv/kernel"][0] + params[p + "residual_block_out_conv/bias"]
            x = (o + residual) * SQRT_HALF if hp.residual_legacy else (o + residual)
            if hp.legacy:
                skips = s if skips is None else (skips + s) * SQRT_HALF
            else:
                skips = s if skips is None else (skips + s)
        y = F.relu(skips)
        y = F.relu(y @ params["final_convolution_1/kernel"][0] + params["final_convolution_1/bias"])
        y = y @ params["final_convolution_2/kernel"][0] + params["final_convolution_2/bias"]
        raws.append(y)
        if is_scalar_input(hp.input_type) and hp.out_channels == 2:          # single Gaussian head (wavenet.py:853-856, gaussian.py:39-52)
            nz = normal[:, t:t + 1] if normal is not None else torch.randn(B, 1)
            smp = sample_from_gaussian(y.unsqueeze(-1), hp.log_scale_min_gauss, nz)   # [B, 1]
            nxt = smp.unsqueeze(-1)
            outs.append(smp)
        elif is_scalar_input(hp.input_type):
            um = u_mix[:, t:t + 1, :] if u_mix is not None else torch.rand(B, 1, hp.out_channels // 3).clamp(1e-5, 1 - 1e-5)
            ul = u_logistic[:, t:t + 1] if u_logistic is not None else torch.rand(B, 1).clamp(1e-5, 1 - 1e-5)
            smp = sample_from_discretized_mix_logistic(y.unsqueeze(-1), hp.log_scale_min, um, ul)  # [B, 1]
            nxt = smp.unsqueeze(-1)
            outs.append(smp)
        else:
            # tf.multinomial treats its argument as LOGITS (wavenet.py:865); the synthesis / eval graphs call
            # incremental(softmax=False) (wavenet.py:376,448) so the raw network output is sampled correctly.
            probs_as_logits = F.softmax(y, dim=-1) if softmax else y
            if u_cat is not None:
                cdf = torch.cumsum(F.softmax(probs_as_logits, dim=-1), dim=-1)
                idx = (cdf < u_cat[:, t:t + 1]).sum(-1).clamp(max=hp.quantize_channels - 1)
            else:
                idx = torch.multinomial(F.softmax(probs_as_logits, dim=-1), 1).squeeze(-1)
            oh = F.one_hot(idx, hp.quantize_channels).float()
            nxt = oh.unsqueeze(1)
            outs.append(oh)
        if test_inputs is not None:
            nxt = test_inputs[:, t:t + 1, :]
        cur = nxt
    return torch.stack(outs, dim=1), torch.stack(raws, dim=1)


# ---------------------------------------------------------------------------------------------------------
# optimizer — wavenet.py:522-628
# ---------------------------------------------------------------------------------------------------------
def learning_rate(hp, global_step):
    if hp.wavenet_lr_schedule == "noam":
        step = float(global_step + 1)
        w = hp.wavenet_warmup
        return max(hp.wavenet_learning_rate * w ** 0.5 * min(step * w ** -1.5, step ** -0.5), 1e-4)
    assert hp.wavenet_lr_schedule == "exponential"
    return hp.wavenet_learning_rate * hp.wavenet_decay_rate ** (global_step / hp.wavenet_decay_steps)


def clip_gradients(grads, hp):
    out = {}
    for k, g in grads.items():
        n = torch.sqrt((g * g).sum())
        g = g * hp.wavenet_gradient_max_norm / torch.clamp(n, min=hp.wavenet_gradient_max_norm)
        out[k] = torch.clamp(g, -hp.wavenet_gradient_max_value, hp.wavenet_gradient_max_value)
    return out


def adam_step(params, grads, state, hp, global_step):
    """tf.train.AdamOptimizer + EMA(0.9999) (wavenet.py:549,613). state = {'m','v','ema','t'}."""
    lr = learning_rate(hp, global_step)
    b1, b2, eps = hp.wavenet_adam_beta1, hp.wavenet_adam_beta2, hp.wavenet_adam_epsilon
    state["t"] = state.get("t", 0) + 1
    t = state["t"]
    lr_t = lr * math.sqrt(1 - b2 ** t) / (1 - b1 ** t)
    if hp.wavenet_clip_gradients:
        grads = clip_gradients(grads, hp)
    for k in params:
        m = state.setdefault("m", {}).setdefault(k, torch.zeros_like(params[k]))
        v = state.setdefault("v", {}).setdefault(k, torch.zeros_like(params[k]))
        e = state.setdefault("ema", {}).setdefault(k, params[k].clone())
        g = grads[k]
        m.mul_(b1).add_(g, alpha=1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        params[k] = params[k] - lr_t * m / (v.sqrt() + eps)
        e.sub_((1 - hp.wavenet_ema_decay) * (e - params[k]))
    return lr


def train_step(params, x, c, y, lengths, hp, c_is_upsampled=False, dropout_masks=None):
    """One teacher-forced forward + loss + autograd backward. Returns (loss, grads, y_hat).
    dropout_masks: optional per-layer [B, R, T] masks already scaled by 1/keep (modules.py:483-484 draws them with tf.layers.dropout)."""
    ps = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    y_hat = step(x, c, ps, hp, dropout_masks=dropout_masks, c_is_upsampled=c_is_upsampled)
    loss = loss_fn(y_hat, y, lengths, hp)
    used = [k for k in ps]
    gr = torch.autograd.grad(loss, [ps[k] for k in used], allow_unused=True)
    grads = {k: (g if g is not None else torch.zeros_like(ps[k])) for k, g in zip(used, gr)}
    return loss.detach(), grads, y_hat.detach()
