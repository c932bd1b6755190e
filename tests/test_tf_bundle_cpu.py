"""TF checkpoint-V2 reader / writer (t2_tf_bundle.py): CRC-32C known answers, table + bundle round trips, a hand-assembled
prefix-compressed block, snappy blocks, the reference variable-name map."""
import os
import struct

import numpy as np
import pytest

import t2_tf_bundle as tb


def test_crc32c_known_answers():
    assert tb.crc32c(b"123456789") == 0xE3069283                       # the CRC-32C check value (RFC 3720 B.4 family)
    assert tb.crc32c(bytes(32)) == 0x8A9136AA                          # RFC 3720 B.4: 32 bytes of zeros
    assert tb.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43                 # 32 bytes of 0xff
    assert tb.crc32c(bytes(range(32))) == 0x46DD794E                   # 0..31 ascending
    assert tb.unmask_crc(tb.mask_crc(0x12345678)) == 0x12345678
    # leveldb's crc32c_test: mask(crc("foo")) differs from crc and double masking is not idempotent
    c = tb.crc32c(b"foo")
    assert tb.mask_crc(c) != c and tb.mask_crc(tb.mask_crc(c)) != c


def test_crc32c_lane_path_matches_serial():
    rng = np.random.default_rng(0)
    data = rng.integers(0, 256, 4096 * 64 + 1237, dtype=np.uint8).tobytes()
    table = tb._crc_table().tolist()
    reg = 0xFFFFFFFF
    for b in data:
        reg = table[(reg ^ b) & 0xFF] ^ (reg >> 8)
    assert tb.crc32c(data) == reg ^ 0xFFFFFFFF
    # incremental form: crc(a + b) == crc(b, crc(a))
    assert tb.crc32c(data[1000:], tb.crc32c(data[:1000])) == tb.crc32c(data)


def test_table_round_trip_many_blocks(tmp_path):
    items = [(b"", b"hdr")] + [(("scope/var_%05d/kernel" % i).encode(), os.urandom(1 + i % 37)) for i in range(3000)]
    p = str(tmp_path / "t.index")
    tb.write_table(p, items, block_size=2048)                          # forces many data blocks + restart points
    assert tb.read_table(p) == items
    raw = bytearray(open(p, "rb").read())
    raw[100] ^= 0x40
    open(p, "wb").write(raw)
    with pytest.raises(ValueError):
        tb.read_table(p)


def test_hand_assembled_block_and_footer(tmp_path):
    """a table assembled byte by byte from the LevelDB format description (not through write_table)"""
    def entry(shared, suffix, value):
        return bytes([shared, len(suffix), len(value)]) + suffix + value
    block = entry(0, b"", b"H") + entry(0, b"abc/kernel", b"v1") + entry(4, b"bias", b"v2")        # "abc/bias" < "abc/kernel"? no:
    # keys must ascend: abc/bias then abc/kernel
    block = entry(0, b"", b"H") + entry(0, b"abc/bias", b"v2") + entry(4, b"kernel", b"v1")
    block += struct.pack("<I", 0) + struct.pack("<I", 1)
    def with_trailer(b):
        return b + b"\x00" + struct.pack("<I", tb.mask_crc(tb.crc32c(b + b"\x00")))
    data = with_trailer(block)
    meta_off = len(data)
    meta = struct.pack("<I", 0) + struct.pack("<I", 1)
    data += with_trailer(meta)
    idx_off = len(data)
    handle = bytes([0, len(block)])
    idx = bytes([0, 1, len(handle)]) + b"b" + handle + struct.pack("<I", 0) + struct.pack("<I", 1)
    data += with_trailer(idx)
    footer = bytes([meta_off, len(meta), idx_off, len(idx)])
    footer += bytes(40 - len(footer)) + struct.pack("<Q", 0xDB4775248B80FB57)
    p = str(tmp_path / "h.index")
    open(p, "wb").write(data + footer)
    assert tb.read_table(p) == [(b"", b"H"), (b"abc/bias", b"v2"), (b"abc/kernel", b"v1")]


def test_snappy_block_decoding():
    lit = b"tacotron-"
    # literal(9) + copy(offset 9, len 9) with a 2-byte offset + literal "x"
    comp = bytes([19]) + bytes([(len(lit) - 1) << 2]) + lit + bytes([((9 - 1) << 2) | 2, 9, 0]) + bytes([0 << 2]) + b"x"
    assert tb._snappy_decompress(comp) == lit + lit + b"x"
    # overlapping copy (run-length): "ab" then copy offset 2 length 6 -> "abababab"
    comp = bytes([8, (2 - 1) << 2]) + b"ab" + bytes([((6 - 4) << 2) | 1 | (0 << 5), 2])
    assert tb._snappy_decompress(comp) == b"abababab"


def test_bundle_round_trip_and_checksums(tmp_path):
    rng = np.random.default_rng(1)
    tensors = {"Tacotron_model/inference/inputs_embedding": rng.standard_normal((66, 512)).astype(np.float32),
               "Tacotron_model/inference/decoder/decoder_LSTM/multi_rnn_cell/cell_0/decoder_LSTM_1/bias": np.zeros(4096, np.float32),
               "global_step": np.asarray(1234, dtype=np.int32),
               "scalar64": np.asarray(7, dtype=np.int64),
               "empty": np.zeros((0, 3), np.float32)}
    prefix = str(tmp_path / "taco_pretrained" / "tacotron_model.ckpt-1234")
    tb.write_bundle(prefix, tensors)
    assert os.path.isfile(prefix + ".index") and os.path.isfile(prefix + ".data-00000-of-00001")
    back = tb.read_bundle(prefix)
    assert set(back) == set(tensors)
    for k in tensors:
        assert back[k].dtype == tensors[k].dtype and back[k].shape == tensors[k].shape and np.array_equal(back[k], tensors[k])
    ent = tb.list_bundle(prefix)
    assert ent["global_step"]["dtype"] == tb.DT_INT32 and ent["global_step"]["shape"] == ()
    # BundleEntryProto bytes of a known entry: dtype=1, shape{dim{size:66} dim{size:512}}, offset, size, masked crc (fixed32)
    e = tb._encode_entry(tb.DT_FLOAT, (66, 512), 16384, 135168, 0xAABBCCDD)
    assert e == bytes([0x08, 0x01, 0x12, 0x09, 0x12, 0x02, 0x08, 0x42, 0x12, 0x03, 0x08, 0x80, 0x04,
                       0x20, 0x80, 0x80, 0x01, 0x28, 0x80, 0xA0, 0x08, 0x35, 0xDD, 0xCC, 0xBB, 0xAA])
    assert tb._encode_header() == bytes([0x08, 0x01, 0x1A, 0x02, 0x08, 0x01])
    # flipped data byte -> checksum error
    d = prefix + ".data-00000-of-00001"
    raw = bytearray(open(d, "rb").read())
    raw[5] ^= 1
    open(d, "wb").write(raw)
    with pytest.raises(ValueError):
        tb.read_bundle(prefix)
    tb.write_checkpoint_state(os.path.dirname(prefix), os.path.basename(prefix))
    assert tb.read_checkpoint_state(os.path.dirname(prefix)) == "tacotron_model.ckpt-1234"


def test_reference_variable_names():
    t = tb.tacotron_tf_name
    P = "Tacotron_model/inference/"
    assert t("inputs_embedding") == P + "inputs_embedding"
    assert t("encoder_convolutions/conv_layer_2/kernel") == P + "encoder_convolutions/conv_layer_2_encoder_convolutions/conv1d/kernel"
    assert t("postnet_convolutions/conv_layer_5/moving_variance") == (
        P + "postnet_convolutions/conv_layer_5_postnet_convolutions/batch_normalization/moving_variance")
    assert t("encoder_LSTM/bw/bias") == P + "encoder_LSTM/bidirectional_rnn/bw/encoder_bw_LSTM/bias"
    assert t("attention/memory_layer/kernel") == P + "memory_layer/kernel"
    assert t("attention/attention_bias") == P + "decoder/Location_Sensitive_Attention/attention_bias"
    assert t("attention/location_features_convolution/kernel") == P + "decoder/Location_Sensitive_Attention/location_features_convolution/kernel"
    assert t("decoder_prenet/dense_2/bias") == P + "decoder/decoder_prenet/dense_2/bias"
    assert t("decoder_LSTM/cell_2/kernel") == P + "decoder/decoder_LSTM/multi_rnn_cell/cell_1/decoder_LSTM_2/kernel"
    assert t("stop_token_projection/kernel") == P + "decoder/stop_token_projection/projection_stop_token_projection/kernel"
    assert t("postnet_projection/bias") == P + "postnet_projection/projection_postnet_projection/bias"
    # CBHG post-processing net (predict_linear)
    cb = {"CBHG_postnet/conv_bank/conv1d_3/kernel": "CBHG_postnet/conv_bank/conv1d_3/conv1d/kernel",
          "CBHG_postnet/conv_bank/conv1d_3/moving_mean": "CBHG_postnet/conv_bank/conv1d_3/batch_normalization/moving_mean",
          "CBHG_postnet/proj2/gamma": "CBHG_postnet/proj2/batch_normalization/gamma",
          "CBHG_postnet/dense/kernel": "CBHG_postnet/dense/kernel",
          "CBHG_postnet/highwaynet_4/T/bias": "CBHG_postnet/CBHG_postnet_highwaynet_4/T/bias",
          "CBHG_postnet/backward_RNN/candidate/kernel": "CBHG_postnet/bidirectional_rnn/bw/CBHG_postnet_backward_RNN/candidate/kernel",
          "CBHG_postnet/forward_RNN/gates/bias": "CBHG_postnet/bidirectional_rnn/fw/CBHG_postnet_forward_RNN/gates/bias",
          "cbhg_linear_specs_projection/kernel": "cbhg_linear_specs_projection/projection_cbhg_linear_specs_projection/kernel"}
    for ours, theirs in cb.items():
        assert t(ours) == P + theirs and tb.engine_name(P + theirs) == ours, ours
    w = tb.wavenet_tf_name
    Q = "WaveNet_model/inference/"
    assert w("input_convolution/kernel") == Q + "input_convolution/kernel"
    assert w("ResidualConv1DGLU_7/residual_block_cin_conv/bias") == (
        Q + "ResidualConv1DGLU_7/residual_block_cin_conv_ResidualConv1DGLU_7/bias")
    assert w("local_conditioning_upsampling_2/kernel", "2D") == Q + "ConvTranspose2D_layer_1/kernel"
    assert w("local_conditioning_upsampling_1/bias") == Q + "SubPixelConvolution_layer_0/bias"
    assert w("final_convolution_2/kernel") == Q + "final_convolution_2/kernel"


class _FakeEngine(object):
    """host-only stand-in with the engine attributes export_tf / import_tf use (the real engines need a GPU)"""

    def __init__(self, names_shapes, seed):
        import torch
        self.tensors, off = [], 0
        for n, s in names_shapes:
            self.tensors.append((n, off, s, True))
            off += int(np.prod(s))
        self.n_params, self.device, self.global_step = off, torch.device("cpu"), 0
        g = torch.Generator().manual_seed(seed)
        self.params = torch.randn(off, generator=g)
        self.m, self.v = torch.randn(off, generator=g), torch.rand(off, generator=g)

    def unflatten(self, buf):
        return {n: buf[o:o + int(np.prod(s))].reshape(s).clone() for n, o, s, _ in self.tensors}

    def export_params(self):
        return self.unflatten(self.params)

    def load_params(self, params):
        for n, o, s, _ in self.tensors:
            self.params[o:o + int(np.prod(s))] = params[n].reshape(-1)


def test_export_import_engine_round_trip(tmp_path):
    import torch
    shapes = [("inputs_embedding", (66, 8)), ("encoder_convolutions/conv_layer_1/kernel", (5, 8, 8)),
              ("decoder_LSTM/cell_1/kernel", (24, 32)), ("attention/attention_bias", (8,)),
              ("stop_token_projection/kernel", (12, 1))]
    a, b = _FakeEngine(shapes, 1), _FakeEngine(shapes, 2)
    a.global_step = 4321
    prefix = str(tmp_path / "tacotron_model.ckpt-4321")
    names = tb.export_tf(prefix, "Tacotron", a)
    assert "Tacotron_model/inference/decoder/decoder_LSTM/multi_rnn_cell/cell_0/decoder_LSTM_1/kernel/Adam_1" in names
    loaded, missing = tb.import_tf(prefix, "Tacotron", b)
    assert not missing and len(loaded) == len(shapes)
    assert torch.equal(a.params, b.params) and torch.equal(a.m, b.m) and torch.equal(a.v, b.v) and b.global_step == 4321
    # a checkpoint saved under a different outer scope still resolves through the unique-suffix rule
    t = {k.replace("Tacotron_model/", "model/Tacotron_model/"): v for k, v in tb.read_bundle(prefix).items()}
    tb.write_bundle(str(tmp_path / "other.ckpt-1"), t)
    c = _FakeEngine(shapes, 3)
    tb.import_tf(str(tmp_path / "other.ckpt-1"), "Tacotron", c)
    assert torch.equal(a.params, c.params)
    # missing variables are reported
    t.pop("model/Tacotron_model/inference/inputs_embedding")
    tb.write_bundle(str(tmp_path / "part.ckpt-1"), t)
    with pytest.raises(KeyError):
        tb.import_tf(str(tmp_path / "part.ckpt-1"), "Tacotron", _FakeEngine(shapes, 4))


def test_t2_checkpoint_tf_format(tmp_path, monkeypatch):
    """t2_checkpoint.save(fmt='tf') -> latest() -> load(): the run loops' checkpoint path in the reference's own file format"""
    import torch
    import t2_checkpoint
    shapes = [("input_convolution/kernel", (1, 16, 8)), ("ResidualConv1DGLU_0/residual_block_causal_conv/kernel", (3, 8, 16)),
              ("ResidualConv1DGLU_0/residual_block_out_conv/bias", (8,)), ("local_conditioning_upsampling_1/kernel", (3, 3, 1, 4))]
    a = _FakeEngine(shapes, 5)
    a.ema = a.params * 0.5
    a.hp = type("HP", (), {"upsample_type": "2D"})()
    for step in (10, 20, 30):
        a.global_step = step
        path = t2_checkpoint.save(str(tmp_path), "wavenet_model.ckpt", a, keep=2, fmt="tf")
    assert path.endswith("wavenet_model.ckpt-30") and not os.path.exists(str(tmp_path / "wavenet_model.ckpt-10.index"))
    assert t2_checkpoint.latest(str(tmp_path)) == path
    assert "WaveNet_model/inference/ConvTranspose2D_layer_0/kernel/ExponentialMovingAverage" in tb.list_bundle(path)
    variables, state = t2_checkpoint.load(path)
    assert state["global_step"] == 30 and set(variables) == {s[0] for s in shapes}
    for n, o, s, _ in a.tensors:
        k = int(np.prod(s))
        assert torch.equal(variables[n].reshape(-1), a.params[o:o + k])
        assert torch.equal(state["ema"][n].reshape(-1), a.ema[o:o + k])
        assert torch.equal(state["adam_v"][n].reshape(-1), a.v[o:o + k])
    # the native format still works next to it and `latest` follows whichever was written last
    monkeypatch.setenv("T2_CHECKPOINT_FORMAT", "npz")
    a.global_step = 40
    p2 = t2_checkpoint.save(str(tmp_path), "wavenet_model.ckpt", a)
    assert p2.endswith(".npz") and t2_checkpoint.latest(str(tmp_path)) == p2


def test_every_model_variable_has_an_invertible_reference_name():
    """all 162 Tacotron (with the CBHG head) and all WaveNet variables map to a reference name and back"""
    from hparams import hparams
    from oracle import tacotron as ot, wavenet as ow
    names = list(ot.param_shapes(hparams))
    assert len(names) == 162 and len({tb.tacotron_tf_name(n) for n in names}) == 162
    for n in names:
        assert tb.engine_name(tb.tacotron_tf_name(n)) == n, n
    hp = hparams.copy()
    hp.parse("out_channels=30,upsample_type=2D")
    for n in ow.param_shapes(hp):
        assert tb.engine_name(tb.wavenet_tf_name(n, "2D")) == n, n


def test_table_round_trip_property(tmp_path):
    """random sorted key sets (shared prefixes, empty values, keys longer than a block) at random block sizes survive write -> read"""
    from hypothesis import given, settings, strategies as st

    keys = st.lists(st.binary(min_size=0, max_size=40), min_size=1, max_size=120, unique=True).map(sorted)

    @settings(max_examples=40, deadline=None)
    @given(keys, st.integers(min_value=16, max_value=4096), st.randoms(use_true_random=False))
    def run(ks, block, rnd):
        items = [(k, bytes(rnd.getrandbits(8) for _ in range(rnd.randint(0, 50)))) for k in ks]
        p = str(tmp_path / "prop.index")
        tb.write_table(p, items, block_size=block)
        assert tb.read_table(p) == items
    run()


def test_crc32c_and_mask_against_tensorboards_implementation():
    """an implementation this repo did not write: tensorboard's TensorFlow stub carries the CRC-32C + masking TFRecord / table files use"""
    pw = pytest.importorskip("tensorboard.compat.tensorflow_stub.pywrap_tensorflow")
    import t2_tf_bundle as tb
    rng = np.random.default_rng(0)
    for n in (0, 1, 2, 3, 7, 8, 9, 63, 64, 65, 1000, 4097, 100003):
        data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert tb.crc32c(data) == pw.crc32c(data)
        assert tb.mask_crc(tb.crc32c(data)) == pw.masked_crc32c(data)
        assert tb.unmask_crc(pw.masked_crc32c(data)) == pw.crc32c(data)


def _bundle_message_classes():
    """BundleHeaderProto / BundleEntryProto (tensorflow/core/protobuf/tensor_bundle.proto) built with protobuf's descriptor API on top of
    tensorboard's GENERATED TensorShapeProto / DataType / VersionDef classes (real TensorFlow schemas shipped in the image)"""
    pytest.importorskip("tensorboard.compat.proto.tensor_shape_pb2")
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    from tensorboard.compat.proto import tensor_shape_pb2, types_pb2, versions_pb2
    pkg = tensor_shape_pb2.DESCRIPTOR.package                             # "tensorboard" in this build
    fd = descriptor_pb2.FileDescriptorProto(name="t2_test_tensor_bundle.proto", package="t2test", syntax="proto3")
    fd.dependency.extend([tensor_shape_pb2.DESCRIPTOR.name, types_pb2.DESCRIPTOR.name, versions_pb2.DESCRIPTOR.name])
    F = descriptor_pb2.FieldDescriptorProto
    h = fd.message_type.add(name="BundleHeaderProto")
    en = h.enum_type.add(name="Endianness")
    en.value.add(name="LITTLE", number=0)
    en.value.add(name="BIG", number=1)
    h.field.add(name="num_shards", number=1, type=F.TYPE_INT32, label=F.LABEL_OPTIONAL)
    h.field.add(name="endianness", number=2, type=F.TYPE_ENUM, type_name=".t2test.BundleHeaderProto.Endianness", label=F.LABEL_OPTIONAL)
    h.field.add(name="version", number=3, type=F.TYPE_MESSAGE, type_name=".%s.VersionDef" % pkg, label=F.LABEL_OPTIONAL)
    e = fd.message_type.add(name="BundleEntryProto")
    e.field.add(name="dtype", number=1, type=F.TYPE_ENUM, type_name=".%s.DataType" % pkg, label=F.LABEL_OPTIONAL)
    e.field.add(name="shape", number=2, type=F.TYPE_MESSAGE, type_name=".%s.TensorShapeProto" % pkg, label=F.LABEL_OPTIONAL)
    e.field.add(name="shard_id", number=3, type=F.TYPE_INT32, label=F.LABEL_OPTIONAL)
    e.field.add(name="offset", number=4, type=F.TYPE_INT64, label=F.LABEL_OPTIONAL)
    e.field.add(name="size", number=5, type=F.TYPE_INT64, label=F.LABEL_OPTIONAL)
    e.field.add(name="crc32c", number=6, type=F.TYPE_FIXED32, label=F.LABEL_OPTIONAL)
    pool = descriptor_pool.Default()
    try:
        filed = pool.AddSerializedFile(fd.SerializeToString())
    except Exception:                                                    # already registered by an earlier test in this process
        filed = pool.FindFileByName(fd.name)
    get = getattr(message_factory, "GetMessageClass", None)
    mk = (lambda d: get(d)) if get else (lambda d: message_factory.MessageFactory(pool).GetPrototype(d))
    return mk(filed.message_types_by_name["BundleHeaderProto"]), mk(filed.message_types_by_name["BundleEntryProto"]), types_pb2


def test_hand_encoded_bundle_protos_against_the_protobuf_library():
    """the hand-written protobuf encoder / decoder of t2_tf_bundle.py against Google's protobuf runtime: our bytes parse into the expected
    messages, and messages serialised by the runtime decode to the same entries here (both directions, several dtypes / shapes / offsets)"""
    Header, Entry, types_pb2 = _bundle_message_classes()
    import t2_tf_bundle as tb
    hdr = Header()
    hdr.ParseFromString(tb._encode_header(1))
    assert hdr.num_shards == 1 and hdr.endianness == 0 and hdr.version.producer == 1
    assert hdr.SerializeToString() == tb._encode_header(1)
    assert tb._decode_header(Header(num_shards=3, endianness=1).SerializeToString()) == {"num_shards": 3, "endianness": 1}
    assert (tb.DT_FLOAT, tb.DT_DOUBLE, tb.DT_INT32, tb.DT_UINT8, tb.DT_INT16, tb.DT_INT8, tb.DT_INT64, tb.DT_BOOL) == tuple(
        getattr(types_pb2, n) for n in ("DT_FLOAT", "DT_DOUBLE", "DT_INT32", "DT_UINT8", "DT_INT16", "DT_INT8", "DT_INT64", "DT_BOOL"))
    cases = [(tb.DT_FLOAT, (5, 512, 512), 0, 5 * 512 * 512 * 4, 0x12345678), (tb.DT_INT64, (), 1 << 33, 8, 0xFFFFFFFF),
             (tb.DT_FLOAT, (0, 80), 12, 0, 1), (tb.DT_INT32, (1,), 300, 4, 0x80000000), (tb.DT_FLOAT, (31, 1, 32), 77, 3968, 7)]
    for dtype, shape, offset, size, crc in cases:
        raw = tb._encode_entry(dtype, shape, offset, size, crc)
        m = Entry()
        m.ParseFromString(raw)
        assert (m.dtype, tuple(d.size for d in m.shape.dim), m.shard_id, m.offset, m.size, m.crc32c) == (dtype, shape, 0, offset, size, crc)
        assert m.SerializeToString() == raw                              # byte-identical to the runtime's canonical serialisation
        back = tb._decode_entry(m.SerializeToString())
        assert (back["dtype"], back["shape"], back["offset"], back["size"], back["crc32c"]) == (dtype, shape, offset, size, crc)
    m = Entry(dtype=tb.DT_FLOAT, shard_id=2, offset=5, size=6, crc32c=9)
    m.shape.dim.add(size=4)
    assert tb._decode_entry(m.SerializeToString())["shard_id"] == 2
