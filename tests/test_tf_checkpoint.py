"""TensorFlow checkpoint files without TensorFlow (syntex_b200/po/tf_bundle.py, po/checkpoint.py; SURVEY.md section 8
row f-4): the reference trains with tensorpack's ModelSaver and restores with SaverRestore
(po/progressive_model.py:354-356), i.e. tf.train.Saver V2 "tensor bundles". No TensorFlow exists in this environment,
so the reader is pinned against byte strings built here BY HAND from the published formats (LevelDB table format,
Snappy format description, tensor_bundle.proto) and against published CRC-32C test vectors, independently of the writer;
the writer against the reader."""
import os
import struct

import numpy as np
import pytest
import torch

import syntex_oracle as so
from syntex_b200.po import checkpoint, tf_bundle as tb


def test_crc32c_known_answers():
    """RFC 3720 (iSCSI) appendix B.4 vectors + the classic check value."""
    vectors = [(b"123456789", 0xE3069283), (bytes(32), 0x8A9136AA), (bytes([0xFF] * 32), 0x62A8AB43),
               (bytes(range(32)), 0x46DD794E), (bytes(range(31, -1, -1)), 0x113FDB5C)]
    for data, want in vectors:
        assert tb.crc32c_python(data) == want
    big = np.random.RandomState(0).bytes(70001)                      # above the threshold of the native helper
    assert tb.crc32c(big) == tb.crc32c_python(big)
    assert tb.crc32c(big[5000:], tb.crc32c(big[:5000])) == tb.crc32c_python(big)     # continuation
    for c in (0, 1, 0x12345678, 0xFFFFFFFF):
        assert tb.unmask_crc(tb.mask_crc(c)) == c
    assert tb.mask_crc(0) == 0xa282ead8


def test_varints():
    for v, enc in ((0, b"\x00"), (1, b"\x01"), (127, b"\x7f"), (128, b"\x80\x01"), (300, b"\xac\x02"),
                   (1 << 32, b"\x80\x80\x80\x80\x10")):
        assert tb.put_varint(v) == enc
        assert tb.get_varint(enc + b"\x55", 0) == (v, len(enc))


def _block(entries_bytes, restarts):
    return entries_bytes + b"".join(struct.pack("<I", r) for r in restarts) + struct.pack("<I", len(restarts))


def _with_trailer(block, ctype=0):
    return block + bytes([ctype]) + struct.pack("<I", tb.mask_crc(tb.crc32c_python(block + bytes([ctype]))))


def test_reader_parses_a_hand_built_table(tmp_path):
    """Two data blocks (prefix-compressed keys, a restart in the middle), an empty metaindex block, an index block and
    the footer, laid out byte by byte as table_format.md describes."""
    # block 0: "", "apple" -> "1", "apply" -> "22" (shares "appl"), restart, "banana" -> "333"
    b0 = (b"\x00\x00\x02" + b"hd" +
          b"\x00\x05\x01" + b"apple" + b"1" +
          b"\x04\x01\x02" + b"y" + b"22")
    r1 = len(b0)
    b0 += b"\x00\x06\x03" + b"banana" + b"333"
    block0 = _block(b0, [0, r1])
    # block 1: "cherry" -> 300 bytes
    val = bytes(range(256)) + bytes(44)
    block1 = _block(b"\x00\x06" + tb.put_varint(len(val)) + b"cherry" + val, [0])
    meta = _block(b"", [0])
    out = b""
    handles = []
    for blk in (block0, block1, meta):
        handles.append(tb.put_varint(len(out)) + tb.put_varint(len(blk)))
        out += _with_trailer(blk)
    # index: separator keys >= the last key of each block (any such key is legal: use a shortened one for block 0)
    idx_entries = (b"\x00\x01" + tb.put_varint(len(handles[0])) + b"c" + handles[0] +
                   b"\x00\x06" + tb.put_varint(len(handles[1])) + b"cherry" + handles[1])
    r2 = len(b"\x00\x01" + tb.put_varint(len(handles[0])) + b"c" + handles[0])
    index = _block(idx_entries, [0, r2])
    ih = tb.put_varint(len(out)) + tb.put_varint(len(index))
    out += _with_trailer(index)
    footer = handles[2] + ih
    footer += bytes(40 - len(footer)) + struct.pack("<Q", 0xdb4775248b80fb57)
    path = str(tmp_path / "hand.index")
    with open(path, "wb") as f:
        f.write(out + footer)
    assert tb.read_table(path) == [(b"", b"hd"), (b"apple", b"1"), (b"apply", b"22"), (b"banana", b"333"),
                                   (b"cherry", val)]
    bad = bytearray(out + footer)
    bad[7] ^= 1
    with open(path, "wb") as f:
        f.write(bytes(bad))
    with pytest.raises(ValueError):
        tb.read_table(path)


def test_snappy_decoder_on_hand_encoded_streams():
    # literal "abc", copy (1-byte offset form) offset 3 length 9 - overlapping its own output -, literal "X"
    s = tb.put_varint(13) + bytes([2 << 2]) + b"abc" + bytes([((9 - 4) << 2) | 1, 3]) + bytes([0]) + b"X"
    assert tb.snappy_decompress(s) == b"abcabcabcabcX"
    # a 100-byte literal (length in one extra byte), then a copy with a 2-byte offset: offset 100, length 64
    lit = bytes(range(100))
    s = tb.put_varint(164) + bytes([60 << 2, 99]) + lit + bytes([((64 - 1) << 2) | 2, 100, 0])
    assert tb.snappy_decompress(s) == lit + lit[:64]
    with pytest.raises(ValueError):
        tb.snappy_decompress(tb.put_varint(5) + bytes([((4 - 4) << 2) | 1, 9]))      # copy from before the start


def test_snappy_compressed_index_block_is_read(tmp_path):
    """TensorFlow compresses table blocks with Snappy when it was built with it: block type 1."""
    entries = b"\x00\x01\x01" + b"k" + b"v"
    blk = _block(entries, [0])
    comp = tb.put_varint(len(blk)) + bytes([(len(blk) - 1) << 2]) + blk          # one literal = a valid Snappy stream
    meta = _block(b"", [0])
    out = _with_trailer(comp, 1)
    h0 = tb.put_varint(0) + tb.put_varint(len(comp))
    hm = tb.put_varint(len(out)) + tb.put_varint(len(meta))
    out += _with_trailer(meta)
    index = _block(b"\x00\x01" + tb.put_varint(len(h0)) + b"k" + h0, [0])
    hi = tb.put_varint(len(out)) + tb.put_varint(len(index))
    out += _with_trailer(index)
    footer = hm + hi
    footer += bytes(40 - len(footer)) + struct.pack("<Q", tb.TABLE_MAGIC)
    path = str(tmp_path / "snappy.index")
    with open(path, "wb") as f:
        f.write(out + footer)
    assert tb.read_table(path) == [(b"k", b"v")]


def test_entry_proto_bytes():
    """BundleEntryProto for a float32 [3] tensor at offset 8, 12 bytes: field by field."""
    crc = 0x3f5009e9
    want = (b"\x08\x01"                      # 1: dtype = DT_FLOAT
            b"\x12\x04\x12\x02\x08\x03"      # 2: shape { 2: dim { 1: size = 3 } }
            b"\x20\x08"                      # 4: offset = 8
            b"\x28\x0c"                      # 5: size = 12
            b"\x35" + struct.pack("<I", crc))  # 6: crc32c, fixed32
    assert tb.encode_entry(1, (3,), 8, 12, crc) == want
    e = tb.decode_entry(want)
    assert (e["dtype"], e["shape"], e["offset"], e["size"], e["crc32c"], e["shard_id"]) == (1, [3], 8, 12, crc, 0)
    assert tb.decode_entry(tb.encode_entry(9, (), 0, 8, 1))["shape"] == []           # scalar: empty shape message


def test_table_and_bundle_round_trip(tmp_path):
    rng = np.random.RandomState(1)
    items = sorted((("key/%05d/%s" % (i, "x" * (i % 50))).encode(), rng.bytes(i % 97)) for i in range(2000))
    path = str(tmp_path / "t.index")
    tb.write_table(path, items, block_size=1024)
    assert tb.read_table(path) == items
    with pytest.raises(ValueError):
        tb.write_table(path, [(b"b", b""), (b"a", b"")])
    tensors = {"global_step": np.array(77, dtype=np.int64), "w/W": rng.rand(3, 3, 5, 7).astype(np.float32),
               "idx": np.arange(12, dtype=np.int32).reshape(3, 4), "flag": np.array([True, False]),
               "big": rng.rand(200, 300).astype(np.float32), "d": rng.rand(4), "empty": np.zeros((0, 3), np.float32)}
    prefix = str(tmp_path / "model-77")
    tb.write_bundle(prefix, tensors)
    assert os.path.getsize(prefix + tb.DATA_SUFFIX) == sum(np.asarray(v).nbytes for v in tensors.values())
    got = tb.read_bundle(prefix, verify=True)
    assert set(got) == set(tensors)
    for k, v in tensors.items():
        assert got[k].dtype == v.dtype and got[k].shape == v.shape and np.array_equal(got[k], v), k
    assert tb.list_variables(prefix)["w/W"] == (np.dtype(np.float32), (3, 3, 5, 7))
    assert set(tb.read_bundle(prefix, names={"idx"})) == {"idx"}
    with open(prefix + tb.DATA_SUFFIX, "r+b") as f:           # a flipped data byte is caught by the tensor's checksum
        f.seek(100)
        b = f.read(1)
        f.seek(100)
        f.write(bytes([b[0] ^ 0x10]))
    with pytest.raises(ValueError):
        tb.read_bundle(prefix, verify=True)


def _model(raw_weights, seed, **extra):
    from syntex_b200.po.model import SynTexModelDesc
    from syntex_b200.po.progressive_model import ProgressiveSynTex
    tex = so.OracleTexture(so.load_vgg_weights(raw_weights), list(SynTexModelDesc.DEFAULT_COEFS.keys()), torch.float32)
    torch.manual_seed(seed)
    return ProgressiveSynTex(dict({"vgg19_npy_path": raw_weights, "image_size": 32}, **extra), texture_fn=tex)


@pytest.mark.parametrize("extra", [{}, {"pre_act": True}])
def test_generator_as_a_tensorflow_checkpoint(tmp_path, raw_weights, extra):
    """Variable names and layouts of the reference graph (po/progressive_model.py:145-206): scopes
    syn/stage<s>/<layer>/..., Conv2D kernels W in HWIO, InstanceNorm gamma / beta, INReLU's norm under the convolution."""
    m1 = _model(raw_weights, 0, **extra)
    folder = str(tmp_path / "train_log")
    prefix = checkpoint.export_tf_checkpoint(folder, m1, 4000)
    assert os.path.basename(prefix) == "model-4000" and checkpoint.latest_checkpoint(folder) == prefix
    assert 'model_checkpoint_path: "model-4000"' in open(os.path.join(folder, "checkpoint")).read()
    names = tb.list_variables(prefix)
    sd = m1.state_dict()
    assert len(names) == len(sd) + 1 and names["global_step"] == (np.dtype(np.int64), ())
    for n in ("syn/stage0/conv1_1/grad_conv1/W", "syn/stage0/conv1_1/grad_conv1/inorm/gamma",
              "syn/stage0/conv1_1/grad_conv2/W", "syn/stage0/conv1_1/res0/conv0/W",
              "syn/stage0/conv1_1/res1/conv0/inorm/beta", "syn/stage0/conv1_1/res1/conv1/W", "syn/stage0/convlast/W",
              "syn/stage0/convlast/b", "syn/stage1/pool1/grad_conv1/W", "syn/stage1/conv1_1/up/W",
              "syn/stage4/pool4/res0/inorm/gamma"):
        assert n in names, n
    assert ("syn/stage0/conv1_1/post_inorm/gamma" in names) == (not extra)
    assert ("syn/stage1/conv1_1/inorm/gamma" in names) == bool(extra)          # INReLU "pre_inrelu" of the pre-act variant
    assert names["syn/stage1/conv1_1/up/W"][1] == (5, 5, 64, 64) and names["syn/stage0/convlast/W"][1] == (3, 3, 64, 3)
    data = tb.read_bundle(prefix, names={"syn/stage0/convlast/W", "syn/stage1/pool1/grad_conv1/W"})
    w = sd["syn.0.convlast.conv.weight"].numpy()                                 # [out, in, kh, kw]
    assert np.array_equal(data["syn/stage0/convlast/W"][1, 2, 5, 0], w[0, 5, 1, 2])
    assert np.array_equal(data["syn/stage0/convlast/W"], w.transpose(2, 3, 1, 0))
    m2 = _model(raw_weights, 1, **extra)
    assert not torch.equal(m2.state_dict()["syn.0.convlast.conv.weight"], sd["syn.0.convlast.conv.weight"])
    assert checkpoint.restore_checkpoint(folder, m2) == 4000                     # found by its .index companion
    for k, v in sd.items():
        assert torch.equal(m2.state_dict()[k], v), k


def test_single_generator_uses_the_same_scopes(tmp_path, raw_weights):
    """po/single_model.py:142-190 builds its one stage under the same scope names."""
    from syntex_b200.po.model import SynTexModelDesc
    from syntex_b200.po.single_model import SingleSynTex
    tex = so.OracleTexture(so.load_vgg_weights(raw_weights), list(SynTexModelDesc.DEFAULT_COEFS.keys()), torch.float32)
    torch.manual_seed(0)
    m1 = SingleSynTex({"vgg19_npy_path": raw_weights, "image_size": 32}, texture_fn=tex)
    prefix = checkpoint.export_tf_checkpoint(str(tmp_path), m1, 5)
    names = tb.list_variables(prefix)
    assert names["syn/stage0/pool3/grad_conv1/W"][1] == (3, 3, 256, 256)
    assert names["syn/stage0/pool3/up/W"][1] == (5, 5, 512, 256)                 # HWIO: 512 channels in, 256 out
    torch.manual_seed(1)
    m2 = SingleSynTex({"vgg19_npy_path": raw_weights, "image_size": 32}, texture_fn=tex)
    assert checkpoint.restore_checkpoint(prefix, m2) == 5
    assert all(torch.equal(a, b) for a, b in zip(m1.state_dict().values(), m2.state_dict().values()))


@pytest.mark.parametrize("which,extra", [("improved", {"norm_type": "instance"}),
                                         ("improved", {"norm_type": "batch", "pre_act": True}),
                                         ("improved", {"norm_type": "none"}), ("adaptive", {"norm_type": "batch"})])
def test_improved_and_adaptive_generators(tmp_path, raw_weights, which, extra):
    """po/improved_model.py:24-56, :131-330 (po/adaptive_model.py: the same scopes): the activation's norm is `inorm` or
    `bnorm` by --norm-type (BatchNorm also stores mean/EMA and variance/EMA), the up-sampling convolution is
    `up/conv/W`, the norm after the merge is called `post_inorm` whatever its type."""
    from syntex_b200.po.model import SynTexModelDesc
    from syntex_b200.po.adaptive_model import AdaptiveSynTex
    from syntex_b200.po.improved_model import ImprovedSynTex
    cls = ImprovedSynTex if which == "improved" else AdaptiveSynTex
    tex = so.OracleTexture(so.load_vgg_weights(raw_weights), list(SynTexModelDesc.DEFAULT_COEFS.keys()), torch.float32)
    torch.manual_seed(0)
    m1 = cls(dict({"vgg19_npy_path": raw_weights, "image_size": 32}, **extra), texture_fn=tex)
    with torch.no_grad():                                   # running statistics away from their initial values
        for k, v in m1.state_dict().items():
            if k.endswith("running_mean") or k.endswith("running_var"):
                v.copy_(torch.rand_like(v) + 0.5)
    prefix = checkpoint.export_tf_checkpoint(str(tmp_path), m1, 9)
    names = tb.list_variables(prefix)
    nt = extra["norm_type"]
    scope = {"instance": "inorm", "batch": "bnorm"}.get(nt)
    assert "syn/stage4/pool4/grad_conv1/W" in names or "syn/stage0/pool4/grad_conv1/W" in names
    stage = "stage4" if "syn/stage4/pool4/grad_conv1/W" in names else "stage0"
    assert "syn/%s/pool3/up/conv/W" % stage in names
    if scope:
        assert "syn/%s/pool4/grad_conv1/%s/gamma" % (stage, scope) in names
        assert "syn/%s/pool4/res0/conv0/%s/beta" % (stage, scope) in names
        assert ("syn/%s/pool4/res1/%s/gamma" % (stage, scope)) in names
        if not extra.get("pre_act"):
            assert "syn/%s/pool4/post_inorm/gamma" % stage in names
        else:
            assert "syn/%s/%s/gamma" % (stage, scope) in names and "syn/%s/pool3/%s/beta" % (stage, scope) in names
    else:
        assert not any("norm" in n for n in names)
    if any(k.endswith("running_mean") for k in m1.state_dict()):     # (the improved model normalises with batch
        assert "syn/%s/pool4/grad_conv1/bnorm/mean/EMA" % stage in names     # statistics only: no moving averages)
        assert "syn/%s/pool4/grad_conv1/bnorm/variance/EMA" % stage in names
    else:
        assert which == "improved" or nt != "batch"
    assert not any("num_batches_tracked" in n for n in names)
    torch.manual_seed(1)
    m2 = cls(dict({"vgg19_npy_path": raw_weights, "image_size": 32}, **extra), texture_fn=tex)
    assert checkpoint.restore_checkpoint(prefix, m2) == 9
    for (k, a), b in zip(m1.state_dict().items(), m2.state_dict().values()):
        if not k.endswith("num_batches_tracked"):
            assert torch.equal(a, b), k


def test_restore_ignores_optimizer_slots_and_reports_missing_variables(tmp_path, raw_weights):
    m1 = _model(raw_weights, 0)
    tensors = {checkpoint.tf_variable_name(k): checkpoint._to_tf(k, v) for k, v in m1.state_dict().items()}
    tensors["global_step"] = np.array(12, dtype=np.int64)
    tensors["syn/stage0/convlast/W/Adam"] = np.zeros((3, 3, 64, 3), np.float32)      # what a training checkpoint also holds
    tensors["syn/stage0/convlast/W/Adam_1"] = np.zeros((3, 3, 64, 3), np.float32)
    tensors["beta1_power"] = np.array(0.9, np.float32)
    prefix = str(tmp_path / "model-12")
    tb.write_bundle(prefix, tensors)
    m2 = _model(raw_weights, 1)
    assert checkpoint.restore_tf_checkpoint(prefix, m2) == 12
    assert torch.equal(m2.state_dict()["syn.2.res.pool2.1.conv1.conv.weight"],
                       m1.state_dict()["syn.2.res.pool2.1.conv1.conv.weight"])
    del tensors["syn/stage3/pool3/grad_conv2/W"]
    tb.write_bundle(prefix, tensors)
    with pytest.raises(KeyError):
        checkpoint.restore_tf_checkpoint(prefix, m2)
    with pytest.raises(KeyError):
        checkpoint.tf_variable_name("some.other.parameter")


def test_table_round_trip_property(tmp_path):
    """Random key sets (shared prefixes, empty values, binary bytes) through writer and reader, several block sizes."""
    from hypothesis import given, settings, strategies as st
    path = str(tmp_path / "p.index")

    @settings(max_examples=60, deadline=None)
    @given(st.dictionaries(st.binary(min_size=0, max_size=24), st.binary(min_size=0, max_size=200), max_size=80),
           st.sampled_from([64, 256, 4096]))
    def check(d, block_size):
        items = sorted(d.items())
        tb.write_table(path, items, block_size=block_size)
        assert tb.read_table(path) == items

    check()
