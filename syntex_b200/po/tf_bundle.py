"""TensorFlow checkpoint files ("tensor bundles", what tf.train.Saver V2 - and so tensorpack's ModelSaver, which the
reference trains with: po/progressive_model.py:354-356 - writes) read and written WITHOUT TensorFlow.

A checkpoint `prefix` is two files:
  prefix.index                  an SSTable (the LevelDB table format) mapping variable names to BundleEntryProto
                                records; the empty key holds the BundleHeaderProto
  prefix.data-00000-of-00001    the tensors' bytes, little endian, back to back at the offsets the entries give

Only what a Saver produces for dense variables is handled: one shard, no slices, numeric dtypes. Everything here is
host-side file I/O; nothing on the GPU path imports it.

Formats restated from the published sources, not copied: LevelDB's table_format.md (blocks of prefix-compressed
entries + restart array + 5-byte trailer, 48-byte footer with the magic 0xdb4775248b80fb57), Snappy's
format_description.txt (index blocks may be Snappy-compressed when TensorFlow was built with it),
tensorflow/core/protobuf/tensor_bundle.proto, tensor_shape.proto and types.proto (field numbers below), and the
masked CRC-32C of LevelDB's crc32c.h.
"""
import os
import struct

import numpy as np

TABLE_MAGIC = 0xdb4775248b80fb57
DATA_SUFFIX = ".data-00000-of-00001"

# types.proto DataType <-> numpy
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_,
           17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
_DTYPE_ENUM = {np.dtype(v): k for k, v in _DTYPES.items()}


# ---------------------------------------------------------------------------------------------- CRC-32C (Castagnoli)
def _make_crc_table():
    tab = []
    for n in range(256):
        c = n
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        tab.append(c)
    return tab


_CRC_TABLE = _make_crc_table()


def crc32c_python(data, crc=0):
    """CRC-32C of `data` (bytes-like), byte loop; `crc` continues a previous value."""
    c = crc ^ 0xFFFFFFFF
    tab = _CRC_TABLE
    for b in bytes(data):
        c = tab[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


_native_crc = None


def crc32c(data, crc=0):
    """CRC-32C of `data` (bytes-like); crc32c(b"123456789") == 0xE3069283. Large inputs go through the host helper
    stx_crc32c of the C library when it can be loaded (it needs no GPU); the byte loop is about a second per megabyte."""
    global _native_crc
    data = bytes(data)
    if len(data) >= 4096 and _native_crc is not False:
        if _native_crc is None:
            try:
                from .. import _lib
                _native_crc = _lib.lib().stx_crc32c
            except Exception:
                _native_crc = False
        if _native_crc:
            return int(_native_crc(data, len(data), crc)) & 0xFFFFFFFF
    return crc32c_python(data, crc)


def mask_crc(crc):
    """LevelDB's masked CRC (stored form): rotate right by 15 and add a constant."""
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xa282ead8) & 0xFFFFFFFF


def unmask_crc(masked):
    rot = (masked - 0xa282ead8) & 0xFFFFFFFF
    return ((rot >> 17) | (rot << 15)) & 0xFFFFFFFF


# ---------------------------------------------------------------------------------------------- varints / protobuf
def put_varint(v):
    v &= (1 << 64) - 1
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def get_varint(buf, pos):
    shift = result = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 63:
            raise ValueError("varint too long")


def _pb_fields(buf):
    """(field number, wire type, value) of every field of a serialised message; length-delimited values as bytes."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = get_varint(buf, pos)
        field, wire = key >> 3, key & 7
        if wire == 0:
            val, pos = get_varint(buf, pos)
        elif wire == 1:
            val, pos = struct.unpack_from("<Q", buf, pos)[0], pos + 8
        elif wire == 2:
            ln, pos = get_varint(buf, pos)
            val, pos = bytes(buf[pos:pos + ln]), pos + ln
        elif wire == 5:
            val, pos = struct.unpack_from("<I", buf, pos)[0], pos + 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wire)
        yield field, wire, val


def _pb_varint(field, v):
    return put_varint(field << 3) + put_varint(v)


def _pb_bytes(field, b):
    return put_varint((field << 3) | 2) + put_varint(len(b)) + b


def encode_header(num_shards=1):
    """BundleHeaderProto: 1 num_shards, 2 endianness (0 = little), 3 version (VersionDef: 1 producer)."""
    return _pb_varint(1, num_shards) + _pb_bytes(3, _pb_varint(1, 1))


def encode_entry(dtype_enum, shape, offset, size, crc):
    """BundleEntryProto: 1 dtype, 2 shape (TensorShapeProto: repeated 2 dim {1 size}), 3 shard_id, 4 offset, 5 size,
    6 crc32c (fixed32, masked)."""
    shp = b"".join(_pb_bytes(2, _pb_varint(1, int(d))) for d in shape)
    out = _pb_varint(1, dtype_enum) + _pb_bytes(2, shp)
    if offset:
        out += _pb_varint(4, offset)
    out += _pb_varint(5, size) + put_varint((6 << 3) | 5) + struct.pack("<I", crc)
    return out


def decode_entry(buf):
    e = {"dtype": 0, "shape": [], "shard_id": 0, "offset": 0, "size": 0, "crc32c": None, "slices": 0}
    for field, _, val in _pb_fields(buf):
        if field == 1:
            e["dtype"] = val
        elif field == 2:
            for f2, _, v2 in _pb_fields(val):
                if f2 == 2:
                    size = 0
                    for f3, _, v3 in _pb_fields(v2):
                        if f3 == 1:
                            size = v3 - (1 << 64) if v3 >= (1 << 63) else v3
                    e["shape"].append(size)
        elif field == 3:
            e["shard_id"] = val
        elif field == 4:
            e["offset"] = val
        elif field == 5:
            e["size"] = val
        elif field == 6:
            e["crc32c"] = val
        elif field == 7:
            e["slices"] += 1
    return e


# ---------------------------------------------------------------------------------------------- Snappy (decode only)
def snappy_decompress(buf):
    n, pos = get_varint(buf, 0)
    out = bytearray()
    end = len(buf)
    while pos < end:
        tag = buf[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:                                    # literal
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(buf[pos:pos + nb], "little")
                pos += nb
            ln += 1
            out += buf[pos:pos + ln]
            pos += ln
            continue
        if kind == 1:
            ln = ((tag >> 2) & 7) + 4
            off = ((tag >> 5) << 8) | buf[pos]
            pos += 1
        elif kind == 2:
            ln = (tag >> 2) + 1
            off = buf[pos] | (buf[pos + 1] << 8)
            pos += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 4], "little")
            pos += 4
        if off == 0 or off > len(out):
            raise ValueError("corrupt Snappy stream")
        for _ in range(ln):                              # may overlap its own output
            out.append(out[-off])
    if len(out) != n:
        raise ValueError("Snappy length mismatch")
    return bytes(out)


# ---------------------------------------------------------------------------------------------- SSTable
def _read_block(buf, offset, size, verify=True):
    body = buf[offset:offset + size]
    ctype = buf[offset + size]
    stored = struct.unpack_from("<I", buf, offset + size + 1)[0]
    if verify and unmask_crc(stored) != crc32c(bytes(body) + bytes([ctype])):
        raise ValueError("SSTable block checksum mismatch at offset %d" % offset)
    if ctype == 1:
        body = snappy_decompress(body)
    elif ctype != 0:
        raise ValueError("unknown block compression type %d" % ctype)
    return bytes(body)


def _block_entries(block):
    num_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    limit = len(block) - 4 - 4 * num_restarts
    pos, key = 0, b""
    while pos < limit:
        shared, pos = get_varint(block, pos)
        non_shared, pos = get_varint(block, pos)
        vlen, pos = get_varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        yield key, block[pos:pos + vlen]
        pos += vlen


def read_table(path, verify=True):
    """All (key, value) pairs of an SSTable file, in key order."""
    with open(path, "rb") as f:
        buf = f.read()
    if len(buf) < 48 or struct.unpack_from("<Q", buf, len(buf) - 8)[0] != TABLE_MAGIC:
        raise ValueError("%s is not an SSTable (bad magic)" % path)
    footer = buf[len(buf) - 48:]
    _, p = get_varint(footer, 0)            # metaindex handle
    _, p = get_varint(footer, p)
    ioff, p = get_varint(footer, p)
    isize, p = get_varint(footer, p)
    out = []
    for _, handle in _block_entries(_read_block(buf, ioff, isize, verify)):
        boff, q = get_varint(handle, 0)
        bsize, q = get_varint(handle, q)
        out.extend(_block_entries(_read_block(buf, boff, bsize, verify)))
    return out


def _build_block(entries, restart_interval):
    out = bytearray()
    restarts = []
    last = b""
    for i, (key, value) in enumerate(entries):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            m = min(len(last), len(key))
            while shared < m and last[shared] == key[shared]:
                shared += 1
        out += put_varint(shared) + put_varint(len(key) - shared) + put_varint(len(value)) + key[shared:] + value
        last = key
    if not restarts:
        restarts.append(0)
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def write_table(path, items, block_size=4096):
    """items: (key bytes, value bytes) in strictly increasing key order. Uncompressed blocks."""
    items = list(items)
    for a, b in zip(items, items[1:]):
        if not a[0] < b[0]:
            raise ValueError("SSTable keys must be strictly increasing")
    chunks, cur, cur_size = [], [], 0
    for kv in items:
        cur.append(kv)
        cur_size += len(kv[0]) + len(kv[1]) + 3
        if cur_size >= block_size:
            chunks.append(cur)
            cur, cur_size = [], 0
    if cur or not chunks:
        chunks.append(cur)
    out = bytearray()

    def emit(block):
        off = len(out)
        out.extend(block)
        out.append(0)                                                    # no compression
        out.extend(struct.pack("<I", mask_crc(crc32c(block + b"\x00"))))
        return put_varint(off) + put_varint(len(block))

    index = []
    for ch in chunks:
        handle = emit(_build_block(ch, 16))
        index.append((ch[-1][0] if ch else b"", handle))                   # separator: the block's last key
    meta_handle = emit(_build_block([], 1))
    index_handle = emit(_build_block(index, 1))
    footer = meta_handle + index_handle
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC)
    out.extend(footer)
    with open(path, "wb") as f:
        f.write(bytes(out))


# ---------------------------------------------------------------------------------------------- bundles
def is_bundle(prefix):
    return os.path.isfile(prefix + ".index")


def list_variables(prefix):
    """{name: (numpy dtype, shape)} of a checkpoint, without touching the data file."""
    out = {}
    for key, value in read_table(prefix + ".index"):
        if key == b"":
            continue
        e = decode_entry(value)
        out[key.decode("utf-8")] = (np.dtype(_DTYPES[e["dtype"]]), tuple(e["shape"]))
    return out


def read_bundle(prefix, names=None, verify=False):
    """{variable name: numpy array}. names: restrict to these. verify: check every tensor's CRC-32C (slow in pure
    Python: ~1 s per MB)."""
    entries = read_table(prefix + ".index")
    shards = 1
    for key, value in entries:
        if key == b"":
            for field, _, val in _pb_fields(value):
                if field == 1:
                    shards = val
                if field == 2 and val != 0:
                    raise ValueError("big-endian checkpoints are not supported")
    out = {}
    files = {}
    try:
        for key, value in entries:
            if key == b"":
                continue
            name = key.decode("utf-8")
            if names is not None and name not in names:
                continue
            e = decode_entry(value)
            if e["slices"]:
                raise ValueError("variable %s is stored in slices (partitioned variable): not supported" % name)
            if e["dtype"] not in _DTYPES:
                raise ValueError("variable %s: unsupported dtype enum %d" % (name, e["dtype"]))
            dt = np.dtype(_DTYPES[e["dtype"]])
            sid = e["shard_id"]
            if sid not in files:
                files[sid] = open("%s.data-%05d-of-%05d" % (prefix, sid, shards), "rb")
            f = files[sid]
            f.seek(e["offset"])
            raw = f.read(e["size"])
            count = int(np.prod(e["shape"])) if e["shape"] else 1
            if len(raw) != e["size"] or count * dt.itemsize != e["size"]:
                raise ValueError("variable %s: size mismatch (entry %d bytes, shape %s)" % (name, e["size"], e["shape"]))
            if verify and e["crc32c"] is not None and unmask_crc(e["crc32c"]) != crc32c(raw):
                raise ValueError("variable %s: data checksum mismatch" % name)
            out[name] = np.frombuffer(raw, dtype=dt.newbyteorder("<")).astype(dt).reshape(e["shape"])
    finally:
        for f in files.values():
            f.close()
    return out


def write_bundle(prefix, tensors):
    """tensors: {name: array-like}. Writes prefix.index and prefix.data-00000-of-00001 as a Saver (V2) would."""
    items = [(b"", encode_header(1))]
    offset = 0
    with open(prefix + DATA_SUFFIX, "wb") as f:
        for name in sorted(tensors, key=lambda s: s.encode("utf-8")):
            arr = np.asarray(tensors[name])
            arr = arr if arr.flags.c_contiguous else arr.copy()     # (ascontiguousarray would turn a scalar into [1])
            if arr.dtype not in _DTYPE_ENUM:
                raise ValueError("variable %s: unsupported dtype %s" % (name, arr.dtype))
            raw = arr.astype(arr.dtype.newbyteorder("<")).tobytes()
            f.write(raw)
            items.append((name.encode("utf-8"),
                          encode_entry(_DTYPE_ENUM[arr.dtype], arr.shape, offset, len(raw), mask_crc(crc32c(raw)))))
            offset += len(raw)
    write_table(prefix + ".index", items)


if __name__ == "__main__":       # python -m syntex_b200.po.tf_bundle <checkpoint prefix>: list the variables of a checkpoint
    import sys
    if len(sys.argv) != 2:
        raise SystemExit("usage: python -m syntex_b200.po.tf_bundle <checkpoint prefix, e.g. train_log/model-4000>")
    total = 0
    for name, (dt, shape) in sorted(list_variables(sys.argv[1]).items()):
        n = int(np.prod(shape)) if shape else 1
        total += n * dt.itemsize
        print("%-72s %-8s %s" % (name, dt.name, list(shape)))
    print("%d bytes of tensors" % total)
