"""tf_checkpoint.py -- the reference's checkpoint FORMAT (model.py:450-503: Keras `Model.save_weights("…/cp-0003.ckpt")`),
read and written without TensorFlow.

`save_weights` with a non-.h5 path writes a TF "tensor bundle":

    cp-0003.ckpt.index                 a leveldb-format table (tensorflow/core/lib/io/table*): key "" -> BundleHeaderProto,
                                       every tensor key -> BundleEntryProto {dtype, shape, shard_id, offset, size, crc32c}
    cp-0003.ckpt.data-00000-of-00001   the tensors' raw little-endian bytes, back to back
    checkpoint                         text proto naming the latest prefix (what tf.train.latest_checkpoint reads)

and names the variables of a functional Keras model by object-graph path:
`layer_with_weights-<i>/<kernel|bias|gamma|beta>/.ATTRIBUTES/VARIABLE_VALUE`, i counting the layers that own weights in
model order (Conv2D / Conv2DTranspose: kernel, bias; tfa InstanceNormalization: gamma, beta), plus one string tensor
`_CHECKPOINTABLE_OBJECT_GRAPH` holding the serialized TrackableObjectGraph.  The table blocks are written uncompressed
(BundleWriter sets kNoCompression); a snappy block (type 1) in a foreign file is reported, not guessed at.

TensorFlow 2.1 is not installable in this environment and the reference ships no checkpoint, so the container layout is
validated by round trips and by the format's own checksums (every block and every tensor carries a masked CRC32C, which the
reader verifies).  The checksum layer itself (crc32c, mask) and the protobuf wire reader ARE checked against bytes TF wrote:
the TFRecord frames of a TensorBoard log the reference ships use the same masked CRC-32C
(tests/golden/reference_tfevents_records.bin, tests/test_tf_checkpoint.py).  The layout constants below cite the TF sources
they restate.
"""
from __future__ import annotations

import os
import re
import struct

import numpy as np

_MAGIC = 0xdb4775248b80fb57          # table/format.h kTableMagicNumber
_DT = {1: np.dtype("<f4"), 2: np.dtype("<f8"), 3: np.dtype("<i4"), 9: np.dtype("<i8")}  # types.proto DataType
_DT_FLOAT, _DT_STRING = 1, 7
OBJECT_GRAPH_KEY = "_CHECKPOINTABLE_OBJECT_GRAPH"
_SUFFIX = "/.ATTRIBUTES/VARIABLE_VALUE"


# ---------------------------------------------------------------------------------------------- crc32c (Castagnoli)
def _make_table():
    t = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        t.append(c)
    return t


_TABLE = _make_table()
_NP_TABLE = np.array(_TABLE, dtype=np.uint32)


def crc32c(data, crc=0):
    """CRC-32C of bytes-like `data`.  Large buffers go through libsggan's host helper when the library is present; the
    pure-Python table loop is the definition."""
    data = memoryview(data).cast("B")
    if len(data) > 4096:
        try:
            from . import _lib as L
            import ctypes as C
            buf = (C.c_char * len(data)).from_buffer_copy(data)
            return int(L.lib().sggan_crc32c(buf, len(data), crc)) & 0xFFFFFFFF
        except Exception:
            pass
    c = crc ^ 0xFFFFFFFF
    for b in data.tobytes():
        c = _TABLE[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def _mask(crc):   # lib/hash/crc32c.h Mask
    return (((crc >> 15) | (crc << 17)) + 0xa282ead8) & 0xFFFFFFFF


def _unmask(m):
    rot = (m - 0xa282ead8) & 0xFFFFFFFF
    return ((rot >> 17) | (rot << 15)) & 0xFFFFFFFF


# ---------------------------------------------------------------------------------------------- varints / protobuf wire format
def _varint(n):
    out = bytearray()
    n &= (1 << 64) - 1
    while True:
        b = n & 0x7F
        n >>= 7
        out.append(b | (0x80 if n else 0))
        if not n:
            return bytes(out)


def _read_varint(buf, pos):
    shift = val = 0
    while True:
        b = buf[pos]
        pos += 1
        val |= (b & 0x7F) << shift
        if not b & 0x80:
            return val, pos
        shift += 7


def _pb_fields(buf):
    """Yield (field number, wire type, value) of one serialized message."""
    pos = 0
    while pos < len(buf):
        tag, pos = _read_varint(buf, pos)
        f, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _read_varint(buf, pos)
        elif wt == 1:
            v = buf[pos:pos + 8]
            pos += 8
        elif wt == 2:
            n, pos = _read_varint(buf, pos)
            v = buf[pos:pos + n]
            pos += n
        elif wt == 5:
            v = buf[pos:pos + 4]
            pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        yield f, wt, v


def _pb(field, wt, payload):
    tag = _varint((field << 3) | wt)
    if wt == 0:
        return tag + _varint(payload)
    if wt == 2:
        return tag + _varint(len(payload)) + payload
    if wt == 5:
        return tag + struct.pack("<I", payload)
    raise ValueError(wt)


def _entry_proto(dtype, shape, offset, size, crc):
    """tensor_bundle.proto BundleEntryProto: dtype=1, shape=2 (TensorShapeProto.dim=2 {size=1}), shard_id=3, offset=4, size=5,
    crc32c=6 (fixed32).  proto3: zero-valued scalars are omitted."""
    shp = b"".join(_pb(2, 2, _pb(1, 0, int(d)) if d else b"") for d in shape)
    out = _pb(1, 0, dtype) + _pb(2, 2, shp)
    if offset:
        out += _pb(4, 0, offset)
    out += _pb(5, 0, size) + _pb(6, 5, crc)
    return out


def _parse_entry(buf):
    e = {"dtype": 0, "shape": [], "shard_id": 0, "offset": 0, "size": 0, "crc32c": 0}
    for f, _, v in _pb_fields(buf):
        if f == 1:
            e["dtype"] = v
        elif f == 2:
            for f2, _, dim in _pb_fields(v):
                if f2 == 2:
                    size = 0
                    for f3, _, x in _pb_fields(dim):
                        if f3 == 1:
                            size = x
                    e["shape"].append(size)
        elif f == 3:
            e["shard_id"] = v
        elif f == 4:
            e["offset"] = v
        elif f == 5:
            e["size"] = v
        elif f == 6:
            e["crc32c"] = struct.unpack("<I", v)[0]
    return e


# ---------------------------------------------------------------------------------------------- leveldb-format table
def _block(entries, restart_interval=16):
    """table/block_builder.cc: prefix-compressed entries + restart array."""
    out, restarts, last, n = bytearray(), [], b"", 0
    for k, v in entries:
        shared = 0
        if n % restart_interval == 0:
            restarts.append(len(out))
        else:
            m = min(len(last), len(k))
            while shared < m and last[shared] == k[shared]:
                shared += 1
        out += _varint(shared) + _varint(len(k) - shared) + _varint(len(v)) + k[shared:] + v
        last = k
        n += 1
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def _parse_block(buf):
    nrest = struct.unpack("<I", buf[-4:])[0]
    end = len(buf) - 4 - 4 * nrest
    pos, key, out = 0, b"", []
    while pos < end:
        shared, pos = _read_varint(buf, pos)
        non_shared, pos = _read_varint(buf, pos)
        vlen, pos = _read_varint(buf, pos)
        key = key[:shared] + bytes(buf[pos:pos + non_shared])
        pos += non_shared
        out.append((key, bytes(buf[pos:pos + vlen])))
        pos += vlen
    return out


def _with_trailer(block):
    """table/format: block | 1-byte compression type (0 = none) | masked crc32c of block + type."""
    return block + b"\x00" + struct.pack("<I", _mask(crc32c(block + b"\x00")))


def _handle(offset, size):
    return _varint(offset) + _varint(size)


def _read_block(f, offset, size):
    f.seek(offset)
    raw = f.read(size + 5)
    block, ctype, crc = raw[:size], raw[size], struct.unpack("<I", raw[size + 1:size + 5])[0]
    if _unmask(crc) != crc32c(raw[:size + 1]):
        raise ValueError("table block at %d: checksum mismatch" % offset)
    if ctype != 0:
        raise ValueError("table block at %d is compressed (type %d, snappy); only uncompressed tables are supported" % (offset, ctype))
    return block


def write_table(path, items, block_size=1 << 18):
    """items: sorted list of (key bytes, value bytes)."""
    with open(path, "wb") as f:
        index, cur, cur_bytes, pos = [], [], 0, 0

        def flush():
            nonlocal cur, cur_bytes, pos
            if not cur:
                return
            blk = _block(cur)
            f.write(_with_trailer(blk))
            index.append((cur[-1][0], _handle(pos, len(blk))))  # a key >= every key of the block
            pos += len(blk) + 5
            cur, cur_bytes = [], 0

        for k, v in items:
            cur.append((k, v))
            cur_bytes += len(k) + len(v) + 12
            if cur_bytes >= block_size:
                flush()
        flush()
        meta = _block([])
        f.write(_with_trailer(meta))
        meta_h = _handle(pos, len(meta))
        pos += len(meta) + 5
        idx = _block(index, restart_interval=1)
        f.write(_with_trailer(idx))
        idx_h = _handle(pos, len(idx))
        footer = meta_h + idx_h
        footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", _MAGIC)
        f.write(footer)


def read_table(path):
    with open(path, "rb") as f:
        f.seek(0, os.SEEK_END)
        n = f.tell()
        if n < 48:
            raise ValueError("%s: too short for a table" % path)
        f.seek(n - 48)
        footer = f.read(48)
        if struct.unpack("<Q", footer[40:])[0] != _MAGIC:
            raise ValueError("%s: not a TF checkpoint index (bad magic)" % path)
        _, p = _read_varint(footer, 0)       # metaindex offset
        _, p = _read_varint(footer, p)       # metaindex size
        ioff, p = _read_varint(footer, p)
        isz, p = _read_varint(footer, p)
        out = []
        for _, h in _parse_block(_read_block(f, ioff, isz)):
            off, q = _read_varint(h, 0)
            sz, _ = _read_varint(h, q)
            out += _parse_block(_read_block(f, off, sz))
        return out


# ---------------------------------------------------------------------------------------------- the bundle
def _object_graph(layers):
    """trackable_object_graph.proto for a functional Keras model: node 0 = the model with one child per weighted layer
    (`layer_with_weights-i`), every layer node with one child per variable, every variable node with the attribute that
    points at its checkpoint key.  layers: [(layer name, [(local name, checkpoint key), …]), …]."""
    nodes = []
    root_children = []
    layer_nodes = []
    nid = 1
    for i, (lname, variables) in enumerate(layers):
        root_children.append(("layer_with_weights-%d" % i, nid))
        children = []
        lid = nid
        nid += 1
        var_nodes = []
        for local, key in variables:
            children.append((local, nid))
            var_nodes.append((local, "%s/%s" % (lname, local), key))
            nid += 1
        layer_nodes.append((lid, children, var_nodes))

    def obj(children=(), attrs=()):
        out = b""
        for local, cid in children:   # TrackableObject.children = 1 {node_id = 1, local_name = 2}
            out += _pb(1, 2, (_pb(1, 0, cid) if cid else b"") + _pb(2, 2, local.encode()))
        for name, full, key in attrs:  # attributes = 2 {name = 1, full_name = 2, checkpoint_key = 3}
            out += _pb(2, 2, _pb(1, 2, name.encode()) + _pb(2, 2, full.encode()) + _pb(3, 2, key.encode()))
        return out

    nodes.append(obj(root_children))
    for _, children, var_nodes in layer_nodes:
        nodes.append(obj(children))
        for _, full, key in var_nodes:
            nodes.append(obj(attrs=[("VARIABLE_VALUE", full, key)]))
    return b"".join(_pb(1, 2, n) for n in nodes)  # TrackableObjectGraph.nodes = 1


def variable_keys(kinds):
    """Checkpoint keys of a network given per-layer kinds, e.g. ["conv", "norm", "conv", …] -> flat list in OUR variable order
    (kernel, bias | gamma, beta) and the layer table for the object graph."""
    keys, layers = [], []
    counts = {}
    for i, kind in enumerate(kinds):
        names = ("gamma", "beta") if kind == "norm" else ("kernel", "bias")
        base = {"conv": "conv2d", "deconv": "conv2d_transpose", "norm": "instance_normalization"}[kind]
        c = counts.get(base, 0)
        counts[base] = c + 1
        lname = base if c == 0 else "%s_%d" % (base, c)
        vs = [(n, "layer_with_weights-%d/%s%s" % (i, n, _SUFFIX)) for n in names]
        keys += [k for _, k in vs]
        layers.append((lname, vs))
    return keys, layers


def save(prefix, arrays, kinds):
    """Write `arrays` (our flat variable list, Keras creation order) as the TF checkpoint `prefix` (e.g. …/cp-0003.ckpt)."""
    keys, layers = variable_keys(kinds)
    if len(keys) != len(arrays):
        raise ValueError("%d arrays for %d checkpoint keys" % (len(arrays), len(keys)))
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    graph = _object_graph(layers)
    # string tensor encoding (tensor_bundle.cc WriteStringTensor): varint64 lengths | masked crc32c of the lengths | the
    # bytes; the entry's crc32c runs over the lengths (uint32 values), that 4-byte checksum and the bytes
    lens = _varint(len(graph))
    crc = crc32c(struct.pack("<I", len(graph)))          # the lengths as uint32 values, not their varint bytes
    len_ck = struct.pack("<I", _mask(crc))
    crc = crc32c(graph, crc32c(len_ck, crc))
    sdata = lens + len_ck + graph
    entries = {OBJECT_GRAPH_KEY.encode(): (_DT_STRING, [], sdata, _mask(crc))}
    for k, a in zip(keys, arrays):
        a = np.ascontiguousarray(np.asarray(a, dtype="<f4"))
        raw = a.tobytes()
        entries[k.encode()] = (_DT_FLOAT, list(a.shape), raw, _mask(crc32c(raw)))
    items, offset = [], 0
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        for k in sorted(entries):
            dt, shape, raw, crc = entries[k]
            f.write(raw)
            items.append((k, _entry_proto(dt, shape, offset, len(raw), crc)))
            offset += len(raw)
    # BundleHeaderProto: num_shards = 1, endianness = 2 (LITTLE = 0: omitted), version = 3 {producer = 1}
    header = _pb(1, 0, 1) + _pb(3, 2, _pb(1, 0, 1))
    write_table(prefix + ".index", [(b"", header)] + items)
    state = os.path.join(os.path.dirname(os.path.abspath(prefix)), "checkpoint")
    name = os.path.basename(prefix)
    with open(state, "w") as f:   # CheckpointState text proto (training/checkpoint_state.proto)
        f.write('model_checkpoint_path: "%s"\nall_model_checkpoint_paths: "%s"\n' % (name, name))


def load_entries(prefix):
    """-> {key: numpy array} of every numeric tensor in the bundle (checksums verified) + raw bytes of string tensors."""
    items = read_table(prefix + ".index")
    out, files = {}, {}
    nshards = 1
    for k, v in items:
        if k == b"":
            for f, _, x in _pb_fields(v):
                if f == 1:
                    nshards = x
                if f == 2 and x != 0:
                    raise ValueError("big-endian bundle")
            continue
        e = _parse_entry(v)
        path = "%s.data-%05d-of-%05d" % (prefix, e["shard_id"], nshards)
        if path not in files:
            files[path] = open(path, "rb")
        fh = files[path]
        fh.seek(e["offset"])
        raw = fh.read(e["size"])
        if len(raw) != e["size"]:
            raise ValueError("%s: truncated data for %r" % (path, k))
        key = k.decode()
        if e["dtype"] == _DT_STRING:
            out[key] = raw
            continue
        if _unmask(e["crc32c"]) != crc32c(raw):
            raise ValueError("checksum mismatch for tensor %r" % key)
        if e["dtype"] not in _DT:
            raise ValueError("tensor %r: unsupported dtype %d" % (key, e["dtype"]))
        out[key] = np.frombuffer(raw, dtype=_DT[e["dtype"]]).reshape(e["shape"]).copy()
    for fh in files.values():
        fh.close()
    return out


def load(prefix, kinds):
    """The flat variable list (our order) out of a checkpoint written by Keras `save_weights` for the same architecture."""
    ent = load_entries(prefix)
    keys, _ = variable_keys(kinds)
    missing = [k for k in keys if k not in ent]
    if missing:
        have = sorted(k for k in ent if k.endswith(_SUFFIX))
        raise KeyError("checkpoint %s lacks %d of %d variables (first: %s); it holds %s…" %
                       (prefix, len(missing), len(keys), missing[0], have[:3]))
    return [ent[k] for k in keys]


def latest_checkpoint(directory):
    """tf.train.latest_checkpoint: the prefix named by the directory's `checkpoint` state file (None if absent)."""
    state = os.path.join(directory, "checkpoint")
    if not os.path.exists(state):
        return None
    m = re.search(r'^model_checkpoint_path:\s*"(.*)"', open(state).read(), re.M)
    if not m:
        return None
    p = m.group(1)
    p = p if os.path.isabs(p) else os.path.join(directory, p)
    return p if os.path.exists(p + ".index") else None
