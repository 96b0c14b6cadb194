"""The reference's checkpoint format (model.py:450-503: Keras save_weights -> TF tensor bundle), read and written without
TensorFlow (sg-gan-tf2_b200/tf_checkpoint.py).  TF itself cannot be installed here, so what is checked is the format's own
published constants (CRC-32C test vectors, table magic, masked checksums on every block and tensor), round trips, and that
damage is detected.  Host logic only: no GPU."""
import importlib
import os
import struct

import numpy as np
import pytest


@pytest.fixture(scope="module")
def T():
    return importlib.import_module("sg-gan-tf2_b200.tf_checkpoint")


def _arrays(kinds, c=8, seed=0):
    rng = np.random.RandomState(seed)
    out = []
    for k in kinds:
        out += [rng.rand(c).astype(np.float32), rng.rand(c).astype(np.float32)] if k == "norm" else \
            [rng.rand(3, 3, c, c).astype(np.float32), rng.rand(c).astype(np.float32)]
    return out


def test_crc32c_known_answers(T, L):
    # RFC 3720 B.4 test vectors
    assert T.crc32c(b"\x00" * 32) == 0x8A9136AA
    assert T.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert T.crc32c(bytes(range(32))) == 0x46DD794E
    assert T.crc32c(b"123456789") == 0xE3069283
    big = (bytes(range(256)) * 64)[3:]                      # > 4096 bytes: libsggan's slicing-by-8 host helper
    ref = 0xFFFFFFFF
    for b in big:
        ref = T._TABLE[(ref ^ b) & 0xFF] ^ (ref >> 8)
    assert T.crc32c(big) == ref ^ 0xFFFFFFFF
    assert T.crc32c(big[7000:], T.crc32c(big[:7000])) == T.crc32c(big)   # extension
    assert T._unmask(T._mask(0x12345678)) == 0x12345678


def test_table_round_trip_many_blocks(T, tmp_path):
    items = [(b"", b"header")] + [(("key/%05d/suffix" % i).encode(), os.urandom(1 + i % 37)) for i in range(500)]
    p = str(tmp_path / "t.index")
    T.write_table(p, items, block_size=2048)                 # dozens of data blocks, prefix compression + restarts
    assert T.read_table(p) == items
    raw = open(p, "rb").read()
    assert struct.unpack("<Q", raw[-8:])[0] == 0xdb4775248b80fb57
    bad = bytearray(raw)
    bad[100] ^= 1
    open(p, "wb").write(bytes(bad))
    with pytest.raises(ValueError, match="checksum"):
        T.read_table(p)


def test_bundle_round_trip_and_names(T, tmp_path):
    kinds = ["conv", "norm"] * 3 + ["conv", "norm", "conv", "norm"] * 2 + ["deconv", "norm"] * 2 + ["conv"]
    arrs = _arrays(kinds)
    prefix = str(tmp_path / "gen" / "cp-0007.ckpt")
    T.save(prefix, arrs, kinds)
    assert sorted(os.listdir(tmp_path / "gen")) == ["checkpoint", "cp-0007.ckpt.data-00000-of-00001", "cp-0007.ckpt.index"]
    assert T.latest_checkpoint(str(tmp_path / "gen")) == prefix
    back = T.load(prefix, kinds)
    assert len(back) == len(arrs) and all(a.dtype == np.float32 and np.array_equal(a, b) for a, b in zip(back, arrs))
    ent = T.load_entries(prefix)
    # Keras object-graph names: one `layer_with_weights-i` per weighted layer, in model order
    assert "layer_with_weights-0/kernel/.ATTRIBUTES/VARIABLE_VALUE" in ent
    assert "layer_with_weights-1/gamma/.ATTRIBUTES/VARIABLE_VALUE" in ent and "layer_with_weights-1/beta/.ATTRIBUTES/VARIABLE_VALUE" in ent
    assert "layer_with_weights-%d/bias/.ATTRIBUTES/VARIABLE_VALUE" % (len(kinds) - 1) in ent
    graph = ent[T.OBJECT_GRAPH_KEY]
    assert b"layer_with_weights-3" in graph and b"VARIABLE_VALUE" in graph and b"conv2d_transpose/kernel" in graph
    # a flipped bit in the data file is caught by the tensor checksum
    data = prefix + ".data-00000-of-00001"
    good = open(data, "rb").read()
    raw = bytearray(good)
    raw[len(raw) // 2] ^= 0x10
    open(data, "wb").write(bytes(raw))
    with pytest.raises(ValueError, match="checksum"):
        T.load(prefix, kinds)
    open(data, "wb").write(good)
    # a checkpoint of another architecture is refused by name, not silently mis-assigned
    with pytest.raises(KeyError):
        T.load(prefix, kinds + ["conv"])


def test_latest_checkpoint_absent(T, tmp_path):
    assert T.latest_checkpoint(str(tmp_path)) is None
    (tmp_path / "checkpoint").write_text('model_checkpoint_path: "cp-0001.ckpt"\n')
    assert T.latest_checkpoint(str(tmp_path)) is None      # named but not there


def test_checksums_and_wire_format_against_bytes_written_by_tensorflow(T):
    """tests/golden/reference_tfevents_records.bin: TFRecord frames copied verbatim from a TensorBoard log the reference's
    trainer wrote (make_tfrecord_fixture.py).  TFRecord uses the same masked CRC-32C as the checkpoint format, so every
    frame TF wrote must verify with our crc32c / mask, and our protobuf wire reader must recover the scalar the reference
    logged (its first 'Generator Loss')."""
    data = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_tfevents_records.bin"), "rb").read()
    pos, n, scalars = 0, 0, {}
    while pos < len(data):
        (ln,) = struct.unpack("<Q", data[pos:pos + 8])
        (c_len,) = struct.unpack("<I", data[pos + 8:pos + 12])
        rec = data[pos + 12:pos + 12 + ln]
        (c_rec,) = struct.unpack("<I", data[pos + 12 + ln:pos + 16 + ln])
        assert T._mask(T.crc32c(data[pos:pos + 8])) == c_len and T._unmask(c_len) == T.crc32c(data[pos:pos + 8])
        assert T._mask(T.crc32c(rec)) == c_rec
        for f, _, v in T._pb_fields(rec):                      # Event: summary = 5 -> Summary.value = 1
            if f != 5:
                continue
            for f2, _, val in T._pb_fields(v):
                tag = num = None
                for f3, _, x in T._pb_fields(val):             # Value: tag = 1, tensor = 8 -> TensorProto
                    if f3 == 1:
                        tag = bytes(x).decode()
                    if f3 == 8:
                        for f4, w4, y in T._pb_fields(x):      # tensor_content = 4 (raw bytes) or float_val = 5
                            if f4 in (4, 5) and len(bytes(y)) >= 4:
                                num = struct.unpack("<f", bytes(y)[:4])[0]
                if tag is not None and num is not None:
                    scalars.setdefault(tag, []).append(num)
        pos += 16 + ln
        n += 1
    assert n >= 40 and pos == len(data)
    assert abs(scalars["Generator Loss"][0] - 5.78389835357666) < 1e-6
    assert len(scalars["Discriminator Loss"]) == 11 and abs(scalars["Mean IoU"][0] - 0.2222737967967987) < 1e-7
