"""A fixture written by TensorFlow itself, taken from the reference's repository (run in the build container):

    python tests/golden/make_tfrecord_fixture.py     # writes tests/golden/reference_tfevents_records.bin

/root/reference/logs/20200820-181339/train/events.out.tfevents.* is a TensorBoard log the reference's trainer wrote
(model.py:262-275, tf.summary).  Its container is TFRecord -- [uint64 length][masked crc32c(length)][payload][masked
crc32c(payload)] -- i.e. the SAME masked CRC-32C that the checkpoint format (tensor bundle, model.py:463-466) puts on every
table block and tensor.  The small records (scalar summaries; the image summaries are skipped) are copied verbatim, framing
included, so that tests/test_tf_checkpoint.py can check sg-gan-tf2_b200/tf_checkpoint.py's crc32c / mask / protobuf wire
reader against bytes produced by TF, not by us.
"""
import glob
import os
import struct

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = sorted(glob.glob("/root/reference/logs/20200820-181339/train/events.out.tfevents.*"))[0]


def main():
    data = open(SRC, "rb").read()
    pos, out, n = 0, bytearray(), 0
    while pos + 12 <= len(data):
        (ln,) = struct.unpack("<Q", data[pos:pos + 8])
        rec = data[pos:pos + 16 + ln]
        if ln <= 400:
            out += rec
            n += 1
        pos += 16 + ln
    open(os.path.join(HERE, "reference_tfevents_records.bin"), "wb").write(bytes(out))
    print("%d records, %d bytes from %s" % (n, len(out), SRC))


if __name__ == "__main__":
    main()
