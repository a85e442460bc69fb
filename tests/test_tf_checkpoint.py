"""TensorBundle / object-graph checkpoint container (viewformer_b200/tf_checkpoint.py): pure-Python reader + writer (CPU).
No TensorFlow exists in this image, so the reader is exercised against the writer and against hand-built format details
(prefix-compressed multi-block SSTable, masked crc32c, DT_STRING object graph, attribute-path resolution)."""
import os
import struct

import numpy as np
import pytest
import torch

from viewformer_b200 import tf_checkpoint as tfc
from viewformer_b200.config import MIGTConfig


def test_crc32c_known_answers():
    assert tfc.crc32c(b"123456789") == 0xE3069283          # the standard CRC-32C check value
    assert tfc.crc32c(b"") == 0
    assert tfc.masked_crc(b"123456789") == ((((0xE3069283 >> 15) | (0xE3069283 << 17)) + 0xA282EAD8) & 0xFFFFFFFF)


def test_roundtrip_many_keys_and_prefix_compression(tmp_path):
    rng = np.random.default_rng(0)
    tensors = {f"h/{i}/attn/c_attn/weight": rng.standard_normal((6, 18)).astype(np.float32) for i in range(40)}
    tensors.update({f"h/{i}/attn/c_attn/bias": rng.standard_normal((1, 18)).astype(np.float32) for i in range(40)})
    tensors.update({f"h/{i}/ln_1/gamma": rng.standard_normal(6).astype(np.float32) for i in range(40)})
    tensors["wte/weight"] = rng.standard_normal((10, 6)).astype(np.float32)
    tensors["step"] = np.asarray(7, dtype=np.int64)
    prefix = str(tmp_path / "model")
    tfc.write_checkpoint(prefix, tensors)
    assert os.path.exists(prefix + ".index") and os.path.exists(prefix + ".data-00000-of-00001")
    with open(prefix + ".index", "rb") as f:
        raw = f.read()
    assert struct.unpack("<Q", raw[-8:])[0] == 0xDB4775248B80FB57
    ck = tfc.Checkpoint(prefix)
    assert len(ck.entries) == len(tensors) + 1 and ck.num_shards == 1          # + the object graph
    nodes = ck.object_graph()
    for path, arr in tensors.items():
        key = ck.resolve(path, nodes)
        assert key == path + tfc.VAR_SUFFIX
        got = ck.tensor(key, verify_crc=True)
        assert got.dtype == arr.dtype and got.shape == arr.shape and np.array_equal(got, arr)
    with pytest.raises(KeyError):
        ck.resolve("h/0/attn/nope", nodes)
    # corrupt one data byte: the per-tensor crc catches it
    with open(prefix + ".data-00000-of-00001", "r+b") as f:
        f.seek(5)
        b = f.read(1)
        f.seek(5)
        f.write(bytes([b[0] ^ 0xFF]))
    first = min(ck.entries, key=lambda k: (ck.entries[k]["offset"], k == tfc.OBJECT_GRAPH_KEY))
    with pytest.raises(ValueError, match="crc32c"):
        ck.tensor(first, verify_crc=True)


def test_migt_state_dict_roundtrip_through_tf_container(tmp_path):
    """MIGT.save_weights / load_weights surface (Keras API of train_transformer.py:106-129) without touching the device."""
    from viewformer_b200.migt import MIGT
    from oracle import synth
    cfg = MIGTConfig(n_layer=2, n_head=2, d_model=32, sequence_size=4, n_embeddings=64, token_image_size=2)
    sd = synth.make_migt_state_dict(cfg, 3)
    m = MIGT(cfg)
    m._sd = {k: v.clone() for k, v in sd.items()}          # host copy only: no device needed for the container round trip
    prefix = str(tmp_path / "ckpt" / "model")
    m.save_weights(prefix)
    got = tfc.load_state_dict(prefix, m.expected_keys())
    assert set(got) == set(sd)
    for k in sd:
        assert torch.equal(got[k], sd[k]), k
    # a checkpoint key spelled differently from the attribute path still resolves through the object graph
    ck = tfc.Checkpoint(prefix)
    assert ck.resolve("h/1/mlp/c_fc/weight") == "h/1/mlp/c_fc/weight" + tfc.VAR_SUFFIX
    with pytest.raises(RuntimeError, match="Missing keys"):
        tfc.load_state_dict(prefix, ["h.5.ln_1.gamma"])
    # the file is laid out like the reference's Keras model tracks its variables (attribute names, tfc.object_paths): wpe hangs off the
    # root, the pose classifier is a child of pose_criterion
    assert ck.resolve("wpe") == "wpe" + tfc.VAR_SUFFIX and ck.resolve("pose_criterion/pose_classifier/c_fc/weight")
    with pytest.raises(KeyError):
        ck.resolve("pose_classifier/c_fc/weight")
    # ... while files written with the older literal layout (wpe/embeddings, pose_classifier at the root) — or a Keras file seen through
    # add_weight's dependency name 'embeddings' — still load
    old = str(tmp_path / "old" / "model")
    tfc.write_checkpoint(old, {k.replace(".", "/"): v.numpy() for k, v in sd.items()})
    got = tfc.load_state_dict(old, m.expected_keys())
    assert all(torch.equal(got[k], sd[k]) for k in sd)
    alt = str(tmp_path / "alt" / "model")
    tfc.write_checkpoint(alt, {("embeddings" if k == "wpe.embeddings" else tfc.object_paths(k)[0]): v.numpy() for k, v in sd.items()})
    assert torch.equal(tfc.load_state_dict(alt, ["wpe.embeddings"])["wpe.embeddings"], sd["wpe.embeddings"])


def test_reader_on_a_table_assembled_byte_by_byte_from_the_leveldb_format(tmp_path):
    """Known-answer test that does not use this module's writer: a two-data-block table typed out from leveldb/doc/table_format.md
    (entry = varint shared | varint non_shared | varint value_len | key delta | value; restart array; 5-byte block trailer; index block
    of separator keys -> BlockHandle; 48-byte footer with the magic) holding literal BundleHeaderProto / BundleEntryProto bytes of
    tensorflow/core/protobuf/tensor_bundle.proto.  The block CRCs come from TensorFlow's own crc32c shipped inside tensorboard when it is
    importable (else from this module), so reader, framing and checksum are pinned independently of `_build_block` / `_emit_block`."""
    try:
        from tensorboard.compat.tensorflow_stub.pywrap_tensorflow import masked_crc32c as mcrc
    except Exception:
        mcrc = tfc.masked_crc
    w = np.asarray([1.5, -2.0], np.float32).tobytes()                    # "a/x" : f32[2]
    b = np.asarray([7], np.int64).tobytes()                              # "a/y" : i64[1]
    z = np.asarray([0.25], np.float32).tobytes()                         # "a/z" : f32[1]
    data = w + b + z
    header = bytes([0x08, 0x01, 0x1A, 0x02, 0x08, 0x01])                 # num_shards = 1; version { producer = 1 }
    e_x = bytes([0x08, 0x01, 0x12, 0x04, 0x12, 0x02, 0x08, 0x02, 0x28, 0x08, 0x35]) + struct.pack("<I", mcrc(w))      # DT_FLOAT, shape{dim{size 2}}, size 8, crc
    e_y = bytes([0x08, 0x09, 0x12, 0x04, 0x12, 0x02, 0x08, 0x01, 0x20, 0x08, 0x28, 0x08, 0x35]) + struct.pack("<I", mcrc(b))  # DT_INT64, [1], offset 8, size 8
    e_z = bytes([0x08, 0x01, 0x12, 0x04, 0x12, 0x02, 0x08, 0x01, 0x20, 0x10, 0x28, 0x04, 0x35]) + struct.pack("<I", mcrc(z))  # DT_FLOAT, [1], offset 16, size 4
    kx, ky = b"a/x" + tfc.VAR_SUFFIX.encode(), b"a/y" + tfc.VAR_SUFFIX.encode()
    kz = b"a/z" + tfc.VAR_SUFFIX.encode()
    assert kx[:2] == ky[:2] and len(kx) == len(ky) == 30

    def trailer(block):
        return b"\x00" + struct.pack("<I", mcrc(block + b"\x00"))

    # data block 0: the header under the empty key, then kx in full (restart interval 16 -> one restart point at 0)
    blk0 = (bytes([0, 0, len(header)]) + header
            + bytes([0, len(kx), len(e_x)]) + kx + e_x
            + struct.pack("<II", 0, 1))
    # data block 1: ky in full (a new block starts with a full key), then kz PREFIX-COMPRESSED against it: shares "a/", differs from byte 2 on
    blk1 = (bytes([0, len(ky), len(e_y)]) + ky + e_y
            + bytes([2, len(kz) - 2, len(e_z)]) + kz[2:] + e_z
            + struct.pack("<II", 0, 1))
    off0, off1 = 0, len(blk0) + 5
    meta = struct.pack("<II", 0, 1)                                      # empty metaindex block
    moff = off1 + len(blk1) + 5
    # index block: separator >= last key of the block it points to ("a/x/.ATTRIBUTES/VARIABLE_VALUE" < "a/x0" is false -> use the keys' successors)
    sep0, sep1 = b"a/x/~", b"b"
    assert kx <= sep0 < ky < kz <= sep1
    h0, h1 = bytes([off0, len(blk0)]), bytes([off1, len(blk1)])
    assert max(off0, off1, len(blk0), len(blk1)) < 128                   # single-byte varints
    index = (bytes([0, len(sep0), len(h0)]) + sep0 + h0
             + bytes([0, len(sep1), len(h1)]) + sep1 + h1
             + struct.pack("<III", 0, len(sep0) + len(h0) + 3, 2))       # restart interval 1: one restart per entry
    ioff = moff + len(meta) + 5
    assert ioff < 16384
    vi = lambda v: bytes([v]) if v < 128 else bytes([(v & 0x7F) | 0x80, v >> 7])
    footer = vi(moff) + vi(len(meta)) + vi(ioff) + vi(len(index))
    footer += b"\x00" * (40 - len(footer)) + bytes([0x57, 0xFB, 0x80, 0x8B, 0x24, 0x75, 0x47, 0xDB])     # kTableMagicNumber, little endian
    table = blk0 + trailer(blk0) + blk1 + trailer(blk1) + meta + trailer(meta) + index + trailer(index) + footer
    prefix = str(tmp_path / "typed")
    open(prefix + ".index", "wb").write(table)
    open(prefix + ".data-00000-of-00001", "wb").write(data)

    raw = tfc.read_index(prefix, verify=True)
    assert list(raw) == ["", kx.decode(), ky.decode(), kz.decode()] and raw[""] == header and raw[kx.decode()] == e_x and raw[kz.decode()] == e_z
    ck = tfc.Checkpoint(prefix)
    assert ck.num_shards == 1
    assert np.array_equal(ck.tensor(kx.decode(), verify_crc=True), np.asarray([1.5, -2.0], np.float32))
    got = ck.tensor(ky.decode(), verify_crc=True)
    assert got.dtype == np.int64 and got.tolist() == [7]
    assert ck.tensor(kz.decode(), verify_crc=True).tolist() == [0.25]
    sd = tfc.load_state_dict(prefix, ["a.x", "a.y"])                     # no object graph in this file: keys resolve by the naming convention
    assert sd["a.x"].tolist() == [1.5, -2.0]
    # a flipped bit inside the second data block is caught by the block checksum (and only when asked, as in LevelDB)
    bad = bytearray(table)
    bad[off1 + 10] ^= 0x20
    open(prefix + ".index", "wb").write(bytes(bad))
    with pytest.raises(ValueError, match="crc32c mismatch"):
        tfc.read_index(prefix, verify=True)
    # the module's own writer produces tables the typed-out rules accept: same trailer rule on every block it emits
    p2 = str(tmp_path / "own")
    tfc.write_checkpoint(p2, {"a/x": np.asarray([1.5, -2.0], np.float32), "a/y": np.asarray([7], np.int64)})
    own = tfc.read_index(p2, verify=True)
    assert own[kx.decode()] == e_x and tfc.parse_entry(own[ky.decode()])["size"] == 8
