"""Container formats checked against code and schemas WRITTEN BY THE TENSORFLOW AUTHORS that happen to be in this image: the
`tensorboard` package carries (a) TensorFlow's TFRecord writer / reader with its masked CRC-32C (tensorboard/summary/writer/
record_writer.py, tensorboard/compat/tensorflow_stub/pywrap_tensorflow.py) and (b) protoc-generated modules of TensorFlow's own
.proto files (tensorboard/compat/proto/: trackable_object_graph_pb2, tensor_shape_pb2, types_pb2, versions_pb2).

tests/test_formats_vs_protobuf.py restates the schemas' field numbers by hand; here nothing is restated: the TFRecord files of the
token datasets (viewformer/data/tfrecord_dataset.py:134-197, commands/generate_codes.py:20-98) and the object graph / entry protos of
the TF checkpoints (viewformer/utils/tensorflow.py:20-63) are exchanged with TensorFlow-authored code in both directions.  The same
masked CRC-32C guards every block of the LevelDB table that holds the checkpoint index, so pinning it pins the block trailers too.
"""
import struct

import numpy as np
import pytest

from viewformer_b200 import data as D
from viewformer_b200 import tf_checkpoint as tfc

pytest.importorskip("tensorboard")
from tensorboard.compat.proto import tensor_shape_pb2, trackable_object_graph_pb2, types_pb2, versions_pb2  # noqa: E402
from tensorboard.compat.tensorflow_stub import errors as tb_errors  # noqa: E402
from tensorboard.compat.tensorflow_stub import pywrap_tensorflow as tb  # noqa: E402
from tensorboard.summary.writer.record_writer import RecordWriter  # noqa: E402


def _tb_records(path):
    r, out = tb.PyRecordReader_New(str(path)), []
    while True:
        try:
            r.GetNext()
        except tb_errors.OutOfRangeError:
            return out
        out.append(r.record())


def test_crc32c_and_mask_equal_tensorflows():
    assert tfc.crc32c(b"123456789") == 0xE3069283 == tb.crc32c(b"123456789")        # the CRC-32C (Castagnoli) check value
    rng = np.random.default_rng(0)
    for n in (0, 1, 3, 4, 7, 8, 9, 63, 64, 65, 1000, 70001):
        b = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert tfc.crc32c(b) == tb.crc32c(b)
        assert tfc.masked_crc(b) == tb.masked_crc32c(b)
    # incremental form (the table reader checks block contents + the 1-byte compression type in two pieces)
    a, b = bytes(range(200)), b"\x00"
    assert tfc.crc32c(b, tfc.crc32c(a)) == tb.crc32c(a + b)


def test_tfrecords_written_here_are_read_by_tensorflows_reader(tmp_path):
    rng = np.random.default_rng(1)
    recs = [b"", b"x", rng.integers(0, 256, 5000, dtype=np.uint8).tobytes(),
            D.encode_example(dict(codes=rng.integers(0, 1024, (20, 8, 8)).astype(np.int64), cameras=rng.standard_normal((20, 7)).astype(np.float32)))]
    p = tmp_path / "a.tfrecord"
    with D.TFRecordWriter(str(p)) as w:
        for r in recs:
            w.write(r)
    assert _tb_records(p) == recs                 # TensorFlow's reader verifies both CRCs of every record


def test_tfrecords_written_by_tensorflows_writer_are_read_here(tmp_path):
    rng = np.random.default_rng(2)
    recs = [rng.integers(0, 256, n, dtype=np.uint8).tobytes() for n in (0, 1, 11, 4096, 100003)]
    p = tmp_path / "b.tfrecord"
    w = RecordWriter(open(p, "wb"))
    for r in recs:
        w.write(r)
    w.close()
    assert list(D.read_tfrecords(str(p), verify=True)) == recs
    raw = bytearray(open(p, "rb").read())
    raw[(12 + 0 + 4) + (12 + 1 + 4) + 12 + 5] ^= 1   # one bit inside the third record's 11-byte payload
    open(p, "wb").write(bytes(raw))
    with pytest.raises(ValueError, match="corrupt"):
        list(D.read_tfrecords(str(p), verify=True))
    with pytest.raises(tb_errors.DataLossError):  # and TensorFlow's reader agrees on where the damage is
        _tb_records(p)


def test_token_dataset_shards_are_valid_tfrecord_files_for_tensorflow(tmp_path):
    """write_token_dataset -> every shard opens with TensorFlow's record reader, and the Example payloads decode to the scenes."""
    rng = np.random.default_rng(3)
    scenes = [dict(codes=rng.integers(0, 1024, (n, 8, 8)).astype(np.int64), cameras=rng.standard_normal((n, 7)).astype(np.float32))
              for n in (7, 12, 9, 20, 5)]
    D.write_token_dataset(str(tmp_path), "train", scenes, token_image_size=8, scenes_per_shard=2)
    files = sorted(tmp_path.glob("*.tfrecord"))
    assert len(files) == 3
    got = [D.decode_example(r) for f in files for r in _tb_records(f)]
    assert len(got) == len(scenes)
    for g, s in zip(got, scenes):
        assert np.array_equal(np.asarray(g["codes"]).reshape(s["codes"].shape), s["codes"])
        assert np.array_equal(np.asarray(g["cameras"], np.float32).reshape(s["cameras"].shape), s["cameras"])


def test_dtype_codes_are_tensorflows_enum():
    want = {np.float32: types_pb2.DT_FLOAT, np.float64: types_pb2.DT_DOUBLE, np.int32: types_pb2.DT_INT32, np.uint8: types_pb2.DT_UINT8,
            np.int16: types_pb2.DT_INT16, np.int8: types_pb2.DT_INT8, np.int64: types_pb2.DT_INT64, np.bool_: types_pb2.DT_BOOL,
            np.float16: types_pb2.DT_HALF}
    assert {np.dtype(v): k for k, v in tfc._DTYPES.items()} == {np.dtype(k): v for k, v in want.items()}
    assert types_pb2.DT_STRING == 7               # the object-graph entry's dtype in write_checkpoint / Checkpoint.tensor


def _graph_blob(ck_prefix):
    raw = tfc.read_index(ck_prefix)
    e = tfc.parse_entry(raw[tfc.OBJECT_GRAPH_KEY])
    data = open(ck_prefix + ".data-00000-of-00001", "rb").read()
    blob = data[e["offset"]:e["offset"] + e["size"]]
    ln, pos = tfc._varint(blob, 0)
    assert struct.unpack_from("<I", blob, pos)[0] == tb.masked_crc32c(blob[:pos]) and len(blob) == pos + 4 + ln
    assert (e["crc32c"] or 0) == tb.masked_crc32c(blob)
    return blob[pos + 4:], raw, data


def test_object_graph_written_here_parses_with_tensorflows_generated_proto(tmp_path):
    rng = np.random.default_rng(4)
    tensors = {"wte/weight": rng.standard_normal((11, 8)).astype(np.float32), "h/0/attn/c_attn/weight": rng.standard_normal((8, 24)).astype(np.float32),
               "h/0/attn/c_attn/bias": rng.standard_normal((1, 24)).astype(np.float32), "ln_f/beta": rng.standard_normal(8).astype(np.float32),
               "optimizer/iter": np.asarray(5, np.int64)}
    prefix = str(tmp_path / "model")
    tfc.write_checkpoint(prefix, tensors)
    blob, raw, data = _graph_blob(prefix)
    g = trackable_object_graph_pb2.TrackableObjectGraph()
    g.ParseFromString(blob)
    assert g.SerializeToString() == blob          # canonical: TensorFlow's generated code re-serialises the bytes unchanged
    for path, arr in tensors.items():
        node = g.nodes[0]
        for part in path.split("/"):
            (nid,) = [c.node_id for c in node.children if c.local_name == part]
            node = g.nodes[nid]
        (a,) = node.attributes
        assert (a.name, a.checkpoint_key) == ("VARIABLE_VALUE", path + tfc.VAR_SUFFIX)
        # the entry's shape sub-message with TensorFlow's TensorShapeProto, its payload CRC with TensorFlow's CRC
        e = tfc.parse_entry(raw[a.checkpoint_key])
        shape_bytes = [v for fn, _, v in tfc._fields(raw[a.checkpoint_key]) if fn == 2]
        shp = tensor_shape_pb2.TensorShapeProto()
        shp.ParseFromString(shape_bytes[0] if shape_bytes else b"")
        assert [d.size for d in shp.dim] == list(arr.shape) and not shp.unknown_rank
        assert e["crc32c"] == tb.masked_crc32c(data[e["offset"]:e["offset"] + e["size"]])
    ver = [v for fn, _, v in tfc._fields(raw[""]) if fn == 3]
    vd = versions_pb2.VersionDef()
    vd.ParseFromString(ver[0])
    assert vd.producer == 1 and vd.min_consumer == 0


def test_object_graph_serialised_by_tensorflows_proto_is_resolved_here(tmp_path):
    """A Keras-style graph built with TensorFlow's generated classes — including the fields this reader skips (full_name, slot variables,
    registered_saver, has_checkpoint_values) and an alias edge — spliced into a checkpoint written here; Checkpoint.resolve must walk it."""
    rng = np.random.default_rng(5)
    tensors = {"h/0/mlp/c_fc/weight": rng.standard_normal((8, 32)).astype(np.float32), "h/0/mlp/c_fc/bias": rng.standard_normal((1, 32)).astype(np.float32),
               "ln_f/gamma": rng.standard_normal(8).astype(np.float32)}
    prefix = str(tmp_path / "keras")
    tfc.write_checkpoint(prefix, tensors)
    blob, _, _ = _graph_blob(prefix)
    G = trackable_object_graph_pb2.TrackableObjectGraph
    g = G()
    g.ParseFromString(blob)
    ids = {}
    for path in tensors:                                                                # decorate what Keras would decorate
        nid = 0
        for part in path.split("/"):
            (nid,) = [c.node_id for c in g.nodes[nid].children if c.local_name == part]
        ids[path] = nid
        g.nodes[nid].attributes[0].full_name = "migt/" + path + ":0"
        g.nodes[nid].has_checkpoint_values.value = True
    (h0,) = [c.node_id for c in g.nodes[[c.node_id for c in g.nodes[0].children if c.local_name == "h"][0]].children if c.local_name == "0"]
    g.nodes[0].children.add(node_id=h0, local_name="layer_with_weights-0")              # alias edge
    opt = g.nodes.add()
    opt.registered_saver.name = "optimizer"
    opt.slot_variables.add(original_variable_node_id=ids["ln_f/gamma"], slot_name="m", slot_variable_node_id=len(g.nodes) - 1)
    g.nodes[0].children.add(node_id=len(g.nodes) - 1, local_name="optimizer")
    new_blob = g.SerializeToString()
    # splice: append the new string tensor to the data shard and point the graph entry at it (same framing as TensorFlow's DT_STRING)
    lens = tfc._put_varint(len(new_blob))
    sraw = lens + struct.pack("<I", tb.masked_crc32c(lens)) + new_blob
    raw = tfc.read_index(prefix)
    data = open(prefix + ".data-00000-of-00001", "rb").read()
    raw[tfc.OBJECT_GRAPH_KEY] = tfc._msg(tfc._f_varint(1, types_pb2.DT_STRING), tfc._f_bytes(2, b""), tfc._f_varint(4, len(data)),
                                        tfc._f_varint(5, len(sraw)), tfc._f_fixed32(6, tb.masked_crc32c(sraw)))
    open(prefix + ".data-00000-of-00001", "wb").write(data + sraw)
    items = sorted((k.encode(), v) for k, v in raw.items())
    with open(prefix + ".index", "wb") as f:
        off, size = tfc._emit_block(f, tfc._build_block(items, restart_interval=3))
        moff, msize = tfc._emit_block(f, tfc._build_block([]))
        ioff, isize = tfc._emit_block(f, tfc._build_block([(items[-1][0], tfc._put_varint(off) + tfc._put_varint(size))], restart_interval=1))
        footer = tfc._put_varint(moff) + tfc._put_varint(msize) + tfc._put_varint(ioff) + tfc._put_varint(isize)
        f.write(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", tfc._MAGIC))
    ck = tfc.Checkpoint(prefix)
    nodes = ck.object_graph()
    for path, arr in tensors.items():
        assert np.array_equal(ck.tensor(ck.resolve(path, nodes), verify_crc=True), arr)
    assert ck.resolve("layer_with_weights-0/mlp/c_fc/bias", nodes) == "h/0/mlp/c_fc/bias" + tfc.VAR_SUFFIX
    sd = tfc.load_state_dict(prefix, ["h.0.mlp.c_fc.weight", "h.0.mlp.c_fc.bias", "ln_f.gamma"])
    assert np.array_equal(sd["h.0.mlp.c_fc.weight"].numpy(), tensors["h/0/mlp/c_fc/weight"])
