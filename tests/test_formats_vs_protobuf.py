"""Wire formats checked against an INDEPENDENT parser: the `google.protobuf` runtime with message classes built here from TensorFlow's
published schemas (tensorflow/core/example/{example,feature}.proto, core/protobuf/tensor_bundle.proto, core/framework/
{tensor_shape,types,versions}.proto, core/protobuf/trackable_object_graph.proto — field numbers restated below; no TensorFlow in
this image).  Both directions: bytes serialised by protobuf are read by viewformer_b200's hand-written codecs, and bytes written by
viewformer_b200 are parsed by protobuf.  This pins the protobuf layer of the tf.train.Example token datasets
(viewformer/data/loaders, commands/generate_codes.py) and of the TensorBundle checkpoint index (utils/tensorflow.py:20-63) to
third-party code; the LevelDB table framing around the index stays covered by tests/test_tf_checkpoint.py."""
import struct

import numpy as np
import pytest

from viewformer_b200 import data as D
from viewformer_b200 import tf_checkpoint as tfc

pb = pytest.importorskip("google.protobuf")
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory  # noqa: E402

T = descriptor_pb2.FieldDescriptorProto


def _field(msg, name, number, ftype, label=T.LABEL_OPTIONAL, type_name=None, packed=None, oneof=None):
    f = msg.field.add(name=name, number=number, type=ftype, label=label)
    if type_name:
        f.type_name = type_name
    if packed is not None:
        f.options.packed = packed
    if oneof is not None:
        f.oneof_index = oneof
    return f


@pytest.fixture(scope="module")
def tf_messages():
    fd = descriptor_pb2.FileDescriptorProto(name="vf_test_tf_schemas.proto", package="vftf", syntax="proto3")
    R = T.LABEL_REPEATED
    # ---- feature.proto / example.proto
    m = fd.message_type.add(name="BytesList"); _field(m, "value", 1, T.TYPE_BYTES, R)
    m = fd.message_type.add(name="FloatList"); _field(m, "value", 1, T.TYPE_FLOAT, R, packed=True)
    m = fd.message_type.add(name="Int64List"); _field(m, "value", 1, T.TYPE_INT64, R, packed=True)
    m = fd.message_type.add(name="Feature")
    m.oneof_decl.add(name="kind")
    _field(m, "bytes_list", 1, T.TYPE_MESSAGE, type_name=".vftf.BytesList", oneof=0)
    _field(m, "float_list", 2, T.TYPE_MESSAGE, type_name=".vftf.FloatList", oneof=0)
    _field(m, "int64_list", 3, T.TYPE_MESSAGE, type_name=".vftf.Int64List", oneof=0)
    m = fd.message_type.add(name="Features")
    e = m.nested_type.add(name="FeatureEntry"); e.options.map_entry = True
    _field(e, "key", 1, T.TYPE_STRING); _field(e, "value", 2, T.TYPE_MESSAGE, type_name=".vftf.Feature")
    _field(m, "feature", 1, T.TYPE_MESSAGE, R, type_name=".vftf.Features.FeatureEntry")
    m = fd.message_type.add(name="Example"); _field(m, "features", 1, T.TYPE_MESSAGE, type_name=".vftf.Features")
    # ---- tensor_shape.proto / versions.proto / tensor_bundle.proto
    m = fd.message_type.add(name="TensorShapeProto")
    d = m.nested_type.add(name="Dim"); _field(d, "size", 1, T.TYPE_INT64); _field(d, "name", 2, T.TYPE_STRING)
    _field(m, "dim", 2, T.TYPE_MESSAGE, R, type_name=".vftf.TensorShapeProto.Dim"); _field(m, "unknown_rank", 3, T.TYPE_BOOL)
    m = fd.message_type.add(name="VersionDef")
    _field(m, "producer", 1, T.TYPE_INT32); _field(m, "min_consumer", 2, T.TYPE_INT32); _field(m, "bad_consumers", 3, T.TYPE_INT32, R)
    m = fd.message_type.add(name="BundleHeaderProto")
    _field(m, "num_shards", 1, T.TYPE_INT32); _field(m, "endianness", 2, T.TYPE_INT32)       # enum LITTLE = 0 / BIG = 1, same wire type
    _field(m, "version", 3, T.TYPE_MESSAGE, type_name=".vftf.VersionDef")
    m = fd.message_type.add(name="BundleEntryProto")
    _field(m, "dtype", 1, T.TYPE_INT32)                                                       # enum DataType
    _field(m, "shape", 2, T.TYPE_MESSAGE, type_name=".vftf.TensorShapeProto")
    _field(m, "shard_id", 3, T.TYPE_INT32); _field(m, "offset", 4, T.TYPE_INT64); _field(m, "size", 5, T.TYPE_INT64)
    _field(m, "crc32c", 6, T.TYPE_FIXED32)
    # ---- trackable_object_graph.proto
    m = fd.message_type.add(name="TrackableObjectGraph")
    o = m.nested_type.add(name="TrackableObject")
    r = o.nested_type.add(name="ObjectReference"); _field(r, "node_id", 1, T.TYPE_INT32); _field(r, "local_name", 2, T.TYPE_STRING)
    s = o.nested_type.add(name="SerializedTensor")
    _field(s, "name", 1, T.TYPE_STRING); _field(s, "full_name", 2, T.TYPE_STRING); _field(s, "checkpoint_key", 3, T.TYPE_STRING)
    _field(s, "optional_restore", 4, T.TYPE_BOOL)
    v = o.nested_type.add(name="SlotVariableReference")
    _field(v, "original_variable_node_id", 1, T.TYPE_INT32); _field(v, "slot_name", 2, T.TYPE_STRING); _field(v, "slot_variable_node_id", 3, T.TYPE_INT32)
    P = ".vftf.TrackableObjectGraph.TrackableObject."
    _field(o, "children", 1, T.TYPE_MESSAGE, R, type_name=P + "ObjectReference")
    _field(o, "attributes", 2, T.TYPE_MESSAGE, R, type_name=P + "SerializedTensor")
    _field(o, "slot_variables", 3, T.TYPE_MESSAGE, R, type_name=P + "SlotVariableReference")
    _field(m, "nodes", 1, T.TYPE_MESSAGE, R, type_name=".vftf.TrackableObjectGraph.TrackableObject")
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    names = ["Example", "Features", "Feature", "BundleHeaderProto", "BundleEntryProto", "TrackableObjectGraph"]
    return {n: message_factory.GetMessageClass(pool.FindMessageTypeByName("vftf." + n)) for n in names}


def test_example_written_by_protobuf_is_read_by_decode_example(tf_messages):
    rng = np.random.default_rng(0)
    codes = rng.integers(-5, 1 << 40, size=300).astype(np.int64)
    cams = rng.standard_normal(7 * 20).astype(np.float32)
    frames = [bytes(rng.integers(0, 256, size=n, dtype=np.uint8)) for n in (0, 1, 200, 70000)]
    ex = tf_messages["Example"]()
    ex.features.feature["codes"].int64_list.value.extend(int(c) for c in codes)
    ex.features.feature["cameras"].float_list.value.extend(float(c) for c in cams)
    ex.features.feature["frames"].bytes_list.value.extend(frames)
    ex.features.feature["empty"].int64_list.SetInParent()
    got = D.decode_example(ex.SerializeToString())
    assert np.array_equal(got["codes"], codes) and got["codes"].dtype == np.int64
    assert np.array_equal(got["cameras"], cams) and got["cameras"].dtype == np.float32
    assert got["frames"] == frames
    assert got["empty"].size == 0


def test_example_written_by_encode_example_is_parsed_by_protobuf(tf_messages):
    rng = np.random.default_rng(1)
    feats = dict(codes=rng.integers(-(1 << 62), 1 << 62, size=(5, 4, 4)).astype(np.int64), cameras=rng.standard_normal((5, 7)).astype(np.float32),
                 frames=[b"\x89PNG\r\n\x1a\n" + bytes(100), b"\xff\xd8jpeg"])
    ex = tf_messages["Example"]()
    ex.ParseFromString(D.encode_example(feats))
    f = ex.features.feature
    assert sorted(f.keys()) == ["cameras", "codes", "frames"]
    assert f["codes"].WhichOneof("kind") == "int64_list" and list(f["codes"].int64_list.value) == feats["codes"].reshape(-1).tolist()
    assert f["cameras"].WhichOneof("kind") == "float_list"
    assert np.array_equal(np.asarray(f["cameras"].float_list.value, np.float32), feats["cameras"].reshape(-1))
    assert f["frames"].WhichOneof("kind") == "bytes_list" and list(f["frames"].bytes_list.value) == feats["frames"]
    # protobuf's own re-serialisation of the parsed message has the same length (canonical packed encoding, no stray fields)
    assert len(ex.SerializeToString()) == len(D.encode_example(feats))


def test_bundle_entry_written_by_protobuf_is_read_by_parse_entry(tf_messages):
    E = tf_messages["BundleEntryProto"]
    for dtype, shape, shard, off, size, crc in [(1, (768, 2304), 0, 0, 768 * 2304 * 4, 0xDEADBEEF), (9, (), 3, (1 << 40) + 5, 8, 1),
                                                (7, (), 0, 17, 12345, 0xFFFFFFFF), (19, (1, 0, 3), 1, 1, 0, 0)]:
        m = E(dtype=dtype, shard_id=shard, offset=off, size=size, crc32c=crc)
        m.shape.SetInParent()
        for s in shape:
            m.shape.dim.add(size=s)
        e = tfc.parse_entry(m.SerializeToString())
        assert (e["dtype"], tuple(e["shape"]), e["shard_id"], e["offset"], e["size"]) == (dtype, shape, shard, off, size)
        assert (e["crc32c"] or 0) == crc           # proto3 omits a zero fixed32


def test_checkpoint_index_written_here_is_parsed_by_protobuf(tf_messages, tmp_path):
    rng = np.random.default_rng(2)
    tensors = {"wte/weight": rng.standard_normal((11, 8)).astype(np.float32), "h/0/attn/c_attn/bias": rng.standard_normal((1, 24)).astype(np.float32),
               "h/0/ln_1/gamma": rng.standard_normal(8).astype(np.float32), "h/1/ln_1/gamma": rng.standard_normal(8).astype(np.float16),
               "optimizer/iter": np.asarray(12345678901, np.int64)}
    prefix = str(tmp_path / "model")
    tfc.write_checkpoint(prefix, tensors)
    raw = tfc.read_index(prefix)
    hdr = tf_messages["BundleHeaderProto"]()
    hdr.ParseFromString(raw[""])
    assert hdr.num_shards == 1 and hdr.endianness == 0 and hdr.version.producer == 1
    data = open(prefix + ".data-00000-of-00001", "rb").read()
    codes = {np.dtype(np.float32): 1, np.dtype(np.float16): 19, np.dtype(np.int64): 9}
    spans = []
    for path, arr in tensors.items():
        e = tf_messages["BundleEntryProto"]()
        e.ParseFromString(raw[path + tfc.VAR_SUFFIX])
        assert e.dtype == codes[arr.dtype] and [d.size for d in e.shape.dim] == list(arr.shape) and e.shard_id == 0 and e.size == arr.nbytes
        assert data[e.offset:e.offset + e.size] == arr.tobytes() and e.crc32c == tfc.masked_crc(arr.tobytes())
        spans.append((e.offset, e.size))
    g = tf_messages["BundleEntryProto"]()
    g.ParseFromString(raw[tfc.OBJECT_GRAPH_KEY])
    assert g.dtype == 7 and len(g.shape.dim) == 0                                  # scalar DT_STRING
    spans.append((g.offset, g.size))
    spans.sort()
    assert spans[0][0] == 0 and all(a[0] + a[1] == b[0] for a, b in zip(spans, spans[1:])) and spans[-1][0] + spans[-1][1] == len(data)
    # DT_STRING payload: varint length, masked crc32c of the length bytes, then the serialized TrackableObjectGraph
    blob = data[g.offset:g.offset + g.size]
    ln, pos = tfc._varint(blob, 0)
    assert struct.unpack_from("<I", blob, pos)[0] == tfc.masked_crc(blob[:pos]) and len(blob) == pos + 4 + ln
    graph = tf_messages["TrackableObjectGraph"]()
    graph.ParseFromString(blob[pos + 4:])
    for path in tensors:                                                            # walk it the way Keras does, with protobuf's view of the bytes
        node = graph.nodes[0]
        for part in path.split("/"):
            nxt = [c.node_id for c in node.children if c.local_name == part]
            assert len(nxt) == 1
            node = graph.nodes[nxt[0]]
        assert [(a.name, a.checkpoint_key) for a in node.attributes] == [("VARIABLE_VALUE", path + tfc.VAR_SUFFIX)]


def test_checkpoint_with_protobuf_serialised_messages_is_read_by_checkpoint(tf_messages, tmp_path):
    """An index whose values are all produced by protobuf — header with a full VersionDef, entries in protobuf's field order, an object
    graph carrying the fields a Keras checkpoint has and this reader must skip (full_name, optional_restore, slot variables of an
    optimizer, children listed before attributes, a node reachable by two names) — behind the LevelDB table framing of this package."""
    rng = np.random.default_rng(3)
    G = tf_messages["TrackableObjectGraph"]
    g = G()
    names = {}                                      # path tuple -> node id

    def node(path):
        if path not in names:
            names[path] = len(g.nodes)
            g.nodes.add()
            if path:
                parent = node(path[:-1])
                g.nodes[parent].children.add(node_id=names[path], local_name=path[-1])
        return names[path]

    node(())
    tensors = {("h", "0", "mlp", "c_fc", "weight"): rng.standard_normal((8, 32)).astype(np.float32),
               ("h", "0", "mlp", "c_fc", "bias"): rng.standard_normal((1, 32)).astype(np.float32),
               ("ln_f", "gamma"): rng.standard_normal(8).astype(np.float32),
               ("optimizer", "iter"): np.asarray(77, np.int64)}
    data, entries = bytearray(), {}
    for path, arr in tensors.items():
        nid = node(path)
        key = "/".join(path) + tfc.VAR_SUFFIX
        g.nodes[nid].attributes.add(name="VARIABLE_VALUE", full_name="migt/" + "/".join(path) + ":0", checkpoint_key=key, optional_restore=False)
        e = tf_messages["BundleEntryProto"](dtype={np.dtype(np.float32): 1, np.dtype(np.int64): 9}[arr.dtype], offset=len(data), size=arr.nbytes,
                                            crc32c=tfc.masked_crc(arr.tobytes()))
        e.shape.SetInParent()
        for s in arr.shape:
            e.shape.dim.add(size=s)
        entries[key] = e.SerializeToString()
        data += arr.tobytes()
    # Adam slots of one variable + an alias edge ("layer_with_weights-0" -> the same node as h/0), as Keras writes
    m_id = node(("optimizer", "m_slot"))
    g.nodes[names[("optimizer",)]].slot_variables.add(original_variable_node_id=names[("ln_f", "gamma")], slot_name="m", slot_variable_node_id=m_id)
    g.nodes[0].children.add(node_id=names[("h", "0")], local_name="layer_with_weights-0")
    blob = g.SerializeToString()
    lens = tfc._put_varint(len(blob))
    sraw = lens + struct.pack("<I", tfc.masked_crc(lens)) + blob
    e = tf_messages["BundleEntryProto"](dtype=7, offset=len(data), size=len(sraw), crc32c=tfc.masked_crc(sraw))
    e.shape.SetInParent()
    entries[tfc.OBJECT_GRAPH_KEY] = e.SerializeToString()
    data += sraw
    hdr = tf_messages["BundleHeaderProto"](num_shards=1, endianness=0)
    hdr.version.producer = 1
    hdr.version.min_consumer = 0
    prefix = str(tmp_path / "keras")
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(data))
    items = sorted([(b"", hdr.SerializeToString())] + [(k.encode(), v) for k, v in entries.items()])
    with open(prefix + ".index", "wb") as f:
        off, size = tfc._emit_block(f, tfc._build_block(items, restart_interval=2))
        moff, msize = tfc._emit_block(f, tfc._build_block([]))
        ioff, isize = tfc._emit_block(f, tfc._build_block([(items[-1][0], tfc._put_varint(off) + tfc._put_varint(size))], restart_interval=1))
        footer = tfc._put_varint(moff) + tfc._put_varint(msize) + tfc._put_varint(ioff) + tfc._put_varint(isize)
        f.write(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", 0xDB4775248B80FB57))
    ck = tfc.Checkpoint(prefix)
    nodes = ck.object_graph()
    for path, arr in tensors.items():
        key = ck.resolve("/".join(path), nodes)
        assert key == "/".join(path) + tfc.VAR_SUFFIX and np.array_equal(ck.tensor(key, verify_crc=True), arr)
    assert ck.resolve("layer_with_weights-0/mlp/c_fc/weight", nodes) == "h/0/mlp/c_fc/weight" + tfc.VAR_SUFFIX       # alias edge
    sd = tfc.load_state_dict(prefix, ["h.0.mlp.c_fc.weight", "h.0.mlp.c_fc.bias", "ln_f.gamma"])
    assert np.array_equal(sd["ln_f.gamma"].numpy(), tensors[("ln_f", "gamma")])
    with pytest.raises(RuntimeError, match="Missing keys"):
        tfc.load_state_dict(prefix, ["h.1.mlp.c_fc.weight"])
