"""Property tests (hypothesis) of the hand-written wire codecs: varints, tf.train.Example, TFRecord framing and the checkpoint table —
arbitrary keys / shapes / values must survive a write -> read round trip (with every CRC verified)."""
import struct

import numpy as np
import pytest

hyp = pytest.importorskip("hypothesis")
from hypothesis import given, settings, strategies as st, HealthCheck  # noqa: E402

from viewformer_b200 import data as D  # noqa: E402
from viewformer_b200 import tf_checkpoint as tfc  # noqa: E402

FAST = settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])


@FAST
@given(st.integers(min_value=0, max_value=(1 << 64) - 1))
def test_varint_roundtrip(v):
    b = tfc._put_varint(v)
    assert 1 <= len(b) <= 10 and all(x & 0x80 for x in b[:-1]) and not b[-1] & 0x80
    assert tfc._varint(b + b"\xff", 0) == (v, len(b))


@FAST
@given(codes=st.lists(st.integers(min_value=-(1 << 63), max_value=(1 << 63) - 1), max_size=40),
       cams=st.lists(st.floats(width=32, allow_nan=False), max_size=40),
       frames=st.lists(st.binary(max_size=50), max_size=4))
def test_example_roundtrip(codes, cams, frames):
    feats = dict(codes=np.asarray(codes, np.int64), cameras=np.asarray(cams, np.float32), frames=frames)
    buf = D.encode_example(feats)
    got = D.decode_example(buf)
    assert np.array_equal(np.asarray(got["codes"], np.int64), feats["codes"])
    assert np.array_equal(np.asarray(got["cameras"], np.float32), feats["cameras"])
    assert list(got["frames"]) == frames


@FAST
@given(st.lists(st.binary(max_size=300), max_size=8))
def test_tfrecord_roundtrip(tmp_path_factory, recs):
    p = str(tmp_path_factory.mktemp("rec") / "x.tfrecord")
    with D.TFRecordWriter(p) as w:
        for r in recs:
            w.write(r)
    assert list(D.read_tfrecords(p, verify=True)) == recs
    raw = open(p, "rb").read()
    assert len(raw) == sum(16 + len(r) for r in recs)


_name = st.text(alphabet="abcdefgh_0123456789", min_size=1, max_size=6)
_path = st.lists(_name, min_size=1, max_size=4).map("/".join)


@settings(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(paths=st.lists(_path, min_size=1, max_size=90, unique=True), seed=st.integers(0, 1 << 30))
def test_checkpoint_table_roundtrip_with_arbitrary_keys(tmp_path_factory, paths, seed):
    """Any set of attribute paths (shared prefixes, one path a prefix of another's component, > 64 keys -> several data blocks)."""
    # a path that is a strict prefix of another ('a' and 'a/b') would make one node both a variable and a container: legal in the object
    # graph, kept here on purpose
    rng = np.random.default_rng(seed)
    dts = [np.float32, np.float16, np.int64, np.int32, np.uint8, np.bool_, np.float64]
    tensors = {}
    for i, p in enumerate(paths):
        shape = tuple(int(x) for x in rng.integers(0, 4, size=int(rng.integers(0, 4))))
        dt = dts[i % len(dts)]
        tensors[p] = (rng.standard_normal(shape) * 5).astype(dt) if dt != np.bool_ else rng.integers(0, 2, shape).astype(np.bool_)
    prefix = str(tmp_path_factory.mktemp("ck") / "m")
    tfc.write_checkpoint(prefix, tensors)
    raw = tfc.read_index(prefix, verify=True)
    keys = list(raw)
    assert keys == sorted(keys, key=lambda k: k.encode()) and keys[0] == ""          # the table is sorted bytewise; header under the empty key
    ck = tfc.Checkpoint(prefix)
    nodes = ck.object_graph()
    for p, arr in tensors.items():
        got = ck.tensor(ck.resolve(p, nodes), verify_crc=True)
        assert got.dtype == arr.dtype and got.shape == arr.shape and np.array_equal(got, arr)
    assert struct.unpack("<Q", open(prefix + ".index", "rb").read()[-8:])[0] == 0xDB4775248B80FB57
