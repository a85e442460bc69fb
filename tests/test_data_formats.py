"""TFRecord / tf.train.Example codec, token-dataset loader and the generate-codes scene re-batching logic (CPU)."""
import numpy as np
import pytest
import torch

from viewformer_b200 import data as D


def test_example_codec_roundtrip_and_known_bytes(tmp_path):
    feats = dict(codes=np.arange(-3, 200, dtype=np.int64), cameras=np.linspace(-1, 1, 14, dtype=np.float32), frames=[b"\x89PNG...", b"jpeg"])
    buf = D.encode_example(feats)
    got = D.decode_example(buf)
    assert np.array_equal(got["codes"], feats["codes"]) and np.array_equal(got["cameras"], feats["cameras"]) and got["frames"] == feats["frames"]
    # known answer: Example{features{feature{key:"a" value{int64_list{value:[1]}}}}} as protoc encodes it (packed int64)
    assert D.encode_example(dict(a=np.asarray([1], np.int64))) == bytes.fromhex("0a0c0a0a0a0161120 51a030a0101".replace(" ", ""))
    path = str(tmp_path / "x.tfrecord")
    with D.TFRecordWriter(path) as w:
        w.write(buf)
        w.write(b"")
    recs = list(D.read_tfrecords(path, verify=True))
    assert recs == [buf, b""]
    raw = bytearray(open(path, "rb").read())
    raw[20] ^= 0xFF
    open(path, "wb").write(bytes(raw))
    with pytest.raises(ValueError):
        list(D.read_tfrecords(path, verify=True))


def test_token_dataset_windows_and_sharding(tmp_path):
    rng = np.random.default_rng(0)
    scenes = [dict(cameras=rng.standard_normal((t, 7)).astype(np.float32) + 100 * i, codes=np.full((t, 2, 2), i, np.int64))
              for i, t in enumerate([9, 4, 13, 3, 8, 8, 8, 21])]
    D.write_token_dataset(str(tmp_path), "train", scenes, token_image_size=2, scenes_per_shard=3)
    seq = 4
    batches = list(D.load_token_dataset(str(tmp_path), batch_size=2, sequence_size=seq, token_image_size=2, split="train", repeat=1, seed=3))
    assert all(p.shape == (2, seq, 7) and t.shape == (2, seq, 2, 2) for p, t in batches)
    n_windows = sum(t // seq for t in [9, 4, 13, 3, 8, 8, 8, 21])
    assert len(batches) == n_windows // 2
    for p, t in batches:
        for b in range(2):
            sid = int(t[b, 0, 0, 0])
            assert (t[b] == sid).all() and torch.allclose(p[b].mean(), torch.tensor(100.0 * sid), atol=3.0)     # a window never mixes scenes
            assert len({tuple(v.tolist()) for v in p[b]}) == seq                                             # distinct views
    capped = list(D.load_token_dataset(str(tmp_path), 1, seq, 2, split="train", repeat=1, max_samples_per_environment=1))
    assert len(capped) == sum(1 for t in [9, 4, 13, 3, 8, 8, 8, 21] if t >= seq)
    r0 = list(D.load_token_dataset(str(tmp_path), 2, seq, 2, split="train", repeat=1, rank=0, world=2))
    r1 = list(D.load_token_dataset(str(tmp_path), 2, seq, 2, split="train", repeat=1, rank=1, world=2))
    s0 = {int(t[b, 0, 0, 0]) for _, t in r0 for b in range(t.shape[0])}
    s1 = {int(t[b, 0, 0, 0]) for _, t in r1 for b in range(t.shape[0])}
    assert not (s0 & s1) and all(p.shape[0] == 1 for p, _ in r0 + r1)        # disjoint file shards, local batch = global / world


def test_latent_code_transformer_rebatches_across_scenes():
    class FakeCodebook:
        class config:
            image_size, stride, batch_size = 8, 4, 5
        device = "cpu"
        calls = []

        def encode_u8(self, x):
            self.calls.append(len(x))
            return x[:, :2, :2, 0].to(torch.int64)          # "codes" = top-left pixels: traceable back to the frame

    import viewformer_b200._lib as L
    orig = L.resize_u8
    L.resize_u8 = lambda x, size, method=None: x            # no device in this test
    try:
        cb = FakeCodebook()
        tr = D.LatentCodeTransformer(cb, batch_size=5)
        assert tr.update_dataset_info({})["token_image_size"] == 2
        scenes = []
        for i, t in enumerate([3, 7, 1, 4]):
            fr = np.zeros((t, 8, 8, 3), np.uint8)
            fr[:, :2, :2, 0] = (10 * i + np.arange(t))[:, None, None]
            scenes.append(dict(frames=fr, cameras=np.full((t, 7), i, np.float32)))
        out = list(tr("train", iter(scenes)))
        assert [len(o["codes"]) for o in out] == [3, 7, 1, 4]
        for i, o in enumerate(out):
            assert (o["cameras"] == i).all() and np.array_equal(o["codes"][:, 0, 0], 10 * i + np.arange(len(o["codes"])))
        assert cb.calls == [5, 5, 5]                          # 15 frames encoded in full batches of 5 across scene boundaries
    finally:
        L.resize_u8 = orig


def test_frames_feature_png_roundtrip(tmp_path):
    """The raw-dataset side of generate-codes (commands/generate_codes.py:58-66): scenes carry `frames` as a bytes_list of encoded
    images; decode_frames turns them into the uint8 [T,H,W,3] array LatentCodeTransformer takes (PNG is lossless, so exact)."""
    import io
    PIL = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(4)
    frames = rng.integers(0, 256, (3, 20, 24, 3), dtype=np.uint8)
    blobs = []
    for f in frames:
        buf = io.BytesIO()
        PIL.fromarray(f).save(buf, format="PNG")
        blobs.append(buf.getvalue())
    cams = rng.standard_normal((3, 7)).astype(np.float32)
    path = str(tmp_path / "raw.tfrecord")
    with D.TFRecordWriter(path) as w:
        w.write(D.encode_example(dict(frames=blobs, cameras=cams)))
    (rec,) = list(D.read_tfrecords(path, verify=True))
    ex = D.decode_example(rec)
    assert list(ex["frames"]) == blobs
    got = D.decode_frames(ex["frames"])
    assert got.dtype == np.uint8 and got.shape == frames.shape and np.array_equal(got, frames)
    assert np.array_equal(np.asarray(ex["cameras"], np.float32).reshape(3, 7), cams)
