"""TFRecord / tf.train.Example codec, token-dataset loader and the generate-codes scene re-batching logic (CPU)."""
import os

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
    capped = list(D.load_token_dataset(str(tmp_path), 1, seq, 2, split="train", repeat=1, max_windows_per_environment=1))
    assert len(capped) == sum(1 for t in [9, 4, 13, 3, 8, 8, 8, 21] if t >= seq)
    # the reference's own knob: take(k) on a one-sample dataset (tfrecord_dataset.py:177-181) -> k >= 1 keeps everything, 0 nothing
    assert len(list(D.load_token_dataset(str(tmp_path), 1, seq, 2, split="train", repeat=1, max_samples_per_environment=1))) == n_windows
    assert len(list(D.load_token_dataset(str(tmp_path), 1, seq, 2, split="train", repeat=1, max_samples_per_environment=0))) == 0
    # the training transform (train_transformer.py:113: partial(process_batch, augment=config.augment_poses)) is applied per window
    import functools
    rel = list(D.load_token_dataset(str(tmp_path), 2, seq, 2, split="train", repeat=1, seed=3, transform=functools.partial(D.process_batch, augment="relative")))
    assert len(rel) == len(batches)
    for p, t in rel:
        assert p.dtype == torch.float32 and p.shape == (2, seq, 7) and t.dtype == torch.int64
        assert float(p[:, 0, :3].abs().max()) < 1e-5 and torch.allclose(p[:, 0, 3:], torch.tensor([1.0, 0, 0, 0]).expand(2, 4), atol=1e-5)   # first view = identity
        assert float((p[..., 3:].norm(dim=-1) - 1).abs().max()) < 1e-5 and bool((p[..., 3] >= 0).all())
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


def _load_reference_generate_codes():
    """commands/generate_codes.py loaded as shipped, with stand-ins for what is not on this path: `webdataset.filters` (three generators
    restating the documented behaviour of map_ / unbatched_ / batched_ with the default collation), aparse.click, viewformer.data
    .transform_dataset; TensorFlow is made un-importable for the duration so that the script's own `except ImportError` branch runs."""
    import importlib.util
    import sys
    import types
    from oracle import ref_loader
    ref_loader.load_reference_modules()
    import typing
    ap = sys.modules["aparse"]
    if not hasattr(ap, "click"):                          # the same stand-ins oracle/ref_loader.py::load_reference_evaluate installs
        ap.click = types.SimpleNamespace(command=lambda *a, **k: (lambda f: f))
    if not hasattr(ap, "ConditionalType"):
        ap.ConditionalType = lambda name, table, default=None: typing.Any
    utils = sys.modules["viewformer.utils"]
    if not hasattr(utils, "SplitIndices"):
        utils.SplitIndices = object
    data = sys.modules.get("viewformer.data")
    if data is None:
        data = types.ModuleType("viewformer.data")
        data.__path__ = []
        sys.modules["viewformer.data"] = data
    data.transform_dataset = lambda *a, **k: None

    def collate(samples):
        cols = list(zip(*samples))
        out = []
        for c in cols:
            if isinstance(c[0], (int, float)):
                out.append(np.array(list(c)))
            elif isinstance(c[0], torch.Tensor):
                out.append(torch.stack(list(c)))
            elif isinstance(c[0], np.ndarray):
                out.append(np.array(list(c)))
            else:
                out.append(list(c))
        return out

    def map_(data, f):
        for s in data:
            yield f(s)

    def unbatched_(data):
        for s in data:
            for i in range(len(s[0])):
                yield tuple(x[i] for x in s)

    def batched_(data, batchsize=20, partial=True):
        batch = []
        for s in data:
            if len(batch) >= batchsize:
                yield collate(batch)
                batch = []
            batch.append(s)
        if batch and (len(batch) == batchsize or partial):
            yield collate(batch)

    wds = types.ModuleType("webdataset")
    wds.filters = types.SimpleNamespace(map_=map_, unbatched_=unbatched_, batched_=batched_)
    saved_wds, saved_tf = sys.modules.get("webdataset"), sys.modules.pop("tensorflow", None)
    sys.modules["webdataset"] = wds
    sys.modules["tensorflow"] = None                      # `import tensorflow` -> ImportError, as on a torch-only machine
    try:
        path = os.path.join(ref_loader.REFERENCE_ROOT, "viewformer", "commands", "generate_codes.py")
        spec = importlib.util.spec_from_file_location("viewformer_commands_generate_codes_ref", path)
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
    finally:
        if saved_tf is not None:
            sys.modules["tensorflow"] = saved_tf
        else:
            sys.modules.pop("tensorflow", None)
    return m, (lambda: sys.modules.__setitem__("webdataset", saved_wds) if saved_wds is not None else sys.modules.pop("webdataset", None))


def test_latent_code_transformer_equals_the_reference_class():
    """commands/generate_codes.py:20-78 (the reference's LatentCodeTransformer, run as shipped with a stand-in for webdataset's three
    filters) against data.LatentCodeTransformer on the same scenes with the same stand-in codebook: same scenes out, same codes, same
    cameras, same encoder batch sizes.  Equal-length scenes: with ragged scenes the reference mis-assigns the carried-over frames (it
    assumes they belong to a scene as long as the next batch's first one, generate_codes.py:63) — this implementation keeps a queue."""
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference sources not present (GPU box)")
    ref_mod, restore = _load_reference_generate_codes()

    class Cfg:
        image_size, stride, batch_size = 8, 4, 5

    class RefCodebook(torch.nn.Module):
        config, calls = Cfg, []

        def encode(self, x):                                 # NCHW float in [-1, 1], as _convert_image_type hands it over
            self.calls.append(len(x))
            px = ((x[:, 0, :2, :2] + 1) / 2 * 255).round().to(torch.int64)
            return None, None, px

    class OurCodebook:
        config, device, calls = Cfg, "cpu", []

        def encode_u8(self, x):                              # NHWC uint8
            self.calls.append(len(x))
            return x[:, :2, :2, 0].to(torch.int64)

    scenes = []
    for i in range(4):
        fr = np.zeros((6, 8, 8, 3), np.uint8)
        fr[:, :2, :2, 0] = (10 * i + np.arange(6))[:, None, None]
        scenes.append(dict(frames=fr, cameras=np.full((6, 7), i, np.float32)))
    import viewformer_b200._lib as L
    orig = L.resize_u8
    L.resize_u8 = lambda x, size, method=None: x            # frames already have the codebook's size; no device in this test
    try:
        rcb, ocb = RefCodebook(), OurCodebook()
        ref_tr = ref_mod.LatentCodeTransformer(rcb, batch_size=5, device="cpu")
        our_tr = D.LatentCodeTransformer(ocb, batch_size=5)
        assert ref_tr.update_dataset_info({}) == our_tr.update_dataset_info({}) == {"token_image_size": 2}
        for feats in (None, ["frames", "cameras"], ["frames", "cameras-gqn"]):
            assert ref_tr.output_features(feats) == our_tr.output_features(feats)
        want = list(ref_tr("train", iter([dict(s) for s in scenes])))
        got = list(our_tr("train", iter([dict(s) for s in scenes])))
    finally:
        L.resize_u8 = orig
        restore()
    assert len(want) == len(got) == 4 and rcb.calls == ocb.calls == [5, 5, 5, 5, 4]
    for w, g in zip(want, got):
        assert np.array_equal(np.asarray(w["codes"]), np.asarray(g["codes"])) and np.array_equal(np.asarray(w["cameras"]), np.asarray(g["cameras"]))
