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
