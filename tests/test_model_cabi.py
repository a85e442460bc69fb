"""Model-level C-ABI (include/vf_b200_model.h -> libvf_b200_model.so): symbol check on the CPU, and on the GPU a C program with no Python
of its own (tests/host/cabi_host.c, gcc) that creates both models, encodes / decodes / runs the transformer / the KV-cache query through
the library — its outputs must equal what the Python classes give for the same configuration and seed."""
import ctypes
import os
import re
import shutil
import struct
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "viewformer_b200", "libvf_b200_model.so")


def _ensure_lib():
    from viewformer_b200 import build
    return build.build_model_abi(verbose=False)


def test_model_abi_exports_every_declared_symbol():
    lib = ctypes.CDLL(_ensure_lib())
    hdr = open(os.path.join(ROOT, "include", "vf_b200_model.h")).read()
    names = set(re.findall(r"\b(vf_[a-z_]+)\s*\(", hdr))
    assert len(names) >= 13
    for n in sorted(names):
        assert hasattr(lib, n), f"{n} declared in vf_b200_model.h but not exported"


@pytest.mark.gpu
def test_c_host_drives_models_through_the_model_abi(tmp_path):
    from oracle import synth
    from viewformer_b200 import VQGAN, MIGT
    _ensure_lib()
    gcc = shutil.which("gcc")
    cuda = "/usr/local/cuda"
    exe = str(tmp_path / "cabi_host")
    r = subprocess.run([gcc, "-O1", "-o", exe, os.path.join(ROOT, "tests", "host", "cabi_host.c"), "-I", os.path.join(ROOT, "include"),
                        "-I", os.path.join(cuda, "include"), "-L", os.path.join(ROOT, "viewformer_b200"), "-lvf_b200_model",
                        "-L", os.path.join(cuda, "lib64"), "-lcudart", f"-Wl,-rpath,{os.path.join(ROOT, 'viewformer_b200')}",
                        f"-Wl,-rpath,{os.path.join(cuda, 'lib64')}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    n, B, T = 5, 2, 4
    imgs = synth.make_images_uint8(1, n, size=32, seed=21)[0].contiguous()
    ids = synth.make_codes(B, T, n_embed=64, side=8, seed=22).to(torch.int32)
    ids[:, -1] = 64                                              # mask token in the view to generate
    from oracle import migt_oracle as mo
    poses = mo.normalize_cameras(mo.to_relative_cameras(synth.make_cameras(B, T, seed=23))[0]).float().contiguous()
    scenes = synth.make_images_uint8(B, T, size=32, seed=24).contiguous()
    world_cams = synth.make_cameras(B, T, seed=25).float().contiguous()
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        f.write(struct.pack("<3i", n, B, T))
        f.write(imgs.numpy().tobytes()); f.write(ids.numpy().tobytes()); f.write(poses.numpy().tobytes())
        f.write(scenes.numpy().tobytes()); f.write(world_cams.numpy().tobytes())
    env = dict(os.environ, VF_PYTHON_EXECUTABLE=sys.executable, VF_B200_ROOT=ROOT)
    r = subprocess.run([exe, fin, fout], capture_output=True, text=True, env=env, timeout=900)
    print(r.stdout[-500:], r.stderr[-1500:])
    assert r.returncode == 0, r.stderr[-3000:]
    raw = open(fout, "rb").read()
    o = 0
    codes = np.frombuffer(raw, np.int64, n * 64, o).reshape(n, 8, 8); o += n * 64 * 8
    dec = np.frombuffer(raw, np.uint8, n * 32 * 32 * 3, o).reshape(n, 32, 32, 3); o += n * 32 * 32 * 3
    last = np.frombuffer(raw, np.int64, B * 64, o).reshape(B, 8, 8); o += B * 64 * 8
    qcodes = np.frombuffer(raw, np.int64, B * 64, o).reshape(B, 8, 8); o += B * 64 * 8
    gen = np.frombuffer(raw, np.uint8, B * 32 * 32 * 3, o).reshape(B, 32, 32, 3)
    # the same models through the Python classes
    vq = VQGAN(precision="fp32", ch=32, ch_mult=[1, 2, 2], attn_resolutions=[8], image_size=32, embed_dim=16, z_channels=16, n_embed=64).init_weights(3)
    tr = MIGT(precision="fp32", n_layer=2, n_head=4, d_model=256, sequence_size=4, n_loss_skip=1, n_embeddings=64, token_image_size=8,
              localization_weight="0").init_weights(4)
    want_codes = vq.encode_u8(imgs.cuda()).cpu().numpy()
    assert np.array_equal(codes, want_codes)
    assert np.array_equal(dec, vq.decode_code_u8(torch.from_numpy(want_codes).cuda()).cpu().numpy())
    want_last = tr.generate_codes(ids[:, :-1].cuda(), poses.cuda()).cpu().numpy()
    assert np.array_equal(last, want_last)
    cache = tr.prefill_context(ids[:, :-1].cuda(), poses[:, :-1].cuda())
    assert np.array_equal(qcodes, tr.query(cache, poses[:, -1].cuda()).cpu().numpy())
    from viewformer_b200 import generate_batch_predictions
    want_gen = generate_batch_predictions(tr, vq, scenes, world_cams)["generated_images"]
    assert np.array_equal(gen, torch.as_tensor(want_gen).cpu().numpy())
