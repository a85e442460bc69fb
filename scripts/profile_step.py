"""One generate() step between cudaProfilerStart/Stop (run under `ncu --profile-from-start off`)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scenes", type=int, default=32)
ap.add_argument("--precision", default="mixed")
ap.add_argument("--part", default="all", choices=["all", "encode", "migt", "decode"])
a = ap.parse_args()

from bench import synth_inputs, model_pair  # noqa: E402
from viewformer_b200 import generate_batch_predictions  # noqa: E402
from viewformer_b200.config import VQGANConfig, MIGTConfig  # noqa: E402

dev = torch.device("cuda", 0)
cb, tr = model_pair(a.precision, VQGANConfig(), MIGTConfig(localization_weight="0"), dev)
images, cams = synth_inputs(a.scenes, 1234)
images, cams = images.to(dev), cams.to(dev)


def step():
    if a.part == "all":
        return generate_batch_predictions(tr, cb, images, cams)
    if a.part == "encode":
        return cb.encode_u8(images[:, :9].reshape(-1, 128, 128, 3).contiguous())
    if a.part == "migt":
        codes = torch.randint(0, 1024, (a.scenes, 9, 8, 8), device=dev)
        return tr.generate_codes(codes, cams)
    codes = torch.randint(0, 1024, (a.scenes, 8, 8), device=dev)
    return cb.decode_code_u8(codes)


step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
