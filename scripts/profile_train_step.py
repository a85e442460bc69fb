"""One full-size codebook training step (32 images, BASELINE configs[3] per-GPU shape) between cudaProfilerStart/Stop
(run under `ncu --profile-from-start off`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from viewformer_b200 import VQGAN
from viewformer_b200.config import VQGANConfig
from viewformer_b200.train import VQGANTrainer

n = int(os.environ.get("VF_TRAIN_IMAGES", "32"))
cfg = VQGANConfig(perceptual_weight=0.0)
tr = VQGANTrainer(VQGAN(cfg, precision="fp32", device="cuda:0").init_weights(0))
x = torch.rand((n, 3, 128, 128), generator=torch.Generator().manual_seed(0)) * 2 - 1
tr.training_step(x)
torch.cuda.synchronize()
torch.cuda.profiler.start()
tr.training_step(x)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
