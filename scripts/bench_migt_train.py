"""Full-size transformer training step timing (MIGTConfig defaults: 12 layers, d = 768; B scenes x 20 views x 64 tokens)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import synth, migt_oracle as mo
from viewformer_b200 import MIGT
from viewformer_b200.config import MIGTConfig
from viewformer_b200.train_migt import MIGTTrainer

B, T = int(os.environ.get("VF_B", "4")), 20
cfg = MIGTConfig()
model = MIGT(cfg, precision="fp32").init_weights(0)
tr = MIGTTrainer(model)
codes = synth.make_codes(B, T, n_embed=cfg.n_embeddings, seed=1)
cams = mo.normalize_cameras(mo.to_relative_cameras(synth.make_cameras(B, T, seed=2))[0])
for _ in range(2):
    tr.forward_backward(cams, codes); tr.optimizer_step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
n = 3
for _ in range(n):
    loss = tr.forward_backward(cams, codes); tr.optimizer_step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"[migt train step, full size] B={B} T={T}: {ms:.1f} ms/step -> {B * T * 64 / ms * 1e3:.0f} tokens/s; loss {float(loss):.4f}")
