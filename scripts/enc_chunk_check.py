import os, sys, torch
sys.path.insert(0, os.getcwd())
from viewformer_b200 import VQGAN
from viewformer_b200.config import VQGANConfig
cb = VQGAN(VQGANConfig(), precision="bf16").init_weights(0)
g = torch.Generator().manual_seed(1)
img = torch.randint(0, 256, (90, 128, 128, 3), generator=g, dtype=torch.uint8).cuda()
cb.encoder_chunk = 0
c0 = cb.encode_u8(img)
for ch in (18, 36):
    cb.encoder_chunk = ch
    c1 = cb.encode_u8(img)
    print(f"chunk {ch}: codes equal to unchunked: {bool(torch.equal(c0, c1))} ({int((c0 != c1).sum())} differ of {c0.numel()})")
