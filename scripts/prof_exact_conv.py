"""The exact (split-fp16) wide conv at the encoder's dominant shape: 288 images, 128 -> 128 channels @ 128x128 (for ncu captures and timing)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from viewformer_b200 import _lib as L

L.load(True)
dev = "cuda"
n = int(sys.argv[1]) if len(sys.argv) > 1 else 288
x = torch.randn((n, 128, 128, 128), device=dev)
xs = L.groupnorm(x, None, None, swish=False, out_dtype=torch.float16, normalize=False)
w = torch.randn((128 * 9, 128), device=dev) / 34.0
ws = L.split_f16x2(w).reshape(128, 18 * 128)
b = torch.zeros(128, device=dev)
o = torch.empty((n, 128, 128, 128), device=dev)
res = torch.randn((n, 128, 128, 128), device=dev)
for r in (None, res):
    for _ in range(3):
        L.tc_conv(xs, ws, b, out=o, residual=r, gn_groups=32)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        L.tc_conv(xs, ws, b, out=o, residual=r, gn_groups=32)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    fl = 2.0 * n * 128 * 128 * 128 * 9 * 128
    print(f"exact wide conv n={n} res={r is not None}: {ms:.3f} ms  algorithmic {fl / ms / 1e9:.1f} TFLOP/s  executed fp16 MMA {3 * fl / ms / 1e9:.1f} TFLOP/s")
