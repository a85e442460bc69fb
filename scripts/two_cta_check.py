"""cta_group::2 path of the 128x128 tcgen05 kernel (VF_TC_2CTA=1: CTA pairs, M = 256 per MMA, the B tile split across the pair) against an
fp64 reference: an un-batched bf16 GEMM the wide kernel does not take (N = 192) and a 3x3 conv on a 16x16 map (not wide-eligible).
Prints one line per case; tests/test_kernels_gpu.py runs it in a subprocess with the flag on and off and compares."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from viewformer_b200 import _lib as L

g = torch.Generator().manual_seed(5)
M, N, K = 1024, 192, 512
a = (torch.randn(M, K, generator=g) * 0.5).bfloat16().cuda()
b = (torch.randn(N, K, generator=g) * 0.5).bfloat16().cuda()
bias = torch.randn(N, generator=g).cuda()
out = torch.empty(M, N, device="cuda")
L.tc_gemm(a, b, out, M=M, N=N, K=K, lda=K, ldb=K, ldc=N, bias=bias, bias_mode=L.BIAS_N)
want = a.double().cpu() @ b.double().cpu().t() + bias.double().cpu()
print(f"gemm max_err {float((out.double().cpu() - want).abs().max()):.3e} checksum {float(out.double().sum()):.9e}")

n, hw, cin, cout = 8, 16, 128, 256
x = torch.randn(n, hw, hw, cin, generator=g).bfloat16().cuda()
w = (torch.randn(cout, 9 * cin, generator=g) / (9 * cin) ** 0.5).bfloat16().cuda()
bc = torch.randn(cout, generator=g).cuda()
y = L.tc_conv(x, w, bc)
wt = w.float().reshape(cout, 3, 3, cin).permute(0, 3, 1, 2).double().cpu()
ref = torch.nn.functional.conv2d(x.double().cpu().permute(0, 3, 1, 2), wt, bc.double().cpu(), padding=1).permute(0, 2, 3, 1)
print(f"conv max_err {float((y.double().cpu() - ref).abs().max()):.3e} checksum {float(y.double().sum()):.9e}")
torch.cuda.synchronize()
