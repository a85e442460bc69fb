"""cta_group::2 paths against an fp64 reference — VF_TC_2CTA=1: the 128x128 tcgen05 kernel on CTA pairs (M = 256 per MMA, the B tile
split across the pair): an un-batched bf16 GEMM the wide kernel does not take (N = 192) and a 3x3 conv on a 16x16 map; VF_TC_WIDE2=1: the
wide GEMM on CTA pairs (256 features x 256 rows per pair).  Prints one line per case; tests/test_kernels_gpu.py runs it in a subprocess
with the flags on and off and compares."""
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
# un-batched linear the wide GEMM takes (N % 256 == 0): VF_TC_WIDE2=1 runs it on CTA pairs (256 features x 256 rows per pair)
M2, N2, K2 = 1280, 768, 768
a2 = (torch.randn(M2, K2, generator=g) * 0.5).bfloat16().cuda()
b2 = (torch.randn(N2, K2, generator=g) * 0.5).bfloat16().cuda()
bias2 = torch.randn(N2, generator=g).cuda()
res2 = torch.randn(M2, N2, generator=g).cuda()
out2 = torch.empty(M2, N2, device="cuda")
L.tc_gemm(a2, b2, out2, M=M2, N=N2, K=K2, lda=K2, ldb=K2, ldc=N2, bias=bias2, bias_mode=L.BIAS_N, residual=res2)
want2 = a2.double().cpu() @ b2.double().cpu().t() + bias2.double().cpu() + res2.double().cpu()
print(f"wide max_err {float((out2.double().cpu() - want2).abs().max()):.3e} checksum {float(out2.double().sum()):.9e}")
torch.cuda.synchronize()
