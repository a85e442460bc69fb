"""How does tcgen05.mma accumulate?  Products of bf16 (8-bit significands) are exact in fp32, so every difference between the
TMEM fp32 accumulator and an fp64 sum of the same products is ACCUMULATION rounding.  Reports, per K, the mean signed relative
error (a bias toward zero = truncation) and the rms relative error of |y|, next to a sequential fp32 RN chain (CUDA-core path).

Also evaluates split-operand schemes fed as a K-concatenated bf16 GEMM (no kernel change):
    x3  : [a_h, a_h, a_l] . [w_h, w_l, w_h]                      (2-piece split, 3 products)
    x6  : [a_h, a_h, a_m, a_m, a_h, a_l] . [w_h, w_m, w_h, w_m, w_l, w_h]   (3-piece split, 6 products; 24 significand bits)
against the fp64 product of the ORIGINAL fp32 operands, and the fp32 CUDA-core GEMM for comparison.
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from viewformer_b200 import _lib as L

L.load(True)
dev = "cuda"
torch.manual_seed(0)


def tc(A, B):
    M, K = A.shape
    N = B.shape[0]
    out = torch.empty((M, N), device=dev)
    L.tc_gemm(A.contiguous(), B.contiguous(), out, M=M, N=N, K=K, lda=K, ldb=K, ldc=N)
    return out


def simt(A, B):
    M, K = A.shape
    N = B.shape[0]
    out = torch.empty((M, N), device=dev)
    L.simt_gemm(A.contiguous(), B.contiguous(), out, M=M, N=N, K=K, a_strides=(K, 1), b_strides=(1, K), ldc=N)
    return out


def stats(name, got, ref):
    ref = ref.double()
    scale = ref.abs().mean()
    err = (got.double() - ref)
    # bias toward zero: mean of err * sign(ref) relative to mean |ref|
    bias = float((err * torch.sign(ref)).mean() / scale)
    rms = float(err.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt())
    mx = float(err.abs().max() / scale)
    print(f"{name:58s} bias {bias:+.3e}   rms {rms:.3e}   max {mx:.3e}")


def split(x, pieces):
    out, r = [], x.clone()
    for _ in range(pieces):
        p = r.bfloat16()
        out.append(p)
        r = r - p.float()
    return out


M, N = 512, 256
for K in (64, 1152, 6912):
    A = torch.randn((M, K), device=dev).bfloat16()
    B = (torch.randn((N, K), device=dev) / K ** 0.5).bfloat16()
    ref = A.double() @ B.double().t()
    stats(f"bf16 exact products, tcgen05 f32 accumulate, K={K}", tc(A, B), ref)
    stats(f"   same operands as f32 on the CUDA-core FFMA path, K={K}", simt(A.float(), B.float()), ref)
    # positive-only operands: no cancellation, accumulator grows monotonically -> a truncating adder shows a clean negative bias
    Ap, Bp = A.abs(), B.abs()
    refp = Ap.double() @ Bp.double().t()
    stats(f"   all-positive operands, tcgen05, K={K}", tc(Ap, Bp), refp)
    stats(f"   all-positive operands, FFMA,    K={K}", simt(Ap.float(), Bp.float()), refp)

print()
K = 1152
A = torch.randn((M, K), device=dev)
B = torch.randn((N, K), device=dev) / K ** 0.5
ref = A.double() @ B.double().t()
stats("fp32 FFMA (simt_gemm)", simt(A, B), ref)
stats("tf32 tcgen05 (kind::tf32, fp32 operands)", tc(A, B), ref)
stats("bf16 tcgen05", tc(A.bfloat16(), B.bfloat16()), ref)
a2, b2 = split(A, 2), split(B, 2)
stats("bf16 x3 (K-concatenated)", tc(torch.cat([a2[0], a2[0], a2[1]], 1), torch.cat([b2[0], b2[1], b2[0]], 1)), ref)
# small terms FIRST so that they are not absorbed by a large accumulator
stats("bf16 x3, small terms first", tc(torch.cat([a2[1], a2[0], a2[0]], 1), torch.cat([b2[0], b2[1], b2[0]], 1)), ref)
a3, b3 = split(A, 3), split(B, 3)
ah, am, al = a3
bh, bm, bl = b3
stats("bf16 x6 (K-concatenated)", tc(torch.cat([ah, ah, am, am, ah, al], 1), torch.cat([bh, bm, bh, bm, bl, bh], 1)), ref)
stats("bf16 x6, small terms first", tc(torch.cat([al, ah, am, am, ah, ah], 1), torch.cat([bh, bl, bm, bh, bm, bh], 1)), ref)
# chunked: each product group accumulated on its own (separate launches), summed in fp32 RN on the CUDA cores
parts = [tc(ah, bh), tc(ah, bm), tc(am, bh), tc(am, bm), tc(ah, bl), tc(al, bh)]
s = parts[5] + parts[4]
s = s + parts[3]
s = s + (parts[2] + parts[1])
s = s + parts[0]
stats("bf16 x6, per-product accumulators summed in fp32 RN", s, ref)
# K-chunked hh term: 9 chunks of 128 (one per filter tap) each from a zero accumulator, summed RN
acc = torch.zeros((M, N), device=dev)
for c in range(0, K, 128):
    sl = slice(c, c + 128)
    t = tc(al[:, sl], bh[:, sl]) + tc(ah[:, sl], bl[:, sl])
    t = t + tc(am[:, sl], bm[:, sl])
    t = t + (tc(am[:, sl], bh[:, sl]) + tc(ah[:, sl], bm[:, sl]))
    t = t + tc(ah[:, sl], bh[:, sl])
    acc = acc + t
stats("bf16 x6, K chunks of 128 from zero accumulators, RN sums", acc, ref)
