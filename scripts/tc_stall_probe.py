"""Where does the MMA issuer wait?  clock64 counters around its mbarrier waits (profiling aid; VF_TC_2CTA / VF_TC_HALO select the mode)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from viewformer_b200 import _lib as L
lib = L.load(True)
def run(name, fn):
    for _ in range(2): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"    [{name}] {e0.elapsed_time(e1) / 5 * 1e3:.1f} us per launch (CUDA events, 5 launches)")
    buf = torch.zeros((148, 8), dtype=torch.int64, device="cuda")
    lib.vf_tc_debug_counters(ctypes.c_void_p(buf.data_ptr()))
    torch.cuda.synchronize(); fn(); torch.cuda.synchronize()
    lib.vf_tc_debug_counters(ctypes.c_void_p(0))
    t = buf.double().cpu(); t = t[t[:, 3] > 0]
    tot, ops, tm, tiles = t[:, 0].mean().item(), t[:, 1].mean().item(), t[:, 2].mean().item(), t[:, 3].mean().item()
    full = buf.double().cpu()
    prod = full[full[:, 4] > 0]
    pw = 100 * (prod[:, 5] / prod[:, 4])
    print(f"    producers: {prod.shape[0]} CTAs, waiting on empty barriers {pw.mean().item():5.1f}% of their time (min {pw.min().item():.1f} max {pw.max().item():.1f})")
    print(f"{name:40s} issuers {t.shape[0]:3d} tiles/issuer {tiles:6.1f} | cycles/tile {tot/tiles:8.0f} | wait operands {100*ops/tot:5.1f}% | wait tmem_empty {100*tm/tot:5.1f}% | issuing {100*(tot-ops-tm)/tot:5.1f}%")
flags = int(os.environ.get("VF_TC_DBG_FLAGS", "0"))
lib.vf_tc_debug_flags(flags)
print(f"debug flags = {flags} (1 no epilogue, 2 no B loads, 4 no A loads)")
n, hw, c = 288, 128, 128
x = torch.randn((n, hw, hw, c), device="cuda").bfloat16(); w = (torch.randn((c, 9 * c), device="cuda") / 30).bfloat16()
b = torch.zeros(c, device="cuda"); res = torch.randn((n, hw, hw, c), device="cuda"); out = torch.empty((n, hw, hw, c), device="cuda")
run("conv 128->128 @128^2 no residual", lambda: L.tc_conv(x, w, b, out=out))
run("conv 128->128 @128^2 + residual", lambda: L.tc_conv(x, w, b, out=out, residual=res))
mr = torch.zeros((n, 32, 2), device="cuda"); mr[..., 1] = 1.0
ga, be = torch.ones(c, device="cuda"), torch.zeros(c, device="cuda")
run("conv + residual + norm-on-load", lambda: L.tc_conv(x, w, b, out=out, residual=res, norm=(mr, ga, be, 32, True)))
run("conv + norm-on-load (no residual)", lambda: L.tc_conv(x, w, b, out=out, norm=(mr, ga, be, 32, True)))
run("conv + residual + norm-on-load, packed tanh swish", lambda: L.tc_conv(x, w, b, out=out, residual=res, norm=(mr, ga, be, 32, 2)))
run("conv + residual + norm-on-load, no swish", lambda: L.tc_conv(x, w, b, out=out, residual=res, norm=(mr, ga, be, 32, 0)))
A = torch.randn((20480, 3072), device="cuda").bfloat16(); B = torch.randn((768, 3072), device="cuda").bfloat16(); o2 = torch.empty((20480, 768), device="cuda")
run("gemm 20480x768x3072", lambda: L.tc_gemm(A, B, o2, M=20480, N=768, K=3072, lda=3072, ldb=3072, ldc=768))

bias768 = torch.zeros(768, device="cuda"); res768 = torch.randn((20480, 768), device="cuda")
run("fc2 + bias + residual", lambda: L.tc_gemm(A, B, o2, M=20480, N=768, K=3072, lda=3072, ldb=3072, ldc=768, bias=bias768, bias_mode=1, residual=res768))
X = torch.randn((20480, 768), device="cuda").bfloat16(); Wqk = torch.randn((1536, 768), device="cuda").bfloat16(); oqk = torch.empty((20480, 1536), device="cuda", dtype=torch.bfloat16)
run("qk -> bf16", lambda: L.tc_gemm(X, Wqk, oqk, M=20480, N=1536, K=768, lda=768, ldb=768, ldc=1536))
oqf = torch.empty((20480, 1536), device="cuda")
run("qk -> f32", lambda: L.tc_gemm(X, Wqk, oqf, M=20480, N=1536, K=768, lda=768, ldb=768, ldc=1536))
