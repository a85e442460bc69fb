"""Fused block-causal attention microbench: the MIGT shapes of the bench step (B=32, 10 views) and of BASELINE config 5 (20 views, full
forward and the KV-cache decode call).  Prints launch time, useful tensor FLOP/s (one Q K^T + one P V over the visible key tiles) and
the share of the measured dense bf16 peak."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from viewformer_b200 import _lib as L


def visible_tile_flops(S, blk, first=0):
    # per (batch, head): every 128-query tile times the keys its last view sees, 2 GEMMs of 2*128*keys*64 flops
    fl = 0
    for q0 in range((first // 128) * 128, S, 128):
        last = min(q0 + 128, S) - 1
        keys = min(S, (last // blk + 1) * blk)
        fl += 2 * 2 * 128 * keys * 64
    return fl


def timeit(fn, iters=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--once", action="store_true", help="one launch per shape (for ncu)")
    a = ap.parse_args()
    peak = 1691.2
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["bf16_tflops"])
    except Exception:
        pass
    H, d, blk = 12, 768, 64
    for name, B, T, first, skip in [("bench step B32 T10", 32, 10, 0, -1), ("config 5 forward B16 T20", 16, 20, 0, -1),
                                    ("config 5 decode B128, 19 ctx + query sharing a tile", 128, 20, 19 * 64, -1),
                                    ("config 5 decode B128, 19 ctx + empty slot + query (MIGT layout)", 128, 21, 20 * 64, 19)]:
        S = T * blk
        g = torch.Generator().manual_seed(S)
        qk = (torch.randn(B, S, 2 * d, generator=g) * 0.6).bfloat16().cuda()
        vt = torch.randn(B, d, S, generator=g).bfloat16().cuda()
        out = torch.empty(B * S, d, dtype=torch.bfloat16, device="cuda")
        fn = lambda: L.attn_block_causal(qk, vt, B, S, H, d, blk, first_query=first, out=out, skip_view=skip)
        if a.once:
            fn(); torch.cuda.synchronize(); continue
        ms = timeit(fn)
        fl = visible_tile_flops(S, blk, first) * B * H
        if skip >= 0:            # one 64-query view against 20 visible key tiles (the skipped slot is not work)
            fl = 2 * 2 * 64 * (19 + 1) * 64 * 64 * B * H
        print(f"[attn] {name}: {ms * 1e3:.1f} us  useful {fl / ms / 1e9:.1f} TFLOP/s = {fl / ms / 1e9 / peak * 100:.1f}% of the measured bf16 peak ({peak:.0f})")


if __name__ == "__main__":
    main()
