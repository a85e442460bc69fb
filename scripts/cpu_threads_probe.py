import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import synth, vqgan_oracle as vo, migt_oracle as mo
from viewformer_b200.config import VQGANConfig, MIGTConfig
vcfg, tcfg = VQGANConfig(), MIGTConfig(localization_weight="0")
vsd, tsd = synth.make_vqgan_state_dict(vcfg, 0), synth.make_migt_state_dict(tcfg, 0)
x = torch.rand(10, 3, 128, 128) * 2 - 1
codes = synth.make_codes(1, 10)
cams = synth.make_cameras(1, 10)
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for nt in (8, 16, 32, 64, 128):
    torch.set_num_threads(nt)
    with torch.no_grad():
        vo.encode(vsd, vcfg, x[:2])
        t0 = time.perf_counter(); vo.encode(vsd, vcfg, x); t1 = time.perf_counter()
        mo.forward(tsd, tcfg, dict(input_ids=codes, poses=cams), use_localization=False); t2 = time.perf_counter()
        vo.decode_code(vsd, vcfg, codes[:, 0]); t3 = time.perf_counter()
    print(f"threads={nt:4d} encode10={t1-t0:.2f}s migt={t2-t1:.2f}s decode1={t3-t2:.2f}s")
