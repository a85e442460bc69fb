#!/bin/bash
# attention v2 (single pass, P through TMEM): correctness in both P modes, microbench
set -u
mkdir -p gpurun_out
python -m viewformer_b200.build > gpurun_out/build.log 2>&1
for mode in 0 1; do
  echo "=== attention tests VF_ATTN_PSMEM=$mode"
  VF_ATTN_PSMEM=$mode timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider -s -k "attention" > gpurun_out/tests_attn_$mode.log 2>&1; echo "rc=$?"
  grep -aE "^\[fused|passed|failed|^E |Error|timeout" gpurun_out/tests_attn_$mode.log | cut -c1-220 | tail -16
  echo "=== bench VF_ATTN_PSMEM=$mode"
  VF_ATTN_PSMEM=$mode timeout 300 python scripts/bench_attn.py 2>&1 | tail -4
done
