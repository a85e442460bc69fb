#!/bin/bash
# fused VQ (fixed-point keys, set-restricted exact pass) + attention v2 (persistent): tests + microbench
set -u
mkdir -p gpurun_out
python -m viewformer_b200.build > gpurun_out/build.log 2>&1
echo "=== kernel tests (vq, attention)"
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider -s -k "attention or vq or lookup or smoke" > gpurun_out/tests_q.log 2>&1; echo "rc=$?"
grep -aE "^\[|passed|failed|^E |Error|timeout" gpurun_out/tests_q.log | cut -c1-220 | tail -30
echo "=== bench vq"; timeout 300 python scripts/bench_vq.py fused 2>&1 | tail -2
echo "=== bench attn"; timeout 300 python scripts/bench_attn.py 2>&1 | tail -4
