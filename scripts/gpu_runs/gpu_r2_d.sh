#!/bin/bash
set -u
mkdir -p gpurun_out
python -m viewformer_b200.build > gpurun_out/build.log 2>&1
echo "=== fused vq tests"; timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider -k "fused and vq" -s > gpurun_out/tests_vq.log 2>&1; echo "rc=$?"; grep -aE "vq_lookup_fused|passed|failed|^E |rror|timeout" gpurun_out/tests_vq.log | tail -30
echo "=== bench vq"; timeout 300 python scripts/bench_vq.py 2>&1 | tee gpurun_out/bench_vq.log | tail -12
