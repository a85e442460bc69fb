#!/bin/bash
set -u
mkdir -p gpurun_out
python -m viewformer_b200.build > gpurun_out/build.log 2>&1
echo "=== eval tests"; timeout 600 python -m pytest tests/test_eval_gpu.py -m gpu -q --tb=short -p no:cacheprovider -s > gpurun_out/tests_eval.log 2>&1; echo "rc=$?"; grep -aE "^\[|passed|failed|^E " gpurun_out/tests_eval.log | tail -30
echo "=== fused vq tests"; timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider -k "fused and vq" > gpurun_out/tests_vq.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/tests_vq.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/vq_launches.csv python scripts/bench_vq.py fused > gpurun_out/vq_prof.log 2>&1
grep -a "vq_\|rescue" gpurun_out/vq_launches.csv | awk -F'","' '{print $5, $(NF)}' | tail -2
echo "=== bench vq"; timeout 300 python scripts/bench_vq.py fused 2>&1 | tee gpurun_out/bench_vq.log | grep fused
