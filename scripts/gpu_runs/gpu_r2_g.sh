#!/bin/bash
set -u
mkdir -p gpurun_out
python -m viewformer_b200.build > gpurun_out/build.log 2>&1
echo "=== fused vq tests"; timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider -k "fused and vq" -s > gpurun_out/tests_vq.log 2>&1; echo "rc=$?"; grep -aE "vq_lookup_fused|passed|failed|^E |rror|timeout" gpurun_out/tests_vq.log | tail -30
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/vq_launches.csv python scripts/bench_vq.py fused > gpurun_out/vq_prof.log 2>&1
grep -a "vq_\|rescue" gpurun_out/vq_launches.csv | awk -F'","' '{print $5, $(NF)}' | tail -2
echo "=== bench vq"; timeout 300 python scripts/bench_vq.py 2>&1 | tee gpurun_out/bench_vq.log | grep fused
echo "=== all tests"; timeout 1500 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/tests_g.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/tests_g.log
