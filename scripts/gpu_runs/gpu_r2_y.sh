#!/bin/bash
set -u
mkdir -p gpurun_out
python -m viewformer_b200.build > gpurun_out/build.log 2>&1
echo "=== multi-end attention tests"
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider -s -k "attention" > gpurun_out/tests_y1.log 2>&1; echo "rc=$?"
grep -aE "^\[fused|passed|failed|^E |Error|timeout" gpurun_out/tests_y1.log | cut -c1-220 | tail -24
echo "=== all GPU tests"
timeout 2400 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/tests_y.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/tests_y.log | cut -c1-300
