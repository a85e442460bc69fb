#!/bin/bash
set -u
mkdir -p gpurun_out
python -m viewformer_b200.build > gpurun_out/build.log 2>&1
echo "=== training tests"
timeout 1200 python -m pytest tests/test_train_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider -s > gpurun_out/tests_s.log 2>&1; echo "rc=$?"
grep -aE "^\[|passed|failed|^E |Error|timeout" gpurun_out/tests_s.log | cut -c1-260 | tail -30
