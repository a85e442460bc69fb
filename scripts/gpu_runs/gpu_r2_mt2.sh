#!/bin/bash
set -u
mkdir -p gpurun_out
python -m viewformer_b200.build > gpurun_out/build.log 2>&1
echo "=== migt training tests"
timeout 1500 python -m pytest tests/test_train_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider -s -k "migt" > gpurun_out/tests_mt.log 2>&1; echo "rc=$?"
grep -aE "^\[migt|passed|failed|^E |Error|timeout|vf_" gpurun_out/tests_mt.log | cut -c1-260 | tail -14
timeout 900 python scripts/bench_migt_train.py 2>&1 | tail -1
VF_TRAIN_TC=0 timeout 900 python scripts/bench_migt_train.py 2>&1 | tail -1
