#!/bin/bash
set -u
mkdir -p gpurun_out
python -m viewformer_b200.build > gpurun_out/build.log 2>&1
echo "=== migt train test"; timeout 900 python -m pytest tests/test_train_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider -s -k migt > gpurun_out/tests_train_migt.log 2>&1; echo "rc=$?"; grep -aE "^\[|passed|failed|^E |Error" gpurun_out/tests_train_migt.log | tail -20; tail -25 gpurun_out/tests_train_migt.log | cut -c1-300
