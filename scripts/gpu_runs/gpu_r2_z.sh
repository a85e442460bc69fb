#!/bin/bash
set -u
mkdir -p gpurun_out
python -m viewformer_b200.build > gpurun_out/build.log 2>&1
echo "=== wgrad tc kernel test"
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider -s -k "weight_gradient" > gpurun_out/tests_z1.log 2>&1; echo "rc=$?"
grep -aE "^\[conv_wgrad|passed|failed|^E |Error|timeout|vf_" gpurun_out/tests_z1.log | cut -c1-260 | tail -12
echo "=== training tests"
timeout 1800 python -m pytest tests/test_train_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider -s > gpurun_out/tests_z.log 2>&1; echo "rc=$?"
grep -aE "^\[full|^\[train|passed|failed|^E |Error|timeout" gpurun_out/tests_z.log | cut -c1-260 | tail -12
echo "=== full-size step timing"
timeout 600 python scripts/train_dp_check.py 2>&1 | grep -aE "^\[" | tail -2
echo "=== training step launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_train.csv python scripts/profile_train_step.py > gpurun_out/prof_train.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_train.csv > gpurun_out/launches_train_summary.md 2>&1; head -16 gpurun_out/launches_train_summary.md
