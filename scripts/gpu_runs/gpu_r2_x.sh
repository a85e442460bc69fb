#!/bin/bash
set -u
mkdir -p gpurun_out
python -m viewformer_b200.build > gpurun_out/build.log 2>&1
echo "=== training tests (tensor-core convs in the trainer)"
timeout 1800 python -m pytest tests/test_train_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider -s > gpurun_out/tests_x.log 2>&1; echo "rc=$?"
grep -aE "^\[|passed|failed|^E |Error|timeout" gpurun_out/tests_x.log | cut -c1-300 | tail -24
echo "=== full-size step timing"
timeout 600 python scripts/train_dp_check.py 2>&1 | grep -aE "^\[" | tail -3
VF_TRAIN_TC=0 timeout 600 python scripts/train_dp_check.py 2>&1 | grep -aE "^\[train step" | tail -1
echo "=== training step launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_train.csv python scripts/profile_train_step.py > gpurun_out/prof_train.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_train.csv > gpurun_out/launches_train_summary.md 2>&1; head -14 gpurun_out/launches_train_summary.md
