#!/bin/bash
set -u
mkdir -p gpurun_out
python -m viewformer_b200.build > gpurun_out/build.log 2>&1
echo "=== model C-ABI + full-size training tests"
timeout 1800 python -m pytest tests/test_model_cabi.py tests/test_train_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider -s -k "c_host or full_size" > gpurun_out/tests_w.log 2>&1; echo "rc=$?"
grep -aE "^\[|cabi_host|passed|failed|^E |Error|timeout" gpurun_out/tests_w.log | cut -c1-300 | tail -20
echo "=== training step launch list (32 images, fp32)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_train.csv python scripts/profile_train_step.py > gpurun_out/prof_train.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_train.csv > gpurun_out/launches_train_summary.md 2>&1; head -18 gpurun_out/launches_train_summary.md
