#!/bin/bash
set -u
mkdir -p gpurun_out
python -m viewformer_b200.build > gpurun_out/build.log 2>&1
echo "=== train test"; timeout 900 python -m pytest tests/test_train_gpu.py -m gpu -q -x --tb=long -p no:cacheprovider -s > gpurun_out/tests_train.log 2>&1; echo "rc=$?"; grep -aE "^\[|passed|failed|^E |Error" gpurun_out/tests_train.log | tail -30; tail -30 gpurun_out/tests_train.log | cut -c1-250
echo "=== dp check (1 GPU)"; timeout 900 python scripts/train_dp_check.py 2>&1 | tail -5
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/vq_launches.csv python scripts/bench_vq.py fused > gpurun_out/vq_prof.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vq_rescue -s 2 -c 1 -o gpurun_out/prof_vq_rescue -f python scripts/bench_vq.py fused > gpurun_out/prof_vq_rescue.log 2>&1
