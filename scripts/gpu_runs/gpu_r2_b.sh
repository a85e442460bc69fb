#!/bin/bash
# round-2 pass b: exact tensor-core conv tests, BASELINE-config parity tests, bench
set -u
mkdir -p gpurun_out
python -m viewformer_b200.build > gpurun_out/build.log 2>&1
echo "=== exact kernel tests"; timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider -k "exact or split" -s > gpurun_out/tests_exact.log 2>&1; echo "rc=$?"; grep -E "^\[|passed|failed|Error|error" gpurun_out/tests_exact.log | tail -40
echo "=== baseline config tests"; timeout 900 python -m pytest tests/test_baseline_configs_gpu.py -m gpu -q --tb=short -p no:cacheprovider -s > gpurun_out/tests_baseline.log 2>&1; echo "rc=$?"; grep -E "^\[|passed|failed|^E " gpurun_out/tests_baseline.log | tail -40
echo "=== bench"; timeout 900 python bench.py --steps 5 --warmup 3 --also bf16 > gpurun_out/bench_b.log 2> gpurun_out/bench_b.err; echo "bench rc=$?"; tail -c 5000 gpurun_out/bench_b.log; tail -5 gpurun_out/bench_b.err
