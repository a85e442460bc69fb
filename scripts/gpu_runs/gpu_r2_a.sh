#!/bin/bash
# round-2 first GPU pass: existing parity tests, accumulation probe, bench (all precisions side by side + parity block)
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python -m viewformer_b200.build > gpurun_out/build.log 2>&1
echo "=== probe"; timeout 300 python scripts/acc_rounding_probe.py > gpurun_out/acc_probe.log 2>&1; echo "probe rc=$?"; cat gpurun_out/acc_probe.log
echo "=== bench"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_a.log 2> gpurun_out/bench_a.err; echo "bench rc=$?"; tail -c 6000 gpurun_out/bench_a.log; tail -5 gpurun_out/bench_a.err
echo "=== tests"; timeout 1500 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/tests_a.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/tests_a.log
