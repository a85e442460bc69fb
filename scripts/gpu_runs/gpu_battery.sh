#!/bin/bash
# Runs on the GPU box (under gpurun): GPU parity tests, smoke, bench.  Logs -> gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python -m viewformer_b200.build > gpurun_out/build.log 2>&1
if [ "${RUN_TESTS:-1}" = "1" ]; then
echo "=== kernels" ; timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -rA --tb=short -p no:cacheprovider > gpurun_out/test_kernels.log 2>&1; echo "kernels rc=$?"
grep -E "passed|failed" gpurun_out/test_kernels.log | tail -2; grep -B2 -A12 "^E " gpurun_out/test_kernels.log | head -60
echo "=== models" ; timeout 1500 python -m pytest tests/test_models_gpu.py -m gpu -q -rA --tb=short -p no:cacheprovider > gpurun_out/test_models.log 2>&1; echo "models rc=$?"
grep -E "passed|failed" gpurun_out/test_models.log | tail -2; grep -B2 -A12 "^E " gpurun_out/test_models.log | head -60
echo "=== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
fi
if [ "${RUN_PHASES:-0}" = "1" ]; then
echo "=== tc phase times"; timeout 300 python scripts/tc_phase_times.py 2>&1 | tee gpurun_out/tc_phases.log
fi
if [ "${RUN_BENCH:-1}" = "1" ]; then
echo "=== bench"; timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -3 gpurun_out/bench.log
fi
