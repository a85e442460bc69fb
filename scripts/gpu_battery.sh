#!/bin/bash
# Runs on the GPU box (under gpurun): GPU parity tests, smoke, bench.  Logs -> gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python -m viewformer_b200.build > gpurun_out/build.log 2>&1
echo "=== kernels" ; timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -rA --tb=short -p no:cacheprovider > gpurun_out/test_kernels.log 2>&1; echo "kernels rc=$?"
tail -5 gpurun_out/test_kernels.log
echo "=== models" ; timeout 1500 python -m pytest tests/test_models_gpu.py -m gpu -q -rA --tb=short -p no:cacheprovider > gpurun_out/test_models.log 2>&1; echo "models rc=$?"
tail -5 gpurun_out/test_models.log
echo "=== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
if [ "${RUN_BENCH:-1}" = "1" ]; then
echo "=== bench"; timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -3 gpurun_out/bench.log
fi
