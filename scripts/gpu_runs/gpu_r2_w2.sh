#!/bin/bash
set -u
mkdir -p gpurun_out
python -m viewformer_b200.build > gpurun_out/build.log 2>&1
echo "=== kernel tests (gemm / linear / migt)"
timeout 1200 python -m pytest tests/test_kernels_gpu.py tests/test_models_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/tests_w2.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/tests_w2.log | cut -c1-300
echo "=== linears microbench"
for f in 0 1; do echo "VF_TC_WIDE2=$f"; VF_TC_WIDE2=$f timeout 600 python scripts/bench_kernels.py 2>&1 | grep -aiE "migt|lm head" | head -8; done
