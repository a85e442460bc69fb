#!/bin/bash
set -u
mkdir -p gpurun_out
python -m viewformer_b200.build > gpurun_out/build.log 2>&1
echo "=== attention + kv-cache tests"
timeout 1500 python -m pytest tests/test_kernels_gpu.py tests/test_baseline_configs_gpu.py tests/test_models_gpu.py tests/test_model_cabi.py -m gpu -q -x --tb=short -p no:cacheprovider -s -k "attention or c5 or kv or cache or c_host" > gpurun_out/tests_dec.log 2>&1; echo "rc=$?"
grep -aE "^\[fused attention decode|^\[C5|passed|failed|^E |Error|timeout" gpurun_out/tests_dec.log | cut -c1-260 | tail -16
echo "=== kvcache bench"
timeout 600 python bench.py --workload kvcache --precision bf16 --scenes 128 --steps 5 --warmup 3 2>&1 | tail -1 | cut -c1-400
timeout 300 python scripts/bench_attn.py 2>&1 | tail -3
