#!/bin/bash
set -u
mkdir -p gpurun_out
python -m viewformer_b200.build > gpurun_out/build.log 2>&1
echo "=== model C-ABI test"
timeout 1200 python -m pytest tests/test_model_cabi.py -m gpu -q -x --tb=short -p no:cacheprovider -s > gpurun_out/tests_v.log 2>&1; echo "rc=$?"
grep -aE "cabi_host|passed|failed|^E |Error|timeout" gpurun_out/tests_v.log | cut -c1-300 | tail -20
