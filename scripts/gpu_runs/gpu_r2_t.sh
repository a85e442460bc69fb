#!/bin/bash
set -u
mkdir -p gpurun_out
python -m viewformer_b200.build > gpurun_out/build.log 2>&1
echo "=== training tests + 2cta"
timeout 1200 python -m pytest tests/test_train_gpu.py tests/test_kernels_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider -s -k "train or commit or dynamic or cta_pair" > gpurun_out/tests_t.log 2>&1; echo "rc=$?"
grep -aE "^\[|passed|failed|^E |Error|timeout" gpurun_out/tests_t.log | cut -c1-260 | tail -30
echo "=== 2cta microbench (MIGT linears)"
for f in 0 1; do echo "VF_TC_2CTA=$f"; VF_TC_2CTA=$f timeout 600 python scripts/bench_kernels.py 2>&1 | grep -aiE "linear|c_fc|c_proj|qk|lm head|fc2|gemm" | head -12; done
