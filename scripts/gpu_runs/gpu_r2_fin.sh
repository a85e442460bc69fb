#!/bin/bash
set -u
mkdir -p gpurun_out
python -m viewformer_b200.build > gpurun_out/build.log 2>&1
echo "=== model C-ABI test"
timeout 1200 python -m pytest tests/test_model_cabi.py -m gpu -q -x --tb=short -p no:cacheprovider -s > gpurun_out/tests_fin.log 2>&1; echo "rc=$?"
grep -aE "cabi_host|passed|failed|^E |Error|timeout" gpurun_out/tests_fin.log | cut -c1-300 | tail -8
echo "=== bench"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_fin.log 2> gpurun_out/bench_fin.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_fin.log').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'parity', d['parity']['code_mismatches'])
r=d['roofline']; print({k:r[k] for k in ('achieved','frac','achieved_executed_mma','frac_executed_mma','launch_ms')})
print(d['roofline_vq_lookup']['frac'], d['gpu_launches'])
PY
