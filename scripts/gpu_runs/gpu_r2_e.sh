#!/bin/bash
set -u
mkdir -p gpurun_out
python -m viewformer_b200.build > gpurun_out/build.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/vq_launches.csv python scripts/bench_vq.py fused > gpurun_out/vq_prof.log 2>&1
grep -a "vq_\|rescue" gpurun_out/vq_launches.csv | awk -F'","' '{print $5, $(NF)}' | tail -12
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vq_lookup_fused -s 2 -c 1 -o gpurun_out/prof_vq_fused -f python scripts/bench_vq.py fused > gpurun_out/prof_vq_fused.log 2>&1
echo "capture rc=$?"
echo "=== exact tests"; timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_baseline_configs_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider -k "exact or split or c2 or batch_invariance" > gpurun_out/tests_e.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/tests_e.log
timeout 300 python scripts/prof_exact_conv.py 2>&1 | tee gpurun_out/exact_conv_timing2.log
echo "=== bench"; timeout 900 python bench.py --steps 5 --warmup 3 --also bf16 --no-cpu-baseline > gpurun_out/bench_e.log 2> gpurun_out/bench_e.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_e.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','parity','value_by_precision')})
PY
tail -3 gpurun_out/bench_e.err
