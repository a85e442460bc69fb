#!/bin/bash
set -u
mkdir -p gpurun_out
python -m viewformer_b200.build > gpurun_out/build.log 2>&1
echo "=== all tests"; timeout 1800 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/tests_m.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/tests_m.log | cut -c1-300
echo "=== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log | cut -c1-300
timeout 300 python scripts/prof_exact_conv.py 2>&1 | tee gpurun_out/exact_conv_timing4.log
echo "=== bench"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_m.log 2> gpurun_out/bench_m.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_m.log').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','parity','value_by_precision','roofline','roofline_vq_lookup','cpu_baseline','clocks'): print(k, d.get(k))
PY
tail -3 gpurun_out/bench_m.err
echo "=== kvcache"; timeout 600 python bench.py --workload kvcache --precision bf16 --scenes 128 --steps 5 --warmup 3 2>&1 | tail -2
echo "=== launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_mixed.csv python scripts/profile_step.py --precision mixed > gpurun_out/prof_step.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_mixed.csv > gpurun_out/launches_mixed_summary.md 2>&1; head -24 gpurun_out/launches_mixed_summary.md
