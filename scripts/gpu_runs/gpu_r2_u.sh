#!/bin/bash
# full validation of the current tree: all GPU tests, smoke, bench (mixed), kvcache, launch list
set -u
mkdir -p gpurun_out
python -m viewformer_b200.build > gpurun_out/build.log 2>&1
echo "=== all tests"; timeout 2400 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/tests_u.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/tests_u.log | cut -c1-300
echo "=== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log | cut -c1-300
echo "=== bench"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_u.log 2> gpurun_out/bench_u.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_u.log').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','parity','value_by_precision','roofline','roofline_vq_lookup','cpu_baseline','clocks','gpu_launches'): print(k, d.get(k))
PY
tail -3 gpurun_out/bench_u.err
echo "=== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-700
echo "=== kvcache"; timeout 600 python bench.py --workload kvcache --precision bf16 --scenes 128 --steps 5 --warmup 3 2>&1 | tail -1 | tee gpurun_out/kvcache_u.json | cut -c1-500
echo "=== train workload"; timeout 600 python bench.py --workload train --scenes 32 --steps 5 --warmup 3 2>&1 | tail -1 | tee gpurun_out/train_u.json | cut -c1-500
echo "=== launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_mixed.csv python scripts/profile_step.py --precision mixed > gpurun_out/prof_step.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_mixed.csv > gpurun_out/launches_mixed_summary.md 2>&1; head -16 gpurun_out/launches_mixed_summary.md
