#!/bin/bash
set -u
mkdir -p gpurun_out
python -m viewformer_b200.build > gpurun_out/build.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:vq_ --csv --log-file gpurun_out/vq_durations.csv python scripts/bench_vq.py fused > /dev/null 2>&1
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/vq_durations.csv',errors='ignore')))
hdr=None; t=collections.defaultdict(list)
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r))
        if d.get('Metric Name')=='gpu__time_duration.sum': t[d['Kernel Name'][:50]].append(float(d['Metric Value'].replace(',','')))
for k,v in t.items(): print(k, len(v), 'median us', sorted(v)[len(v)//2]/1e3)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vq_lookup_fused -s 2 -c 1 -o gpurun_out/prof_vq_fused4 -f python scripts/bench_vq.py fused > gpurun_out/prof_vq_fused.log 2>&1; echo "vq ncu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_block_causal -c 1 -o gpurun_out/prof_attn_v3 -f python scripts/bench_attn.py --once > gpurun_out/prof_attn.log 2>&1; echo "attn ncu rc=$?"
