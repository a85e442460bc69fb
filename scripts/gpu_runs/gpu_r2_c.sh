#!/bin/bash
# round-2 pass c: launch list of the mixed-mode step, timing + full ncu capture of the exact wide conv
set -u
mkdir -p gpurun_out
python -m viewformer_b200.build > gpurun_out/build.log 2>&1
timeout 600 python scripts/prof_exact_conv.py 2>&1 | tee gpurun_out/exact_conv_timing.log
for kc in 1 9; do VF_EXACT_KC=$kc timeout 300 python scripts/prof_exact_conv.py 2>&1 | sed "s/^/kc=$kc /" | tee -a gpurun_out/exact_conv_timing.log; done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_mixed.csv python scripts/profile_step.py --precision mixed > gpurun_out/prof_step.log 2>&1
echo "launch list rc=$?"
python scripts/summarize_launches.py gpurun_out/launches_mixed.csv > gpurun_out/launches_mixed_summary.md 2>&1; head -40 gpurun_out/launches_mixed_summary.md
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_conv3x3_wide -c 1 -o gpurun_out/prof_exact_conv -f python scripts/prof_exact_conv.py > gpurun_out/prof_exact_conv.log 2>&1
echo "full capture rc=$?"
ncu -i gpurun_out/prof_exact_conv.ncu-rep --page raw --csv > gpurun_out/prof_exact_conv_raw.csv 2>/dev/null
grep -E "gpu__time_duration.sum|sm__pipe_tc|tensor|dram__bytes_read.sum,|dram__bytes_write.sum,|sm__throughput|smsp__inst_executed.sum," gpurun_out/prof_exact_conv_raw.csv | head -30
