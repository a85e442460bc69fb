#!/bin/bash
# launch list of one full step + one full ncu capture of the top conv kernel
set -u
mkdir -p gpurun_out
python -m viewformer_b200.build > gpurun_out/build.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py > gpurun_out/prof_step.log 2>&1
echo "launch list rc=$?"
python scripts/summarize_launches.py gpurun_out/launches.csv > gpurun_out/launches_summary.md 2>&1; head -24 gpurun_out/launches_summary.md
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tc_gemm -s 1 -c 1 -o gpurun_out/prof_conv -f python scripts/profile_step.py --part encode > gpurun_out/prof_conv.log 2>&1
echo "full capture rc=$?"
