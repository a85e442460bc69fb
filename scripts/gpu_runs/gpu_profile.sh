#!/bin/bash
# launch list of one full step + full ncu captures of the dominant kernel (conv 128->128 @128^2, 288 images)
set -u
mkdir -p gpurun_out
python -m viewformer_b200.build > gpurun_out/build.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py > gpurun_out/prof_step.log 2>&1
echo "launch list rc=$?"
python scripts/summarize_launches.py gpurun_out/launches.csv > gpurun_out/launches_summary.md 2>&1; head -24 gpurun_out/launches_summary.md
# first wide-conv launch of bench_kernels.py = the roofline kernel of bench.py (no residual); 2nd shape (+3 warmups +10 reps later) has the residual
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_conv3x3_wide -c 1 -o gpurun_out/prof_conv_nores -f python scripts/bench_kernels.py > gpurun_out/prof_conv_nores.log 2>&1
echo "full capture (no residual) rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_conv3x3_wide -s 13 -c 1 -o gpurun_out/prof_conv_res -f python scripts/bench_kernels.py > gpurun_out/prof_conv_res.log 2>&1
echo "full capture (residual) rc=$?"
