#!/bin/bash
# ncu captures: attention v2 (bench shape), fused VQ lookup (main + rescue) at M = 2^20
set -u
mkdir -p gpurun_out
python -m viewformer_b200.build > gpurun_out/build.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_block_causal -c 1 -o gpurun_out/prof_attn_v2 -f python scripts/bench_attn.py --once > gpurun_out/prof_attn.log 2>&1; echo "attn ncu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vq_lookup_fused -s 2 -c 1 -o gpurun_out/prof_vq_fused3 -f python scripts/bench_vq.py fused > gpurun_out/prof_vq_fused.log 2>&1; echo "vq ncu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vq_rescue -s 2 -c 1 -o gpurun_out/prof_vq_rescue3 -f python scripts/bench_vq.py fused > gpurun_out/prof_vq_rescue.log 2>&1; echo "rescue ncu rc=$?"
python scripts/bench_vq.py fused 2>&1 | tail -2
ls -la gpurun_out/*.ncu-rep
