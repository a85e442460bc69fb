#!/bin/bash
set -u
mkdir -p gpurun_out
python -m viewformer_b200.build > gpurun_out/build.log 2>&1
export NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 scripts/train_dp_check.py > gpurun_out/train_dp_n8_tc.raw 2>&1; echo "rc=$?"
grep -aE "^\[|Init COMPLETE" gpurun_out/train_dp_n8_tc.raw | cut -c1-300 | tee gpurun_out/train_dp_n8_tc.log | grep -a "^\[" 
