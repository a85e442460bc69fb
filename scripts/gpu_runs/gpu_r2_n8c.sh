#!/bin/bash
# final-tree refresh of the 8-GPU evidence: generate() bench and the training workload
set -u
mkdir -p gpurun_out
python -m viewformer_b200.build > gpurun_out/build.log 2>&1
export NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT
run8() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
run8 29621 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_n8_final.raw 2>&1; echo "rc=$?"
grep -a '^{"metric' gpurun_out/bench_n8_final.raw | tail -1 > gpurun_out/bench_n8_final.json
grep -aE "Init COMPLETE|nranks" gpurun_out/bench_n8_final.raw | head -8 | cut -c1-200 > gpurun_out/bench_n8_final.nccl
python -c "
import json; d=json.load(open('gpurun_out/bench_n8_final.json')); print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','e2e','clocks')})"
run8 29622 bench.py --gpus 8 --workload train --scenes 32 --steps 5 --warmup 3 > gpurun_out/train_n8_final.raw 2>&1; echo "rc=$?"
grep -a '^{"metric' gpurun_out/train_n8_final.raw | tail -1 | tee gpurun_out/train_n8_final.json | cut -c1-400
