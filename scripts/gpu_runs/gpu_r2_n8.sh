#!/bin/bash
# multi-GPU evidence: bench at N=1,2,4,8 (mixed), KV-cache workload at 8, data-parallel codebook training check at 8
set -u
mkdir -p gpurun_out
NG=${NG:-8}
python -m viewformer_b200.build > gpurun_out/build.log 2>&1
export NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT
run() { n=$1; shift; if [ "$n" = 1 ]; then timeout 600 python "$@"; else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) "$@"; fi; }
for n in 1 2 4 8; do
  [ $n -gt $NG ] && continue
  echo "=== bench N=$n"
  run $n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/bench_n$n.raw 2>&1; echo "rc=$?"
  grep -a '^{"metric' gpurun_out/bench_n$n.raw | tail -1 > gpurun_out/bench_n$n.json
  grep -aE "Init COMPLETE|NVLS|nranks" gpurun_out/bench_n$n.raw | head -12 | cut -c1-220 > gpurun_out/bench_n$n.nccl
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_n$n.json').read())
    print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','e2e','clocks')})
except Exception as e: print('parse failed', e)
PY
  head -3 gpurun_out/bench_n$n.nccl
done
echo "=== kvcache N=$NG"
run $NG bench.py --gpus $NG --workload kvcache --precision bf16 --scenes 128 --steps 5 --warmup 3 > gpurun_out/kvcache_n$NG.raw 2>&1; echo "rc=$?"
grep -a '^{"metric' gpurun_out/kvcache_n$NG.raw | tail -1 | tee gpurun_out/kvcache_n$NG.json | cut -c1-600
echo "=== config 3 shape: 8 scenes/GPU (64 scenes over 8 GPUs), mixed"
run $NG bench.py --gpus $NG --scenes 8 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg3_n$NG.raw 2>&1; echo "rc=$?"
grep -a '^{"metric' gpurun_out/bench_cfg3_n$NG.raw | tail -1 | tee gpurun_out/bench_cfg3_n$NG.json | cut -c1-500
echo "=== config 5 shape: 16 scenes/GPU (128 over 8 GPUs) KV-cache"
run $NG bench.py --gpus $NG --workload kvcache --precision bf16 --scenes 16 --steps 5 --warmup 3 > gpurun_out/kvcache16_n$NG.raw 2>&1; echo "rc=$?"
grep -a '^{"metric' gpurun_out/kvcache16_n$NG.raw | tail -1 | tee gpurun_out/kvcache16_n$NG.json | cut -c1-500
echo "=== kvcache N=1"
run 1 bench.py --gpus 1 --workload kvcache --precision bf16 --scenes 128 --steps 5 --warmup 3 2>/dev/null | grep -a '^{"metric' | tail -1 | tee gpurun_out/kvcache_n1.json | cut -c1-400
echo "=== config 3 (co3d 288 images encode+decode; bench generate workload is config 2) N=$NG train dp check"
run $NG scripts/train_dp_check.py > gpurun_out/train_dp_n$NG.raw 2>&1; echo "rc=$?"
grep -aE "^\[|Init COMPLETE" gpurun_out/train_dp_n$NG.raw | cut -c1-300 | tee gpurun_out/train_dp_n$NG.log | head -20
