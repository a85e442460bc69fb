#!/bin/bash
set -u
python -m viewformer_b200.build > /dev/null 2>&1
timeout 300 python scripts/bench_attn.py 2>&1 | tail -4
timeout 600 python bench.py --workload kvcache --precision bf16 --scenes 128 --steps 10 --warmup 3 2>&1 | tail -1 | cut -c1-260
