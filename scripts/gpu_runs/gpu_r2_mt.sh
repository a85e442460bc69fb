#!/bin/bash
python -m viewformer_b200.build > /dev/null 2>&1
timeout 900 python scripts/bench_migt_train.py 2>&1 | tail -2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_migt_train.csv -c 3000 python scripts/bench_migt_train.py > /dev/null 2>&1
python scripts/summarize_launches.py gpurun_out/launches_migt_train.csv 2>&1 | head -14
