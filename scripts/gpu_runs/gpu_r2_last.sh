#!/bin/bash
set -u
mkdir -p gpurun_out
python -m viewformer_b200.build > gpurun_out/build.log 2>&1
timeout 240 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/tests_last.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/tests_last.log | cut -c1-200
