#!/bin/bash
# profiles/final_extra.sh <tag> — the at-scale FULL property test and the ncu launch list of the bench command
tag=${1:-x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "at_scale or full_mode" > gpurun_out/${tag}_full_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${tag}_full_tests.log
tail -3 gpurun_out/${tag}_full_tests.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_under_ncu.log 2>&1; echo "ncu rc=$?"

