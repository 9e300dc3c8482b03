#!/bin/bash
# profiles/quick_check.sh <tag> — build check, GPU parity tests and one bench line (run under gpurun, 1 GPU)
set -u
OUT=gpurun_out/${1:-quick}; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 $OUT/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_ref32.json 2> $OUT/bench_ref32.err; echo "bench rc=$?"; tail -c 600 $OUT/bench_ref32.err
python -c "
import json; d=json.load(open('$OUT/bench_ref32.json')); print('value %.3e ms %.2f e2e %.3e'%(d['value'],d['ms_per_step'],d['e2e']['value'])); print(d['roofline'].get('executed'))"
