#!/bin/bash
# profiles/final_check.sh <tag> — the driver's round-end sequence on one box: GPU suite, smoke, reference arm, default bench line
tag=${1:-x}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${tag}_gpu_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${tag}_gpu_tests.log
tail -3 gpurun_out/${tag}_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py --impl reference > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err; echo "reference rc=$?"
timeout 900 python bench.py > gpurun_out/${tag}_bench_default.json 2> gpurun_out/${tag}_bench_default.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/${tag}_bench_default.json')); print('value %.4g e2e %.4g ms %.3f frac %.3f'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'])); print(json.dumps(d.get('full_mode'))); r=json.load(open('gpurun_out/${tag}_bench_reference.json')); print('reference %.4g'%r['value'])"
