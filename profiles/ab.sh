#!/bin/bash
# profiles/ab.sh <tag> <lib.so> [<lib.so> ...] — A/B the headline bench line over differently tuned builds of librvs_cuda.so (RVS_CUDA_LIB).
# Parity first (the bench's own 2^14-configuration oracle sample must be bit-identical), then the timing.
TAG=${1:-ab}; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
for lib in "$@"; do
  name=$(basename $lib .so)
  RVS_CUDA_LIB=$PWD/$lib RVS_BENCH_CONFIGS=${AB_CONFIGS:-4194304} timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $OUT/$name.json 2> $OUT/$name.err
  python -c "
import json; d=json.load(open('$OUT/$name.json')); w=d['roofline']['executed_work']; p=d['parity_sample']
print('%-14s ms/1e6 %.4f value %.3e e2e %.3e parity_exact %s surv %.1f cls %.1f edges %.1f'%('$name', d['ms_per_step']/d['config']['configs_per_step']*1e6*0+d['roofline']['avg_launch_ms']/d['roofline']['configs_per_launch']*1e6, d['value'], d['e2e']['value'], p['bit_identical'], w['scan_survivors']/d['roofline']['configs_per_launch'], w['classified']/d['roofline']['configs_per_launch'], w['edges']/d['roofline']['configs_per_launch']))" 2>&1 | tail -1
done
