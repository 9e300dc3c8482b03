#!/bin/bash
# profiles/ab.sh <tag> <lib.so> [<lib.so> ...] — A/B the default bench line over differently tuned builds of librvs_cuda.so (RVS_CUDA_LIB)
TAG=${1:-ab}; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
for lib in "$@"; do
  name=$(basename $lib .so)
  RVS_CUDA_LIB=$PWD/$lib timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/$name.json 2> $OUT/$name.err
  python -c "
import json; d=json.load(open('$OUT/$name.json')); print('$name', 'ms %.3f value %.3e e2e %.3e'%(d['ms_per_step'],d['value'],d['e2e']['value']), d['roofline'].get('executed',{}).get('scan_survivors'))"
done
