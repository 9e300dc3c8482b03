#!/bin/bash
# profiles/full_split_ab3.sh <tag> — whole GPU suite on the tree with the FULL-mode split, the obstacle-range split of
# The variants/*.so it loads through RVS_CUDA_LIB are builds of the same tree with one macro changed, made beforehand with e.g.
#   make -C rpmplv2_b200/csrc BUILD=_build_x OUT=../../variants/librvs_full_ring32.so EXTRA=-DRVS_FULL_RING=32
# (RVS_FULL_MINBLOCKS, RVS_FULL_RING, RVS_FULL_STORE16, RVS_FULL_PARTS_TARGET, RVS_SMALL_BATCH_FULL); variants/ is scratch and not committed.
# nearest_points_kernel (RVS_FULL_PARTS_TARGET 2220 / 1 = off / 4440), a default bench line
tag=${1:-x}
mkdir -p gpurun_out
for lib in rpmplv2_b200/librvs_cuda.so variants/librvs_full_parts1.so variants/librvs_full_parts4k.so; do
  echo "== $lib"
  RVS_CUDA_LIB=$PWD/$lib timeout 300 python profiles/full_probe.py 2>&1 | tee gpurun_out/${tag}_full_probe_$(basename $lib .so).txt | grep "FULL " | grep -v unsorted
done
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${tag}_gpu_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${tag}_gpu_tests.log
tail -3 gpurun_out/${tag}_gpu_tests.log
timeout 600 python bench.py --no-extras > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/${tag}_bench.json')); print('value %.4g e2e %.4g ms %.3f frac %.3f'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac']))"
