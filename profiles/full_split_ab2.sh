#!/bin/bash
# profiles/full_split_ab2.sh <tag> — nearest_points_kernel with deferred-edge rings of 16 / 8 / 32 entries, then one ncu --set full capture of it
# The variants/*.so it loads through RVS_CUDA_LIB are builds of the same tree with one macro changed, made beforehand with e.g.
#   make -C rpmplv2_b200/csrc BUILD=_build_x OUT=../../variants/librvs_full_ring32.so EXTRA=-DRVS_FULL_RING=32
# (RVS_FULL_MINBLOCKS, RVS_FULL_RING, RVS_FULL_STORE16, RVS_FULL_PARTS_TARGET, RVS_SMALL_BATCH_FULL); variants/ is scratch and not committed.
tag=${1:-x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "full" > gpurun_out/${tag}_full_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${tag}_full_tests.log
tail -2 gpurun_out/${tag}_full_tests.log
for lib in rpmplv2_b200/librvs_cuda.so variants/librvs_full_mb6.so variants/librvs_full_store8.so; do
  echo "== $lib"
  RVS_CUDA_LIB=$PWD/$lib timeout 300 python profiles/full_probe.py 2>&1 | tee gpurun_out/${tag}_full_probe_$(basename $lib .so).txt | grep "FULL "
done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:nearest_points --launch-skip 1 --launch-count 1 -o gpurun_out/${tag}_nearest_points python profiles/full_ncu_driver.py > gpurun_out/${tag}_ncu.log 2>&1
ncu -i gpurun_out/${tag}_nearest_points.ncu-rep --page details > gpurun_out/${tag}_ncu_details_nearest_points.txt 2>&1
ls -la gpurun_out/${tag}_nearest_points.ncu-rep
