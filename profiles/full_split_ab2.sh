#!/bin/bash
# profiles/full_split_ab2.sh <tag> — nearest_points_kernel with deferred-edge rings of 16 / 8 / 32 entries, then one ncu --set full capture of it
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
