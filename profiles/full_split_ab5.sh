#!/bin/bash
# profiles/full_split_ab5.sh <tag> — nearest_points_kernel at 5 (96 registers) / 4 (101) / 6 (80) CTAs per SM after the obstacle-range split
# (variants/*.so: see full_split_ab2.sh)
tag=${1:-x}
mkdir -p gpurun_out
for lib in rpmplv2_b200/librvs_cuda.so variants/librvs_full_mb4.so variants/librvs_full_mb6.so; do
  echo "== $lib"
  RVS_CUDA_LIB=$PWD/$lib timeout 300 python profiles/full_probe.py 2>&1 | tee gpurun_out/${tag}_full_probe_$(basename $lib .so).txt | grep "FULL " | grep -v "unsorted\|N 2048\|N 16384\|M 6 "
done
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "full_mode or at_scale" 2>&1 | tail -1
