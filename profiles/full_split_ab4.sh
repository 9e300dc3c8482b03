#!/bin/bash
# profiles/full_split_ab4.sh <tag> — nearest_points_kernel: streaming stores (st.global.cs), rings of 4 entries
tag=${1:-x}
mkdir -p gpurun_out
for lib in rpmplv2_b200/librvs_cuda.so variants/librvs_full_stcs.so variants/librvs_full_ring4.so; do
  echo "== $lib"
  RVS_CUDA_LIB=$PWD/$lib timeout 300 python profiles/full_probe.py 2>&1 | tee gpurun_out/${tag}_full_probe_$(basename $lib .so).txt | grep "FULL " | grep -v "unsorted\|N 2048\|N 16384\|M 6 "
done
