#!/bin/bash
# profiles/full_split_ab4.sh <tag> — nearest_points_kernel: streaming stores (st.global.cs), rings of 4 entries
# The variants/*.so it loads through RVS_CUDA_LIB are builds of the same tree with one macro changed, made beforehand with e.g.
#   make -C rpmplv2_b200/csrc BUILD=_build_x OUT=../../variants/librvs_full_ring32.so EXTRA=-DRVS_FULL_RING=32
# (RVS_FULL_MINBLOCKS, RVS_FULL_RING, RVS_FULL_STORE16, RVS_FULL_PARTS_TARGET, RVS_SMALL_BATCH_FULL); variants/ is scratch and not committed.
tag=${1:-x}
mkdir -p gpurun_out
for lib in rpmplv2_b200/librvs_cuda.so variants/librvs_full_stcs.so variants/librvs_full_ring4.so; do
  echo "== $lib"
  RVS_CUDA_LIB=$PWD/$lib timeout 300 python profiles/full_probe.py 2>&1 | tee gpurun_out/${tag}_full_probe_$(basename $lib .so).txt | grep "FULL " | grep -v "unsorted\|N 2048\|N 16384\|M 6 "
done
