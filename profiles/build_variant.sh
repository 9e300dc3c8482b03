#!/bin/bash
# profiles/build_variant.sh <name> "<-D flags>" — an experimental build of librvs_cuda.so into variants/librvs_<name>.so (git-ignored; travels to the
# GPU box) for A/B runs with profiles/ab.sh (RVS_CUDA_LIB selects the library).
NAME=$1; shift
make -C /root/repo/rpmplv2_b200/csrc -j4 OUT=/root/repo/variants/librvs_$NAME.so BUILD=/tmp/_build_$NAME EXTRA="$*" 2>&1 | grep -E "error|warning: v" | head
ls -la /root/repo/variants/librvs_$NAME.so | awk '{print $5, $9}'
