#!/bin/bash
# profiles/full_split_ab.sh <tag> — FULL mode of large batches: three-launch split (rvs_distance_full.cuh) against distance_kernel<FULL>
tag=${1:-x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "full or small_batch or work_queue_distance" > gpurun_out/${tag}_full_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${tag}_full_tests.log
tail -3 gpurun_out/${tag}_full_tests.log
RVS_FULL_SPLIT=1 timeout 300 python profiles/full_probe.py > gpurun_out/${tag}_full_probe_split.txt 2>&1
RVS_FULL_SPLIT=0 timeout 300 python profiles/full_probe.py > gpurun_out/${tag}_full_probe_nosplit.txt 2>&1
RVS_FULL_SPLIT=1 RVS_SORT_CONFIGS=0 timeout 300 python profiles/full_probe.py > gpurun_out/${tag}_full_probe_split_nosort.txt 2>&1
paste -d'\n' gpurun_out/${tag}_full_probe_split.txt gpurun_out/${tag}_full_probe_nosplit.txt | grep "FULL " 
