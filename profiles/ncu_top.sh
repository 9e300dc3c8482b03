#!/bin/bash
# profiles/ncu_top.sh <tag> [precision] — one `ncu --set full` capture of the dominant distance kernel inside a short bench run (1 GPU).
TAG=${1:-run}; PREC=${2:-ref32}
OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:distance_ -s 3 -c 1 -f -o "$OUT/prof_distance_$PREC" \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --precision $PREC > "$OUT/prof_bench_$PREC.log" 2>&1
ls -la "$OUT"
