#!/bin/bash
# profiles/run_profile.sh <tag> — the per-round GPU evidence run (B200_PROFILING.md recipe); run under gpurun on ONE GPU:
#   gpurun --timeout 2400 -- 'bash profiles/run_profile.sh r2x'
# Writes everything under gpurun_out/<tag>/ ; summaries worth keeping are copied to profiles/r2/ by hand (profiles/ncu_extract.py).
set -u
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit,memory.total --format=csv > "$OUT/gpu.csv" 2>&1
nproc > "$OUT/nproc.txt"
# 1. parity tests proper
timeout 1200 python -m pytest tests -m gpu -q > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?" | tee -a "$OUT/pytest_gpu.log"
# 2. smoke
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1; echo "smoke rc=$?" | tee -a "$OUT/smoke.log"
# 3. bench lines (never under a profiler)
timeout 900 python bench.py --steps 20 --warmup 3 > "$OUT/bench_ref32.json" 2> "$OUT/bench_ref32.err"; echo "bench rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --precision fast32 --no-cpu-baseline --no-extras > "$OUT/bench_fast32.json" 2> "$OUT/bench_fast32.err"
timeout 600 python bench.py --steps 5 --warmup 3 --precision f64 --no-cpu-baseline --no-extras > "$OUT/bench_f64.json" 2> "$OUT/bench_f64.err"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > "$OUT/bench_reference.json" 2> "$OUT/bench_reference.err"
timeout 600 python bench.py --scaling strong --steps 3 --warmup 3 --no-cpu-baseline --no-extras > "$OUT/bench_strong_1gpu.json" 2> "$OUT/bench_strong_1gpu.err"
# 4. ncu launch list of the same bench command (cold-cache, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches.csv" \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > "$OUT/launches_bench.log" 2>&1
# 5. one full capture of the dominant kernel on the bench's own workload (10^7 configurations per launch)
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:distance_wq -s 3 -c 1 -f -o "$OUT/prof_distance" \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > "$OUT/prof_bench.log" 2>&1
ls -la "$OUT"
tail -c 1500 "$OUT/bench_ref32.json"
