set -u
OUT=gpurun_out/plan2; mkdir -p $OUT
timeout 900 python -m pytest tests/test_planners.py tests/test_dropin.py tests/test_spines.py -m gpu -x -q 2>&1 | tail -3
for ch in 8 0; do
  RVS_PLAN_CONNECT_CHUNK=$ch timeout 900 python bench_planning.py --runs 64 --skip "M=1000" > $OUT/p64_chunk$ch.jsonl 2> $OUT/p64_chunk$ch.err
  python - <<PY
import json
for l in open("$OUT/p64_chunk$ch.jsonl"):
    try: r = json.loads(l)
    except Exception: continue
    print("chunk $ch", r["scenario"][:28], "lockstep ms/run %.3f calls %d" % (r["gpu_lockstep"]["per_run_ms"], r["gpu_lockstep"]["backend_calls"]), "single ms/run %.2f" % r["gpu_single"]["per_run_ms"], "cpu1 %.2f cpu_all %.3f" % (r["cpu_1thread"]["per_run_ms"], r["cpu_all"]["per_run_ms"]), "same", r.get("identical_runs_gpu_vs_cpu"))
PY
done
