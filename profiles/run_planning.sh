OUT=gpurun_out/r1t; mkdir -p $OUT
timeout 600 python -m pytest tests/test_planners.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench_planning.py --runs 64 --skip "M=1000" > $OUT/planning_p64.jsonl 2> $OUT/p64.err
timeout 600 python bench_planning.py --runs 1024 --skip "M=1000" > $OUT/planning_p1024.jsonl 2> $OUT/p1024.err
python - <<PY
import json
for f in ("$OUT/planning_p64.jsonl", "$OUT/planning_p1024.jsonl"):
    for l in open(f):
        try: r = json.loads(l)
        except Exception: continue
        print(f.split('/')[-1], r["scenario"], "lockstep ms/run %.3f" % r["gpu_lockstep"]["per_run_ms"], "us/call %.1f" % r["gpu_lockstep"]["us_per_backend_call"], "cpu_all ms/run %.3f" % r.get("cpu_all", {}).get("per_run_ms", float("nan")), "cpu1 %.3f" % r.get("cpu_1thread", {}).get("per_run_ms", float("nan")), "same", r.get("identical_runs_gpu_vs_cpu"))
PY
