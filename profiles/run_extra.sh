#!/bin/bash
# profiles/run_extra.sh <tag> — supplementary measurements (1 GPU): the S1/S2 sweep matrix incl. MotionValidity rows, and the planning
# benchmark at P = 64 and P = 1024 lockstep runs.
TAG=${1:-extra}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python bench_sweep.py --out $OUT/sweep.md > $OUT/sweep.jsonl 2> $OUT/sweep.err; echo "sweep rc=$?"
timeout 900 python bench_planning.py --runs 64 > $OUT/planning_p64.jsonl 2> $OUT/planning_p64.err; echo "planning64 rc=$?"
timeout 600 python bench_planning.py --runs 1024 --skip "M=1000" > $OUT/planning_p1024.jsonl 2> $OUT/planning_p1024.err; echo "planning1024 rc=$?"
tail -n 12 $OUT/sweep.md; tail -c 600 $OUT/sweep.err
python - <<PY
import json
for f in ("$OUT/planning_p64.jsonl", "$OUT/planning_p1024.jsonl"):
    for l in open(f):
        try: r = json.loads(l)
        except Exception: continue
        print(f.split('/')[-1], r["scenario"], "lockstep ms/run %.3f" % r["gpu_lockstep"]["per_run_ms"], "cpu_all ms/run %.3f" % r.get("cpu_all", {}).get("per_run_ms", float("nan")), "cpu1 %.3f" % r.get("cpu_1thread", {}).get("per_run_ms", float("nan")), "same", r.get("identical_runs_gpu_vs_cpu"))
PY
# round 2: BASELINE config 4 (DRGBT-shaped replanning loop, per 50 ms iteration), RRTx tree growth, FULL-mode probe
timeout 600 python bench_drgbt.py --iters 300 --obstacles 5 > $OUT/drgbt_m5.json 2> $OUT/drgbt_m5.err; echo "drgbt5 rc=$?"
timeout 600 python bench_drgbt.py --iters 300 --obstacles 50 > $OUT/drgbt_m50.json 2> $OUT/drgbt_m50.err; echo "drgbt50 rc=$?"
timeout 600 python profiles/rrtx_probe.py > $OUT/rrtx_probe.jsonl 2> $OUT/rrtx_probe.err; echo "rrtx rc=$?"
timeout 300 python profiles/full_probe.py > $OUT/full_probe.txt 2>&1
python - <<PY
import json
for f in ("$OUT/drgbt_m5.json", "$OUT/drgbt_m50.json"):
    try:
        d = json.load(open(f))
        for a in d["arms"]: print(f.split('/')[-1], a["arm"][:30], "iter ms median %.3f p95 %.3f within50ms %.3f" % (a["iter_ms"]["median"], a["iter_ms"]["p95"], a["within_50ms"]), a["split_ms_mean"])
        print("speedup", d.get("iteration_speedup_median"), "same run", d.get("same_run"))
    except Exception as e: print(f, "FAILED", e)
for l in open("$OUT/rrtx_probe.jsonl"): print(l.strip()[:400])
PY
