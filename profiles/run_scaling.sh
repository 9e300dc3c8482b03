#!/bin/bash
# profiles/run_scaling.sh <tag> <N> — weak and strong scaling lines at N GPUs of one box (run under `gpurun --gpus N`)
TAG=${1:-scale}; N=${2:-2}
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N"
timeout 600 $RUN --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $OUT/weak_${N}gpu.json 2> $OUT/weak_${N}gpu.err; echo "weak rc=$?"
timeout 600 $RUN --scaling strong --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $OUT/strong_${N}gpu.json 2> $OUT/strong_${N}gpu.err; echo "strong rc=$?"
python - <<PY
import json
for k in ("weak","strong"):
    try:
        d=json.load(open("$OUT/%s_${N}gpu.json"%k)); print(k, "value %.3e ms %.2f e2e %.3e"%(d["value"], d["ms_per_step"], d["e2e"]["value"]), d.get("strong_scaling"))
    except Exception as e: print(k, "FAILED", e); print(open("$OUT/%s_${N}gpu.err"%k).read()[-1500:])
PY
