set -u
OUT=gpurun_out/df1; mkdir -p $OUT
for df in 1 0; do
  RVS_DISTANCE_FIELD=$df RVS_BENCH_CONFIGS=4194304 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $OUT/df$df.json 2> $OUT/df$df.err
  python - <<PY
import json
d=json.load(open('$OUT/df$df.json')); w=d['roofline']['executed_work']; p=d['parity_sample']; n=d['roofline']['configs_per_launch']
print('df=$df ms/1e6 %.4f value %.3e e2e %.3e parity_exact %s surv %.1f cls %.1f edges %.1f'%(d['roofline']['avg_launch_ms']/n*1e6, d['value'], d['e2e']['value'], p['bit_identical'], w['scan_survivors']/n, w['classified']/n, w['edges']/n))
PY
done
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -5
