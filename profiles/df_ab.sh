# profiles/df_ab.sh [tag] — A/B of the two result-neutral switches of the work-queue distance kernel (RVS_DISTANCE_FIELD, RVS_SORT_CONFIGS)
# on the headline bench line; parity first (the bench's own oracle sample must be bit-identical).
set -u
OUT=gpurun_out/${1:-df_ab}; mkdir -p $OUT
for combo in "1 1" "1 0" "0 1" "0 0"; do
  set -- $combo
  RVS_DISTANCE_FIELD=$1 RVS_SORT_CONFIGS=$2 RVS_BENCH_CONFIGS=${AB_CONFIGS:-4194304} timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $OUT/df$1_sort$2.json 2> $OUT/df$1_sort$2.err
  python - <<PY
import json
d=json.load(open('$OUT/df$1_sort$2.json')); w=d['roofline']['executed_work']; p=d['parity_sample']; n=d['roofline']['configs_per_launch']
print('df=$1 sort=$2 ms/1e6 %.4f (step %.4f) value %.3e e2e %.3e parity_exact %s surv %.1f cls %.1f edges %.1f'%(d['roofline']['avg_launch_ms']/n*1e6, d['ms_per_step']/n*1e6, d['value'], d['e2e']['value'], p['bit_identical'], w['scan_survivors']/n, w['classified']/n, w['edges']/n))
PY
done
