# profiles/sort_joints_ab.sh — experiments on the sort key of rvs_sort.cuh (leading joints, configurations per bin); headline bench line, parity first
for sb in ${SORT_AB_BINS:-128 256 512 1024 4096}; do
  RVS_SORT_PER_BIN=$sb RVS_BENCH_CONFIGS=4194304 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/sb$sb.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/sb$sb.json')); n=d['roofline']['configs_per_launch']; print('per bin $sb ms/1e6 %.4f parity %s'%(d['ms_per_step']/n*1e6, d['parity_sample']['bit_identical']))"
done
