# profiles/pipe_chunks_ab.sh — the host-pointer pipeline of rvs_distance: number of chunks (RVS_PIPE_CHUNKS) and ramped chunk sizes (RVS_PIPE_RAMP)
for combo in ${PIPE_AB_COMBOS:-"12 0" "8 1" "10 1" "12 1" "16 1"}; do
  set -- $combo
  RVS_PIPE_CHUNKS=$1 RVS_PIPE_RAMP=$2 timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/pc$1_r$2.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/pc$1_r$2.json')); print('chunks $1 ramp $2 value %.3e e2e %.3e e2e ms/step %.3f parity %s'%(d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['parity_sample']['bit_identical']))"
done
