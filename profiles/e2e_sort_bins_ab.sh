for sb in 32 64 128 256; do
  RVS_SORT_PER_BIN=$sb timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/e2e_sb$sb.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/e2e_sb$sb.json')); print('per bin $sb value %.3e e2e %.3e e2e ms/step %.3f'%(d['value'], d['e2e']['value'], d['e2e']['ms_per_step']))"
done
