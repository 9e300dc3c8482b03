for pc in 4 8 12 16 24; do
  RVS_PIPE_CHUNKS=$pc timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/pc$pc.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/pc$pc.json')); print('chunks $pc value %.3e e2e %.3e e2e ms/step %.3f'%(d['value'], d['e2e']['value'], d['e2e']['ms_per_step']))"
done
