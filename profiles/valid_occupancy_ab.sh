for lib in rpmplv2_b200/librvs_cuda.so variants/librvs_vq4.so variants/librvs_vq6.so variants/librvs_vq7.so; do
  RVS_CUDA_LIB=$PWD/$lib RVS_BENCH_CONFIGS=4194304 timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/vq_$(basename $lib .so).json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/vq_$(basename $lib .so).json')); print('$(basename $lib .so)', 'is_valid ms %.3f cfg/s %.3e'%(d['is_valid']['ms_per_launch'], d['is_valid']['configs_per_s']))"
done
