# A/B of kernel variants on the GPU box: "lib-variant:ENV=VAL" entries (variants built by build_variants.py)
for spec in ${SPECS:-"default:" "default:MC3B_CPT=2" "default:"}; do
  v=${spec%%:*}; e=${spec#*:}
  if [ "$v" != default ]; then export MC3B_LIBPATH=$PWD/variants/libmc3b200_$v.so; else unset MC3B_LIBPATH; fi
  [ -n "$e" ] && export $e
  timeout 120 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -k "grid" 2>&1 | tail -1
  timeout 200 python bench.py --no-cpu > gpurun_out/bench_v.log 2>gpurun_out/bench_v.err
  python -c "
import json;d=json.loads(open('gpurun_out/bench_v.log').read().strip().splitlines()[-1]);print('VARIANT','$spec', round(d['value']), d['ms_per_step'], d['roofline']['ms_per_launch'], round(d['e2e']['value']))"
  [ -n "$e" ] && unset ${e%%=*}
done
