#!/bin/bash
# 1 GPU: k_sinemma (Pe, Po as FP64 tensor-core products): parity + guard tests, A/B against the FMA form
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/s30_*
timeout 900 python -m pytest tests/test_gpu_moment.py -x -q > gpurun_out/s30_pytest_moment.log 2>&1; echo "moment tests (mma) rc=$?" >> gpurun_out/s30_summary.txt
MC3B_MOM_LAYOUT=0 timeout 900 python -m pytest tests/test_gpu_moment.py -x -q > gpurun_out/s30_pytest_moment_l0.log 2>&1; echo "moment tests (fma) rc=$?" >> gpurun_out/s30_summary.txt
ab() { name=$1; shift; env "$@" python bench.py --steps 200 --warmup 5 --no-cpu > gpurun_out/s30_bench_$name.json 2>gpurun_out/s30_bench_$name.err; }
ab mma A=1
ab fma MC3B_MOM_LAYOUT=0
ab mma5 MC3B_LIBPATH=$PWD/variants/libmc3b200_mma5.so
ab mma3 MC3B_LIBPATH=$PWD/variants/libmc3b200_mma3.so
python profiles/gen_breakdown.py > gpurun_out/s30_gen_breakdown.txt 2>&1
python - <<'PY' >> gpurun_out/s30_summary.txt
import json, glob
for f in sorted(glob.glob('gpurun_out/s30_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); r = d['roofline']
        print(f, 'value %.3e' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'kernel ms %.4f' % r['ms_per_launch'], 'e2e %.3e' % d['e2e']['value'], 'hits', r.get('guard_hits'))
    except Exception as e: print(f, 'ERR', e)
PY
cat gpurun_out/s30_summary.txt gpurun_out/s30_gen_breakdown.txt; tail -12 gpurun_out/s30_pytest_moment.log; tail -3 gpurun_out/s30_pytest_moment_l0.log
