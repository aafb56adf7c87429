#!/bin/bash
# 1 GPU: explicit shared-memory prefetch in the moment kernel (distance LDS -> first use: 9 / 12 / 29 instructions)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/s25_*
ab() { name=$1; shift; env "$@" python bench.py --steps 200 --warmup 5 --no-cpu > gpurun_out/s25_bench_$name.json 2>gpurun_out/s25_bench_$name.err; }
ab default A=1
ab pf4 MC3B_LIBPATH=$PWD/variants/libmc3b200_pf4.so
ab pf3 MC3B_LIBPATH=$PWD/variants/libmc3b200_pf3.so
ab pf3plan3 MC3B_LIBPATH=$PWD/variants/libmc3b200_pf3.so MC3B_PLAN_RESIDENT=3
ab pf3a2 MC3B_LIBPATH=$PWD/variants/libmc3b200_pf3a2.so
ab pf4s15 MC3B_LIBPATH=$PWD/variants/libmc3b200_pf4.so MC3B_SCHED=1.5,8 MC3B_PLAN_RESIDENT=4
python - <<'PY' >> gpurun_out/s25_summary.txt
import json, glob
for f in sorted(glob.glob('gpurun_out/s25_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); r = d['roofline']
        print(f, 'value %.3e' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'kernel ms %.4f' % r['ms_per_launch'], 'e2e %.3e' % d['e2e']['value'])
    except Exception as e: print(f, 'ERR', e)
PY
cat gpurun_out/s25_summary.txt
