#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/s2_*
timeout 900 python -m pytest tests/test_gpu_r2.py -x -q > gpurun_out/s2_pytest_r2.log 2>&1; echo "r2 tests rc=$?" >> gpurun_out/s2_summary.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "grid or partial or model_chisq" > gpurun_out/s2_pytest_k.log 2>&1; echo "kernel tests rc=$?" >> gpurun_out/s2_summary.txt
./profiles/probes/fp64_pipes > gpurun_out/s2_fp64_pipes.jsonl 2>&1
python bench.py --steps 200 --warmup 5 --no-cpu > gpurun_out/s2_bench_default.json 2> gpurun_out/s2_bench_default.err
MC3B_NO_USIG=1 python bench.py --steps 200 --warmup 5 --no-cpu > gpurun_out/s2_bench_nousig.json 2>/dev/null
for v in r2m4 r4m4 r2m5 r4m6; do
  MC3B_LIBPATH=$PWD/variants/libmc3b200_$v.so python bench.py --steps 200 --warmup 5 --no-cpu > gpurun_out/s2_bench_$v.json 2>/dev/null
done
MC3B_PLAN_RESIDENT=4 MC3B_LIBPATH=$PWD/variants/libmc3b200_r4m4.so python bench.py --steps 200 --warmup 5 --no-cpu > gpurun_out/s2_bench_r4m4_plan4.json 2>/dev/null
MC3B_PLAN_RESIDENT=8 MC3B_LIBPATH=$PWD/variants/libmc3b200_r4m4.so python bench.py --steps 200 --warmup 5 --no-cpu > gpurun_out/s2_bench_r4m4_plan8.json 2>/dev/null
MC3B_PLAN_RESIDENT=5 MC3B_LIBPATH=$PWD/variants/libmc3b200_r2m5.so python bench.py --steps 200 --warmup 5 --no-cpu > gpurun_out/s2_bench_r2m5_plan5.json 2>/dev/null
python profiles/e2e_breakdown.py 20 > gpurun_out/s2_e2e_brk20.log 2>&1
python - <<'PY' >> gpurun_out/s2_summary.txt
import json, glob
for f in sorted(glob.glob('gpurun_out/s2_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d['roofline']
        print(f, 'value %.3e' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'kernel ms %.4f' % r['ms_per_launch'],
              'frac %.3f' % r['frac'], 'e2e %.3e' % d['e2e']['value'], 'launches', d['gpu_launches'])
    except Exception as e:
        print(f, 'ERR', e)
PY
cat gpurun_out/s2_summary.txt gpurun_out/s2_fp64_pipes.jsonl
tail -5 gpurun_out/s2_pytest_r2.log gpurun_out/s2_pytest_k.log
