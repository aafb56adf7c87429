#!/bin/bash
# 1 GPU: tuning variants of k_sinefold<MOM>, e2e after moving the reference-line fit to the device, full GPU suite
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/s14_*
ab() { name=$1; shift; env "$@" python bench.py --steps 200 --warmup 5 --no-cpu > gpurun_out/s14_bench_$name.json 2>gpurun_out/s14_bench_$name.err; }
ab default A=1
ab waves1 MC3B_WAVES=1
ab sched8 MC3B_SCHED=2,8
ab sched16 MC3B_SCHED=2,16
ab plan4 MC3B_PLAN_RESIDENT=4
ab minb3 MC3B_LIBPATH=$PWD/variants/libmc3b200_minb3.so
ab minb3plan3 MC3B_LIBPATH=$PWD/variants/libmc3b200_minb3.so MC3B_PLAN_RESIDENT=3
ab acc4 MC3B_LIBPATH=$PWD/variants/libmc3b200_acc4.so
ab acc1 MC3B_LIBPATH=$PWD/variants/libmc3b200_acc1.so
ab minb3acc4 MC3B_LIBPATH=$PWD/variants/libmc3b200_minb3acc4.so
ab restart8 MC3B_LIBPATH=$PWD/variants/libmc3b200_restart8.so
ab nstage2 MC3B_LIBPATH=$PWD/variants/libmc3b200_nstage2.so
python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/s14_bench_k20.json 2>/dev/null
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s14_pytest.log 2>&1; echo "gpu tests rc=$?" >> gpurun_out/s14_summary.txt
python - <<'PY' >> gpurun_out/s14_summary.txt
import json, glob
for f in sorted(glob.glob('gpurun_out/s14_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); r = d['roofline']
        print(f, 'value %.3e' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'kernel ms %.4f' % r['ms_per_launch'], 'frac %.3f' % r['frac'], 'e2e %.3e' % d['e2e']['value'], 'hits', r.get('guard_hits'))
    except Exception as e: print(f, 'ERR', e)
PY
cat gpurun_out/s14_summary.txt
tail -12 gpurun_out/s14_pytest.log
