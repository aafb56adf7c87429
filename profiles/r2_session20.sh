#!/bin/bash
# 1 GPU: split-schedule variants for the moment kernel (per-CTA overhead is now ~38 % of warp time)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/s20_*
ab() { name=$1; shift; env "$@" python bench.py --steps 200 --warmup 5 --no-cpu > gpurun_out/s20_bench_$name.json 2>gpurun_out/s20_bench_$name.err; }
ab default A=1
ab eq_w1_r4 MC3B_SCHED=0 MC3B_WAVES=1 MC3B_PLAN_RESIDENT=4
ab eq_w2_r4 MC3B_SCHED=0 MC3B_WAVES=2 MC3B_PLAN_RESIDENT=4
ab eq_w3_r4 MC3B_SCHED=0 MC3B_WAVES=3 MC3B_PLAN_RESIDENT=4
ab s2_16_r4 MC3B_SCHED=2,16 MC3B_PLAN_RESIDENT=4
ab s2_8_r4 MC3B_SCHED=2,8 MC3B_PLAN_RESIDENT=4
ab s15_8_r4 MC3B_SCHED=1.5,8 MC3B_PLAN_RESIDENT=4
ab s13_8_r4 MC3B_SCHED=1.3,8 MC3B_PLAN_RESIDENT=4
ab s3_8_r4 MC3B_SCHED=3,8 MC3B_PLAN_RESIDENT=4
ab s15_4_r4 MC3B_SCHED=1.5,4 MC3B_PLAN_RESIDENT=4
ab s2_4_r4 MC3B_SCHED=2,4 MC3B_PLAN_RESIDENT=4
ab s13_16_r4 MC3B_SCHED=1.3,16 MC3B_PLAN_RESIDENT=4
python - <<'PY' >> gpurun_out/s20_summary.txt
import json, glob
for f in sorted(glob.glob('gpurun_out/s20_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); r = d['roofline']
        print(f, 'value %.3e' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'kernel ms %.4f' % r['ms_per_launch'], 'e2e %.3e' % d['e2e']['value'])
    except Exception as e: print(f, 'ERR', e)
PY
cat gpurun_out/s20_summary.txt
