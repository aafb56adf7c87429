#!/bin/bash
# 1 GPU: does the split plan that suits the moment kernel (4 resident CTAs, f=1.5, min 8 tiles) hurt the general kernels?
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/s26_*
python profiles/model_survey.py > gpurun_out/s26_survey_default.jsonl 2>gpurun_out/s26_survey_default.err
MC3B_SCHED=1.5,8 MC3B_PLAN_RESIDENT=4 python profiles/model_survey.py > gpurun_out/s26_survey_r4s15.jsonl 2>/dev/null
MC3B_SCHED=1.5,8 python profiles/model_survey.py > gpurun_out/s26_survey_r6s15.jsonl 2>/dev/null
MC3B_NO_MOMENT=1 python bench.py --steps 200 --warmup 5 --no-cpu > gpurun_out/s26_bench_pair_default.json 2>/dev/null
MC3B_NO_MOMENT=1 MC3B_SCHED=1.5,8 MC3B_PLAN_RESIDENT=4 python bench.py --steps 200 --warmup 5 --no-cpu > gpurun_out/s26_bench_pair_r4s15.json 2>/dev/null
python - <<'PY' > gpurun_out/s26_summary.txt
import json
runs = {}
for tag in ('default', 'r4s15', 'r6s15'):
    for ln in open(f'gpurun_out/s26_survey_{tag}.jsonl'):
        try: d = json.loads(ln)
        except Exception: continue
        key = (d['kernel'], d['model'], d['dtype'], d['uncertainties'], d['abscissa'])
        runs.setdefault(key, {})[tag] = d['ms_per_launch']
for k, v in runs.items():
    print(' | '.join(k), {t: round(x, 4) for t, x in v.items()})
for f in ('pair_default', 'pair_r4s15'):
    d = json.loads(open(f'gpurun_out/s26_bench_{f}.json').read().strip().splitlines()[-1])
    print(f, 'ms/step %.4f' % d['ms_per_step'], 'kernel %.4f' % d['roofline']['ms_per_launch'])
PY
cat gpurun_out/s26_summary.txt
