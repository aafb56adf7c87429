#!/bin/bash
# 1 GPU: split plan of its own for the moment kernel (MC3B_PLAN_MOMENT): full suite, smoke, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/s34_*
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s34_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/s34_summary.txt
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/s34_pytest.log 2>&1; echo "gpu tests rc=$?" >> gpurun_out/s34_summary.txt
python bench.py --gpus 1 --steps 200 --warmup 5 --no-cpu > gpurun_out/s34_bench_k200.json 2> gpurun_out/s34_bench_k200.err; echo "bench rc=$?" >> gpurun_out/s34_summary.txt
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu > gpurun_out/s34_bench_k20.json 2> /dev/null
python profiles/gapped_bench.py > gpurun_out/s34_gapped.json 2>/dev/null
python - <<'PY' >> gpurun_out/s34_summary.txt
import json
for f in ('k200', 'k20'):
    d = json.loads(open(f'gpurun_out/s34_bench_{f}.json').read().strip().splitlines()[-1]); r = d['roofline']
    print(f, 'value %.4e' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'kernel ms %.4f' % r['ms_per_launch'], 'e2e %.4e' % d['e2e']['value'], 'frac %.3f' % r['frac'])
g = json.loads(open('gpurun_out/s34_gapped.json').read()); print({k: (round(v['ms_per_generation'], 4) if isinstance(v, dict) else v) for k, v in g.items()})
PY
cat gpurun_out/s34_summary.txt; tail -2 gpurun_out/s34_smoke.log; tail -4 gpurun_out/s34_pytest.log
