#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/s7_*
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s7_pytest.log 2>&1; echo "gpu tests rc=$?" >> gpurun_out/s7_summary.txt
python profiles/gen_breakdown.py > gpurun_out/s7_gen_breakdown.txt 2>&1
MC3B_PDL=0 python profiles/gen_breakdown.py > gpurun_out/s7_gen_breakdown_nopdl.txt 2>&1
python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/s7_bench_k20.json 2>gpurun_out/s7_bench_k20.err
python bench.py --steps 200 --warmup 5 --no-cpu > gpurun_out/s7_bench_k200.json 2>/dev/null
MC3B_PDL=0 python bench.py --steps 200 --warmup 5 --no-cpu > gpurun_out/s7_bench_k200_nopdl.json 2>/dev/null
python bench_configs.py config4 > gpurun_out/s7_config4.json 2> gpurun_out/s7_config4.err
MC3B_BA_MODE=1 python bench_configs.py config4 > gpurun_out/s7_config4_ba1.json 2>/dev/null
MC3B_BA_MODE=2 python bench_configs.py config4 > gpurun_out/s7_config4_ba2.json 2>/dev/null
python bench_configs.py config3 --steps 10 > gpurun_out/s7_config3.json 2> gpurun_out/s7_config3.err
python bench_configs.py config1 > gpurun_out/s7_config1.json 2> gpurun_out/s7_config1.err
python - <<'PY' >> gpurun_out/s7_summary.txt
import json, glob
for f in sorted(glob.glob('gpurun_out/s7_bench_k*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); r = d['roofline']
        print(f, 'value %.3e' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'kernel ms %.4f' % r['ms_per_launch'], 'frac %.3f' % r['frac'], 'e2e %.3e' % d['e2e']['value'])
    except Exception as e: print(f, 'ERR', e)
for f in sorted(glob.glob('gpurun_out/s7_config4*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k: round(d[k]['ms'], 4) for k in ('bin_array_unweighted', 'bin_array_weighted', 'time_avg')}, d['time_avg']['max_rel_err_vs_direct'], d['time_avg']['roofline']['frac'])
    except Exception as e: print(f, 'ERR', e)
for f in ('gpurun_out/s7_config3.json', 'gpurun_out/s7_config1.json'):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['value'], d.get('ms_per_step'), d.get('roofline', {}).get('frac'))
    except Exception as e: print(f, 'ERR', e)
PY
cat gpurun_out/s7_summary.txt gpurun_out/s7_gen_breakdown.txt gpurun_out/s7_gen_breakdown_nopdl.txt
tail -15 gpurun_out/s7_pytest.log
