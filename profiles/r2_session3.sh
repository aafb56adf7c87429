#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/s3_*
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/s3_pytest.log 2>&1; echo "gpu tests rc=$?" >> gpurun_out/s3_summary.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/s3_bench_k20.json 2> gpurun_out/s3_bench_k20.err
python bench.py --steps 200 --warmup 5 --no-cpu > gpurun_out/s3_bench_k200.json 2>/dev/null
MC3B_NO_USIG=1 python bench.py --steps 200 --warmup 5 --no-cpu > gpurun_out/s3_bench_nousig.json 2>/dev/null
MC3B_NO_USIG=1 MC3B_OLD_GRID=1 python bench.py --steps 200 --warmup 5 --no-cpu > gpurun_out/s3_bench_nousig_oldgrid.json 2>/dev/null
python profiles/e2e_breakdown.py 20 > gpurun_out/s3_e2e_brk20.log 2>&1
python - <<'PY' >> gpurun_out/s3_summary.txt
import json, glob
for f in sorted(glob.glob('gpurun_out/s3_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d['roofline']
        print(f, 'value %.3e' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'kernel ms %.4f' % r['ms_per_launch'],
              'frac %.3f' % r['frac'], 'e2e %.3e' % d['e2e']['value'], 'launches', d['gpu_launches'], d['clocks'])
    except Exception as e:
        print(f, 'ERR', e)
PY
cat gpurun_out/s3_summary.txt
tail -15 gpurun_out/s3_pytest.log
