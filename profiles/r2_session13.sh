#!/bin/bash
# 1 GPU: sufficient-statistics form of the grid kernel (k_sinefold<MOM>): parity + guard tests, bench A/B, ncu
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/s13_*
timeout 900 python -m pytest tests/test_gpu_moment.py -x -q > gpurun_out/s13_pytest_moment.log 2>&1; echo "moment tests rc=$?" >> gpurun_out/s13_summary.txt
ab() { name=$1; shift; env "$@" python bench.py --steps 200 --warmup 5 --no-cpu > gpurun_out/s13_bench_$name.json 2>gpurun_out/s13_bench_$name.err; }
ab default A=1
ab nomoment MC3B_NO_MOMENT=1
ab nopdl MC3B_FOLD_PDL=0
python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/s13_bench_k20.json 2>/dev/null
python profiles/gen_breakdown.py > gpurun_out/s13_gen_breakdown.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sinefold -s 12 -c 1 -o gpurun_out/r2_sinemom python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/s13_ncu.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s13_pytest.log 2>&1; echo "gpu tests rc=$?" >> gpurun_out/s13_summary.txt
python - <<'PY' >> gpurun_out/s13_summary.txt
import json, glob
for f in sorted(glob.glob('gpurun_out/s13_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); r = d['roofline']
        print(f, 'value %.3e' % d['value'], 'ms/step %.4f' % d['ms_per_step'], r['kernel'], 'kernel ms %.4f' % r['ms_per_launch'], 'frac %.3f' % r['frac'], 'pipe %.3f' % r['fp64_pipe_frac'], 'e2e %.3e' % d['e2e']['value'], 'hits', r.get('guard_hits'))
    except Exception as e: print(f, 'ERR', e)
PY
cat gpurun_out/s13_summary.txt gpurun_out/s13_gen_breakdown.txt
tail -25 gpurun_out/s13_pytest_moment.log; tail -8 gpurun_out/s13_pytest.log
