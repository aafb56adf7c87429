#!/bin/bash
# 1 GPU: the mirrored-pair grid kernel (k_sinefold): parity tests, A/B against k_sinegrid, ncu
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/s11_*
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "fold or grid" > gpurun_out/s11_pytest_fold.log 2>&1; echo "fold tests rc=$?" >> gpurun_out/s11_summary.txt
python bench.py --steps 200 --warmup 5 --no-cpu > gpurun_out/s11_bench_k200.json 2>gpurun_out/s11_bench_k200.err
MC3B_NO_FOLD=1 python bench.py --steps 200 --warmup 5 --no-cpu > gpurun_out/s11_bench_k200_nofold.json 2>/dev/null
python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/s11_bench_k20.json 2>/dev/null
MC3B_WAVES=1 python bench.py --steps 200 --warmup 5 --no-cpu > gpurun_out/s11_bench_k200_waves1.json 2>/dev/null
MC3B_SCHED=2,8 python bench.py --steps 200 --warmup 5 --no-cpu > gpurun_out/s11_bench_k200_sched8.json 2>/dev/null
python profiles/gen_breakdown.py > gpurun_out/s11_gen_breakdown.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sinefold -s 8 -c 1 -o gpurun_out/r2_sinefold python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/s11_ncu.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s11_pytest.log 2>&1; echo "gpu tests rc=$?" >> gpurun_out/s11_summary.txt
python - <<'PY' >> gpurun_out/s11_summary.txt
import json, glob
for f in sorted(glob.glob('gpurun_out/s11_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); r = d['roofline']
        print(f, 'value %.3e' % d['value'], 'ms/step %.4f' % d['ms_per_step'], r['kernel'], 'kernel ms %.4f' % r['ms_per_launch'], 'frac %.3f' % r['frac'], 'pipe %.3f' % r['fp64_pipe_frac'], 'e2e %.3e' % d['e2e']['value'])
    except Exception as e: print(f, 'ERR', e)
PY
cat gpurun_out/s11_summary.txt gpurun_out/s11_gen_breakdown.txt
tail -15 gpurun_out/s11_pytest_fold.log; tail -5 gpurun_out/s11_pytest.log
