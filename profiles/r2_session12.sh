#!/bin/bash
# 1 GPU: k_sinefold with per-chain constants from k_fold_consts; A/B of launch-shape and residency variants
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/s12_*
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "fold or grid" > gpurun_out/s12_pytest_fold.log 2>&1; echo "fold tests rc=$?" >> gpurun_out/s12_summary.txt
ab() { name=$1; shift; env "$@" python bench.py --steps 200 --warmup 5 --no-cpu > gpurun_out/s12_bench_$name.json 2>gpurun_out/s12_bench_$name.err; }
ab default A=1
ab noconsts MC3B_NO_FOLD_CONSTS=1
ab pdl MC3B_FOLD_PDL=1
ab plan4 MC3B_PLAN_RESIDENT=4
ab plan4w1 MC3B_PLAN_RESIDENT=4 MC3B_WAVES=1
ab plan4w3 MC3B_PLAN_RESIDENT=4 MC3B_WAVES=3
ab minb5 MC3B_LIBPATH=$PWD/variants/libmc3b200_minb5.so
ab minb5plan5 MC3B_LIBPATH=$PWD/variants/libmc3b200_minb5.so MC3B_PLAN_RESIDENT=5
ab minb6 MC3B_LIBPATH=$PWD/variants/libmc3b200_minb6.so
ab restart8 MC3B_LIBPATH=$PWD/variants/libmc3b200_restart8.so
ab nstage6 MC3B_LIBPATH=$PWD/variants/libmc3b200_nstage6.so
python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/s12_bench_k20.json 2>/dev/null
python profiles/gen_breakdown.py > gpurun_out/s12_gen_breakdown.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s12_pytest.log 2>&1; echo "gpu tests rc=$?" >> gpurun_out/s12_summary.txt
python - <<'PY' >> gpurun_out/s12_summary.txt
import json, glob
for f in sorted(glob.glob('gpurun_out/s12_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); r = d['roofline']
        print(f, 'value %.3e' % d['value'], 'ms/step %.4f' % d['ms_per_step'], r['kernel'], 'kernel ms %.4f' % r['ms_per_launch'], 'frac %.3f' % r['frac'], 'pipe %.3f' % r['fp64_pipe_frac'], 'e2e %.3e' % d['e2e']['value'])
    except Exception as e: print(f, 'ERR', e)
PY
cat gpurun_out/s12_summary.txt gpurun_out/s12_gen_breakdown.txt
tail -15 gpurun_out/s12_pytest_fold.log; tail -8 gpurun_out/s12_pytest.log
