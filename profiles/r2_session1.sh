#!/bin/bash
# round-2 GPU session 1: correctness of the new paths, A/B of kernel variants, FP64 pipe probe, ncu
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/s1_smi.txt
timeout 900 python -m pytest tests/test_gpu_r2.py -x -q > gpurun_out/s1_pytest_r2.log 2>&1; echo "r2 tests rc=$?" >> gpurun_out/s1_summary.txt
timeout 1200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_mcmc.py tests/test_gpu_api.py -x -q > gpurun_out/s1_pytest_old.log 2>&1; echo "old tests rc=$?" >> gpurun_out/s1_summary.txt
./profiles/probes/fp64_pipes > gpurun_out/s1_fp64_pipes.jsonl 2>&1
python bench.py --steps 200 --warmup 5 --no-cpu > gpurun_out/s1_bench_default.json 2> gpurun_out/s1_bench_default.err
MC3B_NO_FUSE=1 python bench.py --steps 200 --warmup 5 --no-cpu > gpurun_out/s1_bench_nofuse.json 2>/dev/null
MC3B_NO_USIG=1 python bench.py --steps 200 --warmup 5 --no-cpu > gpurun_out/s1_bench_nousig.json 2>/dev/null
for v in r1 r4 r2m5 r4m5 r2m4; do
  MC3B_LIBPATH=$PWD/variants/libmc3b200_$v.so python bench.py --steps 200 --warmup 5 --no-cpu > gpurun_out/s1_bench_$v.json 2>/dev/null
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sinegrid_usig -s 8 -c 1 -o gpurun_out/r2_sinegrid_usig python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/s1_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/s1_ncu_launches.log 2>&1
python - <<'PY' >> gpurun_out/s1_summary.txt
import json, glob
for f in sorted(glob.glob('gpurun_out/s1_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d['roofline']
        print(f, 'value %.3e' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'kernel ms %.4f' % r['ms_per_launch'],
              'frac %.3f' % r['frac'], 'e2e %.3e' % d['e2e']['value'], 'launches', d['gpu_launches'], d['clocks'])
    except Exception as e:
        print(f, 'ERR', e)
PY
cat gpurun_out/s1_summary.txt gpurun_out/s1_fp64_pipes.jsonl
tail -5 gpurun_out/s1_pytest_r2.log gpurun_out/s1_pytest_old.log
