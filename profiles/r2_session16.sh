#!/bin/bash
# 1 GPU: final single-GPU evidence of the round: tests, sanitizer on the new kernels, fp32 line + ncu, launch list, ncu of the moment kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/s16_*
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s16_pytest.log 2>&1; echo "gpu tests rc=$?" >> gpurun_out/s16_summary.txt
python bench.py --steps 200 --warmup 5 --no-cpu > gpurun_out/s16_bench_k200.json 2>gpurun_out/s16_bench_k200.err
python bench.py --steps 20 --warmup 5 > gpurun_out/s16_bench_k20.json 2>gpurun_out/s16_bench_k20.err
python bench.py --dtype f32 --steps 200 --warmup 5 --no-cpu > gpurun_out/s16_bench_f32.json 2>gpurun_out/s16_bench_f32.err
python profiles/gen_breakdown.py > gpurun_out/s16_gen_breakdown.txt 2>&1
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_moment.py tests/test_gpu_kernels.py -x -q -k "guard or moment_kernel_matches_oracle or folded_kernel or fold_data" > gpurun_out/s16_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/s16_summary.txt
timeout 1200 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_moment.py tests/test_gpu_kernels.py -x -q -k "guard or 4096-0.5 or 3001-2.0" > gpurun_out/s16_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/s16_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sinefold -s 12 -c 1 -o gpurun_out/r2_sinemom_final python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/s16_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_model_chisq -s 8 -c 1 -o gpurun_out/r2_model_chisq_f32 python bench.py --dtype f32 --steps 3 --warmup 3 --no-cpu > gpurun_out/s16_ncu_f32.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/s16_ncu_launches.log 2>&1
python - <<'PY' >> gpurun_out/s16_summary.txt
import json, glob
for f in sorted(glob.glob('gpurun_out/s16_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); r = d['roofline']
        print(f, 'value %.3e' % d['value'], 'ms/step %.4f' % d['ms_per_step'], r['kernel'], 'kernel ms %.4f' % r['ms_per_launch'], 'frac %.3f' % r['frac'], 'e2e %.3e' % d['e2e']['value'], 'hits', r.get('guard_hits'), 'cpu', (d.get('cpu_baseline') or {}).get('value'))
    except Exception as e: print(f, 'ERR', e)
PY
cat gpurun_out/s16_summary.txt gpurun_out/s16_gen_breakdown.txt
tail -4 gpurun_out/s16_pytest.log; tail -n 4 gpurun_out/s16_sanitizer_memcheck.log gpurun_out/s16_sanitizer_racecheck.log
