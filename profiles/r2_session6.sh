#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/s6_*
ls -la oracle/_ref oracle/_ref/mc3 > gpurun_out/s6_ls_ref.txt 2>&1
python profiles/gen_breakdown.py > gpurun_out/s6_gen_breakdown.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_r2.py tests/test_gpu_kernels.py tests/test_gpu_api.py -x -q > gpurun_out/s6_pytest.log 2>&1; echo "tests rc=$?" >> gpurun_out/s6_summary.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/s6_bench_k20.json 2>/dev/null
python bench.py --steps 200 --warmup 5 --no-cpu > gpurun_out/s6_bench_k200.json 2>/dev/null
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/s6_bench_reference.json 2> gpurun_out/s6_bench_reference.err
python bench_configs.py config4 > gpurun_out/s6_config4.json 2> gpurun_out/s6_config4.err
MC3B_TILE_CTAS=1 python bench_configs.py config4 > gpurun_out/s6_config4_cta1.json 2>/dev/null
for v in tt4096 tt4096w8 tt8192w8; do
  MC3B_LIBPATH=$PWD/variants/libmc3b200_$v.so python bench_configs.py config4 > gpurun_out/s6_config4_$v.json 2>/dev/null
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_binrms_tile -c 1 -o gpurun_out/r2_binrms_tile python bench_configs.py config4 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dwt_reg_model -s 2 -c 1 -o gpurun_out/r2_dwt_reg_model python bench_configs.py config3 --steps 2 --chains 4096 > /dev/null 2>&1
python - <<'PY' >> gpurun_out/s6_summary.txt
import json, glob
for f in sorted(glob.glob('gpurun_out/s6_bench_k*.json')):
    d = json.loads(open(f).read().strip().splitlines()[-1]); r = d['roofline']
    print(f, 'value %.3e' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'kernel ms %.4f' % r['ms_per_launch'], 'frac %.3f' % r['frac'], 'e2e %.3e' % d['e2e']['value'], d['e2e']['seconds_all_calls'])
for f in sorted(glob.glob('gpurun_out/s6_config4*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k: (round(d[k]['ms'], 4), round(d[k]['roofline']['frac'], 3)) for k in ('bin_array_unweighted', 'bin_array_weighted', 'time_avg')}, d['time_avg']['max_rel_err_vs_direct'])
    except Exception as e:
        print(f, 'ERR', e)
d = json.loads(open('gpurun_out/s6_bench_reference.json').read().strip().splitlines()[-1])
print('reference', d['value'], d['cpu_baseline']['kind'], d['cpu_baseline']['cores'])
PY
cat gpurun_out/s6_summary.txt gpurun_out/s6_gen_breakdown.txt gpurun_out/s6_ls_ref.txt
tail -3 gpurun_out/s6_pytest.log
