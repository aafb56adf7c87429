#!/bin/bash
# 1 GPU: unfused moment form + mc3b_moment_finish (initial population, chisq(), data shard); full suite; e2e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/s23_*
timeout 900 python -m pytest tests/test_gpu_moment.py -x -q > gpurun_out/s23_pytest_moment.log 2>&1; echo "moment tests rc=$?" >> gpurun_out/s23_summary.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "config4_full or fold or grid" > gpurun_out/s23_pytest_k.log 2>&1; echo "kernel tests rc=$?" >> gpurun_out/s23_summary.txt
python profiles/gapped_bench.py > gpurun_out/s23_gapped.json 2> gpurun_out/s23_gapped.err
python bench.py --steps 200 --warmup 5 --no-cpu > gpurun_out/s23_bench_k200.json 2>gpurun_out/s23_bench_k200.err; python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/s23_bench_k20.json 2>/dev/null; python -c "import json;d=json.loads(open('gpurun_out/s23_bench_k20.json').read().strip().splitlines()[-1]);print('K20 e2e %.4e' % d['e2e']['value'], d['e2e']['seconds_all_calls'])" >> gpurun_out/s23_summary.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s23_pytest.log 2>&1; echo "gpu tests rc=$?" >> gpurun_out/s23_summary.txt
cat gpurun_out/s23_summary.txt gpurun_out/s23_gapped.json; tail -3 gpurun_out/s23_gapped.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/s23_bench_k200.json').read().strip().splitlines()[-1]); r = d['roofline']
print('bench value %.3e' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'kernel ms %.4f' % r['ms_per_launch'], 'e2e %.3e' % d['e2e']['value'])
PY
tail -15 gpurun_out/s23_pytest_moment.log; tail -5 gpurun_out/s23_pytest_k.log; tail -5 gpurun_out/s23_pytest.log
