#!/bin/bash
# 1 GPU: what the driver runs at round end -- smoke(), pytest -m gpu, bench.py (both arms) with its own flags
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/s31_*
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s31_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/s31_summary.txt
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/s31_pytest.log 2>&1; echo "gpu tests rc=$?" >> gpurun_out/s31_summary.txt
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/s31_bench_reference.json 2> gpurun_out/s31_bench_reference.err; echo "reference arm rc=$?" >> gpurun_out/s31_summary.txt
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/s31_bench_k20.json 2> gpurun_out/s31_bench_k20.err; echo "bench rc=$?" >> gpurun_out/s31_summary.txt
python - <<'PY' >> gpurun_out/s31_summary.txt
import json
d = json.loads(open('gpurun_out/s31_bench_k20.json').read().strip().splitlines()[-1]); r = d['roofline']
print('ours: value %.4e' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'e2e %.4e' % d['e2e']['value'], r['kernel'], 'frac %.3f' % r['frac'], 'launches', d['gpu_launches'], d['clocks'])
print('cpu_baseline', d['cpu_baseline'])
f = json.loads(open('gpurun_out/s31_bench_reference.json').read().strip().splitlines()[-1])
print('reference: value %.4e' % f['value'], f['cpu_baseline']['cores'], 'cores')
print('ratio %.0f' % (d['value']/f['value']), 'e2e ratio %.0f' % (d['e2e']['value']/f['e2e']['value']))
PY
cat gpurun_out/s31_summary.txt; tail -3 gpurun_out/s31_smoke.log; tail -4 gpurun_out/s31_pytest.log
