#!/bin/bash
# 2 GPUs: multi-GPU tests and bench with the relaxed flag publish and the moment form
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/s15_*
timeout 1500 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/s15_pytest_multi.log 2>&1; echo "multi tests rc=$?" >> gpurun_out/s15_summary.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run() { name=$1; shift; timeout 400 "$@" > gpurun_out/s15_$name.json 2> gpurun_out/s15_$name.err; echo "$name rc=$?" >> gpurun_out/s15_summary.txt; }
run bench_n2 $TR --nproc-per-node 2 --master-port 29531 bench.py --gpus 2 --steps 200 --warmup 5
MC3B_NO_MOMENT=1 run bench_n2_nomoment $TR --nproc-per-node 2 --master-port 29532 bench.py --gpus 2 --steps 200 --warmup 5
run bench_n1 python bench.py --steps 200 --warmup 5 --no-cpu
run config3_n2 $TR --nproc-per-node 2 --master-port 29533 bench_configs.py config3 --steps 10
python - <<'PY' >> gpurun_out/s15_summary.txt
import json, glob
for f in sorted(glob.glob('gpurun_out/s15_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.3e' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'e2e %.3e' % d['e2e']['value'], d.get('multi_gpu_parity'))
    except Exception as e: print(f, 'ERR', e)
PY
cat gpurun_out/s15_summary.txt; tail -5 gpurun_out/s15_pytest_multi.log
