#!/bin/bash
# 2 GPUs: peer pointers staged in shared memory: multi-GPU tests + bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/s17_*
timeout 1500 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/s17_pytest_multi.log 2>&1; echo "multi tests rc=$?" >> gpurun_out/s17_summary.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run() { name=$1; shift; timeout 400 "$@" > gpurun_out/s17_$name.json 2> gpurun_out/s17_$name.err; echo "$name rc=$?" >> gpurun_out/s17_summary.txt; }
run bench_n2 $TR --nproc-per-node 2 --master-port 29541 bench.py --gpus 2 --steps 200 --warmup 5
run bench_n1 python bench.py --steps 200 --warmup 5 --no-cpu
python - <<'PY' >> gpurun_out/s17_summary.txt
import json, glob
for f in sorted(glob.glob('gpurun_out/s17_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.3e' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'e2e %.3e' % d['e2e']['value'], (d.get('multi_gpu_parity') or {}).get('bitwise_equal_to_1gpu'))
    except Exception as e: print(f, 'ERR', e)
PY
cat gpurun_out/s17_summary.txt; tail -5 gpurun_out/s17_pytest_multi.log
