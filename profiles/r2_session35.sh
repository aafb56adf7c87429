#!/bin/bash
# 2 GPUs: bench as the driver launches it, after the plan change
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/s35_*
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 400 $TR --nproc-per-node 2 --master-port 29581 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/s35_bench_n2.json 2> gpurun_out/s35_bench_n2.err; echo "bench n2 rc=$?" >> gpurun_out/s35_summary.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -k "data" > gpurun_out/s35_pytest_multi.log 2>&1; echo "multi data-shard test rc=$?" >> gpurun_out/s35_summary.txt
python - <<'PY' >> gpurun_out/s35_summary.txt
import json
d = json.loads(open('gpurun_out/s35_bench_n2.json').read().strip().splitlines()[-1])
print('n2: value %.4e' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'e2e %.4e' % d['e2e']['value'], d['multi_gpu_parity']['bitwise_equal_to_1gpu'], d['multi_gpu_parity']['oracle_rel_err'])
PY
cat gpurun_out/s35_summary.txt; tail -3 gpurun_out/s35_pytest_multi.log
