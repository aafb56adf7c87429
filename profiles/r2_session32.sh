#!/bin/bash
# 2 GPUs: final build, bench as the driver launches it (both arms) + the 2-GPU chain-partition tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/s32_*
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 400 $TR --nproc-per-node 2 --master-port 29571 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/s32_bench_n2.json 2> gpurun_out/s32_bench_n2.err; echo "bench n2 rc=$?" >> gpurun_out/s32_summary.txt
timeout 400 $TR --nproc-per-node 2 --master-port 29572 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > gpurun_out/s32_bench_n2_ref.json 2> gpurun_out/s32_bench_n2_ref.err; echo "reference n2 rc=$?" >> gpurun_out/s32_summary.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -k "chain_partition or savefile" > gpurun_out/s32_pytest_multi.log 2>&1; echo "multi tests rc=$?" >> gpurun_out/s32_summary.txt
python - <<'PY' >> gpurun_out/s32_summary.txt
import json
d = json.loads(open('gpurun_out/s32_bench_n2.json').read().strip().splitlines()[-1])
print('n2: value %.4e' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'e2e %.4e' % d['e2e']['value'], d['multi_gpu_parity']['bitwise_equal_to_1gpu'])
f = json.loads(open('gpurun_out/s32_bench_n2_ref.json').read().strip().splitlines()[-1]); print('ref n2:', f.get('value'), f.get('unavailable'))
PY
cat gpurun_out/s32_summary.txt; tail -3 gpurun_out/s32_pytest_multi.log
