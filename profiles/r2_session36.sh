#!/bin/bash
# 4 GPUs: weak-scaling point of the final build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/s36_*
timeout 300 python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 4 --master-port 29591 bench.py --gpus 4 --steps 200 --warmup 5 > gpurun_out/s36_bench_n4.json 2> gpurun_out/s36_bench_n4.err; echo "bench n4 rc=$?"
python -c "
import json
d = json.loads(open('gpurun_out/s36_bench_n4.json').read().strip().splitlines()[-1])
print('n4: value %.4e' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'e2e %.4e' % d['e2e']['value'], d['multi_gpu_parity']['bitwise_equal_to_1gpu'])"
