#!/bin/bash
# 8 GPUs, final code of the round (pair + moment kernels, relaxed flag publish, staged peer pointers): bench at N=8 and N=4,
# config 5 (65536 chains, N=1e6) weak and strong over chains
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/s18_*
run() { name=$1; shift; timeout 400 "$@" > gpurun_out/s18_$name.json 2> gpurun_out/s18_$name.err; echo "$name rc=$?" >> gpurun_out/s18_summary.txt; }
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run bench_n8 $TR --nproc-per-node 8 --master-port 29551 bench.py --gpus 8 --steps 200 --warmup 5
run bench_n4 $TR --nproc-per-node 4 --master-port 29552 bench.py --gpus 4 --steps 200 --warmup 5
run bench_n8_k20 $TR --nproc-per-node 8 --master-port 29553 bench.py --gpus 8 --steps 20 --warmup 5
run config5_n8_weak $TR --nproc-per-node 8 --master-port 29554 bench_configs.py config5 --shard chains --scaling weak --steps 20
run config5_n8_strong $TR --nproc-per-node 8 --master-port 29555 bench_configs.py config5 --shard chains --scaling strong --steps 20
run config5_n8_data $TR --nproc-per-node 8 --master-port 29556 bench_configs.py config5 --shard data --scaling weak --steps 20
python - <<'PY' >> gpurun_out/s18_summary.txt
import json, glob
for f in sorted(glob.glob('gpurun_out/s18_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.3e' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'e2e %.3e' % d['e2e']['value'], d.get('multi_gpu_parity', {}).get('bitwise_equal_to_1gpu'))
    except Exception as e: print(f, 'ERR', e)
for f in sorted(glob.glob('gpurun_out/s18_config*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(f, d.get('value'), [(r['sampler'], r['ms_per_step']) for r in d.get('runs', [])])
    except Exception as e: print(f, 'ERR', e)
PY
cat gpurun_out/s18_summary.txt
