#!/bin/bash
# 2 GPUs: multi-GPU tests again after the epilogue / proposal changes, N=2 bench with aligned steps
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/s8_*
timeout 1500 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/s8_pytest_multi.log 2>&1; echo "multi tests rc=$?" >> gpurun_out/s8_summary.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 5 > gpurun_out/s8_bench_n2.json 2> gpurun_out/s8_bench_n2.err; echo "bench n2 rc=$?" >> gpurun_out/s8_summary.txt
MC3B_P2P=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 200 --warmup 5 > gpurun_out/s8_bench_n2_nccl.json 2> gpurun_out/s8_bench_n2_nccl.err; echo "bench n2 nccl rc=$?" >> gpurun_out/s8_summary.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/s8_bench_n2_k20.json 2> gpurun_out/s8_bench_n2_k20.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench_configs.py config3 --steps 10 > gpurun_out/s8_config3_n2.json 2> gpurun_out/s8_config3_n2.err; echo "config3 n2 rc=$?" >> gpurun_out/s8_summary.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 bench_configs.py config5 --shard chains --scaling weak --steps 20 > gpurun_out/s8_config5_n2_chains.json 2> gpurun_out/s8_config5_n2_chains.err; echo "config5 chains rc=$?" >> gpurun_out/s8_summary.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29516 bench_configs.py config5 --shard data --scaling weak --steps 20 > gpurun_out/s8_config5_n2_data.json 2> gpurun_out/s8_config5_n2_data.err; echo "config5 data rc=$?" >> gpurun_out/s8_summary.txt
python - <<'PY' >> gpurun_out/s8_summary.txt
import json, glob
for f in sorted(glob.glob('gpurun_out/s8_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.3e' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'e2e %.3e' % d['e2e']['value'], d.get('multi_gpu_parity'))
    except Exception as e: print(f, 'ERR', e)
for f in sorted(glob.glob('gpurun_out/s8_config*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(f, d.get('value'), d.get('ms_per_step'), d.get('runs'), d.get('roofline'))
    except Exception as e: print(f, 'ERR', e)
PY
cat gpurun_out/s8_summary.txt
tail -5 gpurun_out/s8_pytest_multi.log; tail -3 gpurun_out/s8_config5_n2_chains.err gpurun_out/s8_config3_n2.err
