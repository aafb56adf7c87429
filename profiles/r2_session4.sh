#!/bin/bash
# 2 GPUs: multi-GPU tests at real sizes, N=2 bench with the peer-memory exchange and with NCCL
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/s4_*
nvidia-smi topo -m > gpurun_out/s4_topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_r2.py -x -q -k "hpd or fused or population" > gpurun_out/s4_pytest_r2.log 2>&1; echo "r2 tests rc=$?" >> gpurun_out/s4_summary.txt
timeout 1500 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/s4_pytest_multi.log 2>&1; echo "multi tests rc=$?" >> gpurun_out/s4_summary.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 5 > gpurun_out/s4_bench_n2.json 2> gpurun_out/s4_bench_n2.err; echo "bench n2 rc=$?" >> gpurun_out/s4_summary.txt
MC3B_P2P=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 200 --warmup 5 > gpurun_out/s4_bench_n2_nccl.json 2> gpurun_out/s4_bench_n2_nccl.err; echo "bench n2 nccl rc=$?" >> gpurun_out/s4_summary.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/s4_bench_n2_k20.json 2> gpurun_out/s4_bench_n2_k20.err; echo "bench n2 k20 rc=$?" >> gpurun_out/s4_summary.txt
python - <<'PY' >> gpurun_out/s4_summary.txt
import json, glob
for f in sorted(glob.glob('gpurun_out/s4_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.3e' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'e2e %.3e' % d['e2e']['value'],
              'launches', d['gpu_launches'], d.get('multi_gpu_parity'))
    except Exception as e:
        print(f, 'ERR', e)
PY
cat gpurun_out/s4_summary.txt
tail -25 gpurun_out/s4_pytest_multi.log
tail -5 gpurun_out/s4_pytest_r2.log gpurun_out/s4_bench_n2.err
