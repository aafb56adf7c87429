#!/bin/bash
# 2 GPUs: validate the single-fence exchange before the 8-GPU run
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/s9_*
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -k "chain_partition" > gpurun_out/s9_pytest_multi.log 2>&1; echo "multi tests rc=$?" >> gpurun_out/s9_summary.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 5 > gpurun_out/s9_bench_n2.json 2> gpurun_out/s9_bench_n2.err; echo "bench n2 rc=$?" >> gpurun_out/s9_summary.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench_configs.py config5 --shard chains --scaling strong --n5 8000000 --steps 10 > gpurun_out/s9_config5_n2_strong.json 2> gpurun_out/s9_config5_n2_strong.err; echo "config5 strong rc=$?" >> gpurun_out/s9_summary.txt
python - <<'PY' >> gpurun_out/s9_summary.txt
import json, glob
for f in sorted(glob.glob('gpurun_out/s9_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.3e' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'e2e %.3e' % d['e2e']['value'], d.get('multi_gpu_parity', {}).get('bitwise_equal_to_1gpu'))
    except Exception as e: print(f, 'ERR', e)
for f in sorted(glob.glob('gpurun_out/s9_config*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(f, d.get('value'), [(r['sampler'], r['ms_per_step']) for r in d.get('runs', [])])
    except Exception as e: print(f, 'ERR', e)
PY
cat gpurun_out/s9_summary.txt; tail -3 gpurun_out/s9_pytest_multi.log
