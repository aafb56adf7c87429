#!/bin/bash
# 2 GPUs: final code (unfused moment form in the data shard, tile origins): multi-GPU tests + bench + config 5 data shard
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/s24_*
timeout 1500 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/s24_pytest_multi.log 2>&1; echo "multi tests rc=$?" >> gpurun_out/s24_summary.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run() { name=$1; shift; timeout 400 "$@" > gpurun_out/s24_$name.json 2> gpurun_out/s24_$name.err; echo "$name rc=$?" >> gpurun_out/s24_summary.txt; }
run bench_n2 $TR --nproc-per-node 2 --master-port 29561 bench.py --gpus 2 --steps 200 --warmup 5
run config5_n2_data $TR --nproc-per-node 2 --master-port 29563 bench_configs.py config5 --shard data --scaling weak --steps 20
run config5_n2_weak $TR --nproc-per-node 2 --master-port 29564 bench_configs.py config5 --shard chains --scaling weak --steps 20
python - <<'PY' >> gpurun_out/s24_summary.txt
import json, glob
for f in sorted(glob.glob('gpurun_out/s24_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.3e' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'e2e %.3e' % d['e2e']['value'], (d.get('multi_gpu_parity') or {}).get('bitwise_equal_to_1gpu'))
    except Exception as e: print(f, 'ERR', e)
PY
python - <<'PY' >> gpurun_out/s24_summary.txt
import json, glob
for f in sorted(glob.glob('gpurun_out/s24_config*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(f, d.get('value'), [(r['sampler'], r['ms_per_step'], r['gelman_rubin_max']) for r in d.get('runs', [])], d['roofline']['kernel'], d['roofline']['ms_per_launch'])
    except Exception as e: print(f, 'ERR', e)
PY
cat gpurun_out/s24_summary.txt; tail -5 gpurun_out/s24_pytest_multi.log
