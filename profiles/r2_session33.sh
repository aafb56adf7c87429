#!/bin/bash
# 1 GPU: final ncu evidence: --set full of the generation kernel (FMA form) and of k_sinemma, launch list of the bench command
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/s33_* gpurun_out/r2c_*
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sinefold -s 12 -c 1 -o gpurun_out/r2c_sinemom python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/s33_ncu.log 2>&1
MC3B_MOM_LAYOUT=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sinemma -s 12 -c 1 -o gpurun_out/r2c_sinemma python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/s33_ncu_mma.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/s33_ncu_launches.log 2>&1
ls -la gpurun_out/r2c_* > gpurun_out/s33_summary.txt
cat gpurun_out/s33_summary.txt
