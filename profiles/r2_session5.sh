#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/s5_*
python profiles/gen_breakdown.py > gpurun_out/s5_gen_breakdown.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_r2.py tests/test_gpu_mcmc.py -x -q > gpurun_out/s5_pytest.log 2>&1; echo "tests rc=$?" >> gpurun_out/s5_summary.txt
python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/s5_bench_k20.json 2>/dev/null
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/s5_bench_reference.json 2> gpurun_out/s5_bench_reference.err
python profiles/model_survey.py > gpurun_out/s5_model_survey.jsonl 2> gpurun_out/s5_model_survey.err
python bench_configs.py config4 > gpurun_out/s5_config4.json 2> gpurun_out/s5_config4.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/s5_config4_launches.csv python bench_configs.py config4 --n 100000000 > /dev/null 2>&1
python bench_configs.py config3 --steps 10 > gpurun_out/s5_config3.json 2> gpurun_out/s5_config3.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/s5_config3_launches.csv python bench_configs.py config3 --steps 3 --chains 4096 > /dev/null 2>&1
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_r2.py -x -q -k "grid_usig_with_data or fused_metropolis_equals_separate_kernels or hpd" > gpurun_out/s5_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/s5_summary.txt
timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_r2.py -x -q -k "test_population_uses_uniform_sigma_and_grid or test_uniform_sigma_fp32" > gpurun_out/s5_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/s5_summary.txt
cat gpurun_out/s5_summary.txt gpurun_out/s5_gen_breakdown.txt
tail -3 gpurun_out/s5_pytest.log gpurun_out/s5_sanitizer_memcheck.log gpurun_out/s5_sanitizer_racecheck.log
cat gpurun_out/s5_bench_reference.json | head -c 600
