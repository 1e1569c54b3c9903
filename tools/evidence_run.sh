#!/bin/bash
# One pass over the round's evidence on a 1-GPU box (run through gpurun); outputs under gpurun_out/ev_*.
set -x
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/ev_gpu_tests.log 2>&1; tail -2 gpurun_out/ev_gpu_tests.log
timeout 300 python bench.py > gpurun_out/ev_bench_n1.json 2> gpurun_out/ev_bench_n1.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 72 -c 48 --csv --log-file gpurun_out/ev_launches.csv python tools/prof_step.py 6 > gpurun_out/ev_launches.log 2>&1
timeout 800 ncu --set full --clock-control none --import-source on -s 72 -c 24 -f -o gpurun_out/ev_full python tools/prof_step.py 4 > gpurun_out/ev_full.log 2>&1
timeout 400 python tools/run_configs.py > gpurun_out/ev_configs.jsonl 2> gpurun_out/ev_configs.err
timeout 300 python tools/cpu_baselines.py > gpurun_out/ev_cpu_baselines.json 2> gpurun_out/ev_cpu_baselines.err
timeout 200 python tools/lanes_probe.py > gpurun_out/ev_lanes_probe.json 2>&1
ls -la gpurun_out/ev_*
