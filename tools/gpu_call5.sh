#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/pytest_gpu.txt
timeout 600 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step_umma -s 5 -c 1 -o gpurun_out/prof_umma_v3 python bench.py --steps 8 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
tail -15 gpurun_out/pytest_gpu.txt; cat gpurun_out/bench_ours.json; tail -5 gpurun_out/bench_ours.err
