#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 12 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step_umma -s 5 -c 2 -o gpurun_out/prof_umma_r01 python bench.py --steps 8 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
cat gpurun_out/bench_ours.json; tail -3 gpurun_out/bench_ours.err; grep -E "k_step|k_adam|k_mask" gpurun_out/launches_r01.csv | head -12
