#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -E "assert|FAILED|passed|failed|Error" | head -12 > gpurun_out/pytest_gpu.txt
timeout 600 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
timeout 300 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 6 -c 30 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 6 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1
VMB_GRAPHS=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step_umma -s 4 -c 1 -o gpurun_out/prof_umma_r01 python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
TRACE_PER=56 VMB_LIB=$PWD/vmap_b200/libvmap_b200_trace.so timeout 300 python tools/trace_umma.py > gpurun_out/trace.txt 2>&1
cat gpurun_out/pytest_gpu.txt; cut -c1-300 gpurun_out/bench_ours.json; python -c "import json; d=json.load(open('gpurun_out/bench_ours.json')); print(d['e2e'], d['roofline']['kernel_us'], d['clocks'])"
