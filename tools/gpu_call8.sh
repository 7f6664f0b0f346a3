#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/pytest_gpu.txt
VMB_LIB=$PWD/vmap_b200/libvmap_b200_trace.so timeout 300 python tools/trace_umma.py > gpurun_out/trace.txt 2>&1
timeout 600 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step_umma -s 1 -c 1 -o gpurun_out/prof_umma_v4 python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
tail -5 gpurun_out/pytest_gpu.txt; grep -E "tile 1|coarse|issuer 0" gpurun_out/trace.txt | cut -c1-400; cat gpurun_out/bench_ours.json | cut -c1-1500; tail -3 gpurun_out/bench_ours.err
