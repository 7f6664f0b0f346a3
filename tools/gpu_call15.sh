#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -E "assert|FAILED|passed|failed|Error" | head -20 > gpurun_out/pytest_gpu.txt
timeout 300 python tools/frame_time.py > gpurun_out/frame_time.json 2> gpurun_out/frame_time.err
timeout 600 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
cat gpurun_out/pytest_gpu.txt; cat gpurun_out/frame_time.json; tail -2 gpurun_out/frame_time.err; cut -c1-1700 gpurun_out/bench_ours.json
