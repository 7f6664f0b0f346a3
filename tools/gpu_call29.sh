#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_umma_gpu.py tests/test_step_gpu.py tests/test_ingest_gpu.py -m gpu -q -x -s 2>&1 | grep -E "passed|failed|FAILED|assert|Error|eval_points" | tail -15 > gpurun_out/ev.txt
timeout 300 python tools/ingest_time.py > gpurun_out/ingest_time.json 2> gpurun_out/ingest_time.err
cat gpurun_out/ev.txt; cat gpurun_out/ingest_time.json; tail -3 gpurun_out/ingest_time.err
