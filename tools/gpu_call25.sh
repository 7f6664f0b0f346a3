#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x -s 2>&1 | grep -E "passed|failed|FAILED|assert|Error|GEMM" | tail -25 > gpurun_out/ws.txt
cat gpurun_out/ws.txt
