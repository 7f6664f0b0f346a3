#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x -s 2>&1 | grep -vE "^$|Warning|warn" | tail -25 > gpurun_out/gemm.txt
cat gpurun_out/gemm.txt
