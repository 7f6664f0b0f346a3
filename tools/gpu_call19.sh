#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_layerwise_gpu.py tests/test_gemm_gpu.py -m gpu -q -s 2>&1 | grep -vE "^$|Warning|warn" | grep -E "passed|failed|FAILED|assert|PSNR|iMAP step|GEMM|render rel" | tail -30 > gpurun_out/lw.txt
cat gpurun_out/lw.txt | cut -c1-400
