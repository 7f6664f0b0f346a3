#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_layerwise_gpu.py -m gpu -q -s 2>&1 | grep -E "passed|failed|FAILED|assert|PSNR|iMAP step|GEMM" | tail -20 > gpurun_out/lw.txt
timeout 200 python tools/lw_time.py >> gpurun_out/lw.txt 2>&1
H=128 R=1200 S=14 timeout 200 python tools/lw_time.py >> gpurun_out/lw.txt 2>&1
cat gpurun_out/lw.txt | cut -c1-300
